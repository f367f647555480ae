// Microbenchmark: does the shared-memory atomic rate on sm_100a depend on the lane -> bank PATTERN of a conflict-free ATOMS?
//   P0  fixed      addr = bin*128 + 4*lane                          (lane == bank for ever; r1's m0/m5 pattern)
//   P1  rotating   addr = bin*128 + 4*((lane + step) & 31)          (a different conflict-free permutation every step)
//   P2  hist.cu    addr = bin*128 + A[step] (LaneConst table: two lanes per row, per-row rotation)
// each with 1 ATOMS per step (G only) or 2 (G, H at +32 KB).  Bins come from a 16 B register chunk via PRMT like the real loop.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o atoms_pattern atoms_pattern.cu
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("CUDA error %s at %d\n",cudaGetErrorString(e),__LINE__); return 1;}}while(0)

template <int PAT, int NA, int HOFF = 32768>
__global__ void __launch_bounds__(1024) kern(int iters, const uint4* __restrict__ src, unsigned long long* sink, long long* cyc) {
  extern __shared__ __align__(16) int smem[];      // G[256][32] then H[256][32] (x3 groups for PAT 3)
  constexpr int NGRP = PAT == 3 ? 3 : 1;
  for (int i = threadIdx.x; i < NGRP * 2 * 8192 + 64; i += blockDim.x) smem[i] = 0;
  __syncthreads();
  const unsigned base = (unsigned)__cvta_generic_to_shared(smem);
  const int lane = threadIdx.x & 31;
  unsigned A[16], S[4];
  int rot = lane >> 1, half = lane & 1; unsigned gbase = 0;
  if (PAT == 3) { const int q = lane / 6, c = lane % 6; rot = q < 5 ? 3 * q + (c >> 1) : 15; half = c & 1; gbase = q < 5 ? (unsigned)(c >> 1) * 65536u : 0u; }
  const int qw = rot >> 2, qb = rot & 3;
  for (int jb = 0; jb < 4; ++jb) S[jb] = PAT >= 2 ? (0x4440u | ((jb + qb) & 3)) : (0x4440u | jb);
  for (int jw = 0; jw < 4; ++jw) for (int jb = 0; jb < 4; ++jb) {
    const int step = 4 * jw + jb;
    if (PAT == 0) A[step] = base + 4u * lane;
    else if (PAT == 1) A[step] = base + 4u * ((lane + step) & 31);
    else A[step] = base + gbase + 64u * half + 16u * ((jw + qw) & 3) + 4u * ((jb + qb) & 3);
  }
#pragma unroll
  for (int i = 0; i < 16; ++i) asm volatile("" : "+r"(A[i]));
#pragma unroll
  for (int i = 0; i < 4; ++i) asm volatile("" : "+r"(S[i]));
  const uint4* p = src + (size_t)(blockIdx.x * blockDim.x + threadIdx.x);
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  int g = lane * 3 + 1; unsigned h = lane + 2;
  uint4 w = p[0];
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    uint4 nw = p[(size_t)((it + 1) & 63) * stride];          // L2-resident stream: the atomics are what is measured
    const unsigned ww[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
    for (int jw = 0; jw < 4; ++jw)
#pragma unroll
      for (int jb = 0; jb < 4; ++jb) {
        const unsigned bin = __byte_perm(ww[jw], 0u, S[jb]);
        const unsigned addr = (bin << 7) + A[4 * jw + jb];
        asm volatile("red.shared.add.s32 [%0], %1;" :: "r"(addr), "r"(g) : "memory");
        if (NA == 2) asm volatile("red.shared.add.u32 [%0+%2], %1;" :: "r"(addr), "r"(h), "n"(HOFF) : "memory");
      }
    w = nw;
  }
  long long t1 = clock64();
  __syncthreads();
  unsigned long long acc = 0;
  for (int i = threadIdx.x; i < NGRP * 2 * 8192; i += blockDim.x) acc += (unsigned)smem[i];
  if (acc == 0xdeadbeefULL) sink[0] = acc;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int PAT, int NA, int HOFF = 32768> int run(const char* name, int threads, const uint4* src) {
  int nsm = 148; cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, 0);
  const int iters = 4000; size_t smem = (PAT == 3 ? 3 : 1) * 2 * 8192 * 4 + 256;
  CK(cudaFuncSetAttribute(kern<PAT, NA, HOFF>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  unsigned long long* sink; long long* cyc; CK(cudaMalloc(&sink, 8)); CK(cudaMalloc(&cyc, 8 * nsm));
  kern<PAT, NA, HOFF><<<nsm, threads, smem>>>(iters / 4, src, sink, cyc); CK(cudaDeviceSynchronize());
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0); kern<PAT, NA, HOFF><<<nsm, threads, smem>>>(iters, src, sink, cyc); cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  long long hc[256]; CK(cudaMemcpy(hc, cyc, 8 * nsm, cudaMemcpyDeviceToHost));
  double avg = 0; for (int i = 0; i < nsm; i++) avg += hc[i]; avg /= nsm;
  const double atoms_instr = (double)(threads / 32) * iters * 16 * NA;      // warp-level ATOMS per SM
  printf("%-34s thr=%4d  %7.3f ms  %6.3f ATOMS/clk/SM  %5.2f clk per ATOMS\n", name, threads, ms, atoms_instr / avg, avg / atoms_instr);
  cudaFree(sink); cudaFree(cyc); return 0;
}

int main() {
  const size_t n = (size_t)148 * 1024 * 64;
  uint4* src; CK(cudaMalloc(&src, n * sizeof(uint4)));
  {
    uint32_t* h = (uint32_t*)malloc(n * 16); uint32_t s = 12345u;
    for (size_t i = 0; i < n * 4; ++i) { s = s * 1664525u + 1013904223u; h[i] = s ^ (s >> 13); }
    CK(cudaMemcpy(src, h, n * 16, cudaMemcpyHostToDevice)); free(h);
  }
  if (getenv("ATOMS_PAIR_ONLY")) {          // pair variants only (short list for an ncu capture)
    run<2, 2>("P2 hist.cu lane table    2xATOMS", 768, src);
    run<2, 2, 32768 + 64>("P2 + H plane shifted 16 banks   ", 768, src);
    run<2, 2, 32768 + 4>("P2 + H plane shifted 1 bank     ", 768, src);
    run<3, 2>("P3 gather NG=3 mapping   2xATOMS", 768, src);
    run<3, 2, 32768 + 64>("P3 + H plane shifted 16 banks   ", 768, src);
    run<3, 1>("P3 gather NG=3 mapping   1xATOMS", 768, src);
    run<0, 2, 32768 + 64>("P0 + H plane shifted 16 banks   ", 768, src);
    return 0;
  }
  for (int thr : {256, 768, 1024}) {
    run<0, 1>("P0 fixed lane==bank      1xATOMS", thr, src);
    run<1, 1>("P1 rotating permutation  1xATOMS", thr, src);
    run<2, 1>("P2 hist.cu lane table    1xATOMS", thr, src);
    run<0, 2>("P0 fixed lane==bank      2xATOMS", thr, src);
    run<1, 2>("P1 rotating permutation  2xATOMS", thr, src);
    run<2, 2>("P2 hist.cu lane table    2xATOMS", thr, src);
  }
  return 0;
}
