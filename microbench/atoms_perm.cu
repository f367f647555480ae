// Microbenchmark: which lane -> bank PERMUTATIONS of a conflict-free shared-memory ATOMS run at full rate on sm_100a?
// Every pattern below touches 32 distinct banks per instruction; they differ only in the permutation and in whether it
// changes from step to step.  addr = bin*128 + 4*bank(lane, step); bins from a 16 B register chunk via PRMT (like hist.cu).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o atoms_perm atoms_perm.cu ; run under ncu for the conflict counters.
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("CUDA error %s at %d\n",cudaGetErrorString(e),__LINE__); return 1;}}while(0)

__host__ __device__ inline int bitrev5(int x) { int r = 0; for (int i = 0; i < 5; ++i) r |= ((x >> i) & 1) << (4 - i); return r; }

__host__ __device__ inline int bank_of(int pat, int lane, int step) {
  switch (pat) {
    case 0: return lane;                                              // identity
    case 1: return (lane + step) & 31;                                // rotation, changes every step
    case 2: return lane ^ step;                                       // xor with the step (step < 16)
    case 3: return lane ^ (step * 2 + 1);                             // xor with an odd constant per step (touches bit 0 and bit 4)
    case 4: return 16 * (lane & 1) + ((lane >> 1) ^ step);            // hist.cu's lane = 2*row + half, xor schedule
    case 5: { const int row = lane >> 1, half = lane & 1, qw = row >> 2, qb = row & 3;      // hist.cu today
              return 16 * half + 4 * (((step >> 2) + qw) & 3) + (((step & 3) + qb) & 3); }
    case 6: return 16 * (lane >> 4) + (((lane & 15) + step) & 15);    // lanes in natural order, rotation inside each half
    case 7: return lane ^ 5;                                          // constant xor
    case 8: return (lane * 5) & 31;                                   // constant multiplicative permutation
    case 9: return bitrev5(lane);                                     // constant bit reversal
    case 10: return 16 * (lane & 1) + (lane >> 1);                    // constant: hist.cu's (row, half) bit move
    case 11: return (lane + 1) & 31;                                  // constant rotation by one
    case 12: return (lane & 24) | ((lane + step) & 7);                // rotation inside groups of 8 lanes
    case 13: return (lane & 28) | ((lane + step) & 3);                // rotation inside groups of 4 lanes
    case 14: return lane ^ (step & 3);                                // xor inside groups of 4
    case 15: return lane ^ ((step & 3) << 3);                         // xor of the high bits only
  }
  return lane;
}

template <int NA>
__global__ void __launch_bounds__(1024) kern(int pat, int iters, const uint4* __restrict__ src, unsigned long long* sink, long long* cyc) {
  extern __shared__ __align__(16) int smem[];      // G[256][32] then H[256][32]
  for (int i = threadIdx.x; i < 7 * 8192; i += blockDim.x) smem[i] = 0;
  __syncthreads();
  const unsigned base = (unsigned)__cvta_generic_to_shared(smem);
  const int lane = threadIdx.x & 31;
  if (pat == 24 && blockIdx.x == 0 && threadIdx.x == 0 && iters < 1000) printf("dynamic shared memory starts at shared address 0x%x\n", base);
  unsigned A[16], S[4];
  for (int jb = 0; jb < 4; ++jb) S[jb] = 0x4440u | jb;
  // patterns >= 16: bank-neutral PLANE offsets per lane (do the lanes of one ATOMS have to stay inside one region of shared memory?)
  unsigned plane = 0; int bp = pat;
  if (pat == 16) { plane = (unsigned)(lane % 3) * 65536u; bp = 0; }            // identity banks, three planes 64 KB apart
  if (pat == 17) { plane = (unsigned)(lane & 1) * 65536u; bp = 0; }            // identity banks, two planes 64 KB apart
  if (pat == 18) { plane = (unsigned)(lane & 1) * 32768u; bp = 0; }            // identity banks, two planes 32 KB apart
  if (pat == 19) { plane = (unsigned)(lane & 1) * 8192u; bp = 0; }             // identity banks, two planes 8 KB apart
  if (pat == 20) { plane = (unsigned)((lane % 6) >> 1) * 65536u; bp = 5; }     // hist_gather_kernel<3>: lane = 6q + c, group c >> 1
  if (pat == 21) { bp = 5; }                                                   // hist.cu table + rotated PRMT selectors (bins from rotated bytes)
  if (pat == 22) { plane = (unsigned)(lane >> 4) * 65536u; bp = 0; }           // identity banks, half-warps in different planes
  if (pat == 23) { plane = (unsigned)(lane % 3) * 98304u; bp = 0; }            // three planes 96 KB apart... wraps: uses 64 KB * 3 region
  if (pat == 21) for (int jb = 0; jb < 4; ++jb) S[jb] = 0x4440u | (unsigned)((jb + ((lane >> 1) & 3)) & 3);
  if (pat == 24) { plane = 136u * 1024u; bp = 0; }                            // every lane above 128 KB, identity banks
  if (pat == 25) { plane = 136u * 1024u; bp = 5; }                            // every lane above 128 KB, hist.cu table
  if (pat == 26) { plane = 0x20000u - base; bp = 5; }                          // G below / H (+32 KB) ... plane starts exactly at absolute 128 KB
  if (pat == 27) { plane = 0x20000u - base - 128u; bp = 5; }                   // starts one bin below the boundary
  if (pat == 23) plane = (unsigned)(lane % 3) * 40960u;                        // three planes 40 KB apart
  for (int st = 0; st < 16; ++st) A[st] = base + plane + 4u * (unsigned)bank_of(bp, lane, st);
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
    uint4 nw = p[(size_t)((it + 1) & 63) * stride];
    const unsigned ww[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
    for (int jw = 0; jw < 4; ++jw)
#pragma unroll
      for (int jb = 0; jb < 4; ++jb) {
        const unsigned bin = __byte_perm(ww[jw], 0u, S[jb]);
        const unsigned addr = (bin << 7) + A[4 * jw + jb];
        asm volatile("red.shared.add.s32 [%0], %1;" :: "r"(addr), "r"(g) : "memory");
        if (NA == 2) asm volatile("red.shared.add.u32 [%0+32768], %1;" :: "r"(addr), "r"(h) : "memory");
      }
    w = nw;
  }
  long long t1 = clock64();
  __syncthreads();
  unsigned long long acc = 0;
  for (int i = threadIdx.x; i < 7 * 8192; i += blockDim.x) acc += (unsigned)smem[i];
  if (acc == 0xdeadbeefULL) sink[0] = acc;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int NA> int run(int pat, int threads, const uint4* src) {
  if (pat < 16) for (int st = 0; st < 16; ++st) { unsigned m = 0; for (int l = 0; l < 32; ++l) m |= 1u << bank_of(pat, l, st); if (m != 0xffffffffu) { printf("pattern %d is NOT conflict free at step %d\n", pat, st); return 1; } }
  int nsm = 148; cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, 0);
  const int iters = 2000; size_t smem = 7 * 8192 * 4;
  CK(cudaFuncSetAttribute(kern<NA>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  unsigned long long* sink; long long* cyc; CK(cudaMalloc(&sink, 8)); CK(cudaMalloc(&cyc, 8 * nsm));
  if (!getenv("ATOMS_NO_WARMUP")) { kern<NA><<<nsm, threads, smem>>>(pat, iters / 4, src, sink, cyc); CK(cudaDeviceSynchronize()); }
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0); kern<NA><<<nsm, threads, smem>>>(pat, iters, src, sink, cyc); cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  long long hc[256]; CK(cudaMemcpy(hc, cyc, 8 * nsm, cudaMemcpyDeviceToHost));
  double avg = 0; for (int i = 0; i < nsm; i++) avg += hc[i]; avg /= nsm;
  const double atoms_instr = (double)(threads / 32) * iters * 16 * NA;
  printf("pat %2d  %dxATOMS thr=%4d  %7.3f ms  %6.3f ATOMS/clk/SM\n", pat, NA, threads, ms, atoms_instr / avg);
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
  const int p0 = getenv("ATOMS_PAT0") ? atoi(getenv("ATOMS_PAT0")) : 0;
  for (int pat = p0; pat < 28; ++pat) { run<1>(pat, 768, src); run<2>(pat, 768, src); }
  return 0;
}
