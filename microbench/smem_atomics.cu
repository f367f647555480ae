// Microbenchmark: shared-memory accumulate throughput on sm_100a.
// Decides the histogram-kernel design (see DESIGN.md "why lane<->feature-slot").
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o smem_atomics smem_atomics.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("CUDA error %s at %d\n",cudaGetErrorString(e),__LINE__); return 1;}}while(0)

constexpr int NBIN = 256;
constexpr int SLOTS = 32;

__device__ __forceinline__ uint32_t lcg(uint32_t& s){ s = s*1664525u + 1013904223u; return s>>24; }

// mode 0: conflict-free 2x ATOMS.ADD.32 (lane == slot), bins from LCG
// mode 1: random-bank 2x ATOMS.ADD.32 (lane == row, layout [slot][bin])
// mode 2: conflict-free non-atomic 64-bit RMW (one warp owns the table -> each warp gets own table; needs small tables) 
// mode 3: conflict-free 2x ATOMS + LDS.64 broadcast of (g,h) per row
// mode 4: conflict-free 2x ATOMS + 2x SHFL broadcast
// mode 5: conflict-free 1x ATOMS (constant-hessian variant)
// mode 6: random-bank 1x ATOMS
// mode 7: conflict-free 2x ATOMS, bins read from smem byte tile (LDS.U8) + LDS.64 broadcast gpair
template<int MODE>
__global__ void __launch_bounds__(1024) kern(int iters, unsigned long long* sink, long long* cyc){
  extern __shared__ __align__(16) unsigned char smem_raw[];
  int* G = reinterpret_cast<int*>(smem_raw);                 // [NBIN][SLOTS]
  int* H = G + NBIN*SLOTS;                                   // [NBIN][SLOTS]
  int2* gp = reinterpret_cast<int2*>(H + NBIN*SLOTS);        // 256 rows of (g,h)
  unsigned char* tile = reinterpret_cast<unsigned char*>(gp + 256); // 256 rows x 32 B
  for(int i=threadIdx.x;i<2*NBIN*SLOTS;i+=blockDim.x) G[i]=0;
  for(int i=threadIdx.x;i<256;i+=blockDim.x) gp[i]=make_int2(i*7+1, i+3);
  for(int i=threadIdx.x;i<256*32;i+=blockDim.x) tile[i]=(unsigned char)((i*2654435761u)>>13);
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  uint32_t s = (blockIdx.x*1024 + threadIdx.x)*2654435761u + 12345u;
  int g = lane*3+1, h = lane+2;
  long long t0 = clock64();
  if (MODE==0){
    #pragma unroll 8
    for(int i=0;i<iters;i++){ uint32_t b=lcg(s); atomicAdd(&G[b*SLOTS+lane], g); atomicAdd(&H[b*SLOTS+lane], h); }
  } else if (MODE==1){
    #pragma unroll 8
    for(int i=0;i<iters;i++){ uint32_t b=lcg(s); int slot=i&31; atomicAdd(&G[slot*NBIN+b], g); atomicAdd(&H[slot*NBIN+b], h); }
  } else if (MODE==2){
    // each warp owns a private 64-bit table of 32 bins x 32 slots (8 KB/warp, up to 16 warps => 128 KB): pure RMW rate
    unsigned long long* T = reinterpret_cast<unsigned long long*>(smem_raw) + (warp&15)*32*32;
    unsigned long long v = ((unsigned long long)g<<32) | (unsigned)h;
    #pragma unroll 8
    for(int i=0;i<iters;i++){ uint32_t b=lcg(s)&31; unsigned long long o=T[b*32+lane]; T[b*32+lane]=o+v; }
  } else if (MODE==3){
    #pragma unroll 8
    for(int i=0;i<iters;i++){ uint32_t b=lcg(s); int2 q=gp[(i+warp)&255]; atomicAdd(&G[b*SLOTS+lane], q.x); atomicAdd(&H[b*SLOTS+lane], q.y); }
  } else if (MODE==4){
    #pragma unroll 8
    for(int i=0;i<iters;i++){ uint32_t b=lcg(s); int gg=__shfl_sync(0xffffffffu,g,i&31); int hh=__shfl_sync(0xffffffffu,h,i&31); atomicAdd(&G[b*SLOTS+lane], gg); atomicAdd(&H[b*SLOTS+lane], hh); }
  } else if (MODE==5){
    #pragma unroll 8
    for(int i=0;i<iters;i++){ uint32_t b=lcg(s); atomicAdd(&G[b*SLOTS+lane], g); }
  } else if (MODE==6){
    #pragma unroll 8
    for(int i=0;i<iters;i++){ uint32_t b=lcg(s); int slot=i&31; atomicAdd(&G[slot*NBIN+b], g); }
  } else if (MODE==7){
    #pragma unroll 8
    for(int i=0;i<iters;i++){ int r=(i+warp*8)&255; uint32_t b=tile[r*32+lane]; int2 q=gp[r]; atomicAdd(&G[b*SLOTS+lane], q.x); atomicAdd(&H[b*SLOTS+lane], q.y); }
  }
  long long t1 = clock64();
  __syncthreads();
  unsigned long long acc=0;
  for(int i=threadIdx.x;i<2*NBIN*SLOTS;i+=blockDim.x) acc+= (unsigned)G[i];
  if(acc==0xdeadbeefULL) sink[0]=acc+s;
  if(threadIdx.x==0) cyc[blockIdx.x]=t1-t0;
}

template<int MODE>
int run(const char* name, int threads, int ctas_per_sm, int iters){
  int nsm=148; cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, 0);
  size_t smem = 2*NBIN*SLOTS*4 + 256*8 + 256*32;
  if (MODE==2) smem = 16*32*32*8;
  CK(cudaFuncSetAttribute(kern<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  unsigned long long* sink; long long* cyc; CK(cudaMalloc(&sink,8)); CK(cudaMalloc(&cyc, 8*nsm*ctas_per_sm));
  int grid = nsm*ctas_per_sm;
  kern<MODE><<<grid,threads,smem>>>(iters/4,sink,cyc); CK(cudaDeviceSynchronize());
  cudaEvent_t e0,e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  kern<MODE><<<grid,threads,smem>>>(iters,sink,cyc);
  cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
  float ms; cudaEventElapsedTime(&ms,e0,e1);
  long long hc[2048]; CK(cudaMemcpy(hc,cyc,8*grid,cudaMemcpyDeviceToHost));
  double avg=0; for(int i=0;i<grid;i++) avg+=hc[i]; avg/=grid;
  double updates = (double)grid*threads*iters;   // one (g,h) update per thread-iteration
  double per_clk_sm = (double)ctas_per_sm*threads*iters/avg;
  printf("%-44s thr=%4d cta/sm=%d  %8.3f ms  %7.2f Gupd/s  %6.2f upd/clk/SM  (%.0f cyc, eff %.0f MHz)\n",
         name,threads,ctas_per_sm,ms,updates/ms*1e-6,per_clk_sm,avg,avg/ms*1e-3);
  cudaFree(sink); cudaFree(cyc);
  return 0;
}

int main(){
  int iters=20000;
  for (int thr : {256, 512, 1024}) {
    for (int c : {1, 2, 3}) {
      if (thr*c > 2048) continue;
      run<0>("m0 conflict-free 2xATOMS", thr,c,iters);
    }
  }
  for (int thr : {512, 1024}) for (int c : {1,2}) { if (thr*c>2048) continue;
    run<1>("m1 random-bank 2xATOMS", thr,c,iters);
    run<3>("m3 conflict-free 2xATOMS + LDS.64 bcast", thr,c,iters);
    run<4>("m4 conflict-free 2xATOMS + 2xSHFL", thr,c,iters);
    run<5>("m5 conflict-free 1xATOMS", thr,c,iters);
    run<6>("m6 random-bank 1xATOMS", thr,c,iters);
    run<7>("m7 cf 2xATOMS + LDS.U8 bins + LDS.64 gpair", thr,c,iters);
  }
  for (int thr : {128, 256, 512}) run<2>("m2 conflict-free nonatomic 64b RMW (warp-private)", thr,1,iters);
  return 0;
}
