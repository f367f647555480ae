// Microbenchmark 2: cost of exact 64-bit-equivalent accumulation on top of native 32-bit ATOMS.ADD
//  c0: plain 2xATOMS (no return)                                  [baseline, = m0 of smem_atomics.cu]
//  c1: 2xATOMS with return + carry/borrow detection into HI arrays (exact, P up to 31 bits)
//  c2: plain 2xATOMS + LO->HI normalisation pass every K rows
//  f0: global flush: every CTA RED.ADD.64's a 16K-entry table into one global table (zero-skip off)
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("CUDA error %s at %d\n",cudaGetErrorString(e),__LINE__); return 1;}}while(0)
constexpr int NBIN=256, SLOTS=32, NE=NBIN*SLOTS;
__device__ __forceinline__ uint32_t lcg(uint32_t& s){ s = s*1664525u + 1013904223u; return s>>24; }

template<int MODE>
__global__ void __launch_bounds__(512) kern(int iters, int K, unsigned long long* sink){
  extern __shared__ __align__(16) int sm[];
  int* LG=sm; int* LH=sm+NE; int* HG=sm+2*NE; int* HH=sm+3*NE;
  for(int i=threadIdx.x;i<4*NE;i+=blockDim.x) sm[i]=0;
  __syncthreads();
  const int lane=threadIdx.x&31;
  uint32_t s=(blockIdx.x*1024+threadIdx.x)*2654435761u+12345u;
  if (MODE==0){
    #pragma unroll 8
    for(int i=0;i<iters;i++){ uint32_t b=lcg(s); int g=(int)(s>>9)-(1<<22); int h=(s>>10)&0x3fffff;
      atomicAdd(&LG[b*SLOTS+lane],g); atomicAdd(&LH[b*SLOTS+lane],h);} 
  } else if (MODE==1){
    #pragma unroll 8
    for(int i=0;i<iters;i++){ uint32_t b=lcg(s); int g=(int)(s>>9)-(1<<22); int h=(s>>10)&0x3fffff;
      uint32_t m = g<0 ? 0xffffffffu : 0u;               // per-row in the real kernel (amortised)
      uint32_t thr = g<0 ? (uint32_t)(-g) : ~(uint32_t)g;
      uint32_t og=(uint32_t)atomicAdd(&LG[b*SLOTS+lane],g);
      uint32_t oh=(uint32_t)atomicAdd((unsigned*)&LH[b*SLOTS+lane],(unsigned)h);
      bool cg = ((og^m) > (thr^m));
      bool ch = oh > ~(uint32_t)h;
      if (cg|ch){ if(cg) atomicAdd(&HG[b*SLOTS+lane], g<0?-1:1); if(ch) atomicAdd(&HH[b*SLOTS+lane],1);} 
    }
  } else if (MODE==2){
    int done=0;
    while(done<iters){
      int n=min(K/16,iters-done);    // 16 warps: K rows per CTA between normalisations
      #pragma unroll 8
      for(int i=0;i<n;i++){ uint32_t b=lcg(s); int g=(int)((s>>9)&0x3ffff)-(1<<17); int h=(s>>10)&0x1ffff;
        atomicAdd(&LG[b*SLOTS+lane],g); atomicAdd(&LH[b*SLOTS+lane],h);} 
      done+=n;
      __syncthreads();
      for(int e=threadIdx.x*4;e<2*NE;e+=blockDim.x*4){
        int4 lo=*reinterpret_cast<int4*>(&sm[e]); int4 hi=*reinterpret_cast<int4*>(&sm[2*NE+e]);
        hi.x+=lo.x>>16; lo.x&=0xffff; hi.y+=lo.y>>16; lo.y&=0xffff; hi.z+=lo.z>>16; lo.z&=0xffff; hi.w+=lo.w>>16; lo.w&=0xffff;
        *reinterpret_cast<int4*>(&sm[e])=lo; *reinterpret_cast<int4*>(&sm[2*NE+e])=hi; }
      __syncthreads();
    }
  }
  __syncthreads();
  unsigned long long acc=0; for(int i=threadIdx.x;i<4*NE;i+=blockDim.x) acc+=(unsigned)sm[i];
  if(acc==0xdeadbeefULL) sink[0]=acc;
}

__global__ void __launch_bounds__(512) flush_kern(long long* gh, int reps){
  extern __shared__ __align__(16) int sm[];
  for(int i=threadIdx.x;i<4*NE;i+=blockDim.x) sm[i]=i*7+blockIdx.x;
  __syncthreads();
  for(int r=0;r<reps;r++){
    for(int e=threadIdx.x;e<NE;e+=blockDim.x){
      long long g=((long long)sm[2*NE+e]<<32)+(unsigned)sm[e]; long long h=((long long)sm[3*NE+e]<<32)+(unsigned)sm[NE+e];
      atomicAdd((unsigned long long*)&gh[2*e],(unsigned long long)g); atomicAdd((unsigned long long*)&gh[2*e+1],(unsigned long long)h);
    }
  }
}

template<int MODE> int run(const char* name,int iters,int K){
  size_t smem=4*NE*4; CK(cudaFuncSetAttribute(kern<MODE>,cudaFuncAttributeMaxDynamicSharedMemorySize,(int)smem));
  unsigned long long* sink; CK(cudaMalloc(&sink,8));
  kern<MODE><<<148,512,smem>>>(iters/4,K,sink); CK(cudaDeviceSynchronize());
  cudaEvent_t e0,e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0); kern<MODE><<<148,512,smem>>>(iters,K,sink); cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
  float ms; cudaEventElapsedTime(&ms,e0,e1);
  printf("%-52s K=%5d %8.3f ms %8.2f Gupd/s\n",name,K,ms,148.0*512*iters/ms*1e-6);
  cudaFree(sink); return 0;
}
int main(){
  int iters=20000;
  run<0>("c0 plain 2xATOMS (512 thr x 1 CTA/SM, 128KB smem)",iters,0);
  run<1>("c1 2xATOMS.return + exact carry into HI",iters,0);
  for(int K: {2048,4096,8192,16384}) run<2>("c2 plain 2xATOMS + LO->HI normalisation every K rows",iters,K);
  // flush
  long long* gh; CK(cudaMalloc(&gh,16*NE)); CK(cudaMemset(gh,0,16*NE));
  size_t smem=4*NE*4; CK(cudaFuncSetAttribute(flush_kern,cudaFuncAttributeMaxDynamicSharedMemorySize,(int)smem));
  for (int grid: {148, 296, 592}) for (int reps: {1, 4}) {
    flush_kern<<<grid,512,smem>>>(gh,1); CK(cudaDeviceSynchronize());
    cudaEvent_t e0,e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0); flush_kern<<<grid,512,smem>>>(gh,reps); cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
    float ms; cudaEventElapsedTime(&ms,e0,e1);
    printf("f0 flush RED.64 grid=%4d reps=%d: %8.3f ms  %8.2f G RED/s (%.1f us per 16K-entry (g,h) flush wave)\n",grid,reps,ms,(double)grid*reps*2*NE/ms*1e-6, ms*1e3/reps);
  }
  return 0;
}
