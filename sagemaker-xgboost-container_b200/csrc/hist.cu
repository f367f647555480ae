// hist.cu -- per-node gradient/hessian histogram build: THE hot kernel.
// Replaces upstream xgboost's BuildHist (src/common/hist_util.cc / src/tree/gpu_hist/histogram.cu),
// reached from the container at algorithm_mode/train.py:367-376 (xgb.train -> Booster.update).
//
// Design (DESIGN.md "histogram kernel"; measurements in profiles/microbench_r1.md):
//  * sm_100a has exactly one fast shared-memory atomic: 32-bit integer ATOMS.ADD (float and 64-bit
//    adds compile to ATOMS.CAST.SPIN CAS loops), and it runs 2x faster when the 32 lanes of the
//    instruction hit 32 distinct banks.  So the histogram of one 32-feature group is two int32 planes
//    [256 bins][32 slots]: bank == slot, and every instruction below has lanes on 32 distinct slots.
//  * A warp takes a tile of 16 rows; two lanes share a row, each holding 16 of its 32 bin bytes
//    (one LDG.128 per lane, the row slice is exactly one 32 B sector).  Step j of 16 makes lane
//    (row q, half h) update slot 16h + rot_q(j): the per-row rotation makes the 16 rows of an
//    instruction touch 16 different slots of each half -> conflict-free by construction.
//  * Gradients are rounded to a power-of-two fixed-point grid (|g_q| <= 2^18, h_q <= 2^19) so a
//    window of 4096 rows per CTA cannot overflow int32; between windows, accumulators above 2^24 are
//    spilled to the global int64 histogram with RED.ADD.64 (sparse), and everything is flushed at the
//    end of the CTA's node portion.  Sums are exact integers => bit-reproducible for any grid size,
//    block schedule or GPU count, and the NCCL all-reduce of the int64 histograms is order-independent.
#include "engine.h"
#include "tree.h"

namespace b200 {

constexpr int kHistThreads = 256;
constexpr int kHistWarps = kHistThreads / 32;
constexpr int kTileRows = 16;
constexpr int kWindowRows = 4096;                 // rows per CTA between overflow checks
constexpr int kWindowTiles = kWindowRows / kTileRows;
constexpr int kMinRowsPerCta = 2048;              // do not pay a 16K-entry flush for fewer rows than this
constexpr int kSpillThreshold = 1 << 24;

__device__ __forceinline__ uint4 ldg_nc_v4(const void* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
  return r;
}
__device__ __forceinline__ float2 ldg_nc_f2(const void* p) {
  float2 r;
  asm volatile("ld.global.nc.v2.f32 {%0,%1}, [%2];" : "=f"(r.x), "=f"(r.y) : "l"(p));
  return r;
}
__device__ __forceinline__ void red_shared_s32(unsigned addr, int v) {
  asm volatile("red.shared.add.s32 [%0], %1;" :: "r"(addr), "r"(v) : "memory");
}
__device__ __forceinline__ void red_shared_u32_off(unsigned addr, unsigned v) {
  asm volatile("red.shared.add.u32 [%0+32768], %1;" :: "r"(addr), "r"(v) : "memory");
}
__device__ __forceinline__ void red_global_s64(long long* p, long long v) {
  asm volatile("red.global.add.u64 [%0], %1;" :: "l"(p), "l"(v) : "memory");
}


// Process rows [pa, pb) (segment positions) of one node into the shared-memory planes.
__device__ __forceinline__ void hist_window(const HistArgs& a, const uint8_t* gbins, unsigned smem_g,
                                            unsigned pa, unsigned pb, float sg, float sh,
                                            long long& accG, long long& accH) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int q = lane >> 1, half = lane & 1;
  const int qw = q >> 2, qb = q & 3;
  // per-lane constants of the rotation: byte selectors and slot offsets (bytes) for the 4x4 steps
  unsigned sel[4], offb[4], offw[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    sel[j] = 0x4440u | (unsigned)((j + qb) & 3);
    offb[j] = 4u * (unsigned)((j + qb) & 3);
    offw[j] = smem_g + 64u * (unsigned)half + 16u * (unsigned)((j + qw) & 3);
  }
  const unsigned ntiles = (pb - pa + kTileRows - 1) / kTileRows;
  // software pipeline: registers of the next tile are loaded before the atomics of the current one
  uint4 wn = make_uint4(0, 0, 0, 0); float2 ghn = make_float2(0.f, 0.f);
  unsigned t = warp;
  auto load_tile = [&](unsigned tile, uint4& w, float2& gh) {
    unsigned p = pa + tile * kTileRows + q;
    bool valid = p < pb;
    unsigned r = valid ? (a.ridx ? __ldg(a.ridx + p) : p) : 0u;
    if (valid) {
      w = ldg_nc_v4(gbins + (int64_t)r * kSlots + half * 16);
      gh = ldg_nc_f2(a.gpair + r);
    } else { w = make_uint4(0, 0, 0, 0); gh = make_float2(0.f, 0.f); }
  };
  if (t < ntiles) load_tile(t, wn, ghn);
  for (; t < ntiles; t += kHistWarps) {
    uint4 w = wn; float2 gh = ghn;
    if (t + kHistWarps < ntiles) load_tile(t + kHistWarps, wn, ghn);
    const int gq = __float2int_rn(gh.x * sg);
    const unsigned hq = (unsigned)__float2int_rn(gh.y * sh);
    if (half == 0) { accG += gq; accH += hq; }
    // rotate the four words by qw so that step jw reads word (jw + qw) & 3 from a fixed register
    unsigned w0 = w.x, w1 = w.y, w2 = w.z, w3 = w.w;
    if (qw & 1) { unsigned x = w0; w0 = w1; w1 = w2; w2 = w3; w3 = x; }
    if (qw & 2) { unsigned x = w0; w0 = w2; w2 = x; x = w1; w1 = w3; w3 = x; }
    const unsigned ww[4] = {w0, w1, w2, w3};
#pragma unroll
    for (int jw = 0; jw < 4; ++jw) {
#pragma unroll
      for (int jb = 0; jb < 4; ++jb) {
        unsigned bin = __byte_perm(ww[jw], 0u, sel[jb]);
        unsigned addr = (bin << 7) + offw[jw] + offb[jb];
        red_shared_s32(addr, gq);
        red_shared_u32_off(addr, hq);
      }
    }
  }
}

__global__ void __launch_bounds__(kHistThreads, 3) hist_build_kernel(HistArgs a) {
  extern __shared__ __align__(16) int smem[];         // LG[8192] then LH[8192]
  const int nb = *a.build_count;
  if (nb <= 0) return;
  const unsigned T = a.build_prefix[nb];
  if (T == 0) return;
  if (a.rows_counter && blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0) atomicAdd(a.rows_counter, (unsigned long long)T);
  const unsigned C = gridDim.x;
  unsigned ceff = (T + kMinRowsPerCta - 1) / kMinRowsPerCta;
  ceff = ceff < 1 ? 1 : (ceff > C ? C : ceff);
  if (blockIdx.x >= ceff) return;
  unsigned chunk = ((T + ceff - 1) / ceff + kTileRows - 1) & ~(unsigned)(kTileRows - 1);
  unsigned long long r0l = (unsigned long long)blockIdx.x * chunk;
  if (r0l >= T) return;
  unsigned r0 = (unsigned)r0l;
  unsigned r1 = (unsigned long long)r0 + chunk > T ? T : r0 + chunk;
  const int group = blockIdx.y;
  const uint8_t* gbins = a.bins + (int64_t)group * a.n * kSlots;
  const float sg = a.scales[0], sh = a.scales[1];
  const unsigned smem_g = (unsigned)__cvta_generic_to_shared(smem);

  for (int i = threadIdx.x; i < 2 * kGroupEntries / 4; i += kHistThreads) reinterpret_cast<int4*>(smem)[i] = make_int4(0, 0, 0, 0);
  __syncthreads();

  // first build node whose range contains r0
  int b = 0;
  { int lo = 0, hi = nb; while (lo < hi) { int mid = (lo + hi) >> 1; if (a.build_prefix[mid + 1] > r0) hi = mid; else lo = mid + 1; } b = lo; }

  while (r0 < r1) {
    const unsigned nbeg = a.build_prefix[b], nend_node = a.build_prefix[b + 1];
    const unsigned nend = nend_node < r1 ? nend_node : r1;
    const int nid = a.build_nid[b];
    const unsigned seg = a.seg_begin[nid];
    GH64* out = a.hist_pool + ((int64_t)a.hist_slot[nid] * a.ngroups + group) * kGroupEntries;
    long long accG = 0, accH = 0;
    unsigned pa = seg + (r0 - nbeg), pend = seg + (nend - nbeg);
    while (pa < pend) {
      unsigned pb = pend - pa > (unsigned)kWindowRows ? pa + kWindowRows : pend;
      hist_window(a, gbins, smem_g, pa, pb, sg, sh, accG, accH);
      pa = pb;
      __syncthreads();
      const bool last = pa >= pend;
      // spill (between windows: only accumulators that could overflow in the next window) / final flush
      for (int e = threadIdx.x; e < kGroupEntries; e += kHistThreads) {
        int g = smem[e]; unsigned h = (unsigned)smem[kGroupEntries + e];
        bool sg_ = last ? (g != 0) : (g >= kSpillThreshold || g <= -kSpillThreshold);
        bool sh_ = last ? (h != 0) : (h >= (unsigned)kSpillThreshold);
        if (sg_) { red_global_s64(&out[e].g, (long long)g); smem[e] = 0; }
        if (sh_) { red_global_s64(&out[e].h, (long long)h); smem[kGroupEntries + e] = 0; }
      }
      __syncthreads();
    }
    if (a.accumulate_sum && group == 0) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) { accG += __shfl_xor_sync(0xffffffffu, accG, o); accH += __shfl_xor_sync(0xffffffffu, accH, o); }
      if ((threadIdx.x & 31) == 0 && (accG != 0 || accH != 0)) { red_global_s64(&a.node_sum[nid].g, accG); red_global_s64(&a.node_sum[nid].h, accH); }
    }
    r0 = nend; ++b;
  }
}

void launch_hist_build(const HistArgs& a, int grid_x, cudaStream_t stream) {
  static bool configured = false;
  if (!configured) {
    CUDA_OK(cudaFuncSetAttribute(hist_build_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * kGroupEntries * (int)sizeof(int)));
    configured = true;
  }
  dim3 grid(grid_x, a.ngroups);
  hist_build_kernel<<<grid, kHistThreads, 2 * kGroupEntries * sizeof(int), stream>>>(a); ++g_kernel_launches;
  CUDA_OK(cudaGetLastError());
}

}  // namespace b200
