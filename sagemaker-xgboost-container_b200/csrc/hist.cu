// hist.cu -- per-node gradient/hessian histogram build: THE hot kernel.
// Replaces upstream xgboost's BuildHist (src/common/hist_util.cc / src/tree/gpu_hist/histogram.cu),
// reached from the container at algorithm_mode/train.py:367-376 (xgb.train -> Booster.update).
//
// Design (DESIGN.md "histogram kernel"; measurements in profiles/):
//  * sm_100a has exactly one fast shared-memory atomic: 32-bit integer ATOMS.ADD (float and 64-bit
//    adds compile to ATOMS.CAST.SPIN CAS loops), and it runs 2x faster when the 32 lanes of the
//    instruction hit 32 distinct banks.  So the histogram of one 32-feature group is two int32 planes
//    [256 bins][32 slots]: bank == slot, and every instruction below has lanes on 32 distinct slots.
//  * The binned matrix is row-major, each row = feature groups of 32 B.  A CTA owns a PAIR of adjacent
//    groups (64 B of every row = one full DRAM burst, so gathered rows of deep levels waste nothing)
//    or a single group when the matrix has only one.
//  * A warp takes a tile of 16 rows; two lanes share a row, each holding 16 of a group's 32 bin bytes
//    (one LDG.128 per lane and group).  Step j of 16 makes lane (row q, half h) update slot
//    16h + rot_q(j): the per-row rotation makes the 16 rows of an instruction touch 16 different slots
//    of each half -> conflict-free by construction.
//  * (g,h) pairs are read by POSITION (they travel with the row ids through the partition), so the
//    only gathered loads are the bin slices; row ids and pairs are prefetched two stages ahead.
//  * Gradients are rounded to a power-of-two fixed-point grid (|g_q| <= 2^18, h_q <= 2^19) so a
//    window of 4096 rows per CTA cannot overflow int32; between windows, accumulators above 2^24 are
//    spilled to the global int64 histogram with RED.ADD.64 (sparse), and everything is flushed at the
//    end of the CTA's node portion.  Sums are exact integers => bit-reproducible for any grid size,
//    block schedule or GPU count, and the NCCL all-reduce of the int64 histograms is order-independent.
#include <cstdlib>
#include "engine.h"
#include "tree.h"

namespace b200 {

constexpr int kTileRows = 16;
constexpr int kSuperRows = 2 * kTileRows;
constexpr int kWindowRows = 4096;                 // rows per CTA between overflow checks
constexpr int kMinRowsPerCta = 4096;              // do not pay a flush for fewer rows than this
constexpr int kSpillThreshold = 1 << 24;
constexpr int kPlaneBytes = kGroupEntries * 4;    // 32 KB

__device__ __forceinline__ uint4 ldg_nc_v4(const void* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
  return r;
}
__device__ __forceinline__ float2 ldg_nc_f2(const void* p) {
  float2 r;
  asm volatile("ld.global.nc.L1::no_allocate.v2.f32 {%0,%1}, [%2];" : "=f"(r.x), "=f"(r.y) : "l"(p));
  return r;
}
__device__ __forceinline__ void red_shared_s32(unsigned addr, int v) {
  asm volatile("red.shared.add.s32 [%0], %1;" :: "r"(addr), "r"(v) : "memory");
}
__device__ __forceinline__ void red_shared_u32_h(unsigned addr, unsigned v) {      // hessian plane = gradient plane + 32 KB
  asm volatile("red.shared.add.u32 [%0+32768], %1;" :: "r"(addr), "r"(v) : "memory");
}
__device__ __forceinline__ void red_global_s64(long long* p, long long v) {
  asm volatile("red.global.add.u64 [%0], %1;" :: "l"(p), "l"(v) : "memory");
}

struct LaneConst { unsigned sel[4], offb[4], offw[4]; int qw; int rowlane; int colbyte; };

// 16 conflict-free (g,h) atomic pairs of one lane's 16 bin bytes into the planes at smem byte offset `plane`
__device__ __forceinline__ void accumulate16(const LaneConst& lc, const uint4& w, int gq, unsigned hq) {
  unsigned w0 = w.x, w1 = w.y, w2 = w.z, w3 = w.w;
  if (lc.qw & 1) { unsigned x = w0; w0 = w1; w1 = w2; w2 = w3; w3 = x; }
  if (lc.qw & 2) { unsigned x = w0; w0 = w2; w2 = x; x = w1; w1 = w3; w3 = x; }
  const unsigned ww[4] = {w0, w1, w2, w3};
#pragma unroll
  for (int jw = 0; jw < 4; ++jw) {
#pragma unroll
    for (int jb = 0; jb < 4; ++jb) {
      unsigned bin = __byte_perm(ww[jw], 0u, lc.sel[jb]);
      unsigned addr = (bin << 7) + lc.offw[jw] + lc.offb[jb];
      red_shared_s32(addr, gq);
      red_shared_u32_h(addr, hq);
    }
  }
}

struct Stage { unsigned id; float2 gh; };
template <int NSUB> struct Rows { uint4 w[NSUB]; };         // one 16 B chunk per sub-tile of the 32-position super-tile

// Spill / flush pass over the CTA's accumulators.  Between windows only accumulators that could overflow in the next
// window leave for the global int64 histogram (sparse RED.ADD.64); `last` flushes everything that is non-zero.
template <int NTHREADS>
__device__ __forceinline__ void spill_pass(int* smem, int ng_here, GH64* out, bool last) {
  const int vecs = ng_here * 2 * kGroupEntries / 4;
  for (int v = threadIdx.x; v < vecs; v += NTHREADS) {
    int4 x = reinterpret_cast<int4*>(smem)[v];
    const int plane = (v * 4) / kGroupEntries;              // 0: G of group 0, 1: H of group 0, 2: G of group 1, 3: H of group 1
    const int e0 = v * 4 - plane * kGroupEntries;
    const bool is_h = plane & 1;
    int vals[4] = {x.x, x.y, x.z, x.w};
    bool any = false;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int val = vals[k];
      bool sp = last ? (val != 0) : (is_h ? ((unsigned)val >= (unsigned)kSpillThreshold) : (val >= kSpillThreshold || val <= -kSpillThreshold));
      if (sp) {
        GH64* o = out + (size_t)(plane >> 1) * kGroupEntries + e0 + k;
        red_global_s64(is_h ? &o->h : &o->g, is_h ? (long long)(unsigned)val : (long long)val);
        vals[k] = 0; any = true;
      }
    }
    if (any) reinterpret_cast<int4*>(smem)[v] = make_int4(vals[0], vals[1], vals[2], vals[3]);
  }
}

template <int NG, int NTHREADS>
__global__ void __launch_bounds__(NTHREADS, NG == 2 ? 1 : 3) hist_build_kernel(HistArgs a) {
  constexpr int NWARPS = NTHREADS / 32;
  constexpr int kItersPerWindow = kWindowRows / (kSuperRows * NWARPS);     // super-tiles per warp between overflow checks
  static_assert(kItersPerWindow >= 1, "window too small");
  extern __shared__ __align__(16) int smem[];         // per group: LG[8192] then LH[8192]
  const int nb = *a.build_count;
  if (nb <= 0) return;
  const unsigned T = a.build_prefix[nb];
  if (T == 0) return;
  if (a.rows_counter && blockIdx.x == 0 && a.group_base + blockIdx.y == 0 && threadIdx.x == 0) atomicAdd(a.rows_counter, (unsigned long long)T);
  const unsigned C = gridDim.x;
  unsigned ceff = (T + kMinRowsPerCta - 1) / kMinRowsPerCta;
  ceff = ceff < 1 ? 1 : (ceff > C ? C : ceff);
  if (blockIdx.x >= ceff) return;
  unsigned chunk = ((T + ceff - 1) / ceff + kSuperRows - 1) & ~(unsigned)(kSuperRows - 1);
  unsigned long long r0l = (unsigned long long)blockIdx.x * chunk;
  if (r0l >= T) return;
  unsigned r0 = (unsigned)r0l;
  unsigned r1 = (unsigned long long)r0 + chunk > T ? T : r0 + chunk;
  const int g0 = a.group_base + blockIdx.y * NG;
  const int ng_here = NG;
  const uint8_t* gbins = a.bins + (int64_t)g0 * kSlots;
  const float sg = a.scales[0], sh = a.scales[1];
  const unsigned smem_g = (unsigned)__cvta_generic_to_shared(smem);
  const int64_t row_stride = (int64_t)a.row_stride;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;

  // Lane roles.  NG == 1: two lanes per row (16 B halves of the 32 B slice), 16 rows per LDG.128 instruction.
  // NG == 2: four lanes per row (16 B chunks of the row's 64 B = group A low/high, group B low/high), 8 rows per
  // instruction, so an instruction touches each 64 B burst exactly once (half the L1 wavefronts of two 32 B loads).
  // `rot` is the per-lane rotation of the 16-step slot schedule: the 16 lanes that share a bank range (same `half`)
  // get 16 different rotations, so every ATOMS instruction hits 32 distinct banks.
  constexpr int NSUB = NG == 2 ? 4 : 2;               // sub-tiles per 32-position super-tile
  constexpr int SUBROWS = 32 / NSUB;
  LaneConst lc;
  { int rot, half; unsigned plane;
    if (NG == 2) { const int q8 = lane >> 2, c = lane & 3; rot = 2 * q8 + (c >> 1); half = c & 1; plane = (unsigned)(c >> 1) * 2u * kPlaneBytes; lc.rowlane = q8; lc.colbyte = c * 16; }
    else { rot = lane >> 1; half = lane & 1; plane = 0u; lc.rowlane = lane >> 1; lc.colbyte = half * 16; }
    lc.qw = rot >> 2; const int qb = rot & 3;
#pragma unroll
    for (int j = 0; j < 4; ++j) { lc.sel[j] = 0x4440u | (unsigned)((j + qb) & 3); lc.offb[j] = 4u * (unsigned)((j + qb) & 3);
      lc.offw[j] = smem_g + plane + 64u * (unsigned)half + 16u * (unsigned)((j + lc.qw) & 3); } }

  const int entries = ng_here * 2 * kGroupEntries;
  for (int i = threadIdx.x; i < entries / 4; i += NTHREADS) reinterpret_cast<int4*>(smem)[i] = make_int4(0, 0, 0, 0);
  __syncthreads();

  int b = 0;   // first build node whose range contains r0
  { int lo = 0, hi = nb; while (lo < hi) { int mid = (lo + hi) >> 1; if (a.build_prefix[mid + 1] > r0) hi = mid; else lo = mid + 1; } b = lo; }

  const Stage none{0xffffffffu, make_float2(0.f, 0.f)};
  while (r0 < r1) {
    const unsigned nbeg = a.build_prefix[b], nend_node = a.build_prefix[b + 1];
    const unsigned nend = nend_node < r1 ? nend_node : r1;
    const int nid = a.build_nid[b];
    const unsigned seg = a.seg_begin[nid];
    GH64* out = a.hist_pool + ((int64_t)a.hist_slot[nid] * a.ngroups + g0) * kGroupEntries;
    long long accG = 0, accH = 0;
    const unsigned pa = seg + (r0 - nbeg), pb = seg + (nend - nbeg);
    const unsigned nsuper = (pb - pa + kSuperRows - 1) / kSuperRows;
    const unsigned iters = (nsuper + NWARPS - 1) / NWARPS;          // same for every warp: barriers stay aligned

    // Three-stage software pipeline per warp over super-tiles of 32 positions; it runs THROUGH the overflow-check
    // barriers (the loads of the next super-tiles stay in flight while the CTA spills), so no window restarts cold:
    //   stage A: row ids + (g,h) of super-tile s+2 (coalesced, one position per lane)
    //   stage B: bin slices of super-tile s+1 (LDG.128 per lane and group, addresses from stage A via SHFL)
    //   stage C: conflict-free ATOMS pairs of super-tile s
    auto load_ids = [&](unsigned st) -> Stage {
      Stage s_ = none;
      unsigned p = pa + st * kSuperRows + lane;
      if (st < nsuper && p < pb) { s_.id = a.ridx ? __ldg(a.ridx + p) : p; s_.gh = ldg_nc_f2(a.gpair + p); }
      return s_;
    };
    auto load_rows = [&](unsigned ids, Rows<NSUB>& r) {
#pragma unroll
      for (int t = 0; t < NSUB; ++t) {
        const unsigned rid = __shfl_sync(0xffffffffu, ids, t * SUBROWS + lc.rowlane);
        r.w[t] = rid != 0xffffffffu ? ldg_nc_v4(gbins + (int64_t)rid * row_stride + lc.colbyte) : make_uint4(0, 0, 0, 0);
      }
    };
    Stage cur = load_ids(warp);
    Stage nxt = load_ids(warp + NWARPS);
    Rows<NSUB> rows; load_rows(cur.id, rows);
    for (unsigned it = 0; it < iters; ++it) {
      const unsigned s = warp + it * NWARPS;
      Stage nn = load_ids(s + 2 * NWARPS);
      Rows<NSUB> nrows; load_rows(nxt.id, nrows);
      if (s < nsuper) {
        const int gq_l = __float2int_rn(cur.gh.x * sg);
        const unsigned hq_l = (unsigned)__float2int_rn(cur.gh.y * sh);
        accG += gq_l; accH += hq_l;
#pragma unroll
        for (int t = 0; t < NSUB; ++t) {
          const int gq = __shfl_sync(0xffffffffu, gq_l, t * SUBROWS + lc.rowlane);
          const unsigned hq = __shfl_sync(0xffffffffu, hq_l, t * SUBROWS + lc.rowlane);
          accumulate16(lc, rows.w[t], gq, hq);
        }
      }
      rows = nrows; cur = nxt; nxt = nn;
      if ((it + 1) % kItersPerWindow == 0 && it + 1 < iters) {       // overflow check: at most 4096 rows since the last one
        __syncthreads();
        spill_pass<NTHREADS>(smem, ng_here, out, false);
        __syncthreads();
      }
    }
    __syncthreads();
    spill_pass<NTHREADS>(smem, ng_here, out, true);
    __syncthreads();
    if (a.accumulate_sum && g0 == 0) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) { accG += __shfl_xor_sync(0xffffffffu, accG, o); accH += __shfl_xor_sync(0xffffffffu, accH, o); }
      if (lane == 0 && (accG != 0 || accH != 0)) { red_global_s64(&a.node_sum[nid].g, accG); red_global_s64(&a.node_sum[nid].h, accH); }
    }
    r0 = nend; ++b;
  }
}

int hist_grid_x(int num_sms, int ngroups) {
  if (ngroups == 1) return num_sms * 3;
  const int pairs = ngroups / 2;
  const int x = (num_sms + pairs - 1) / pairs;
  return x > 0 ? x : 1;
}

void hist_configure() {
  static bool configured = false;
  if (configured) return;
  CUDA_OK(cudaFuncSetAttribute(hist_build_kernel<1, 256>, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * kPlaneBytes));
  CUDA_OK(cudaFuncSetAttribute(hist_build_kernel<2, 768>, cudaFuncAttributeMaxDynamicSharedMemorySize, 4 * kPlaneBytes));
  CUDA_OK(cudaFuncSetAttribute(hist_build_kernel<2, 1024>, cudaFuncAttributeMaxDynamicSharedMemorySize, 4 * kPlaneBytes));
  configured = true;
}

void launch_hist_build(const HistArgs& a_in, int grid_x, cudaStream_t stream) {
  HistArgs a = a_in;
  const int pairs = a.ngroups / 2;
  // experiment knob for the next round: 32 warps/SM at 64 registers (small spills) instead of 24 warps at 80
  static const bool wide = getenv("B200XGB_HIST_THREADS") != nullptr && atoi(getenv("B200XGB_HIST_THREADS")) == 1024;
  if (pairs > 0) {
    a.group_base = 0;
    if (wide) hist_build_kernel<2, 1024><<<dim3(grid_x, pairs), 1024, 4 * kPlaneBytes, stream>>>(a);
    else hist_build_kernel<2, 768><<<dim3(grid_x, pairs), 768, 4 * kPlaneBytes, stream>>>(a);
    ++g_kernel_launches;
    CUDA_OK(cudaGetLastError());
  }
  if (a.ngroups & 1) {                       // single (or odd last) group: 64 KB CTAs, three per SM
    a.group_base = a.ngroups - 1;
    hist_build_kernel<1, 256><<<dim3(pairs > 0 ? grid_x : hist_grid_x(148, 1), 1), 256, 2 * kPlaneBytes, stream>>>(a); ++g_kernel_launches;
    CUDA_OK(cudaGetLastError());
  }
}

}  // namespace b200
