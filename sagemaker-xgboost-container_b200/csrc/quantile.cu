// quantile.cu -- feature-quantile cut points on the device.
// Upstream xgboost builds cuts with a weighted GK sketch (src/common/quantile.cc, hist_util.cc); that sketch is
// not restatable bit-exactly, so product and oracle share an EXACT definition instead (oracle/gbt_oracle.c
// cuts_from_distinct): sort each feature, collapse to distinct values with weights, then
//   m <= max_bin : cuts = distinct[1..m-1] U {last + (|last| + 1e-5)}          (identical to upstream)
//   m >  max_bin : cut k = the distinct value following the one whose cumulative weight reaches k*W/max_bin.
// One-time cost per DMatrix; uses CUB device primitives (sort / run-length / scan), not on the per-round path.
#include <cub/cub.cuh>
#include <cmath>
#include <algorithm>
#include "engine.h"
#include "misc.h"

namespace b200 {

__global__ void extract_col_kernel(const float* X, int64_t n, int F, int f, const float* w, float* keys, float* wout) {
  for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < n; r += (int64_t)gridDim.x * blockDim.x) {
    float v = X[r * F + f];
    bool nan = isnan(v);
    keys[r] = nan ? __int_as_float(0x7f800000) : (v == 0.f ? 0.f : v);      // NaN -> +inf: sorts last; -0.0 -> +0.0 so that the
                                                                            // representative of the zero run (a cut value) does not depend on sort order / rank count
    if (wout) wout[r] = nan ? 0.f : (w ? w[r] : 1.f);
  }
}
__global__ void count_valid_kernel(const float* X, int64_t n, int F, int f, unsigned long long* out) {
  unsigned long long c = 0;
  for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < n; r += (int64_t)gridDim.x * blockDim.x) c += isnan(X[r * F + f]) ? 0ull : 1ull;
  for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
  if ((threadIdx.x & 31) == 0 && c) atomicAdd(out, c);
}
struct ToDouble { __host__ __device__ double operator()(float x) const { return (double)x; } };
struct IntToDouble { __host__ __device__ double operator()(int x) const { return (double)x; } };

// pick `cap` summary points: point k = first distinct index whose inclusive cumulative weight >= k*W/cap
__global__ void pick_kernel(const double* cum, int m, int cap, int* idx_out) {
  int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= cap) return;
  double W = cum[m - 1];
  double target = W * (double)(k + 1) / (double)cap;
  int lo = 0, hi = m - 1;          // first i with cum[i] >= target
  while (lo < hi) { int mid = (lo + hi) >> 1; if (cum[mid] >= target) hi = mid; else lo = mid + 1; }
  idx_out[k] = lo;
}
__global__ void gather_kernel(const float* vals, const double* cum, const int* idx, int cnt, float* v_out, double* c_out) {
  int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= cnt) return;
  v_out[k] = vals[idx[k]]; c_out[k] = cum[idx[k]];
}

void compute_summaries_device(const float* dX, int64_t n, int F, const float* dweights, int cap,
                              std::vector<FeatureSummary>* out, cudaStream_t s) {
  out->assign(F, FeatureSummary());
  if (n == 0) return;
  DevBuf<float> keys, keys2, wts, wts2, uniq, wsum;
  DevBuf<int> counts, nruns, idx;
  DevBuf<double> cum, csel; DevBuf<float> vsel;
  DevBuf<unsigned long long> nvalid;
  keys.alloc(n); keys2.alloc(n); uniq.alloc(n); nruns.alloc(1); nvalid.alloc(1); cum.alloc(n);
  idx.alloc(cap); csel.alloc(cap); vsel.alloc(cap);
  const bool weighted = dweights != nullptr;
  if (weighted) { wts.alloc(n); wts2.alloc(n); wsum.alloc(n); } else counts.alloc(n);
  size_t tmp_bytes = 0, need = 0;
  // temp storage: max over the primitives used
  cub::DeviceRadixSort::SortKeys(nullptr, need, keys.p, keys2.p, (int64_t)n, 0, 32, s); tmp_bytes = std::max(tmp_bytes, need);
  if (weighted) { cub::DeviceRadixSort::SortPairs(nullptr, need, keys.p, keys2.p, wts.p, wts2.p, (int64_t)n, 0, 32, s); tmp_bytes = std::max(tmp_bytes, need);
    cub::DeviceReduce::ReduceByKey(nullptr, need, keys2.p, uniq.p, wts2.p, wsum.p, nruns.p, cub::Sum(), (int)n, s); tmp_bytes = std::max(tmp_bytes, need); }
  else { cub::DeviceRunLengthEncode::Encode(nullptr, need, keys2.p, uniq.p, counts.p, nruns.p, (int)n, s); tmp_bytes = std::max(tmp_bytes, need); }
  cub::DeviceScan::InclusiveSum(nullptr, need, cum.p, cum.p, (int)n, s); tmp_bytes = std::max(tmp_bytes, need);
  DevBuf<unsigned char> tmp; tmp.alloc(tmp_bytes + 16);
  const int grid = (int)std::min<int64_t>((n + 255) / 256, 148 * 16);
  for (int f = 0; f < F; ++f) {
    extract_col_kernel<<<grid, 256, 0, s>>>(dX, n, F, f, dweights, keys.p, weighted ? wts.p : nullptr); ++g_kernel_launches;
    CUDA_OK(cudaMemsetAsync(nvalid.p, 0, 8, s));
    count_valid_kernel<<<grid, 256, 0, s>>>(dX, n, F, f, nvalid.p); ++g_kernel_launches;
    size_t tb = tmp_bytes;
    if (weighted) {
      CUDA_OK(cub::DeviceRadixSort::SortPairs(tmp.p, tb, keys.p, keys2.p, wts.p, wts2.p, (int64_t)n, 0, 32, s)); tb = tmp_bytes;
      CUDA_OK(cub::DeviceReduce::ReduceByKey(tmp.p, tb, keys2.p, uniq.p, wts2.p, wsum.p, nruns.p, cub::Sum(), (int)n, s));
    } else {
      CUDA_OK(cub::DeviceRadixSort::SortKeys(tmp.p, tb, keys.p, keys2.p, (int64_t)n, 0, 32, s)); tb = tmp_bytes;
      CUDA_OK(cub::DeviceRunLengthEncode::Encode(tmp.p, tb, keys2.p, uniq.p, counts.p, nruns.p, (int)n, s));
    }
    int m = 0; unsigned long long nv = 0;
    CUDA_OK(cudaMemcpyAsync(&m, nruns.p, sizeof(int), cudaMemcpyDeviceToHost, s));
    CUDA_OK(cudaMemcpyAsync(&nv, nvalid.p, 8, cudaMemcpyDeviceToHost, s));
    CUDA_OK(cudaStreamSynchronize(s));
    if (nv < (unsigned long long)n) m -= 1;              // trailing +inf run holds the missing entries
    FeatureSummary& fs = (*out)[f];
    if (m <= 0) continue;
    // inclusive cumulative weights in double
    tb = tmp_bytes;
    if (weighted) { cub::TransformInputIterator<double, ToDouble, float*> it(wsum.p, ToDouble()); CUDA_OK(cub::DeviceScan::InclusiveSum(tmp.p, tb, it, cum.p, m, s)); }
    else { cub::TransformInputIterator<double, IntToDouble, int*> it(counts.p, IntToDouble()); CUDA_OK(cub::DeviceScan::InclusiveSum(tmp.p, tb, it, cum.p, m, s)); }
    std::vector<float> v; std::vector<double> c;
    if (m <= cap) {
      v.resize(m); c.resize(m);
      CUDA_OK(cudaMemcpyAsync(v.data(), uniq.p, sizeof(float) * m, cudaMemcpyDeviceToHost, s));
      CUDA_OK(cudaMemcpyAsync(c.data(), cum.p, sizeof(double) * m, cudaMemcpyDeviceToHost, s));
      CUDA_OK(cudaStreamSynchronize(s));
    } else {
      // cap-point summary plus the first and last distinct values (so min/max survive)
      pick_kernel<<<(cap + 255) / 256, 256, 0, s>>>(cum.p, m, cap, idx.p); ++g_kernel_launches;
      gather_kernel<<<(cap + 255) / 256, 256, 0, s>>>(uniq.p, cum.p, idx.p, cap, vsel.p, csel.p); ++g_kernel_launches;
      std::vector<float> vs(cap); std::vector<double> cs(cap); float v0; double c0;
      CUDA_OK(cudaMemcpyAsync(vs.data(), vsel.p, sizeof(float) * cap, cudaMemcpyDeviceToHost, s));
      CUDA_OK(cudaMemcpyAsync(cs.data(), csel.p, sizeof(double) * cap, cudaMemcpyDeviceToHost, s));
      CUDA_OK(cudaMemcpyAsync(&v0, uniq.p, sizeof(float), cudaMemcpyDeviceToHost, s));
      CUDA_OK(cudaMemcpyAsync(&c0, cum.p, sizeof(double), cudaMemcpyDeviceToHost, s));
      CUDA_OK(cudaStreamSynchronize(s));
      v.push_back(v0); c.push_back(c0);
      for (int k = 0; k < cap; ++k) if (vs[k] > v.back()) { v.push_back(vs[k]); c.push_back(cs[k]); }
    }
    fs.vals = v; fs.weights.resize(v.size());
    for (size_t i = 0; i < v.size(); ++i) fs.weights[i] = c[i] - (i ? c[i - 1] : 0.0);
  }
}

// Same arithmetic as oracle/gbt_oracle.c cuts_from_distinct (shared definition, independently written here).
void cuts_from_summaries(const std::vector<FeatureSummary>& sums, int max_bin, bool has_missing, HostCuts* out) {
  int nb = std::min(max_bin, 256);
  if (has_missing && nb > 255) nb = 255;
  const int F = (int)sums.size();
  out->ptrs.assign(1, 0); out->vals.clear(); out->mins.assign(F, 0.f);
  for (int f = 0; f < F; ++f) {
    const std::vector<float>& d = sums[f].vals; const std::vector<double>& cw = sums[f].weights;
    const int64_t m = (int64_t)d.size();
    if (m == 0) { out->vals.push_back(1e-5f); out->mins[f] = -1e-5f; out->ptrs.push_back((int)out->vals.size()); continue; }
    if (m <= nb) { for (int64_t i = 1; i < m; ++i) out->vals.push_back(d[i]); }
    else {
      double W = 0; for (int64_t i = 0; i < m; ++i) W += cw[i];
      double cum = 0; int64_t i = 0; float last = d[0];
      for (int k = 1; k < nb; ++k) {
        double target = W * (double)k / (double)nb;
        while (i < m && cum + cw[i] < target) { cum += cw[i]; ++i; }
        int64_t j = i + 1 < m ? i + 1 : m - 1;
        float c = d[j];
        if (c > last) { out->vals.push_back(c); last = c; }
      }
    }
    float lastv = d[m - 1];
    out->vals.push_back(lastv + (std::fabs(lastv) + 1e-5f));
    out->mins[f] = d[0] - (std::fabs(d[0]) + 1e-5f);
    out->ptrs.push_back((int)out->vals.size());
  }
}

void compute_cuts_device(const float* dX, int64_t n, int F, const float* dweights, int max_bin, bool has_missing,
                         HostCuts* out, cudaStream_t s) {
  std::vector<FeatureSummary> sums;
  // single-rank path: exact (cap = everything). The distributed path merges capped summaries first (engine).
  compute_summaries_device(dX, n, F, dweights, (int)std::min<int64_t>(n > 0 ? n : 1, (int64_t)1 << 30), &sums, s);
  cuts_from_summaries(sums, max_bin, has_missing, out);
}

}  // namespace b200
