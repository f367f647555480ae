// csv.cu -- device-side CSV -> float32 matrix for the serving path.
// Replaces the Python split + np.array(...).astype(float) of the container's encoder.csv_to_dmatrix (encoder.py:31-52),
// reached from algorithm_mode/serve_utils.py:121-131 (parse_content_data) for every text/csv invocation request.
//
// Semantics kept from that code path: rows are separated by '\n' (payload already stripped of leading/trailing
// whitespace), fields by one delimiter character, an empty field is NaN, every row must have the same number of
// fields, a field is what Python's float() accepts (decimal, exponent, inf/nan in any case).  Values take the same two
// roundings: text -> nearest double (Clinger's exact fast path: <= 19 significant digits below 2^53 and |exp10| <= 22,
// one IEEE multiply or divide) -> nearest float32 (what xgb.DMatrix does with a float64 array).  A field outside the fast
// path (or malformed) raises a flag and the caller falls back to the host parser, so results never differ.
//
// Kernels: (1) newline positions (flag + CUB select), (2) one thread per row walks its fields.  Text is read once from
// HBM (L1-cached byte loads; rows are short and contiguous per thread).
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cub/cub.cuh>
#include <cstring>
#include "booster.h"

namespace b200 {

struct IsNewline {
  const char* text;
  __host__ __device__ bool operator()(const int64_t& i) const { return text[i] == '\n'; }
};

__device__ __forceinline__ bool is_space(char c) { return c == ' ' || c == '\t' || c == '\r'; }

__device__ __forceinline__ char lower(char c) { return (c >= 'A' && c <= 'Z') ? (char)(c + 32) : c; }

// parse [p, e) as a Python-float literal; returns false when the token is malformed or needs the slow path
__device__ bool parse_field(const char* p, const char* e, float* out) {
  while (p < e && is_space(*p)) ++p;
  while (e > p && is_space(e[-1])) --e;
  if (p == e) { *out = __int_as_float(0x7fc00000); return true; }                 // empty field -> NaN (encoder.py:31-32)
  bool neg = false;
  if (*p == '+' || *p == '-') { neg = *p == '-'; ++p; if (p == e) return false; }
  const int len = (int)(e - p);
  if (len == 3 && lower(p[0]) == 'n' && lower(p[1]) == 'a' && lower(p[2]) == 'n') { *out = __int_as_float(0x7fc00000); return true; }
  if ((len == 3 && lower(p[0]) == 'i' && lower(p[1]) == 'n' && lower(p[2]) == 'f') ||
      (len == 8 && lower(p[0]) == 'i' && lower(p[1]) == 'n' && lower(p[2]) == 'f' && lower(p[3]) == 'i' && lower(p[4]) == 'n' && lower(p[5]) == 'i' &&
       lower(p[6]) == 't' && lower(p[7]) == 'y')) { *out = neg ? __int_as_float(0xff800000) : __int_as_float(0x7f800000); return true; }
  unsigned long long mant = 0; int digits = 0, sig = 0, exp10 = 0; bool any = false, dropped = false;
  while (p < e && *p >= '0' && *p <= '9') {
    any = true;
    if (sig < 19) { mant = mant * 10ull + (unsigned)(*p - '0'); if (mant != 0) ++sig; } else { ++exp10; if (*p != '0') dropped = true; }
    ++p; ++digits;
  }
  if (p < e && *p == '.') {
    ++p;
    while (p < e && *p >= '0' && *p <= '9') {
      any = true;
      if (sig < 19) { mant = mant * 10ull + (unsigned)(*p - '0'); if (mant != 0) ++sig; --exp10; } else if (*p != '0') dropped = true;
      ++p;
    }
  }
  if (!any) return false;
  if (p < e && (*p == 'e' || *p == 'E')) {
    ++p; bool eneg = false;
    if (p < e && (*p == '+' || *p == '-')) { eneg = *p == '-'; ++p; }
    if (p == e) return false;
    int ev = 0;
    while (p < e && *p >= '0' && *p <= '9') { if (ev < 100000) ev = ev * 10 + (*p - '0'); ++p; }
    exp10 += eneg ? -ev : ev;
  }
  if (p != e) return false;                                                       // trailing junk (Python's float() would raise)
  if (mant == 0) { *out = neg ? -0.0f : 0.0f; return true; }
  if (dropped || mant >= (1ull << 53) || exp10 > 22 || exp10 < -22) return false;   // outside the exact fast path: host parser decides
  const double p10[23] = {1e0, 1e1, 1e2, 1e3, 1e4, 1e5, 1e6, 1e7, 1e8, 1e9, 1e10, 1e11, 1e12, 1e13, 1e14, 1e15, 1e16, 1e17, 1e18, 1e19, 1e20, 1e21, 1e22};
  double d = (double)mant;                                                        // exact: mant < 2^53
  d = exp10 >= 0 ? d * p10[exp10] : d / p10[-exp10];                              // one correctly rounded IEEE operation
  *out = (float)(neg ? -d : d);
  return true;
}

// one thread per row: row r covers [start, end) where start = r ? nl[r-1] + 1 : 0 and end = r < n-1 ? nl[r] : len
__global__ void __launch_bounds__(256) csv_parse_kernel(const char* text, int64_t len, const int64_t* nl, int64_t n, int F, char delim, float* out, int* err) {
  const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n) return;
  const char* p = text + (r ? nl[r - 1] + 1 : 0);
  const char* e = text + (r < n - 1 ? nl[r] : len);
  float* o = out + r * F;
  int f = 0;
  const char* tok = p;
  for (const char* q = p;; ++q) {
    if (q == e || *q == delim) {
      if (f < F) { float v; if (!parse_field(tok, q, &v)) { atomicMax(err, 2); v = 0.f; } o[f] = v; }
      ++f; tok = q + 1;
      if (q == e) break;
    }
  }
  if (f != F) atomicMax(err, 1);                                                   // ragged row
}

// newline count: 16 B per thread per step
__global__ void __launch_bounds__(256) count_newlines_kernel(const char* text, int64_t len, unsigned long long* out) {
  unsigned long long c = 0;
  const int64_t nvec = len / 16;
  const uint4* v = reinterpret_cast<const uint4*>(text);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += (int64_t)gridDim.x * blockDim.x) {
    const uint4 w = v[i];
    const unsigned ws[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const unsigned x = ws[k] ^ 0x0a0a0a0au;                                     // zero byte <=> '\n'
      c += __popc(((x - 0x01010101u) & ~x & 0x80808080u));
    }
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) for (int64_t i = nvec * 16; i < len; ++i) c += text[i] == '\n';
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
  if ((threadIdx.x & 31) == 0 && c) atomicAdd(out, c);
}

__global__ void fill_value_kernel(float* X, int64_t count, float v) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (int64_t)gridDim.x * blockDim.x) X[i] = v;
}

// request-sized scratch kept across calls (a serving process parses one body after another): no cudaMalloc per request
struct CsvScratch { DevBuf<char> text; DevBuf<int64_t> nl; DevBuf<unsigned char> tmp; DevBuf<unsigned long long> cnt; DevBuf<int> err; };
static CsvScratch& csv_scratch() { static thread_local CsvScratch s; return s; }

// returns 0 = ok, 1 = ragged rows, 2 = a field outside the exact fast path / malformed (caller falls back to the host parser)
int parse_csv_device(const char* h_text, int64_t len, char delim, int F, int64_t* n_rows_out, DevBuf<float>* X, cudaStream_t s) {
  CsvScratch& sc = csv_scratch();
  static const bool prof = getenv("B200XGB_CSV_PROFILE") != nullptr;          // stage times on stderr (microbench/csv_stages.py)
  auto t_last = std::chrono::steady_clock::now();
  auto lap = [&](const char* what) {
    if (!prof) return;
    cudaStreamSynchronize(s);
    auto now = std::chrono::steady_clock::now();
    fprintf(stderr, "[csv] %-22s %8.3f ms\n", what, std::chrono::duration<double, std::milli>(now - t_last).count());
    t_last = now;
  };
  sc.text.ensure((size_t)len + 16); sc.cnt.ensure(2); sc.err.ensure(1);
  CUDA_OK(cudaMemcpyAsync(sc.text.p, h_text, (size_t)len, cudaMemcpyHostToDevice, s));
  lap("h2d of the text");
  CUDA_OK(cudaMemsetAsync(sc.cnt.p, 0, 16, s));
  CUDA_OK(cudaMemsetAsync(sc.err.p, 0, 4, s));
  count_newlines_kernel<<<148 * 8, 256, 0, s>>>(sc.text.p, len, sc.cnt.p); ++g_kernel_launches;
  CUDA_OK(cudaGetLastError());
  unsigned long long nnl = 0;
  CUDA_OK(cudaMemcpyAsync(&nnl, sc.cnt.p, sizeof(nnl), cudaMemcpyDeviceToHost, s));
  CUDA_OK(cudaStreamSynchronize(s));
  lap("count newlines");
  const int64_t n = (int64_t)nnl + 1;
  *n_rows_out = n;
  sc.nl.ensure((size_t)nnl + 1);
  if (nnl > 0) {                                                                   // positions of the newlines, in order
    cub::CountingInputIterator<int64_t> idx(0);
    IsNewline pred{sc.text.p};
    size_t tmp_bytes = 0;
    long long* d_num = reinterpret_cast<long long*>(sc.cnt.p + 1);
    CUDA_OK(cub::DeviceSelect::If(nullptr, tmp_bytes, idx, sc.nl.p, d_num, len, pred, s));
    sc.tmp.ensure(tmp_bytes);
    CUDA_OK(cub::DeviceSelect::If(sc.tmp.p, tmp_bytes, idx, sc.nl.p, d_num, len, pred, s));
    ++g_kernel_launches;
  }
  lap("newline positions");
  X->alloc((size_t)n * F);
  lap("alloc X");
  csv_parse_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(sc.text.p, len, sc.nl.p, n, F, delim, X->p, sc.err.p); ++g_kernel_launches;
  CUDA_OK(cudaGetLastError());
  int err = 0;
  CUDA_OK(cudaMemcpyAsync(&err, sc.err.p, sizeof(int), cudaMemcpyDeviceToHost, s));
  CUDA_OK(cudaStreamSynchronize(s));
  lap("parse kernel");
  return err;
}

// ---------------------------------------------------------------------------------------------------------------------
// libsvm request bodies ("label idx:val idx:val ..." per line): serve_utils._get_sparse_matrix_from_libsvm
// (algorithm_mode/serve_utils.py:94-118, per-token Python loop -> COO -> csr_matrix -> DMatrix: absent entries are MISSING) and
// encoder.libsvm_to_dmatrix (encoder.py:54-86, script mode: dense zeros, absent entries are 0.0).  Both shift the indices to
// 0-based when the smallest index in the body is >= 1.  One thread per line, two passes over the text: (1) index range and
// entry check, (2) values into the pre-filled matrix.  Anything the two Python routes treat specially -- an index that is not
// plain digits, a value outside the exact fast path or empty, an index repeated inside a line (COO sums it, the dict keeps the
// last), a token with a second ':' -- raises the fallback flag and the caller takes the reference's own host route.
struct LibsvmRange { int min_idx, max_idx; unsigned long long entries; int err; int last_row_entries; };

template <bool kFill>
__global__ void __launch_bounds__(256) libsvm_kernel(const char* text, int64_t len, const int64_t* nl, int64_t n, LibsvmRange* rg, int whitespace_mode,
                                                     float* X, int F, int shift, float absent) {
  const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n) return;
  const char* p = text + (r ? nl[r - 1] + 1 : 0);
  const char* e = text + (r < n - 1 ? nl[r] : len);
  int lo = 0x7fffffff, hi = -1, cnt = 0; bool bad = false;
  const char* q = p;
  while (q < e) {
    // token = run of non-separator characters; serve_utils splits on ' ' only, the encoder on any whitespace
    while (q < e && (*q == ' ' || (whitespace_mode && (*q == '\t' || *q == '\r' || *q == '\f' || *q == '\v')))) ++q;
    const char* t = q;
    while (q < e && !(*q == ' ' || (whitespace_mode && (*q == '\t' || *q == '\r' || *q == '\f' || *q == '\v')))) ++q;
    if (t == q) break;
    const char* c = t; while (c < q && *c != ':') ++c;
    if (c == q) continue;                                                        // no colon: the label (or junk both routes ignore)
    int idx = 0; bool ok = c > t && (c - t) <= 9;
    for (const char* d = t; d < c && ok; ++d) { if (*d < '0' || *d > '9') ok = false; else idx = idx * 10 + (*d - '0'); }
    const char* v = c + 1;
    for (const char* d = v; d < q && ok; ++d) if (*d == ':' || *d == '_') ok = false;       // second colon / digit separators: host decides
    float val = 0.f;
    if (ok) { if (v == q || !parse_field(v, q, &val)) ok = false; else if (v < q && (is_space(*v) || is_space(q[-1]))) ok = false; }
    if (!ok) { bad = true; continue; }
    ++cnt; lo = min(lo, idx); hi = max(hi, idx);
    if (kFill) {
      float* slot = X + r * F + (idx - shift);
      // serve_utils' COO -> CSR conversion SUMS an index repeated inside a line (the encoder's dict keeps the last one, which is
      // what this sequential walk does): with NaN as the fill value a slot that is no longer NaN has been written before
      if (!whitespace_mode && (val != val || *slot == *slot)) bad = true;
      *slot = val;
    }
  }
  if (bad) atomicMax(&rg->err, 2);
  if (!kFill) {
    if (cnt) { atomicMin(&rg->min_idx, lo); atomicMax(&rg->max_idx, hi); atomicAdd(&rg->entries, (unsigned long long)cnt); }
    if (r == n - 1) rg->last_row_entries = cnt;
  }
}

// status: 0 ok, 2 host route needed (see above), 3 no entries at all (both routes special-case it on the host)
int parse_libsvm_device(const char* h_text, int64_t len, int whitespace_mode, float absent, int64_t* n_rows_out, int* F_out, DevBuf<float>* X, cudaStream_t s) {
  CsvScratch& sc = csv_scratch();
  sc.text.ensure((size_t)len + 16); sc.cnt.ensure(2); sc.err.ensure(16);
  CUDA_OK(cudaMemcpyAsync(sc.text.p, h_text, (size_t)len, cudaMemcpyHostToDevice, s));
  CUDA_OK(cudaMemsetAsync(sc.cnt.p, 0, 16, s));
  count_newlines_kernel<<<148 * 8, 256, 0, s>>>(sc.text.p, len, sc.cnt.p); ++g_kernel_launches;
  CUDA_OK(cudaGetLastError());
  unsigned long long nnl = 0;
  CUDA_OK(cudaMemcpyAsync(&nnl, sc.cnt.p, sizeof(nnl), cudaMemcpyDeviceToHost, s));
  CUDA_OK(cudaStreamSynchronize(s));
  const int64_t n = (int64_t)nnl + 1;
  sc.nl.ensure((size_t)nnl + 1);
  if (nnl > 0) {
    cub::CountingInputIterator<int64_t> idx(0);
    IsNewline pred{sc.text.p};
    size_t tmp_bytes = 0;
    long long* d_num = reinterpret_cast<long long*>(sc.cnt.p + 1);
    CUDA_OK(cub::DeviceSelect::If(nullptr, tmp_bytes, idx, sc.nl.p, d_num, len, pred, s));
    sc.tmp.ensure(tmp_bytes);
    CUDA_OK(cub::DeviceSelect::If(sc.tmp.p, tmp_bytes, idx, sc.nl.p, d_num, len, pred, s));
    ++g_kernel_launches;
  }
  static_assert(sizeof(LibsvmRange) <= 16 * sizeof(int), "range block lives in the err scratch");
  LibsvmRange* rg = reinterpret_cast<LibsvmRange*>(sc.err.p);
  LibsvmRange init{0x7fffffff, -1, 0ull, 0, 0};
  CUDA_OK(cudaMemcpyAsync(rg, &init, sizeof init, cudaMemcpyHostToDevice, s));
  const unsigned grid = (unsigned)((n + 255) / 256);
  libsvm_kernel<false><<<grid, 256, 0, s>>>(sc.text.p, len, sc.nl.p, n, rg, whitespace_mode, nullptr, 0, 0, absent); ++g_kernel_launches;
  CUDA_OK(cudaGetLastError());
  LibsvmRange h{};
  CUDA_OK(cudaMemcpyAsync(&h, rg, sizeof h, cudaMemcpyDeviceToHost, s));
  CUDA_OK(cudaStreamSynchronize(s));
  if (h.err != 0) return 2;
  if (h.entries == 0) return 3;
  if (!whitespace_mode && h.last_row_entries == 0) return 2;      // csr_matrix((data, (row, col))) infers its row count: trailing empty lines vanish there
  const int shift = h.min_idx >= 1 ? 1 : 0;
  const int F = h.max_idx - shift + 1;
  B200_CHECK((double)n * (double)F < 4e9, "libsvm body: the dense matrix would exceed 4e9 entries");
  X->alloc((size_t)n * F);
  fill_value_kernel<<<148 * 8, 256, 0, s>>>(X->p, (int64_t)n * F, absent); ++g_kernel_launches;
  libsvm_kernel<true><<<grid, 256, 0, s>>>(sc.text.p, len, sc.nl.p, n, rg, whitespace_mode, X->p, F, shift, absent); ++g_kernel_launches;
  CUDA_OK(cudaGetLastError());
  CUDA_OK(cudaMemcpyAsync(&h, rg, sizeof h, cudaMemcpyDeviceToHost, s));
  CUDA_OK(cudaStreamSynchronize(s));
  if (h.err != 0) return 2;
  *n_rows_out = n; *F_out = F;
  return 0;
}

std::unique_ptr<DMatrix> DMatrix::from_libsvm_text(const char* text, int64_t len, int whitespace_mode, float absent, int* status) {
  auto dm = std::make_unique<DMatrix>();
  cudaStream_t s = engine_stream();
  int64_t n = 0; int F = 0;
  *status = parse_libsvm_device(text, len, whitespace_mode, absent, &n, &F, &dm->X, s);
  if (*status != 0) return nullptr;
  B200_CHECK(n < (int64_t)0x7fffffff, "DMatrix: more than 2^31-1 rows per GPU are not supported");
  dm->n = n; dm->F = F;
  dm->finish_upload(std::nanf(""));
  return dm;
}

// training channels (data_utils.py:289-318: "?format=csv&label_column=0[&weight_column=1]"): the label / weight columns
// leave the parsed matrix on the device, the remaining columns are compacted in place order
__global__ void __launch_bounds__(256) split_columns_kernel(const float* in, int64_t n, int Fin, int label_col, int weight_col, float* out, float* y, float* w) {
  const int Fout = Fin - (label_col >= 0 ? 1 : 0) - (weight_col >= 0 ? 1 : 0);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n * Fin; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / Fin; const int c = (int)(i - r * Fin);
    const float v = in[i];
    if (c == label_col) y[r] = v;
    else if (c == weight_col) w[r] = v;
    else out[r * Fout + c - (label_col >= 0 && c > label_col ? 1 : 0) - (weight_col >= 0 && c > weight_col ? 1 : 0)] = v;
  }
}

std::unique_ptr<DMatrix> DMatrix::from_csv_text_labeled(const char* text, int64_t len, char delim, int label_col, int weight_col, int* status) {
  int Fin = 1;
  for (int64_t i = 0; i < len && text[i] != '\n'; ++i) if (text[i] == delim) ++Fin;
  B200_CHECK(label_col < Fin && weight_col < Fin && (label_col < 0 || label_col != weight_col), "CSV: label_column / weight_column out of range");
  cudaStream_t s = engine_stream();
  DevBuf<float> raw; int64_t n = 0;
  *status = parse_csv_device(text, len, delim, Fin, &n, &raw, s);
  if (*status != 0) return nullptr;
  B200_CHECK(n < (int64_t)0x7fffffff, "DMatrix: more than 2^31-1 rows per GPU are not supported");
  auto dm = std::make_unique<DMatrix>();
  const int Fout = Fin - (label_col >= 0 ? 1 : 0) - (weight_col >= 0 ? 1 : 0);
  dm->n = n; dm->F = Fout;
  dm->X.alloc((size_t)n * std::max(Fout, 0));
  DevBuf<float> dy, dw; dy.alloc(label_col >= 0 ? n : 0); dw.alloc(weight_col >= 0 ? n : 0);
  const int grid = (int)std::min<int64_t>((n * Fin + 255) / 256, 148 * 32);
  split_columns_kernel<<<grid, 256, 0, s>>>(raw.p, n, Fin, label_col, weight_col, dm->X.p, dy.p, dw.p); ++g_kernel_launches;
  CUDA_OK(cudaGetLastError());
  std::vector<float> hy(label_col >= 0 ? n : 0), hw(weight_col >= 0 ? n : 0);
  if (!hy.empty()) CUDA_OK(cudaMemcpyAsync(hy.data(), dy.p, sizeof(float) * n, cudaMemcpyDeviceToHost, s));
  if (!hw.empty()) CUDA_OK(cudaMemcpyAsync(hw.data(), dw.p, sizeof(float) * n, cudaMemcpyDeviceToHost, s));
  CUDA_OK(cudaStreamSynchronize(s));
  dm->finish_upload(std::nanf(""));
  if (!hy.empty()) dm->set_float_info("label", hy.data(), hy.size());
  if (!hw.empty()) dm->set_float_info("weight", hw.data(), hw.size());
  return dm;
}

std::unique_ptr<DMatrix> DMatrix::from_csv_text(const char* text, int64_t len, char delim, int* status) {
  // columns from the first line (the container sniffs the delimiter there too, encoder.py:46-48)
  int F = 1;
  for (int64_t i = 0; i < len && text[i] != '\n'; ++i) if (text[i] == delim) ++F;
  auto dm = std::make_unique<DMatrix>();
  cudaStream_t s = engine_stream();
  int64_t n = 0;
  *status = parse_csv_device(text, len, delim, F, &n, &dm->X, s);
  if (*status != 0) return nullptr;
  B200_CHECK(n < (int64_t)0x7fffffff, "DMatrix: more than 2^31-1 rows per GPU are not supported");
  dm->n = n; dm->F = F;
  dm->finish_upload(std::nanf(""));
  return dm;
}

}  // namespace b200
