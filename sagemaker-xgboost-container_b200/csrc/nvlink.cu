// nvlink.cu -- in-place int64 sum all-reduce of the per-level histograms over NVLink peer memory, one kernel per call.
// "per-node gradient/hessian histograms are all-reduced over NVLink before split finding".  Round 1
// called ncclAllReduce for it; a level's histograms are <= 6.5 MB, where a collective is latency, not bandwidth: this kernel
// does the whole exchange with two flag barriers and one pass of peer loads / stores, and -- being an ordinary kernel -- it is
// captured inside the per-tree CUDA graph (NCCL calls had to cut the graph into segments).
//
// Set-up (PeerReduce::setup, collective): every rank exports its buffers (histogram pool, grow-state block) and a small flag
// page with cudaIpcGetMemHandle, the handles are all-gathered through the existing NCCL communicator, every rank maps its
// peers' buffers (cudaIpcOpenMemHandle, NVLink peer access).  Ranks on different hosts or without peer access agree (an
// all-reduce of a flag) to keep using NCCL.
//
// Kernel (two-shot): [ready barrier] rank r sums slice r of all `world` buffers (volatile peer loads) and stores the sums into
// slice r of EVERY rank's buffer (peer stores) [done barrier].  Barriers are epoch-stamped flags written into the peers' flag
// pages with release stores at system scope; the epoch lives in device memory and is advanced by the last CTA of each call, so
// a captured launch needs no per-call argument.  Sums are integers: the result is identical to NCCL's, bit for bit, at any
// world size (tests/test_multi_gpu.py runs both paths).
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#include <unistd.h>
#include "comm.h"
#include "engine.h"

namespace b200 {

constexpr int kPeerMax = 8;
constexpr int kPeerCtas = 64, kPeerThreads = 512;

struct PeerFlags {                 // one page per rank, written by its peers
  unsigned ready[kPeerMax];        // ready[src]: src's inputs of call `epoch` are complete
  unsigned done[kPeerMax];         // done[src]:  src has stored all its sums of call `epoch`
  unsigned epoch;                  // calls completed so far + 1 == stamp of the call in flight
  unsigned arrived;                // CTAs of this rank that finished their slice in the call in flight
  unsigned timed_out;              // set when a barrier wait exceeded kPeerSpinCycles (a peer died or never launched): the host turns it into an error
};
constexpr long long kPeerSpinCycles = 60ll * 2000000000ll;      // ~60 s at 2 GHz: far beyond any skew between healthy ranks

struct PeerArgs {
  long long* bufs[kPeerMax];       // the same buffer on every rank (own + mapped peers), base pointers
  PeerFlags* flags[kPeerMax];
  size_t offset;                   // element offset of the region inside the buffer
  size_t count;                    // int64 elements to reduce
  int rank, world;
};

__device__ __forceinline__ void st_release_sys(unsigned* p, unsigned v) { asm volatile("st.release.sys.global.u32 [%0], %1;" :: "l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned* p) { unsigned v; asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v; }
__device__ __forceinline__ longlong2 ld_volatile_v2(const long long* p) {
  longlong2 v; asm volatile("ld.volatile.global.v2.s64 {%0, %1}, [%2];" : "=l"(v.x), "=l"(v.y) : "l"(p) : "memory"); return v;
}

// bounded spin on a flag written by a peer: a rank that never arrives must not wedge this GPU for ever
__device__ __forceinline__ void wait_flag(const unsigned* flag, unsigned e, unsigned* timed_out) {
  const long long t0 = clock64();
  while (ld_acquire_sys(flag) < e) {
    if (clock64() - t0 > kPeerSpinCycles) { *timed_out = 1u; break; }
  }
}

__global__ void __launch_bounds__(kPeerThreads) peer_allreduce_kernel(PeerArgs a) {
  PeerFlags* mine = a.flags[a.rank];
  __shared__ unsigned s_epoch;
  if (threadIdx.x == 0) s_epoch = ld_acquire_sys(&mine->epoch);
  __syncthreads();
  const unsigned e = s_epoch;
  // ---- ready barrier: my inputs were produced by earlier kernels of this stream, hence complete; tell everyone, wait for everyone
  if (blockIdx.x == 0 && threadIdx.x < a.world) st_release_sys(&a.flags[threadIdx.x]->ready[a.rank], e);
  if (threadIdx.x < a.world) wait_flag(&mine->ready[threadIdx.x], e, &mine->timed_out);
  __syncthreads();
  // ---- my slice: [lo, hi) in units of (g,h) pairs
  const size_t pairs = a.count / 2;
  const size_t per = (pairs + a.world - 1) / a.world;
  const size_t lo = (size_t)a.rank * per, hi = lo + per < pairs ? lo + per : pairs;
  for (size_t i = lo + (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < hi; i += (size_t)gridDim.x * blockDim.x) {
    long long g = 0, h = 0;
#pragma unroll
    for (int p = 0; p < kPeerMax; ++p) if (p < a.world) { const longlong2 v = ld_volatile_v2(a.bufs[p] + a.offset + 2 * i); g += v.x; h += v.y; }
#pragma unroll
    for (int p = 0; p < kPeerMax; ++p) if (p < a.world) { longlong2* d = reinterpret_cast<longlong2*>(a.bufs[p] + a.offset + 2 * i); *d = make_longlong2(g, h); }
  }
  // ---- done barrier: the last CTA of this rank to finish publishes "rank done", CTA 0 waits for every rank
  __threadfence_system();
  __syncthreads();
  __shared__ bool s_last;
  if (threadIdx.x == 0) {
    s_last = atomicAdd(&mine->arrived, 1u) + 1u == gridDim.x;
    if (s_last) mine->arrived = 0u;                     // calls are serialised on the stream: the next one starts from zero
  }
  __syncthreads();
  if (s_last && threadIdx.x < a.world) st_release_sys(&a.flags[threadIdx.x]->done[a.rank], e);
  if (blockIdx.x == 0) {
    if (threadIdx.x < a.world) wait_flag(&mine->done[threadIdx.x], e, &mine->timed_out);
    __syncthreads();
    // every rank has stored its sums and (to get here) read its inputs: the buffer may be reused; open the next call
    if (threadIdx.x == 0) st_release_sys(&mine->epoch, e + 1u);
  }
}

// -------------------------------------------------------------------------------------------------
struct PeerBuffer { void* base = nullptr; size_t bytes = 0; void* peers[kPeerMax] = {}; };

struct PeerReduceImpl {
  bool active = false; int rank = 0, world = 1;
  std::vector<PeerBuffer> bufs;
  PeerFlags* flags[kPeerMax] = {}; PeerFlags* my_flags = nullptr;
  void close() {
    for (auto& b : bufs) for (int p = 0; p < world; ++p) if (p != rank && b.peers[p]) cudaIpcCloseMemHandle(b.peers[p]);
    for (int p = 0; p < world; ++p) if (p != rank && flags[p]) cudaIpcCloseMemHandle(flags[p]);
    if (my_flags) cudaFree(my_flags);
    bufs.clear(); memset(flags, 0, sizeof flags); my_flags = nullptr; active = false;
  }
};
static PeerReduceImpl g_peer;

void peer_reduce_close() { g_peer.close(); }

// collective over all ranks; bases/sizes: the buffers later passed to peer_allreduce_i64 (pointers inside them)
bool peer_reduce_setup(const std::vector<std::pair<void*, size_t>>& buffers, cudaStream_t s) {
  Comm& comm = Comm::get();
  g_peer.close();
  static const bool disabled = getenv("B200XGB_NO_PEER_REDUCE") != nullptr;
  const int world = comm.world(), rank = comm.rank();
  if (world <= 1) return false;
  int ok = (!disabled && world <= kPeerMax) ? 1 : 0;
  // same host? (IPC handles only work inside one node)
  char host[64] = {0}; gethostname(host, sizeof host - 1);
  struct Rec { char host[64]; cudaIpcMemHandle_t flags; cudaIpcMemHandle_t buf[4]; int nbuf; int ok; };
  Rec mine; memset(&mine, 0, sizeof mine); memcpy(mine.host, host, sizeof host);
  PeerFlags* f = nullptr;
  if (ok) {
    if (cudaMalloc(&f, 4096) != cudaSuccess) { cudaGetLastError(); ok = 0; f = nullptr; }
    else {
      PeerFlags init; memset(&init, 0, sizeof init); init.epoch = 1;
      cudaMemcpyAsync(f, &init, sizeof init, cudaMemcpyHostToDevice, s);
      if (cudaIpcGetMemHandle(&mine.flags, f) != cudaSuccess) { cudaGetLastError(); ok = 0; }
    }
  }
  mine.nbuf = (int)std::min<size_t>(buffers.size(), 4);
  for (int i = 0; ok && i < mine.nbuf; ++i) if (cudaIpcGetMemHandle(&mine.buf[i], buffers[i].first) != cudaSuccess) { cudaGetLastError(); ok = 0; }
  mine.ok = ok;
  DevBuf<unsigned char> dsend, drecv; dsend.alloc(sizeof(Rec)); drecv.alloc(sizeof(Rec) * world);
  CUDA_OK(cudaMemcpyAsync(dsend.p, &mine, sizeof(Rec), cudaMemcpyHostToDevice, s));
  comm.allgather_bytes(dsend.p, drecv.p, sizeof(Rec), s);
  std::vector<Rec> all(world);
  CUDA_OK(cudaMemcpyAsync(all.data(), drecv.p, sizeof(Rec) * world, cudaMemcpyDeviceToHost, s));
  comm.sync_stream(s);
  for (int p = 0; p < world; ++p) if (!all[p].ok || memcmp(all[p].host, host, sizeof host) != 0 || all[p].nbuf != mine.nbuf) ok = 0;
  g_peer.rank = rank; g_peer.world = world; g_peer.my_flags = f;
  if (ok) {
    g_peer.bufs.resize(mine.nbuf);
    for (int i = 0; i < mine.nbuf; ++i) { g_peer.bufs[i].base = buffers[i].first; g_peer.bufs[i].bytes = buffers[i].second; g_peer.bufs[i].peers[rank] = buffers[i].first; }
    g_peer.flags[rank] = f;
    for (int p = 0; ok && p < world; ++p) {
      if (p == rank) continue;
      void* m = nullptr;
      if (cudaIpcOpenMemHandle(&m, all[p].flags, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { cudaGetLastError(); ok = 0; break; }
      g_peer.flags[p] = static_cast<PeerFlags*>(m);
      for (int i = 0; i < mine.nbuf; ++i) {
        if (cudaIpcOpenMemHandle(&m, all[p].buf[i], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { cudaGetLastError(); ok = 0; break; }
        g_peer.bufs[i].peers[p] = m;
      }
    }
  }
  // everybody must agree (a rank that failed to map a peer would otherwise call NCCL while the others spin in the kernel)
  DevBuf<unsigned> agree; agree.alloc(1);
  unsigned v = ok ? 0u : 1u;
  CUDA_OK(cudaMemcpyAsync(agree.p, &v, 4, cudaMemcpyHostToDevice, s));
  comm.allreduce_max_u32(agree.p, 1, s);
  CUDA_OK(cudaMemcpyAsync(&v, agree.p, 4, cudaMemcpyDeviceToHost, s));
  comm.sync_stream(s);
  if (v != 0u) { g_peer.close(); return false; }
  g_peer.active = true;
  return true;
}

bool peer_reduce_active() { return g_peer.active; }

// after a stream wait: did a peer barrier give up?  (cheap: 4 bytes, only while the peer path is active)
void peer_reduce_check() {
  if (!g_peer.active || !g_peer.my_flags) return;
  unsigned v = 0;
  if (cudaMemcpy(&v, &g_peer.my_flags->timed_out, 4, cudaMemcpyDeviceToHost) != cudaSuccess) { cudaGetLastError(); return; }
  if (v) { g_peer.active = false; throw Error("NVLink peer all-reduce timed out waiting for rank(s) that never reached the collective"); }
}

// in-place sum over ranks of `count` int64 at `ptr` (inside a registered buffer); false = not registered / inactive (caller uses NCCL)
bool peer_allreduce_i64(long long* ptr, size_t count, cudaStream_t s) {
  if (!g_peer.active || count == 0 || (count & 1)) return false;
  for (auto& b : g_peer.bufs) {
    char* base = static_cast<char*>(b.base);
    if (reinterpret_cast<char*>(ptr) >= base && reinterpret_cast<char*>(ptr + count) <= base + b.bytes && ((reinterpret_cast<char*>(ptr) - base) & 15) == 0) {
      PeerArgs a; memset(&a, 0, sizeof a);
      for (int p = 0; p < g_peer.world; ++p) { a.bufs[p] = static_cast<long long*>(b.peers[p]); a.flags[p] = g_peer.flags[p]; }
      a.offset = (size_t)(reinterpret_cast<char*>(ptr) - base) / 8; a.count = count; a.rank = g_peer.rank; a.world = g_peer.world;
      const size_t pairs_per_rank = (count / 2 + g_peer.world - 1) / g_peer.world;
      int grid = (int)std::min<size_t>(kPeerCtas, (pairs_per_rank + kPeerThreads - 1) / kPeerThreads);
      if (grid < 1) grid = 1;
      peer_allreduce_kernel<<<grid, kPeerThreads, 0, s>>>(a); ++g_kernel_launches;
      CUDA_OK(cudaGetLastError());
      return true;
    }
  }
  return false;
}

}  // namespace b200
