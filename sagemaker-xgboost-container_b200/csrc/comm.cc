// comm.cc -- see comm.h
#include "comm.h"
#include <dlfcn.h>
#include <nccl.h>
#include <chrono>
#include <cstdlib>
#include <cstring>
#include <thread>
#include "engine.h"

namespace b200 {
namespace {
struct NcclApi {
  void* h = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Broadcast)(const void*, void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  ncclResult_t (*CommGetAsyncError)(ncclComm_t, ncclResult_t*) = nullptr;
  ncclResult_t (*CommAbort)(ncclComm_t) = nullptr;
};
NcclApi& api() {
  static NcclApi a;
  if (a.h) return a;
  // libnccl.so.2 resolves to the copy already loaded in the process (torch's) if there is one
  const char* names[] = {"libnccl.so.2", "libnccl.so"};
  for (const char* n : names) { a.h = dlopen(n, RTLD_NOW | RTLD_GLOBAL); if (a.h) break; }
  B200_CHECK(a.h != nullptr, std::string("cannot load libnccl.so.2: ") + (dlerror() ? dlerror() : "?"));
#define L(sym, field) a.field = reinterpret_cast<decltype(a.field)>(dlsym(a.h, sym)); B200_CHECK(a.field != nullptr, std::string("libnccl: missing symbol ") + sym)
  L("ncclGetUniqueId", GetUniqueId); L("ncclCommInitRank", CommInitRank); L("ncclCommDestroy", CommDestroy);
  L("ncclAllReduce", AllReduce); L("ncclAllGather", AllGather); L("ncclBroadcast", Broadcast); L("ncclGetErrorString", GetErrorString);
  L("ncclCommGetAsyncError", CommGetAsyncError); L("ncclCommAbort", CommAbort);
#undef L
  return a;
}
#define NCCL_OK(expr) do { ncclResult_t r_ = (expr); if (r_ != ncclSuccess) throw Error(std::string("NCCL error: ") + api().GetErrorString(r_)); } while (0)
}  // namespace

Comm& Comm::get() { static Comm c; return c; }

std::string Comm::create_unique_id() {
  ncclUniqueId id; NCCL_OK(api().GetUniqueId(&id));
  return std::string(reinterpret_cast<const char*>(&id), sizeof(id));
}
void Comm::init(const std::string& unique_id, int rank, int world) {
  finalize();
  rank_ = rank; world_ = world;
  if (world <= 1) return;
  B200_CHECK(unique_id.size() == sizeof(ncclUniqueId), "communicator: bad NCCL unique id size");
  ncclUniqueId id; memcpy(&id, unique_id.data(), sizeof(id));
  ncclComm_t c; NCCL_OK(api().CommInitRank(&c, world, id, rank));
  comm_ = c;
}
void Comm::finalize() {
  peer_reduce_close();
  if (comm_) { api().CommDestroy(static_cast<ncclComm_t>(comm_)); comm_ = nullptr; }
  rank_ = 0; world_ = 1;
}
void Comm::allreduce_sum_i64(void* buf, size_t count, cudaStream_t s) {
  if (world_ <= 1 || count == 0) return;
  NCCL_OK(api().AllReduce(buf, buf, count, ncclInt64, ncclSum, static_cast<ncclComm_t>(comm_), s));
}
void Comm::allreduce_sum_f64(void* buf, size_t count, cudaStream_t s) {
  if (world_ <= 1 || count == 0) return;
  NCCL_OK(api().AllReduce(buf, buf, count, ncclFloat64, ncclSum, static_cast<ncclComm_t>(comm_), s));
}
void Comm::allreduce_max_u32(void* buf, size_t count, cudaStream_t s) {
  if (world_ <= 1 || count == 0) return;
  NCCL_OK(api().AllReduce(buf, buf, count, ncclUint32, ncclMax, static_cast<ncclComm_t>(comm_), s));
}
void Comm::allgather_bytes(const void* send, void* recv, size_t bytes_per_rank, cudaStream_t s) {
  if (world_ <= 1) { if (send != recv) CUDA_OK(cudaMemcpyAsync(recv, send, bytes_per_rank, cudaMemcpyDeviceToDevice, s)); return; }
  NCCL_OK(api().AllGather(send, recv, bytes_per_rank, ncclUint8, static_cast<ncclComm_t>(comm_), s));
}
void Comm::sync_stream(cudaStream_t s) {
  if (world_ <= 1 || !comm_) { CUDA_OK(cudaStreamSynchronize(s)); return; }
  static const double limit = [] { const char* e = getenv("B200XGB_COLLECTIVE_TIMEOUT"); double v = e ? atof(e) : 600.0; return v > 0 ? v : 600.0; }();
  const auto t0 = std::chrono::steady_clock::now();
  for (unsigned spin = 0;; ++spin) {
    cudaError_t q = cudaStreamQuery(s);
    if (q == cudaSuccess) { peer_reduce_check(); return; }
    if (q != cudaErrorNotReady) CUDA_OK(q);
    if ((spin & 63) == 63) {
      ncclResult_t ar = ncclSuccess;
      ncclResult_t r = api().CommGetAsyncError(static_cast<ncclComm_t>(comm_), &ar);
      const double waited = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
      if (r != ncclSuccess || (ar != ncclSuccess && ar != ncclInProgress) || waited > limit) {
        const std::string why = r != ncclSuccess ? api().GetErrorString(r) : (ar != ncclSuccess && ar != ncclInProgress) ? api().GetErrorString(ar)
                                : "no progress for " + std::to_string((int)waited) + " s (B200XGB_COLLECTIVE_TIMEOUT)";
        api().CommAbort(static_cast<ncclComm_t>(comm_)); comm_ = nullptr;
        const int r_ = rank_; rank_ = 0; world_ = 1;
        throw Error("NCCL collective failed on rank " + std::to_string(r_) + ": " + why + "; the communicator was aborted");
      }
      std::this_thread::sleep_for(std::chrono::microseconds(50));
    }
  }
}
void Comm::broadcast_bytes(void* buf, size_t bytes, int root, cudaStream_t s) {
  if (world_ <= 1 || bytes == 0) return;
  NCCL_OK(api().Broadcast(buf, buf, bytes, ncclUint8, root, static_cast<ncclComm_t>(comm_), s));
}

}  // namespace b200
