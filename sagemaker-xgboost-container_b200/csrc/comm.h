// comm.h -- in-engine collective used by the tree builder: NCCL over NVLink, one process per GPU.
// Replaces the histogram AllReduce that upstream xgboost performs inside libxgboost (src/collective/*),
// which the container reaches through distributed.py:42-109 (rabit_run) / distributed_gpu_training.py:184-192.
// libnccl is resolved with dlopen at first use so the library also loads on hosts without NCCL / without a GPU.
#pragma once
#include <cuda_runtime.h>
#include <cstddef>
#include <cstdint>
#include <string>
#include <utility>
#include <vector>

namespace b200 {

class Comm {
 public:
  static Comm& get();
  int rank() const { return rank_; }
  int world() const { return world_; }
  bool distributed() const { return world_ > 1; }
  // rank 0 creates the id (128 bytes) and ships it to the others through the Python-side bootstrap
  // (torch.distributed / tracker); every rank then calls init.
  static std::string create_unique_id();
  void init(const std::string& unique_id, int rank, int world);
  void finalize();
  void allreduce_sum_i64(void* buf, size_t count, cudaStream_t s);
  void allreduce_sum_f64(void* buf, size_t count, cudaStream_t s);
  void allreduce_max_u32(void* buf, size_t count, cudaStream_t s);
  void allgather_bytes(const void* send, void* recv, size_t bytes_per_rank, cudaStream_t s);
  void broadcast_bytes(void* buf, size_t bytes, int root, cudaStream_t s);
  // Wait for `s` like cudaStreamSynchronize, but while ranks are connected poll the communicator for asynchronous errors
  // and give up after B200XGB_COLLECTIVE_TIMEOUT seconds (default 600): a failed or hung collective aborts the
  // communicator and surfaces as an Error (-> XGBoostError) instead of blocking the job for ever.
  void sync_stream(cudaStream_t s);
 private:
  int rank_ = 0, world_ = 1;
  void* comm_ = nullptr;
};

// nvlink.cu: single-kernel int64 sum all-reduce over NVLink peer memory (CUDA IPC mappings of the registered buffers).
// setup is collective; peer_allreduce_i64 returns false when the pointer is not inside a registered buffer or peers could
// not be mapped (different hosts, no peer access, B200XGB_NO_PEER_REDUCE): the caller then uses NCCL.
bool peer_reduce_setup(const std::vector<std::pair<void*, size_t>>& buffers, cudaStream_t s);
bool peer_reduce_active();
bool peer_allreduce_i64(long long* ptr, size_t count, cudaStream_t s);
void peer_reduce_close();
void peer_reduce_check();           // throws when a peer barrier timed out (called after stream waits)

}  // namespace b200
