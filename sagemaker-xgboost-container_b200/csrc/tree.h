// tree.h -- kernel argument blocks and launchers of the tree builder (see tree.cu, hist.cu, misc.cu).
#pragma once
#include "engine.h"

namespace b200 {

constexpr unsigned kPartTile = 2048;       // rows per partition tile (256 threads x 8 consecutive rows)

struct TrainParamDev {
  float eta, lambda, alpha, gamma, min_child_weight, max_delta_step;
  int max_depth, max_leaves;
};

struct EvalArgs {
  const GH64* hist_pool; GrowState gs; const int* cut_ptrs; const unsigned char* feat_mask;
  TrainParamDev p; int F, ngroups, tw, ntail, has_missing, level, max_level_nodes;
  const int* monotone;            // per-feature monotone constraint (-1, 0, +1), nullptr = none
  const unsigned char* node_allowed;   // interaction constraints: [cap_nodes][F] flags of the features a node may split on, nullptr = none
  float colsample_bynode; unsigned seed; const int* tree_index;     // per-node feature subset inside feat_mask (the level's set); tree index in device memory (graph replay)
};

struct ApplyArgs {
  GrowState gs; TreeArrays tree; const int* cut_ptrs; const float* cut_vals; const float* min_vals;
  TrainParamDev p; unsigned* scratch; int ngroups /* candidate blocks per node: groups + tail */, level, max_level_nodes, next_base, next_half;
  const int* monotone;            // as in EvalArgs
  // interaction constraints (upstream FeatureInteractionConstraintHost): per node the features used on its path and the features
  // it may split on ([cap_nodes][F] each), the constraint sets as a membership matrix [n_sets][F]
  unsigned char* node_path; unsigned char* node_allowed; const unsigned char* ic_sets; int n_ic_sets, F;
};

struct PartArgs {
  GrowState gs; TreeArrays tree; const uint8_t* bins_col; int64_t n; const unsigned* ridx_cur; unsigned* ridx_next;
  const float2* gp_cur; float2* gp_next;      // (g,h) pairs travel with the row ids (position order)
  const unsigned* tl_cur; unsigned* tl_next;  // ... and so do the 4 tail bin bytes of a row (nullptr when there is no 4-wide tail)
  int has_missing, level, max_level_nodes;
};

struct HistArgs {
  const uint8_t* bins;          // main: row-major [n][ngroups*32 B]
  const uint8_t* bins_tail;     // tail: row-major [n][tw B], nullptr when tw == 0
  const unsigned* tail_pos;     // tw == 4 only: the rows' tail words by POSITION (they travel with the row ids); nullptr = gather from bins_tail
  int64_t n;
  int row_stride;               // ngroups * 32
  const uint8_t* bins_gather;   // rows for the gathered passes (BinnedMatrix::bins_gather) and their stride
  int gather_stride;
  int tw;                       // tail width in bytes (0, 4, 8)
  const float2* gpair;          // (g, h) by POSITION in the row-id buffer (== by row at the root)
  const unsigned* ridx;         // row ids by segment position; nullptr = identity (root)
  const int* build_count;       // number of nodes to build
  const int* build_nid;         // their node ids
  const unsigned* build_prefix; // exclusive prefix of their row counts, [count] = total
  const unsigned* seg_begin;    // per nid
  const int* hist_slot;         // per nid
  const float* scales;          // sg, sh
  GH64* hist_pool;              // slot stride = hist_slot_entries(ngroups, tw)
  GH64* node_sum;               // per nid, accumulated only when accumulate_sum
  int ngroups;
  int ng_chunk;                 // groups per blockIdx.y chunk (set by the launcher)
  int accumulate_sum;
  int g_only;                   // constant-hessian root pass: accumulate G only (the slot already holds the cached H plane)
  int window_rows;              // rows a CTA may accumulate between two int32 overflow checks (engine.h window_rows_for)
  int force_gather;             // tests / profiling: use hist_gather_kernel even for the contiguous root pass
  unsigned long long* rows_counter;   // optional: += rows processed by this launch (profiling)
};

// grow_policy=lossguide (one expansion per iteration; tree.cu)
constexpr int kLgRootSlot = 0, kLgStageSlot = 1, kLgFirstFreeSlot = 2;     // histogram pool slots: root, all-reduce staging, then one per expansion
void launch_apply_lossguide(const ApplyArgs& a, int iter, cudaStream_t s);
void launch_lg_copy_back(const PartArgs& a, unsigned* ridx_dst, float2* gp_dst, unsigned* tl_dst, unsigned max_tiles, cudaStream_t s);
void launch_zero_build_slots(const GrowState& gs, GH64* pool, size_t slot_entries, int max_build, cudaStream_t s);
void launch_lg_stage(const GrowState& gs, GH64* pool, size_t slot_entries, int to_stage, cudaStream_t s);
void launch_hist_build(const HistArgs& a, int num_sms, cudaStream_t stream);
void hist_configure();     // one-time function attributes (must happen outside stream capture)
const char* hist_last_kernel();   // name of the kernel variant the last launch used (profiling / tests)
void launch_init_tree(const GrowState& gs, const TreeArrays& t, unsigned n, int root_slot, int max_level_nodes, cudaStream_t s);
void launch_scales(const GrowState& gs, int grad_bits, cudaStream_t s);
void launch_eval(const EvalArgs& a, int max_nodes_level, cudaStream_t s);
void launch_apply(const ApplyArgs& a, cudaStream_t s);
void launch_partition(const PartArgs& a, unsigned max_tiles, int max_nodes_level, cudaStream_t s);
void launch_update_margin(const TreeArrays& t, const int* n_nodes, const uint8_t* bins_col, int64_t n, int has_missing, float* margin, int K, int k, cudaStream_t s);
void launch_subtract(const GrowState& gs, GH64* pool, size_t slot_entries, int max_build, cudaStream_t s);

}  // namespace b200
