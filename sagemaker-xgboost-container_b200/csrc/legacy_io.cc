// legacy_io.cc -- reader for xgboost's pre-JSON binary model format (host code, no CUDA).
// The container's serving path still meets such files: serve_utils.get_loaded_booster (algorithm_mode/serve_utils.py:171-197)
// tries pickle.load, then Booster.load_model, on whatever the customer's model.tar.gz holds; the reference keeps two fixtures
// of that era (test/resources/models/saved_booster = Booster.save_model of xgboost 1.0, pickled_model = a pickled Booster whose
// state carries the same bytes behind a "CONFIG-offset:" prefix; test_serve_utils.py:77-90, test_multiple_model_endpoint.py).
//
// Layout restated from upstream's published format [UPSTREAM-RECALL dmlc/xgboost v1.x: src/learner.cc LearnerModelParamLegacy +
// LearnerIO::LoadModel, src/gbm/gbtree_model.{h,cc}, include/xgboost/tree_model.h TreeParam / RegTree::Node / RTreeNodeStat];
// every field offset below was checked against the two reference fixtures (tests/test_legacy_model.py):
//   ["binf"]                                   optional 4-byte magic of very old files
//   LearnerModelParamLegacy   136 B            f32 base_score, u32 num_feature, i32 num_class, i32 contain_extra_attrs,
//                                              i32 contain_eval_metrics, u32 major, u32 minor, reserved
//   string name_obj, string name_gbm           u64 length + bytes
//   GBTreeModelParam          160 B            i32 num_trees, i32 num_roots, i32 num_feature, i32 pad, i64 num_pbuffer,
//                                              i32 num_output_group, i32 size_leaf_vector, reserved
//   num_trees x { TreeParam 148 B              i32 num_roots, num_nodes, num_deleted, max_depth, num_feature, size_leaf_vector
//                 num_nodes x Node 20 B        i32 parent (bit 31 = is left child), i32 left, i32 right,
//                                              u32 split index (bit 31 = default left), f32 leaf value | split condition
//                 num_nodes x Stat 16 B        f32 loss_chg, f32 sum_hess, f32 base_weight, i32 leaf_child_cnt
//                 [vector<f32> leaf_vector]    only when size_leaf_vector != 0 }
//   num_trees x i32 tree_info
//   [u64 n, n x (string key, string value)]    when contain_extra_attrs; 1.0 keeps the objective's JSON config under "objective"
//   [string max_delta_step]                    only for count:poisson in files older than 1.0
//   [u64 n, n x string]                        when contain_eval_metrics
// The result is the 3.x model document (json.h), so Booster::model_from_json is the only consumer of tree arrays.
#include <cstring>
#include "booster.h"

namespace b200 {

namespace {
struct Cursor {
  const unsigned char* p; size_t n, off = 0;
  void need(size_t k) const { B200_CHECK(k <= n - off, "legacy binary model: truncated (need " + std::to_string(k) + " bytes at offset " + std::to_string(off) + " of " + std::to_string(n) + ")"); }
  template <typename T> T get() { need(sizeof(T)); T v; memcpy(&v, p + off, sizeof(T)); off += sizeof(T); return v; }
  void skip(size_t k) { need(k); off += k; }
  std::string str() { uint64_t len = get<uint64_t>(); B200_CHECK(len <= n - off, "legacy binary model: string length runs past the end of the buffer"); std::string s((const char*)p + off, (size_t)len); off += (size_t)len; return s; }
};
JPtr S(const std::string& s) { return JValue::Str(s); }
}  // namespace

bool looks_like_legacy_binary(const char* buf, size_t len) {
  if (len >= 4 && !memcmp(buf, "binf", 4)) return true;
  if (len < 136 + 16) return false;
  // a JSON / UBJSON document starts with '{'; the legacy header starts with base_score (a finite float) and small integers
  if (buf[0] == '{') return false;
  uint32_t nf, extra, evalm; int32_t nc; memcpy(&nf, buf + 4, 4); memcpy(&nc, buf + 8, 4); memcpy(&extra, buf + 12, 4); memcpy(&evalm, buf + 16, 4);
  return nc >= 0 && nc < (1 << 20) && extra <= 1 && evalm <= 1;
}

JPtr legacy_binary_to_doc(const char* buf, size_t len) {
  Cursor c{(const unsigned char*)buf, len};
  if (len >= 4 && !memcmp(buf, "binf", 4)) c.skip(4);
  const size_t head = c.off;
  const float base_score = c.get<float>(); const uint32_t num_feature = c.get<uint32_t>(); const int32_t num_class = c.get<int32_t>();
  const int32_t extra_attrs = c.get<int32_t>(), eval_metrics = c.get<int32_t>();
  const uint32_t major = c.get<uint32_t>(), minor = c.get<uint32_t>();
  c.off = head; c.skip(136);
  const std::string name_obj = c.str(), name_gbm = c.str();
  B200_CHECK(name_gbm == "gbtree", "legacy binary model: only gbtree boosters can be loaded (got " + name_gbm + ")");
  const size_t gp = c.off;
  const int32_t num_trees = c.get<int32_t>(); c.get<int32_t>(); c.get<int32_t>(); c.get<int32_t>();
  const int64_t num_pbuffer = c.get<int64_t>(); c.get<int32_t>(); const int32_t gbm_leaf_vec = c.get<int32_t>();
  c.off = gp; c.skip(160);
  B200_CHECK(num_trees >= 0, "legacy binary model: negative tree count");
  (void)num_pbuffer; (void)gbm_leaf_vec;          // prediction-buffer bookkeeping of pre-0.6 files: no bytes follow it in the stream

  JPtr trees = JValue::Array();
  for (int t = 0; t < num_trees; ++t) {
    const size_t tp = c.off;
    const int32_t roots = c.get<int32_t>(), nn = c.get<int32_t>(), ndel = c.get<int32_t>(); c.get<int32_t>(); const int32_t tree_nf = c.get<int32_t>(), leaf_vec = c.get<int32_t>();
    c.off = tp; c.skip(148);
    B200_CHECK(roots == 1, "legacy binary model: trees with several roots are not supported");
    B200_CHECK(nn >= 1 && (size_t)nn <= (len - c.off) / 36, "legacy binary model: node count runs past the end of the buffer");
    std::vector<int32_t> left(nn), right(nn), parent(nn), sidx(nn); std::vector<uint8_t> dleft(nn); std::vector<float> cond(nn), loss(nn), hess(nn), bw(nn);
    for (int i = 0; i < nn; ++i) {
      const int32_t par = c.get<int32_t>(), l = c.get<int32_t>(), r = c.get<int32_t>(); const uint32_t s = c.get<uint32_t>(); const float v = c.get<float>();
      const bool deleted = s == 0xffffffffu;                  // pruned node (RegTree::Node::MarkDelete): unreachable, kept as a zero leaf
      const bool leaf = deleted || l == -1;
      left[i] = leaf ? -1 : l; right[i] = leaf ? -1 : r;
      parent[i] = par == -1 ? 2147483647 : (par & 0x7fffffff);
      sidx[i] = leaf ? 0 : (int32_t)(s & 0x7fffffffu); dleft[i] = leaf ? 0 : (uint8_t)(s >> 31);
      cond[i] = deleted ? 0.0f : v;
      if (!leaf) B200_CHECK(l > 0 && l < nn && r > 0 && r < nn, "legacy binary model: child index out of range");
    }
    for (int i = 0; i < nn; ++i) { loss[i] = c.get<float>(); hess[i] = c.get<float>(); bw[i] = c.get<float>(); c.get<int32_t>(); }
    if (leaf_vec != 0) { const uint64_t k = c.get<uint64_t>(); B200_CHECK(k <= (len - c.off) / 4, "legacy binary model: leaf vector runs past the end of the buffer"); c.skip((size_t)k * 4); }
    JPtr tj = JValue::Object();
    tj->set("base_weights", JValue::F32(bw));
    tj->set("categories", JValue::I32({})); tj->set("categories_nodes", JValue::I32({})); tj->set("categories_segments", JValue::I64({})); tj->set("categories_sizes", JValue::I64({}));
    tj->set("default_left", JValue::U8(dleft)); tj->set("id", JValue::Int(t));
    tj->set("left_children", JValue::I32(left)); tj->set("loss_changes", JValue::F32(loss)); tj->set("parents", JValue::I32(parent));
    tj->set("right_children", JValue::I32(right)); tj->set("split_conditions", JValue::F32(cond)); tj->set("split_indices", JValue::I32(sidx));
    tj->set("split_type", JValue::U8(std::vector<uint8_t>(nn, 0))); tj->set("sum_hessian", JValue::F32(hess));
    JPtr tpj = JValue::Object(); tpj->set("num_deleted", S(std::to_string(ndel))); tpj->set("num_feature", S(std::to_string(tree_nf)));
    tpj->set("num_nodes", S(std::to_string(nn))); tpj->set("size_leaf_vector", S("1"));
    tj->set("tree_param", tpj);
    trees->arr.push_back(tj);
  }
  std::vector<int32_t> tree_info(num_trees);
  for (int t = 0; t < num_trees; ++t) tree_info[t] = c.get<int32_t>();

  JPtr attributes = JValue::Object(); JPtr objective;
  if (extra_attrs != 0) {
    const uint64_t k = c.get<uint64_t>();
    for (uint64_t i = 0; i < k; ++i) {
      const std::string key = c.str(), val = c.str();
      if (key == "objective" && !val.empty() && val[0] == '{') {         // 1.0's binary format parks the objective's JSON config here
        try { objective = JsonReader(val.data(), val.size()).parse(); } catch (...) { objective = nullptr; }
        if (objective && objective->type == JValue::kObject && objective->has("name")) continue;
        objective = nullptr;
      }
      if (key.rfind("SAVED_PARAM_", 0) == 0) continue;                    // 1.0 parks predictor / gpu_id here: run-time settings, not model attributes
      attributes->set(key, S(val));
    }
  }
  JPtr poisson;
  if (major < 1 && name_obj == "count:poisson" && c.off < len) { const std::string mds = c.str(); poisson = JValue::Object(); poisson->set("max_delta_step", S(mds)); }
  if (eval_metrics != 0 && c.off < len) { const uint64_t k = c.get<uint64_t>(); for (uint64_t i = 0; i < k; ++i) c.str(); }
  (void)minor;

  const int K = num_class > 1 ? num_class : 1;
  B200_CHECK(num_trees % K == 0, "legacy binary model: tree count is not a multiple of num_class");
  if (!objective) {
    objective = JValue::Object(); objective->set("name", S(name_obj));
    if (name_obj.rfind("multi:", 0) == 0) { JPtr sp = JValue::Object(); sp->set("num_class", S(std::to_string(K))); objective->set("softmax_multiclass_param", sp); }
    else if (poisson) objective->set("poisson_regression_param", poisson);
  }
  JPtr doc = JValue::Object(); JPtr learner = JValue::Object();
  learner->set("attributes", attributes); learner->set("feature_names", JValue::Array()); learner->set("feature_types", JValue::Array());
  JPtr gb = JValue::Object(); JPtr model = JValue::Object();
  JPtr gmp = JValue::Object(); gmp->set("num_parallel_tree", S("1")); gmp->set("num_trees", S(std::to_string(num_trees))); model->set("gbtree_model_param", gmp);
  std::vector<int32_t> indptr; for (int r = 0; r <= num_trees / K; ++r) indptr.push_back(r * K);
  model->set("iteration_indptr", JValue::I32(indptr)); model->set("tree_info", JValue::I32(tree_info)); model->set("trees", trees);
  gb->set("model", model); gb->set("name", S("gbtree")); learner->set("gradient_booster", gb);
  JPtr lmp = JValue::Object();
  char bs[48]; snprintf(bs, sizeof bs, "%.9g", (double)base_score);
  lmp->set("base_score", S(bs)); lmp->set("boost_from_average", S("1")); lmp->set("num_class", S(std::to_string(num_class > 1 ? num_class : 0)));
  lmp->set("num_feature", S(std::to_string(num_feature))); lmp->set("num_target", S("1"));
  learner->set("learner_model_param", lmp); learner->set("objective", objective);
  doc->set("learner", learner);
  JPtr ver = JValue::Array(); ver->arr = {JValue::Int((int64_t)major), JValue::Int((int64_t)minor), JValue::Int(0)}; doc->set("version", ver);
  return doc;
}

// Booster.__getstate__ of xgboost 1.0 - 1.x ("CONFIG-offset:" + i64 offset of the JSON config, then the binary model at byte 22):
// returns the model section, or {nullptr, 0} when the buffer is not of that form.
std::pair<const char*, size_t> legacy_serialized_model_section(const char* buf, size_t len) {
  static const char kTag[] = "CONFIG-offset:";
  const size_t tl = sizeof(kTag) - 1;
  if (len < tl + 8 || memcmp(buf, kTag, tl) != 0) return {nullptr, 0};
  int64_t off; memcpy(&off, buf + tl, 8);
  B200_CHECK(off >= 0 && (size_t)off <= len - tl - 8, "legacy serialized booster: config offset out of range");
  return {buf + tl + 8, (size_t)off};
}

}  // namespace b200
