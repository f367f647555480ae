"""TEST INFRASTRUCTURE (oracle): an independent numpy reader of xgboost's pre-JSON binary model format, used to check the
product's C++ reader (sagemaker-xgboost-container_b200/csrc/legacy_io.cc) field by field and to let the CPU test engine open
the reference's legacy fixtures.  Never imported by the product.

Format restated from upstream's published layout [UPSTREAM-RECALL dmlc/xgboost v1.x src/learner.cc LearnerModelParamLegacy,
src/gbm/gbtree_model.h GBTreeModelParam, include/xgboost/tree_model.h TreeParam / RegTree::Node / RTreeNodeStat] and pinned
on the two files the reference holds: test/resources/models/saved_booster/xgboost-model (Booster.save_model, xgboost 1.0:
multi:softprob, 3 classes, 4 features, 60 trees) and pickled_model/xgboost-model (pickle of xgboost.core.Booster whose
state["handle"] is "CONFIG-offset:" + i64 + the same 24 603 model bytes + a JSON config) -- the loaders
serve_utils.get_loaded_booster (algorithm_mode/serve_utils.py:171-197) has to cope with.
"""
import json
import struct

import numpy as np

_NODE = np.dtype([("parent", "<i4"), ("left", "<i4"), ("right", "<i4"), ("sindex", "<u4"), ("value", "<f4")])
_STAT = np.dtype([("loss_chg", "<f4"), ("sum_hess", "<f4"), ("base_weight", "<f4"), ("leaf_child_cnt", "<i4")])
_TAG = b"CONFIG-offset:"


def model_section(buf):
    """bytes of a pickled 1.x Booster's state["handle"] -> the binary model inside it (None if `buf` is not of that form)"""
    buf = bytes(buf)
    if not buf.startswith(_TAG):
        return None
    (size,) = struct.unpack_from("<q", buf, len(_TAG))
    return buf[len(_TAG) + 8: len(_TAG) + 8 + size]


def is_legacy(buf):
    buf = bytes(buf[:32])
    return buf.startswith(b"binf") or buf.startswith(_TAG) or (len(buf) >= 20 and buf[:1] != b"{")


def to_document(buf):
    """legacy binary model -> the 3.x model document (same keys as a parsed UBJSON model file)"""
    buf = bytes(buf)
    sect = model_section(buf)
    if sect is not None:
        buf = sect
    p = 4 if buf.startswith(b"binf") else 0
    base_score, num_feature, num_class, extra_attrs, eval_metrics, major, minor = struct.unpack_from("<fIiiiII", buf, p)
    p += 136

    def rstr():
        nonlocal p
        (n,) = struct.unpack_from("<Q", buf, p)
        s = buf[p + 8: p + 8 + n]
        assert len(s) == n, "string runs past the end"
        p += 8 + n
        return s.decode("utf-8", "replace")

    name_obj, name_gbm = rstr(), rstr()
    assert name_gbm == "gbtree", name_gbm
    (num_trees,) = struct.unpack_from("<i", buf, p)
    p += 160
    trees = []
    for t in range(num_trees):
        roots, nn, ndel, _depth, tree_nf, leaf_vec = struct.unpack_from("<6i", buf, p)
        assert roots == 1
        p += 148
        nodes = np.frombuffer(buf, _NODE, nn, p)
        p += nn * _NODE.itemsize
        stats = np.frombuffer(buf, _STAT, nn, p)
        p += nn * _STAT.itemsize
        if leaf_vec != 0:
            (k,) = struct.unpack_from("<Q", buf, p)
            p += 8 + 4 * k
        deleted = nodes["sindex"] == 0xFFFFFFFF
        leaf = deleted | (nodes["left"] == -1)
        trees.append({
            "base_weights": stats["base_weight"].astype(np.float32), "default_left": np.where(leaf, 0, nodes["sindex"] >> 31).astype(np.uint8),
            "id": t, "left_children": np.where(leaf, -1, nodes["left"]).astype(np.int32), "right_children": np.where(leaf, -1, nodes["right"]).astype(np.int32),
            "loss_changes": stats["loss_chg"].astype(np.float32),
            "parents": np.where(nodes["parent"] == -1, 2147483647, nodes["parent"] & 0x7FFFFFFF).astype(np.int32),
            "split_conditions": np.where(deleted, 0, nodes["value"]).astype(np.float32),
            "split_indices": np.where(leaf, 0, nodes["sindex"] & 0x7FFFFFFF).astype(np.int32),
            "split_type": np.zeros(nn, np.uint8), "sum_hessian": stats["sum_hess"].astype(np.float32),
            "tree_param": {"num_deleted": str(ndel), "num_feature": str(tree_nf), "num_nodes": str(nn), "size_leaf_vector": "1"}})
    tree_info = np.frombuffer(buf, "<i4", num_trees, p).astype(np.int32)
    p += 4 * num_trees
    attributes, objective = {}, None
    if extra_attrs:
        (k,) = struct.unpack_from("<Q", buf, p)
        p += 8
        for _ in range(k):
            key, val = rstr(), rstr()
            if key == "objective" and val.startswith("{"):
                objective = json.loads(val)
            elif not key.startswith("SAVED_PARAM_"):
                attributes[key] = val
    K = num_class if num_class > 1 else 1
    if objective is None:
        objective = {"name": name_obj}
        if name_obj.startswith("multi:"):
            objective["softmax_multiclass_param"] = {"num_class": str(K)}
    return {"learner": {"attributes": attributes, "feature_names": [], "feature_types": [],
                        "gradient_booster": {"model": {"gbtree_model_param": {"num_parallel_tree": "1", "num_trees": str(num_trees)},
                                                       "iteration_indptr": np.arange(0, num_trees + 1, K, dtype=np.int32), "tree_info": tree_info, "trees": trees},
                                             "name": "gbtree"},
                        "learner_model_param": {"base_score": repr(float(np.float32(base_score))), "boost_from_average": "1", "num_class": str(num_class if num_class > 1 else 0),
                                                "num_feature": str(num_feature), "num_target": "1"},
                        "objective": objective},
            "version": [int(major), int(minor), 0]}
