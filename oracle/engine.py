"""Oracle-backed engine with the same Python-level interface as sagemaker-xgboost-container_b200/backend.CudaBackend.

TEST INFRASTRUCTURE: lets the CPU test-suite drive the package's host-side logic (DMatrix loaders, train(), callbacks,
checkpoint resume, model IO) and the reference container's own `sagemaker_train` (BASELINE config 1) without a GPU, by
monkey-patching `backend._BACKEND`.  The product never selects this engine: `get_backend()` only ever creates CudaBackend.
"""
import json
import pickle

import numpy as np

from . import gbt_oracle as O
from . import legacy_model
from . import ubjson


class _DM:
    def __init__(self, X):
        self.X = np.ascontiguousarray(X, np.float32)
        self.info = {"label": np.zeros(0, np.float32), "weight": np.zeros(0, np.float32), "base_margin": np.zeros(0, np.float32)}
        self.names = {"feature_name": [], "feature_type": []}


class _Bst:
    def __init__(self):
        self.params = {}
        self.metrics = []
        self.attrs = {}
        self.names = {"feature_name": [], "feature_type": []}
        self.loaded = None          # Model loaded from a buffer (continue training / inference)
        self.trainer = None
        self.trainer_dm = None
        self.num_feature = 0

    def model(self):
        parts = []
        if self.loaded is not None:
            parts.append(self.loaded)
        if self.trainer is not None:
            parts.append(self.trainer.model())
        if not parts:
            m = O.Model()
            for k, dt in (("left", np.int32), ("right", np.int32), ("parent", np.int32), ("split_index", np.int32), ("split_bin", np.int32),
                          ("default_left", np.uint8), ("split_cond", np.float32), ("base_weight", np.float32), ("loss_chg", np.float32),
                          ("sum_hess", np.float32)):
                m[k] = np.zeros(0, dt)
            m["tree_offset"] = np.zeros(1, np.int64); m["tree_info"] = np.zeros(0, np.int32)
            m["base_score"] = float(self.params.get("base_score", 0.5)); m["num_class"] = int(self.params.get("num_class", 1) or 1)
            m["num_feature"] = self.num_feature; m["objective"] = self.objective()
            return m
        if len(parts) == 1:
            return parts[0]
        a, b = parts
        m = O.Model()
        for k in ("left", "right", "parent", "split_index", "split_bin", "default_left", "split_cond", "base_weight", "loss_chg", "sum_hess"):
            m[k] = np.concatenate([a[k], b[k]])
        m["tree_offset"] = np.concatenate([a["tree_offset"], b["tree_offset"][1:] + a["tree_offset"][-1]])
        m["tree_info"] = np.concatenate([a["tree_info"], b["tree_info"]])
        for k in ("base_score", "num_class", "num_feature", "objective"):
            m[k] = a[k]
        return m

    def objective(self):
        o = self.params.get("objective", "reg:squarederror")
        return "reg:squarederror" if o == "reg:linear" else o

    def K(self):
        return max(1, int(self.params.get("num_class", 1) or 1)) if self.objective().startswith("multi") else 1


def _model_to_doc(m, attrs, names):
    K = int(m.get("num_class", 1))
    trees = []
    for t in range(len(m["tree_info"])):
        a, b = int(m["tree_offset"][t]), int(m["tree_offset"][t + 1])
        nn = b - a
        trees.append({
            "base_weights": m["base_weight"][a:b], "categories": np.zeros(0, np.int32), "categories_nodes": np.zeros(0, np.int32),
            "categories_segments": np.zeros(0, np.int64), "categories_sizes": np.zeros(0, np.int64), "default_left": m["default_left"][a:b],
            "id": t, "left_children": m["left"][a:b], "loss_changes": m["loss_chg"][a:b], "parents": m["parent"][a:b],
            "right_children": m["right"][a:b], "split_conditions": m["split_cond"][a:b], "split_indices": m["split_index"][a:b],
            "split_type": np.zeros(nn, np.uint8), "sum_hessian": m["sum_hess"][a:b],
            "tree_param": {"num_deleted": "0", "num_feature": str(m["num_feature"]), "num_nodes": str(nn), "size_leaf_vector": "1"}})
    rounds = len(trees) // max(1, K)
    return {"learner": {"attributes": dict(attrs), "feature_names": list(names["feature_name"]), "feature_types": list(names["feature_type"]),
                        "gradient_booster": {"model": {"gbtree_model_param": {"num_parallel_tree": "1", "num_trees": str(len(trees))},
                                                       "iteration_indptr": np.arange(0, (rounds + 1) * K, K, dtype=np.int32),
                                                       "tree_info": np.asarray(m["tree_info"], np.int32), "trees": trees}, "name": "gbtree"},
                        "learner_model_param": {"base_score": "[%.9E]" % m["base_score"], "boost_from_average": "1",
                                                "num_class": str(K if K > 1 else 0), "num_feature": str(m["num_feature"]), "num_target": "1"},
                        "objective": {"name": m["objective"], "reg_loss_param": {"scale_pos_weight": "1"}}},
            "version": [3, 0, 5]}


def _jsonable(v):
    if isinstance(v, dict):
        return {k: _jsonable(x) for k, x in v.items()}
    if isinstance(v, np.ndarray):
        return v.tolist()
    if isinstance(v, (list, tuple)):
        return [_jsonable(x) for x in v]
    if isinstance(v, (np.integer,)):
        return int(v)
    if isinstance(v, (np.floating,)):
        return float(v)
    return v


class OracleBackend:
    name = "oracle"

    def __init__(self, error_cls=ValueError):
        self.err = error_cls

    # ---- DMatrix
    def dmatrix_from_dense(self, arr, missing):
        X = np.array(arr, dtype=np.float32, copy=True)
        if X.ndim != 2:
            raise ValueError("Expecting 2 dimensional numpy.ndarray, got: %s" % (X.shape,))
        if missing is not None and missing == missing:
            X[X == missing] = np.nan
        return _DM(X)

    def dmatrix_from_csr(self, indptr, indices, data, ncol):
        nrow = len(indptr) - 1
        F = max(int(ncol), int(indices.max()) + 1 if len(indices) else 0)
        X = np.full((nrow, F), np.nan, np.float32)
        for r in range(nrow):
            sl = slice(int(indptr[r]), int(indptr[r + 1]))
            X[r, indices[sl]] = data[sl]
        return _DM(X)

    def dmatrix_free(self, h):
        pass

    def dmatrix_num_row(self, h):
        return h.X.shape[0]

    def dmatrix_num_col(self, h):
        return h.X.shape[1]

    def dmatrix_set_float_info(self, h, field, arr):
        if field not in h.info:
            raise self.err("Unknown float field name: " + field)
        h.info[field] = np.ascontiguousarray(arr, np.float32).reshape(-1)

    def dmatrix_get_float_info(self, h, field):
        return h.info[field].copy()

    def dmatrix_slice(self, h, idx):
        d = _DM(h.X[idx])
        n = h.X.shape[0]
        for k, v in h.info.items():
            if len(v):
                per = len(v) // n
                d.info[k] = v.reshape(n, per)[idx].reshape(-1)
        d.names = {k: list(v) for k, v in h.names.items()}
        return d

    def dmatrix_set_str_info(self, h, field, values):
        h.names[field] = list(values or [])

    def dmatrix_get_str_info(self, h, field):
        return list(h.names[field])

    # ---- Booster
    def booster_create(self, dmat_handles=()):
        return _Bst()

    def booster_free(self, h):
        pass

    def booster_set_param(self, h, k, v):
        if k == "eval_metric":
            if v not in h.metrics:
                h.metrics.append(v)
        else:
            h.params[k] = v

    def _gathered(self, dh):
        """Multi-rank runs on this TEST engine (world_size-2 CPU tests of the launcher, tests/test_multi_gpu_launcher.py): every
        rank assembles the row shards of all ranks and trains / evaluates on the whole matrix, which is what the CUDA engine's
        histogram all-reduce amounts to.  Collective: all ranks must call this for the same matrices in the same order."""
        from sagemaker_xgboost_container_b200 import collective
        world, rank = collective.get_world_size(), collective.get_rank()
        if world <= 1:
            return dh
        if getattr(dh, "_full", None) is None:
            parts = []
            for r in range(world):
                mine = {"X": dh.X.tobytes(), "shape": list(dh.X.shape), "info": {k: v.tobytes() for k, v in dh.info.items()}} if r == rank else None
                parts.append(collective.broadcast(mine, r))
            full = _DM(np.concatenate([np.frombuffer(p["X"], np.float32).reshape(p["shape"]) for p in parts]))
            for k in dh.info:
                full.info[k] = np.concatenate([np.frombuffer(p["info"][k], np.float32) for p in parts])
            dh._full = full
        return dh._full

    def _ensure_trainer(self, h, dh):
        if h.trainer is not None and h.trainer_dm is dh:
            return
        if h.trainer is not None:
            raise self.err("oracle engine: training matrix changed")
        local = dh
        dh = self._gathered(dh)
        y = dh.info["label"]
        if len(y) != dh.X.shape[0]:
            raise self.err("Check failed: preds.size() == info.labels_.size() : labels are not correctly provided")
        params = {k: (float(v) if isinstance(v, str) and k not in ("objective", "tree_method", "grow_policy", "booster") else v) for k, v in h.params.items()}
        params["objective"] = h.objective()
        for k in ("max_depth", "num_class", "max_bin", "seed", "max_leaves"):
            if k in params:
                params[k] = int(float(params[k]))
        if h.loaded is not None:
            params["base_score"] = h.loaded["base_score"]
        w = dh.info["weight"] if len(dh.info["weight"]) else None
        try:
            h.trainer = O.Trainer(params, X=dh.X, y=y, weights=w)
        except ValueError as e:
            raise self.err(str(e))
        h.trainer_dm = local
        h.num_feature = dh.X.shape[1]
        if h.loaded is not None and len(h.loaded["tree_info"]):
            h.trainer.set_margins(O.predict_margin(h.loaded, dh.X))

    def booster_update(self, h, it, dh):
        self._ensure_trainer(h, dh)
        try:
            h.trainer.update()
        except ValueError as e:
            raise self.err("Check failed: " + str(e) if "logistic" in str(e) else str(e))

    def _margin(self, h, dh, tree_begin=0, tree_end=None):
        m = h.model()
        if len(dh.info["base_margin"]):
            bm = dh.info["base_margin"].reshape(dh.X.shape[0], -1)
            out = O.predict_margin(m, dh.X, tree_begin, tree_end, base_margin=0.0) + bm
        else:
            out = O.predict_margin(m, dh.X, tree_begin, tree_end)
        return m, out

    def booster_eval(self, h, it, dhs, names):
        metrics = h.metrics or [{"reg:squarederror": "rmse", "reg:logistic": "rmse", "binary:logistic": "logloss", "binary:logitraw": "logloss"}.get(h.objective(), "mlogloss")]
        msg = "[%d]" % it
        for dh, name in zip(dhs, names):
            dh = self._gathered(dh)
            m, margin = self._margin(h, dh)
            y = dh.info["label"].astype(np.float64)
            w = dh.info["weight"].astype(np.float64) if len(dh.info["weight"]) else np.ones(len(y))
            pred = O.transform(m, margin).astype(np.float64)
            for mn in metrics:
                thr = 0.5
                base = mn
                if mn.startswith("error@"):
                    base, thr = "error", float(mn[6:])
                p = pred[:, 0] if pred.ndim == 2 and pred.shape[1] == 1 else pred
                if base in ("rmse", "mse"):
                    v = np.sum(w * (p - y) ** 2) / w.sum()
                    v = np.sqrt(v) if mn == "rmse" else v
                elif base == "mae":
                    v = np.sum(w * np.abs(p - y)) / w.sum()
                elif base == "logloss":
                    pc = np.clip(p, 1e-16, 1 - 1e-16)
                    v = np.sum(w * -(y * np.log(pc) + (1 - y) * np.log(1 - pc))) / w.sum()
                elif base == "error":
                    v = np.sum(w * np.where(p > thr, 1 - y, y)) / w.sum()
                elif base == "merror":
                    v = np.sum(w * (np.argmax(margin, axis=1) != y.astype(int))) / w.sum()
                elif base == "mlogloss":
                    pk = pred[np.arange(len(y)), y.astype(int)]
                    v = np.sum(w * -np.log(np.maximum(pk, 1e-16))) / w.sum()
                else:
                    raise self.err("Unknown metric function " + mn)
                msg += "\t%s-%s:%.17g" % (name, mn, v)
        return msg

    def booster_predict(self, h, dh, cfg):
        m = h.model()
        K = int(m.get("num_class", 1))
        nt = len(m["tree_info"])
        rounds = nt // max(1, K)
        b, e = int(cfg.get("iteration_begin", 0)), int(cfg.get("iteration_end", 0))
        if e == 0:
            e = rounds
        n = dh.X.shape[0]
        if cfg.get("type", 0) == 6:
            return O.predict_leaf(m, dh.X, b * K, e * K).astype(np.float32)
        _, margin = self._margin(h, dh, b * K, e * K)
        if cfg.get("type", 0) == 1:
            out = margin
        else:
            out = O.transform(m, margin)
        out = np.asarray(out, np.float32)
        if out.ndim == 2 and out.shape[1] == 1 and not cfg.get("strict_shape"):
            out = out[:, 0]
        elif out.ndim == 1 and cfg.get("strict_shape"):
            out = out.reshape(n, 1)
        return out

    def booster_save_raw(self, h, fmt):
        doc = _model_to_doc(h.model(), h.attrs, h.names)
        if fmt == "json":
            return json.dumps(_jsonable(doc)).encode()
        return ubjson.dumps(doc)

    def booster_load_raw(self, h, buf):
        buf = bytes(buf)
        if legacy_model.is_legacy(buf):                  # pre-JSON binary file / pickled 1.x state (serve_utils.py:171-197)
            doc = legacy_model.to_document(buf)
        else:
            doc = json.loads(buf.decode()) if buf[:2] in (b'{"', b"{ ", b"{\n") else ubjson.loads(buf)
        if "Model" in doc:
            doc = doc["Model"]
        m = ubjson.model_from_xgb_json(doc)
        h.loaded, h.trainer, h.trainer_dm = m, None, None
        h.params["objective"] = m["objective"]
        if m["num_class"] > 1:
            h.params["num_class"] = m["num_class"]
        h.num_feature = m["num_feature"]
        h.attrs = dict(doc["learner"].get("attributes", {}))
        h.names = {"feature_name": list(doc["learner"].get("feature_names", [])), "feature_type": list(doc["learner"].get("feature_types", []))}

    def booster_serialize(self, h):
        return ubjson.dumps({"Model": _model_to_doc(h.model(), h.attrs, h.names), "Config": json.loads(self.booster_save_config(h))})

    def booster_unserialize(self, h, buf):
        if legacy_model.is_legacy(bytes(buf)):
            return self.booster_load_raw(h, buf)
        doc = ubjson.loads(bytes(buf))
        self.booster_load_raw(h, ubjson.dumps(doc["Model"]))
        self.booster_load_config(h, json.dumps(_jsonable(doc["Config"])))

    def booster_save_config(self, h):
        K = h.K()
        return json.dumps({"learner": {"objective": {"name": h.objective()}, "learner_model_param": {"num_class": str(K if K > 1 else 0), "num_feature": str(h.num_feature)},
                                       "gradient_booster": {"name": "gbtree", "tree_train_param": {k: str(v) for k, v in h.params.items()}},
                                       "metrics": [{"name": m} for m in h.metrics]}, "version": [3, 0, 5]})

    def booster_load_config(self, h, s):
        doc = json.loads(s)
        h.params.update(doc["learner"]["gradient_booster"].get("tree_train_param", {}))
        h.metrics = [m["name"] for m in doc["learner"].get("metrics", [])]

    def booster_num_features(self, h):
        return h.num_feature

    def booster_boosted_rounds(self, h):
        return len(h.model()["tree_info"]) // h.K()

    def booster_slice(self, h, begin, end, step):
        m = h.model()
        K = h.K()
        keep = [r * K + k for r in range(begin, end, step) for k in range(K)]
        out = _Bst()
        out.params, out.metrics, out.attrs, out.names, out.num_feature = dict(h.params), list(h.metrics), dict(h.attrs), {k: list(v) for k, v in h.names.items()}, h.num_feature
        mm = O.Model()
        offs = [0]
        acc = {k: [] for k in ("left", "right", "parent", "split_index", "split_bin", "default_left", "split_cond", "base_weight", "loss_chg", "sum_hess")}
        for t in keep:
            a, b = int(m["tree_offset"][t]), int(m["tree_offset"][t + 1])
            offs.append(offs[-1] + b - a)
            for k in acc:
                acc[k].append(m[k][a:b])
        for k in acc:
            mm[k] = np.concatenate(acc[k]) if acc[k] else m[k][:0]
        mm["tree_offset"] = np.asarray(offs, np.int64); mm["tree_info"] = np.asarray([m["tree_info"][t] for t in keep], np.int32)
        for k in ("base_score", "num_class", "num_feature", "objective"):
            mm[k] = m[k]
        out.loaded = mm
        return out

    def booster_get_attr(self, h, key):
        return h.attrs.get(key)

    def booster_set_attr(self, h, key, value):
        if value is None:
            h.attrs.pop(key, None)
        else:
            h.attrs[key] = str(value)

    def booster_attr_names(self, h):
        return list(h.attrs)

    def booster_set_str_info(self, h, field, values):
        h.names[field] = list(values or [])

    def booster_get_str_info(self, h, field):
        return list(h.names[field])

    def booster_export_model(self, h):
        return h.model()

    # ---- collective (single process)
    def comm_rank(self):
        return 0

    def comm_world(self):
        return 1

    def synchronize(self):
        pass

    def launch_count(self):
        return 0
