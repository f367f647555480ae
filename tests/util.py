"""Shared helpers of the test-suite: seeded synthetic data (SURVEY.md section 8d recipes) and model comparison."""
import numpy as np


def synth(n, F, seed, kind="reg", K=1, quantised=True, missing_frac=0.0):
    rng = np.random.default_rng(seed)
    X = rng.standard_normal((n, F), dtype=np.float32)
    if quantised:      # 256 levels per feature: cuts are unambiguous for any correct quantile algorithm
        X = (np.round(np.clip(X, -4, 4 - 1 / 32) * 32) / 32).astype(np.float32)
    if kind == "reg":
        beta = (rng.standard_normal(F) / np.sqrt(F)).astype(np.float32)
        y = X @ beta + 0.1 * rng.standard_normal(n).astype(np.float32)
    elif kind == "bin":
        beta = (rng.standard_normal(F) / np.sqrt(F)).astype(np.float32)
        z = X @ beta + 0.5 * rng.standard_normal(n).astype(np.float32)
        y = (1.0 / (1.0 + np.exp(-z)) > rng.random(n)).astype(np.float32)
    elif kind == "multi":
        beta = (rng.standard_normal((F, K)) / np.sqrt(F)).astype(np.float32)
        y = np.argmax(X @ beta + rng.standard_normal((n, K)).astype(np.float32), axis=1).astype(np.float32)
    elif kind in ("pos", "count"):        # positive targets (gamma / tweedie / squaredlogerror) and Poisson counts
        beta = (rng.standard_normal(F) / np.sqrt(F)).astype(np.float32)
        mu = np.exp(0.5 * (X @ beta) + 0.1 * rng.standard_normal(n).astype(np.float32))
        y = rng.poisson(mu).astype(np.float32) if kind == "count" else mu.astype(np.float32)
    else:
        raise ValueError(kind)
    if missing_frac > 0:
        X = X.copy()
        X[rng.random((n, F)) < missing_frac] = np.nan
    return np.ascontiguousarray(X, np.float32), np.ascontiguousarray(y, np.float32)


def assert_same_structure(m_gpu, m_ref):
    assert len(m_gpu["tree_info"]) == len(m_ref["tree_info"]), "number of trees differs"
    np.testing.assert_array_equal(m_gpu["tree_offset"], m_ref["tree_offset"], err_msg="tree sizes differ")
    np.testing.assert_array_equal(m_gpu["tree_info"], m_ref["tree_info"])
    for k in ("left", "right", "parent", "split_index", "default_left"):
        np.testing.assert_array_equal(m_gpu[k], m_ref[k], err_msg="tree array %s differs" % k)
    internal = m_ref["left"] != -1
    np.testing.assert_array_equal(m_gpu["split_cond"][internal], m_ref["split_cond"][internal], err_msg="split thresholds differ")


def max_leaf_diff(m_gpu, m_ref):
    leaf = m_ref["left"] == -1
    return float(np.abs(m_gpu["split_cond"][leaf] - m_ref["split_cond"][leaf]).max())


def first_structural_difference(m_gpu, m_ref):
    nt = min(len(m_gpu["tree_info"]), len(m_ref["tree_info"]))
    for t in range(nt):
        a0, a1 = int(m_gpu["tree_offset"][t]), int(m_gpu["tree_offset"][t + 1])
        b0, b1 = int(m_ref["tree_offset"][t]), int(m_ref["tree_offset"][t + 1])
        if a1 - a0 != b1 - b0:
            return t
        for k in ("left", "split_index"):
            if not np.array_equal(m_gpu[k][a0:a1], m_ref[k][b0:b1]):
                return t
        internal = m_ref["left"][b0:b1] != -1
        if not np.array_equal(m_gpu["split_cond"][a0:a1][internal], m_ref["split_cond"][b0:b1][internal]):
            return t
    return None
