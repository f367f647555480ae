"""CPU-only checks of the oracle pieces added in round 2 (objectives, loss-guided growth, brute-force Shapley values), so that
the GPU parity tests that lean on them compare against something that was itself checked: gradients against numeric derivatives
of the published loss functions, Shapley values against the additivity axiom on the reference's own model fixture, loss-guided
growth against invariants that tie it to depth-wise growth."""
import ctypes as C
import os

import numpy as np
import pytest

from oracle import gbt_oracle as O
from oracle import ubjson
from util import synth

G = os.path.join(os.path.dirname(__file__), "golden")


def _gradient(params, margins, y):
    p = O.make_params(params)
    m = np.ascontiguousarray(margins, np.float32)
    gp = np.zeros((len(m), 2), np.float32)
    rc = O.lib().orc_gradient(C.byref(p), O._p(m), O._p(np.ascontiguousarray(y, np.float32)), None, len(m), O._p(gp))
    assert rc == 0
    return gp


LOSSES = {          # loss(margin m, label y): the objective functions as published in the xgboost documentation
    "reg:squaredlogerror": (lambda m, y, a: 0.5 * (np.log1p(m) - np.log1p(y)) ** 2, {}, "pos"),
    "reg:pseudohubererror": (lambda m, y, a: a ** 2 * (np.sqrt(1 + ((m - y) / a) ** 2) - 1), dict(huber_slope=0.8), "reg"),
    "count:poisson": (lambda m, y, a: np.exp(m) - y * m, {}, "count"),
    "reg:gamma": (lambda m, y, a: y / np.exp(m) + m, {}, "pos"),
    "reg:tweedie": (lambda m, y, a: -y * np.exp((1 - a) * m) / (1 - a) + np.exp((2 - a) * m) / (2 - a), dict(tweedie_variance_power=1.3), "count"),
}


@pytest.mark.parametrize("objective", sorted(LOSSES))
def test_gradients_are_the_derivatives_of_the_published_losses(objective):
    loss, hp, kind = LOSSES[objective]
    aux = hp.get("huber_slope", hp.get("tweedie_variance_power", 0.0))
    _, y = synth(400, 4, 3, kind)
    rng = np.random.default_rng(1)
    m = (rng.random(400) * 1.5 + 0.1).astype(np.float32)          # positive margins: inside squaredlogerror's domain
    gp = _gradient(dict(objective=objective, **hp), m, y)
    m64, y64, eps = m.astype(np.float64), y.astype(np.float64), 1e-4
    g_num = (loss(m64 + eps, y64, aux) - loss(m64 - eps, y64, aux)) / (2 * eps)
    np.testing.assert_allclose(gp[:, 0], g_num, rtol=2e-4, atol=2e-5)
    h_num = (loss(m64 + eps, y64, aux) - 2 * loss(m64, y64, aux) + loss(m64 - eps, y64, aux)) / eps ** 2
    if objective == "count:poisson":             # upstream's hessian is exp(m + max_delta_step), an upper bound of the true one by design
        np.testing.assert_allclose(gp[:, 1], np.exp(m64 + 0.7), rtol=1e-5)
    elif objective == "reg:squaredlogerror":     # clamped at 1e-6 from below
        np.testing.assert_allclose(gp[:, 1], np.maximum(h_num, 1e-6), rtol=5e-3, atol=5e-4)
    else:
        np.testing.assert_allclose(gp[:, 1], h_num, rtol=5e-3, atol=5e-4)


def test_hinge_gradient_and_label_errors():
    y = np.array([0, 1, 1, 0], np.float32); m = np.array([-2.0, 0.3, 1.5, 0.2], np.float32)
    gp = _gradient(dict(objective="binary:hinge"), m, y)
    np.testing.assert_array_equal(gp[:, 0], np.array([0.0, -1.0, 0.0, 1.0], np.float32))     # inside the margin: -y', else 0
    assert gp[0, 1] > 0 and gp[0, 1] < 1e-30 and gp[1, 1] == 1.0
    p = O.make_params(dict(objective="reg:gamma"))
    bad = np.array([1.0, 0.0], np.float32); gpb = np.zeros((2, 2), np.float32)
    assert O.lib().orc_gradient(C.byref(p), O._p(np.zeros(2, np.float32)), O._p(bad), None, 2, O._p(gpb)) == -5


def test_bruteforce_shapley_values_are_additive_on_the_reference_fixture():
    """Efficiency axiom on the reference's own model (50 trees, 8 features, real sum_hessian covers): the contributions of a row
    add up to its margin, the bias column is the cover-weighted mean of the trees plus the base score."""
    doc = ubjson.load(os.path.join(G, "abalone_xgboost-model.ubj"))
    m = ubjson.model_from_xgb_json(doc)
    rng = np.random.default_rng(5)
    X = rng.random((40, 8)).astype(np.float32)
    X[rng.random((40, 8)) < 0.1] = np.nan
    phi = O.shap_bruteforce(m, X)
    assert phi.shape == (40, 1, 9)
    np.testing.assert_allclose(phi.sum(-1), O.predict_margin(m, X), rtol=0, atol=2e-5)
    assert np.ptp(phi[:, 0, 8]) < 1e-12                       # the bias does not depend on the row
    # a feature no tree splits on gets exactly zero
    used = set(int(f) for f, l in zip(m["split_index"], m["left"]) if l != -1)
    for f in range(8):
        if f not in used:
            assert np.all(phi[:, 0, f] == 0.0)


def test_lossguide_bounded_only_by_depth_grows_the_depthwise_tree():
    """Without max_leaves every candidate is eventually expanded: best-first and level order then make the same splits (only the
    node numbering differs), so the margins agree; with max_leaves the leaf count is capped and children are numbered in
    expansion order."""
    X, y = synth(4000, 10, 11, "reg")
    a = O.train(dict(objective="reg:squarederror", max_depth=4), X, y, 3).model()
    b = O.train(dict(objective="reg:squarederror", max_depth=4, grow_policy="lossguide"), X, y, 3).model()
    np.testing.assert_allclose(O.predict_margin(a, X), O.predict_margin(b, X), rtol=0, atol=1e-6)
    c = O.train(dict(objective="reg:squarederror", max_depth=0, max_leaves=9, grow_policy="lossguide"), X, y, 2).model()
    for t in range(2):
        s = slice(c["tree_offset"][t], c["tree_offset"][t + 1])
        left, loss = c["left"][s], c["loss_chg"][s]
        assert int((left == -1).sum()) == 9
        internal = np.flatnonzero(left != -1)
        order = np.argsort(left[internal])                      # expansion order = order of the children's ids
        # best-first: a node is expanded before any node with a smaller gain that was open at the same time; in particular
        # the root comes first and the children ids are consecutive pairs
        assert internal[order][0] == 0 and np.all(np.diff(np.sort(left[internal])) == 2)


def test_oracle_monotone_constraints_give_monotone_models():
    n = 6000
    rng = np.random.default_rng(3)
    X, _ = synth(n, 5, 14, "reg")
    y = (np.sin(2 * X[:, 0]) + 0.5 * X[:, 1] - X[:, 2] ** 2 + 0.2 * rng.standard_normal(n)).astype(np.float32)
    m = O.train(dict(objective="reg:squarederror", max_depth=5, monotone_constraints=(1, -1, 0, 0, 0)), X, y, 10).model()
    for f, sign in ((0, 1), (1, -1)):
        for trial in range(20):
            grid = np.tile(X[rng.integers(n)], (100, 1)).astype(np.float32)
            grid[:, f] = np.linspace(-3.5, 3.5, 100, dtype=np.float32)
            assert np.all(sign * np.diff(O.predict_margin(m, grid).ravel()) >= -1e-6)
    # all-zero constraints are the unconstrained model
    a = O.train(dict(objective="reg:squarederror", max_depth=4), X, y, 3).model()
    b = O.train(dict(objective="reg:squarederror", max_depth=4, monotone_constraints="(0,0,0,0,0)"), X, y, 3).model()
    np.testing.assert_array_equal(a["split_cond"], b["split_cond"])


def test_oracle_interaction_constraints_keep_paths_inside_one_set():
    n = 5000
    rng = np.random.default_rng(8)
    X, _ = synth(n, 6, 15, "reg")
    y = (X[:, 0] * X[:, 1] + X[:, 2] * X[:, 3] + X[:, 4] + 0.1 * rng.standard_normal(n)).astype(np.float32)
    groups = [[0, 1], [2, 3, 4]]
    m = O.train(dict(objective="reg:squarederror", max_depth=5, interaction_constraints=groups), X, y, 6).model()
    for t in range(len(m["tree_info"])):
        a, b = m["tree_offset"][t], m["tree_offset"][t + 1]
        left, right, si = m["left"][a:b], m["right"][a:b], m["split_index"][a:b]
        stack = [(0, frozenset())]
        while stack:
            i, feats = stack.pop()
            if left[i] == -1:
                assert len(feats) <= 1 or any(feats <= set(g) for g in groups), feats
                continue
            f = feats | {int(si[i])}
            stack.append((int(left[i]), f)); stack.append((int(right[i]), f))
    # one set holding every feature constrains nothing
    a = O.train(dict(objective="reg:squarederror", max_depth=4), X, y, 3).model()
    b = O.train(dict(objective="reg:squarederror", max_depth=4, interaction_constraints="[[0,1,2,3,4,5]]"), X, y, 3).model()
    np.testing.assert_array_equal(a["split_cond"], b["split_cond"])
