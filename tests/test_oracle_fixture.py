"""Pin the oracle on the only numeric anchor the reference ships: its UBJSON model fixture
(test/resources/abalone/models/libsvm_pickled/xgboost-model, copied to tests/golden/, see tests/golden/README.md).
CPU-only; checks listed in SURVEY.md section 8(c)."""
import os

import numpy as np
import pytest

from oracle import gbt_oracle as O
from oracle import ubjson

G = os.path.join(os.path.dirname(__file__), "golden")
ETA, LAMBDA, GAMMA, MCW = 0.2, 1.0, 4.0, 6.0     # test/integration/local/test_abalone.py:36-47


@pytest.fixture(scope="module")
def fixture_model():
    doc = ubjson.load(os.path.join(G, "abalone_xgboost-model.ubj"))
    return doc, ubjson.model_from_xgb_json(doc)


def load_libsvm(path, F=8):
    X, y = [], []
    for line in open(path):
        p = line.split()
        y.append(float(p[0]))
        row = [np.nan] * F
        for kv in p[1:]:
            k, v = kv.split(":")
            row[int(k) - 1] = float(v)
        X.append(row)
    return np.array(X, np.float32), np.array(y, np.float32)


def test_fixture_header(fixture_model):
    doc, m = fixture_model
    assert doc["version"] == [3, 2, 0]
    assert m["objective"] == "reg:squarederror" and len(m["tree_info"]) == 50 and m["num_feature"] == 8
    _, y = load_libsvm(os.path.join(G, "abalone", "abalone.train_0"))
    assert abs(m["base_score"] - float(y.mean())) < 1e-5          # base_score = mean label = one Newton stump at margin 0
    p = O.make_params({"objective": "reg:squarederror"})
    assert abs(O.lib().orc_base_score(O.C.byref(p), O._p(y), None, len(y)) - m["base_score"]) < 1e-5


def test_gain_identity_and_leaf_scaling(fixture_model):
    """loss_chg = GL^2/(HL+l) + GR^2/(HR+l) - G^2/(H+l), leaf = eta * (-G/(H+l)); thresholds gamma / min_child_weight."""
    _, m = fixture_model
    worst = 0.0
    n_internal = 0
    for t in range(50):
        T = m.tree(t)
        w = lambda j: T["base_weight"][j] if T["left"][j] != -1 else T["split_cond"][j] / ETA
        for i in range(len(T["left"])):
            if T["left"][i] == -1:
                assert T["sum_hess"][i] >= MCW or len(T["left"]) == 1
                continue
            n_internal += 1
            l, r = T["left"][i], T["right"][i]
            H, HL, HR = T["sum_hess"][i], T["sum_hess"][l], T["sum_hess"][r]
            G, GL, GR = -T["base_weight"][i] * (H + LAMBDA), -w(l) * (HL + LAMBDA), -w(r) * (HR + LAMBDA)
            lc = GL * GL / (HL + LAMBDA) + GR * GR / (HR + LAMBDA) - G * G / (H + LAMBDA)
            worst = max(worst, abs(lc - T["loss_chg"][i]) / T["loss_chg"][i])
            assert T["loss_chg"][i] >= GAMMA               # gamma acts before the split
            assert abs(HL + HR - H) < 1e-3 * H
            assert T["default_left"][i] == 0               # dense data: forward scan only
    assert n_internal == 715
    assert worst < 2e-5


def test_traversal_known_answer(fixture_model):
    _, m = fixture_model
    row = np.array([[2, 0.645, 0.515, 0.15, 1.212, 0.515, 0.2055, 0.385]], np.float32)      # test_abalone.py:24 LIBSVM_SAMPLE
    np.testing.assert_array_equal(O.predict_leaf(m, row)[0, :8], [42, 45, 40, 43, 34, 38, 38, 41])
    assert abs(float(O.predict_margin(m, row)[0, 0]) - 11.100031) < 1e-5


def test_oracle_split_arithmetic_reproduces_fixture_decisions(fixture_model):
    """The ORACLE'S OWN calc_weight / calc_gain / calc_split_gain (exported from gbt_oracle.c, the functions its trainer
    calls) fed with the 715 internal nodes of the reference-held fixture: with (G, H) of a node and of its two children
    taken from the stored sum_hessian and weights, the oracle's loss change must reproduce the stored `loss_changes` and
    its weights the stored `base_weights` / leaf values.  Pins lambda (=1, no 1/2 factor), the gain formula, the
    min_child_weight cut-off and the eta scaling of leaves on reference data through oracle code."""
    _, m = fixture_model
    p = O.make_params(dict(objective="reg:squarederror", eta=ETA, gamma=GAMMA, min_child_weight=MCW, max_depth=5))
    L, ref = O.lib(), O.C.byref(p)
    worst_gain, worst_w, n = 0.0, 0.0, 0
    for t in range(50):
        T = m.tree(t)
        w = lambda j: float(T["base_weight"][j]) if T["left"][j] != -1 else float(T["split_cond"][j]) / ETA     # unscaled weight of node j
        for i in range(len(T["left"])):
            if T["left"][i] == -1:
                continue
            n += 1
            l, r = int(T["left"][i]), int(T["right"][i])
            H, HL, HR = float(T["sum_hess"][i]), float(T["sum_hess"][l]), float(T["sum_hess"][r])
            GL, GR = -w(l) * (HL + LAMBDA), -w(r) * (HR + LAMBDA)
            G = GL + GR                                               # node sum from its children, NOT from its own weight
            lc = L.orc_calc_split_gain(ref, GL, HL, GR, HR) - L.orc_calc_gain(ref, G, H)
            worst_gain = max(worst_gain, abs(lc - float(T["loss_chg"][i])) / float(T["loss_chg"][i]))
            worst_w = max(worst_w, abs(L.orc_calc_weight(ref, G, H) - float(T["base_weight"][i])))      # parent weight from the children's sums
            for c, Gc, Hc in ((l, GL, HL), (r, GR, HR)):
                wc = L.orc_calc_weight(ref, Gc, Hc)
                if T["left"][c] == -1:
                    assert abs(ETA * wc - float(T["split_cond"][c])) <= 2e-6 * max(1.0, abs(wc))            # leaf value = eta * w
                else:
                    assert abs(wc - float(T["base_weight"][c])) <= 2e-6 * max(1.0, abs(wc))
    assert n == 715
    assert worst_gain < 3e-5 and worst_w < 2e-4
    # the min_child_weight rule the fixture obeys (every leaf hessian >= 6) is the oracle's cut-off
    assert L.orc_calc_weight(ref, -10.0, MCW - 0.5) == 0.0 and L.orc_calc_weight(ref, -10.0, MCW) != 0.0


def test_fixture_thresholds_are_oracle_cuts():
    """Cut-membership pin: on abalone features with <= 256 distinct values the oracle's cuts are the distinct values, so
    every split threshold the reference chose on those features must be one of the oracle's cut points (264 of 264)."""
    doc = ubjson.load(os.path.join(G, "abalone_xgboost-model.ubj"))
    m = ubjson.model_from_xgb_json(doc)
    X, _ = load_libsvm(os.path.join(G, "abalone", "abalone.train_0"))
    ptrs, vals, mins, hm = O.make_cuts(X, 256)
    internal = m["left"] != -1
    checked = 0
    for f in range(X.shape[1]):
        if len(np.unique(X[:, f])) > 256:
            continue
        cuts = set(np.float32(v) for v in vals[ptrs[f]:ptrs[f + 1]])
        thr = m["split_cond"][internal & (m["split_index"] == f)]
        assert all(np.float32(t) in cuts for t in thr), "feature %d: a fixture threshold is not an oracle cut" % f
        checked += len(thr)
    assert checked == 264


def test_oracle_trains_abalone_with_fixture_hyperparameters():
    X, y = load_libsvm(os.path.join(G, "abalone", "abalone.train_0"))
    params = dict(objective="reg:squarederror", max_depth=5, eta=ETA, gamma=GAMMA, min_child_weight=MCW)
    t = O.train(params, X, y, 50)
    m = t.model()
    assert len(m["tree_info"]) == 50
    internal = m["left"] != -1
    assert (m["loss_chg"][internal] >= GAMMA).all() and (m["sum_hess"][~internal] >= MCW).all()
    # same root split as the fixture's first tree (feature 7 = shell weight), cache == fresh prediction
    assert m.tree(0)["split_index"][0] == 7
    pred = O.predict_margin(m, X)[:, 0]
    np.testing.assert_array_equal(pred, t.margins()[:, 0])
    assert np.sqrt(np.mean((pred - y) ** 2)) < 2.0


def test_cuts_definition_small_cardinality():
    X = np.array([[1.0, 5.0], [2.0, 5.0], [2.0, 7.0], [4.0, np.nan]], np.float32)
    ptrs, vals, mins, hm = O.make_cuts(X, 256)
    assert hm
    np.testing.assert_array_equal(ptrs, [0, 3, 5])
    np.testing.assert_allclose(vals, [2.0, 4.0, 4.0 + 4.0 + 1e-5, 7.0, 7.0 + 7.0 + 1e-5])
    bins = O.bin_matrix(X, ptrs, vals)
    np.testing.assert_array_equal(bins, [[0, 0], [1, 0], [1, 1], [2, 255]])


@pytest.mark.parametrize("objective,kind", [("reg:squarederror", "reg"), ("binary:logistic", "bin")])
def test_fixed_point_gradient_grid_stays_inside_the_leaf_tolerance(objective, kind):
    """The product rounds gradients to a power-of-two grid (|g_q| <= 2^18, engine.h kGradBits) before the exact integer
    histogram.  The oracle can emulate that rounding (`set_quant_bits`): same trees, leaves within the 1e-5 bar."""
    import sys
    sys.path.insert(0, os.path.dirname(__file__))
    from util import synth
    X, y = synth(60000, 20, 5, kind)
    params = dict(objective=objective, max_depth=6, eta=0.3)
    cuts = O.make_cuts(X, 256)
    bins = O.bin_matrix(X, cuts[0], cuts[1])
    ref = O.Trainer(params, bins=bins, cuts=cuts, y=y)
    fx = O.Trainer(params, bins=bins, cuts=cuts, y=y)
    fx.set_quant_bits(19)               # 18 magnitude bits + sign, like the kernel
    for _ in range(10):
        ref.update(); fx.update()
    a, b = ref.model(), fx.model()
    np.testing.assert_array_equal(a["left"], b["left"])
    np.testing.assert_array_equal(a["split_index"], b["split_index"])
    leaf = a["left"] == -1
    assert np.abs(a["split_cond"][leaf] - b["split_cond"][leaf]).max() < 1e-5
