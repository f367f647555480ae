"""WHOLE-MODEL PIN against a model trained by the real library.

The reference keeps, for its loader tests, a model that xgboost itself trained and saved (test/resources/models/saved_booster/
xgboost-model, xgboost 1.0 binary format; copy under tests/golden/legacy/): `multi:softprob`, 3 classes, 4 features, 60 trees.
Its statistics give the training run away: root cover 66.667 = 150 rows x 2 p (1 - p) at p = 1/3, first split petal-length < 2.45
-- the 150-row iris data (bundled with scikit-learn) -- with eta 0.3 (leaf = 0.3 x weight), lambda 1 (weight = -G / (H + 1)),
max_depth 3, min_child_weight 1, 20 rounds, base_score 0.5.  The thresholds are midpoints between data values, i.e. the `exact`
updater (what tree_method=auto picked for small data in 1.0); on features with fewer than 256 distinct values `hist` enumerates
the same partitions, so everything but the threshold representation must agree.

It does: re-training on iris with those hyperparameters reproduces ALL 60 TREES over 20 dependent boosting rounds -- same
structure, same split features, the same partition of the training rows at every node, loss_chg / sum_hess / leaf values to
float32 round-off.  This pins the training arithmetic (softmax gradients, gain, weight, tie-breaking, the round loop) on
reference-held data produced by the real implementation, for the oracle (CPU test) and for the CUDA path (GPU test)."""
import os

import numpy as np
import pytest

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "legacy", "saved_booster_xgboost-model")
PARAMS = dict(objective="multi:softprob", num_class=3, eta=0.3, max_depth=3, base_score=0.5)
ROUNDS = 20


def _iris():
    from sklearn.datasets import load_iris
    X, y = load_iris(return_X_y=True)
    return X.astype(np.float32), y.astype(np.float32)


def _reference_model():
    from oracle import legacy_model, ubjson
    return ubjson.model_from_xgb_json(legacy_model.to_document(open(GOLD, "rb").read()))


def _assert_same_model(ref, got, X, oracle, leaf_tol, stat_rtol):
    assert len(got["tree_info"]) == len(ref["tree_info"]) == 60
    np.testing.assert_array_equal(got["tree_info"], ref["tree_info"])
    np.testing.assert_array_equal(got["tree_offset"], ref["tree_offset"])            # same node count in every tree
    np.testing.assert_array_equal(got["left"], ref["left"])
    np.testing.assert_array_equal(got["right"], ref["right"])
    internal = ref["left"] != -1
    np.testing.assert_array_equal(got["split_index"][internal], ref["split_index"][internal])
    # thresholds: midpoints (exact) vs cut values (hist) -- the rows must still part the same way at every node of every tree
    np.testing.assert_array_equal(oracle.predict_leaf(got, X), oracle.predict_leaf(ref, X))
    assert np.all(got["split_cond"][internal] >= ref["split_cond"][internal])         # cut value = the upper neighbour of the midpoint
    leaf = ~internal
    assert float(np.abs(got["split_cond"][leaf] - ref["split_cond"][leaf]).max()) <= leaf_tol
    np.testing.assert_allclose(got["loss_chg"][internal], ref["loss_chg"][internal], rtol=stat_rtol, atol=1e-5)
    np.testing.assert_allclose(got["sum_hess"], ref["sum_hess"], rtol=stat_rtol, atol=1e-5)
    np.testing.assert_allclose(oracle.predict_margin(got, X), oracle.predict_margin(ref, X), rtol=0, atol=2e-5)


def test_oracle_reproduces_the_model_real_xgboost_trained_on_iris(oracle):
    X, y = _iris()
    ref = _reference_model()
    assert ref["num_class"] == 3 and ref["num_feature"] == 4 and ref["base_score"] == 0.5
    got = oracle.train(PARAMS, X, y, ROUNDS).model()
    _assert_same_model(ref, got, X, oracle, leaf_tol=2e-6, stat_rtol=2e-5)


def test_the_pin_is_sensitive(oracle):
    """the agreement above is not vacuous: a different lambda / eta / depth gives a different model"""
    X, y = _iris()
    ref = _reference_model()
    for change in (dict(reg_lambda=1.5), dict(eta=0.31), dict(max_depth=4), dict(min_child_weight=2)):
        params = dict(PARAMS, **{("lambda" if k == "reg_lambda" else k): v for k, v in change.items()})
        got = oracle.train(params, X, y, ROUNDS).model()
        same_shape = len(got["left"]) == len(ref["left"]) and np.array_equal(got["left"], ref["left"])
        leaf = ref["left"] == -1
        assert not (same_shape and float(np.abs(got["split_cond"][leaf] - ref["split_cond"][leaf]).max()) <= 1e-4), change


@pytest.mark.gpu
def test_cuda_path_reproduces_the_model_real_xgboost_trained_on_iris(xgb, oracle):
    X, y = _iris()
    ref = _reference_model()
    bst = xgb.train(dict(PARAMS, tree_method="hist"), xgb.DMatrix(X, label=y), num_boost_round=ROUNDS, verbose_eval=False)
    got = xgb.get_backend().booster_export_model(bst.handle)
    got["num_class"], got["num_feature"], got["objective"] = 3, 4, "multi:softprob"
    _assert_same_model(ref, got, X, oracle, leaf_tol=1e-5, stat_rtol=1e-4)
    # and the library's own predictor on the REFERENCE file agrees with the model just trained
    loaded = xgb.Booster()
    loaded.load_model(GOLD)
    d = xgb.DMatrix(X)
    np.testing.assert_array_equal(loaded.predict(d, pred_leaf=True), bst.predict(d, pred_leaf=True))
    np.testing.assert_allclose(loaded.predict(d), bst.predict(d), rtol=0, atol=2e-5)
