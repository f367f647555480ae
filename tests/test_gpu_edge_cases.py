"""Edge cases on the GPU path vs the oracle: tiny and ragged matrices, constant features / labels, an all-missing column,
a single row, weights -- the shapes the reference's own unit tests use (100 x 5 random data, test_checkpointing.py:164-244)."""
import numpy as np
import pytest

from util import assert_same_structure, max_leaf_diff

pytestmark = pytest.mark.gpu


def _be():
    from sagemaker_xgboost_container_b200.backend import get_backend
    return get_backend()


def _check(xgb, oracle, params, X, y, rounds, w=None):
    d = xgb.DMatrix(X, label=y, weight=w)
    bst = xgb.train(params, d, num_boost_round=rounds, verbose_eval=False)
    m = _be().booster_export_model(bst.handle)
    mr = oracle.train(params, X, y, rounds, weights=w).model()
    assert_same_structure(m, mr)
    assert max_leaf_diff(m, mr) <= 1e-5
    np.testing.assert_array_equal(bst.predict(d, pred_leaf=True).astype(np.int32).reshape(len(X), -1), oracle.predict_leaf(mr, X))
    return bst


def test_reference_unit_test_shape_100x5_logistic(xgb, oracle):
    rng = np.random.RandomState(1)
    X = rng.rand(100, 5).astype(np.float32)
    y = rng.randint(2, size=100).astype(np.float32)
    _check(xgb, oracle, {"objective": "binary:logistic", "max_depth": 6}, X, y, 20)


def test_single_row_and_tiny(xgb, oracle):
    _check(xgb, oracle, {"objective": "reg:squarederror", "max_depth": 3}, np.array([[1.0, 2.0, 3.0]], np.float32), np.array([5.0], np.float32), 3)
    rng = np.random.RandomState(2)
    X = rng.rand(17, 3).astype(np.float32)
    _check(xgb, oracle, {"objective": "reg:squarederror", "max_depth": 4, "min_child_weight": 1}, X, (X[:, 0] * 3 + X[:, 1]).astype(np.float32), 5)


def test_constant_feature_constant_label_and_all_missing_column(xgb, oracle):
    rng = np.random.RandomState(3)
    X = rng.rand(500, 6).astype(np.float32)
    X[:, 2] = 7.0                       # constant feature: a single bin
    X[:, 4] = np.nan                    # column without any value
    y = (X[:, 0] > 0.5).astype(np.float32) + X[:, 1]
    _check(xgb, oracle, {"objective": "reg:squarederror", "max_depth": 4}, X, y.astype(np.float32), 6)
    bst = _check(xgb, oracle, {"objective": "reg:squarederror", "max_depth": 4}, X, np.full(500, 2.5, np.float32), 2)
    assert bst.num_boosted_rounds() == 2           # constant labels: root-only stumps, like trees 29/45/47/49 of the fixture


def test_weights_and_33_features(xgb, oracle):
    rng = np.random.RandomState(4)
    X = np.round(rng.randn(3000, 33) * 8).astype(np.float32) / 8          # 33 features: two groups, the second with one slot
    y = (X[:, :3].sum(1) + 0.1 * rng.randn(3000)).astype(np.float32)
    w = rng.rand(3000).astype(np.float32) + 0.5
    _check(xgb, oracle, {"objective": "reg:squarederror", "max_depth": 5, "eta": 0.3}, X, y, 6, w=w)


def test_empty_prediction_batch_and_feature_count(xgb):
    rng = np.random.RandomState(5)
    X = rng.rand(200, 4).astype(np.float32)
    bst = xgb.train({"objective": "reg:squarederror", "max_depth": 2}, xgb.DMatrix(X, label=X[:, 0]), num_boost_round=2, verbose_eval=False)
    assert bst.predict(xgb.DMatrix(np.zeros((0, 4), np.float32))).shape == (0,)
    assert bst.num_features() == 4
