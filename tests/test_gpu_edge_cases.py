"""Edge cases on the GPU path vs the oracle: tiny and ragged matrices, constant features / labels, an all-missing column,
a single row, weights -- the shapes the reference's own unit tests use (100 x 5 random data, test_checkpointing.py:164-244)."""
import numpy as np
import pytest

from util import assert_same_structure, max_leaf_diff, synth

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


def test_c_abi_array_interface_and_uri_constructors(xgb, tmp_path):
    """SURVEY.md 8(b) minimum export set: XGDMatrixCreateFromDense / SetInfoFromInterface / CreateFromURI straight through ctypes,
    the way upstream's Python package (INTEGRATION.md option B) would call them."""
    import ctypes as C
    import json
    be = xgb.get_backend()
    lib = be.lib
    rng = np.random.default_rng(5)
    X64 = np.ascontiguousarray(rng.standard_normal((300, 6)))               # float64 ndarray, like data_utils.py:384 after Parquet
    y = np.ascontiguousarray(rng.random(300).astype(np.float32))

    def aif(a):
        return json.dumps({"data": [a.ctypes.data, True], "shape": list(a.shape), "typestr": a.dtype.str, "version": 3}).encode()
    h = C.c_void_p()
    assert lib.XGDMatrixCreateFromDense(aif(X64), json.dumps({"missing": float("nan"), "nthread": 0}).encode(), C.byref(h)) == 0, lib.XGBGetLastError()
    assert lib.XGDMatrixSetInfoFromInterface(h, b"label", aif(y)) == 0
    np.testing.assert_array_equal(be.dmatrix_get_raw(h).reshape(300, 6), X64.astype(np.float32))
    np.testing.assert_array_equal(be.dmatrix_get_float_info(h, "label"), y)
    be.dmatrix_free(h)
    # URI: a directory with two CSV files (label first) and a libsvm file
    d = tmp_path / "csv"
    d.mkdir()
    A = np.round(rng.standard_normal((50, 4)), 4)
    (d / "a.csv").write_text("\n".join(",".join(repr(float(v)) for v in r) for r in A[:30]) + "\n")
    (d / "b.csv").write_text("\n".join(",".join(repr(float(v)) for v in r) for r in A[30:]) + "\n")
    h2 = C.c_void_p()
    cfg = json.dumps({"uri": "%s?format=csv&label_column=0&delimiter=," % d, "silent": 1}).encode()
    assert lib.XGDMatrixCreateFromURI(cfg, C.byref(h2)) == 0, lib.XGBGetLastError()
    np.testing.assert_array_equal(be.dmatrix_get_raw(h2).reshape(50, 3), A[:, 1:].astype(np.float32))
    np.testing.assert_array_equal(be.dmatrix_get_float_info(h2, "label"), A[:, 0].astype(np.float32))
    be.dmatrix_free(h2)
    f = tmp_path / "d.libsvm"
    f.write_text("1 1:0.5 3:2\n0 2:1.5\n")
    h3 = C.c_void_p()
    assert lib.XGDMatrixCreateFromURI(json.dumps({"uri": "%s?format=libsvm" % f}).encode(), C.byref(h3)) == 0, lib.XGBGetLastError()
    got = be.dmatrix_get_raw(h3).reshape(2, 4)
    assert got[0, 1] == 0.5 and got[0, 3] == 2.0 and got[1, 2] == 1.5 and np.isnan(got[0, 0])
    be.dmatrix_free(h3)


def test_damaged_model_documents_are_refused_before_they_reach_a_kernel(xgb, tmp_path):
    """The predictor walks the tree arrays unchecked, so the loader must refuse what would send it out of bounds or in circles:
    a child that points back at its parent, a child index past the array, a split on a feature the model does not have, a
    tree_info shorter than the tree list, a count field larger than the file (UBJSON)."""
    import json
    X, y = synth(500, 4, 3)
    bst = xgb.train(dict(objective="reg:squarederror", max_depth=3), xgb.DMatrix(X, label=y), num_boost_round=2, verbose_eval=False)
    doc = json.loads(bytes(bst.save_raw("json")).decode())
    tree = doc["learner"]["gradient_booster"]["model"]["trees"][0]
    assert tree["left_children"][0] == 1                                 # a real split at the root

    def damaged(edit):
        d = json.loads(json.dumps(doc))
        edit(d["learner"]["gradient_booster"]["model"])
        return json.dumps(d).encode()

    def cycle(m): m["trees"][0]["left_children"][1] = 0; m["trees"][0]["right_children"][1] = 2
    def past_end(m): m["trees"][0]["right_children"][0] = 10 ** 6
    def bad_feature(m): m["trees"][0]["split_indices"][0] = 4
    def short_info(m): m["tree_info"] = m["tree_info"][:1]
    for edit in (cycle, past_end, bad_feature, short_info):
        with pytest.raises(xgb.XGBoostError, match="model"):
            xgb.Booster(model_file=bytearray(damaged(edit)))
    raw = bytes(bst.save_raw("ubj"))
    k = raw.index(b"[$d#")                                               # first typed float array: give it an absurd length
    with pytest.raises(xgb.XGBoostError, match="count runs past the end"):
        xgb.Booster(model_file=bytearray(raw[:k + 4] + b"L" + (2 ** 40).to_bytes(8, "big") + raw[k + 6:]))
    ok = xgb.Booster(model_file=bytearray(json.dumps(doc).encode()))     # the undamaged document still loads and predicts the same
    np.testing.assert_array_equal(ok.predict(xgb.DMatrix(X)), bst.predict(xgb.DMatrix(X)))
