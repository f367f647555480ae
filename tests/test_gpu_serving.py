"""Serving input path on the device (BASELINE config 5): the CSV request body parsed by csv.cu must
give exactly the float32 matrix the container's own route builds (encoder.csv_to_dmatrix: str.split -> np.array -> float64 ->
DMatrix float32), for every literal form Python's float() accepts; anything outside the exact fast path falls back to it."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _reference_route(payload, delimiter=","):
    rows = [["nan" if x == "" else x for x in line.split(delimiter)] for line in payload.split("\n")]      # encoder.py:31-32,50
    return np.array(rows).astype(float).astype(np.float32)                                                # xgb.DMatrix(float64) stores float32


def _device_matrix(xgb, payload):
    from sagemaker_xgboost_container_b200 import serving
    d = serving.csv_to_dmatrix(payload, dtype=float)
    n, F = d.num_row(), d.num_col()
    be = xgb.get_backend()
    return be.dmatrix_get_raw(d.handle).reshape(n, F), d


def test_csv_payload_device_parse_is_bit_identical_to_the_container_route(xgb):
    rng = np.random.default_rng(7)
    n, F = 20000, 28
    X = rng.standard_normal((n, F)) * np.exp(rng.uniform(-20, 20, size=(n, F)))
    lines = []
    for r in range(n):
        fmt = ["%.6g", "%.17g", "%.3e", "%d", "%.9f"][r % 5]
        vals = [(fmt % (int(v) if fmt == "%d" else v)) for v in X[r]]
        if r % 97 == 0:
            vals[r % F] = ""                     # empty field -> NaN
        if r % 101 == 0:
            vals[(r + 3) % F] = ["nan", "NaN", "inf", "-inf", "+1.5", " 2.5 ", "-0", "1e-45", "3.4028235e38", "1e39"][(r // 101) % 10]
        lines.append(",".join(vals))
    payload = "\n".join(lines)
    got, d = _device_matrix(xgb, payload)
    ref = _reference_route(payload)
    assert got.shape == ref.shape
    assert np.array_equal(got.view(np.uint32)[~np.isnan(ref)], ref.view(np.uint32)[~np.isnan(ref)])          # bit-exact, incl. -0, denormals, inf
    assert np.array_equal(np.isnan(got), np.isnan(ref))


def test_csv_semicolon_single_row_and_single_column(xgb):
    for payload in ("1.5;2;;4", "3.25", "1\n2\n3", "1e3,2e-3\n-4,5"):
        delim = ";" if ";" in payload else ","
        got, _ = _device_matrix(xgb, payload)
        ref = _reference_route(payload, delim)
        np.testing.assert_array_equal(np.nan_to_num(got, nan=-777.0), np.nan_to_num(ref, nan=-777.0))


def test_csv_out_of_fast_path_literals_fall_back_to_the_host_route(xgb):
    payload = "0.1000000000000000055511151231257827021181583404541015625,2\n123456789012345678901234567890,1e400"
    got, _ = _device_matrix(xgb, payload)
    ref = _reference_route(payload)
    np.testing.assert_array_equal(got, ref)


def test_csv_ragged_and_malformed_payloads_raise(xgb):
    from sagemaker_xgboost_container_b200 import serving
    with pytest.raises(ValueError):
        serving.csv_to_dmatrix("1,2,3\n4,5", dtype=float)
    with pytest.raises(ValueError):
        serving.csv_to_dmatrix("1,2\n3,abc", dtype=float)


def test_serving_predict_matches_direct_predict(xgb):
    from sagemaker_xgboost_container_b200 import serving
    from util import synth
    X, y = synth(5000, 12, 3, "bin")
    d = xgb.DMatrix(X, label=y)
    bst = xgb.train(dict(objective="binary:logistic", max_depth=4), d, num_boost_round=5, verbose_eval=False)
    payload = "\n".join(",".join("%.9g" % v for v in row) for row in X[:300])
    dtest = serving.csv_to_dmatrix(payload, dtype=float)
    p = serving.predict(bst, "xgb_format", dtest, "text/csv", objective="binary:logistic")
    np.testing.assert_array_equal(p, bst.predict(xgb.DMatrix(X[:300])))
    ens = serving.predict([bst, bst], ["xgb_format"] * 2, dtest, "text/csv", objective="binary:logistic")
    np.testing.assert_allclose(ens, p, rtol=0, atol=1e-7)


def test_training_csv_channel_is_parsed_on_the_device_like_the_host_loader(xgb, tmp_path):
    """data_utils.py:289-318: a directory of CSV files, label in column 0, optional weight in column 1 (csv_weights=1)."""
    from sagemaker_xgboost_container_b200.data import load_uri
    rng = np.random.default_rng(3)
    d = tmp_path / "train"
    d.mkdir()
    for i in range(3):
        A = np.round(rng.standard_normal((700 + i, 9)) * 10.0 ** rng.integers(-3, 4, size=(1, 9)), 5)
        A[:, 1] = np.abs(A[:, 1]) + 0.5          # weights must be non-negative
        lines = [",".join("" if (r + c) % 53 == 0 and c > 1 else repr(float(v)) for c, v in enumerate(row)) for r, row in enumerate(A)]
        (d / ("part-%d.csv" % i)).write_text("\n".join(lines) + "\n")
    uri = "%s?format=csv&label_column=0&delimiter=,&weight_column=1" % d
    dm = xgb.DMatrix(uri)
    X, y, w = load_uri(uri)                      # host loader (pandas' C parser: its fast strtod may be 1 ulp off in double)
    be = xgb.get_backend()
    got = be.dmatrix_get_raw(dm.handle).reshape(dm.num_row(), dm.num_col())
    assert got.shape == X.shape
    # exact reference: Python's correctly rounded float() of every field, then float32 -- the device parser must match it bit for bit
    rows = [l.split(",") for f in sorted(os.listdir(d)) for l in open(d / f).read().strip().split("\n")]
    ref = np.array([[np.nan if v == "" else float(v) for v in r] for r in rows], dtype=np.float64).astype(np.float32)
    np.testing.assert_array_equal(np.nan_to_num(got, nan=-777.0), np.nan_to_num(ref[:, 2:], nan=-777.0))
    np.testing.assert_array_equal(dm.get_label(), ref[:, 0])
    np.testing.assert_array_equal(dm.get_weight(), ref[:, 1])
    np.testing.assert_allclose(np.nan_to_num(got, nan=-777.0), np.nan_to_num(X, nan=-777.0), rtol=2e-7, atol=0)
    np.testing.assert_allclose(dm.get_label(), y, rtol=2e-7)
