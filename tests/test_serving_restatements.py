"""The GPU serving tests (tests/test_gpu_serving.py) compare the device parsers with small RESTATEMENTS of the container's own
parsing routes, because the reference tree does not travel to the GPU box.  Here, where /root/reference is mounted, the
restatements themselves -- and the host routes serving.py falls back to -- are checked against the reference's real functions
(encoder.csv_to_dmatrix, encoder.libsvm_to_dmatrix, serve_utils._get_sparse_matrix_from_libsvm + xgb.DMatrix) on the very
bodies the GPU tests use, with this package bound as `xgboost` on the CPU test engine."""
import os
import sys

import numpy as np
import pytest

import reference_stubs

pytestmark = pytest.mark.skipif(not reference_stubs.reference_available(), reason="/root/reference is not mounted here")
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture()
def container(monkeypatch):
    import sagemaker_xgboost_container_b200 as xgb
    from sagemaker_xgboost_container_b200 import backend
    from oracle.engine import OracleBackend
    monkeypatch.setattr(backend, "_BACKEND", OracleBackend(error_cls=xgb.XGBoostError))
    reference_stubs.install(xgb)
    from sagemaker_xgboost_container import encoder
    from sagemaker_xgboost_container.algorithm_mode import serve_utils
    return xgb, encoder, serve_utils


def _matrix(d):
    return d.handle.X                                   # oracle engine: the float32 matrix the DMatrix holds (NaN = missing)


def _same(a, b):
    return a.shape == b.shape and np.array_equal(np.isnan(a), np.isnan(b)) and np.array_equal(np.nan_to_num(a), np.nan_to_num(b))


def test_libsvm_restatements_equal_the_reference_functions(container):
    xgb, encoder, serve_utils = container
    import test_gpu_serving as T
    from sagemaker_xgboost_container_b200 import serving
    bodies = [T._libsvm_body(np.random.default_rng(31 + ob), 400, 40, ob) for ob in (True, False)]
    bodies += ["1 1:0.5 1:0.25 2:3", "1 1:0.5 2:3\n0", "1 +1:0.5 2:3", "1 1:0.12345678901234567890123 2:3", "1 1:nan 2:3",
               "1 1:0.5 1:0.25 2:3\n0\n1 4:1", "0 3:1e-3\t7:2 \n\n1 1:5", "1 0:1 5:2\n0 2:3"]
    for body in bodies:
        # sparse route of the algorithm-mode handler (serve_utils.py:94-118 + xgb.DMatrix(csr)): absent entries are missing
        try:
            ref = _matrix(xgb.DMatrix(serve_utils._get_sparse_matrix_from_libsvm(body)))
        except Exception as e:
            with pytest.raises(type(e)):
                T._ref_sparse_route(body)
            with pytest.raises(type(e)):
                serving.sparse_libsvm_to_dmatrix(body)
        else:
            assert _same(T._ref_sparse_route(body), ref), body[:60]
            assert _same(_matrix(serving.sparse_libsvm_to_dmatrix(body)), ref), body[:60]
        # dense route of the script-mode handler (encoder.py:54-86): absent entries are 0.0
        ref = _matrix(encoder.libsvm_to_dmatrix(body))
        assert _same(T._ref_dense_route(body), ref), body[:60]
        assert _same(_matrix(serving.libsvm_to_dmatrix(body)), ref), body[:60]
    assert encoder.libsvm_to_dmatrix("1\n0\n").num_row() == serving.libsvm_to_dmatrix("1\n0\n").num_row() == 0
    with pytest.raises(ValueError):                     # a token with a tab inside: the reference's float() refuses it, so does the mirror
        serve_utils._get_sparse_matrix_from_libsvm("1 1:2\t3:4 5:6")
    with pytest.raises(ValueError):
        serving.sparse_libsvm_to_dmatrix("1 1:2\t3:4 5:6")


def test_csv_restatement_equals_the_reference_function(container):
    xgb, encoder, _ = container
    import test_gpu_serving as T
    from sagemaker_xgboost_container_b200 import serving
    rng = np.random.default_rng(7)
    X = rng.standard_normal((300, 9)) * np.exp(rng.uniform(-20, 20, size=(300, 9)))
    lines = []
    for r in range(300):
        vals = [("%.6g", "%.17g", "%.3e")[r % 3] % v for v in X[r]]
        if r % 11 == 0:
            vals[r % 9] = ""
        if r % 13 == 0:
            vals[(r + 2) % 9] = ["nan", "inf", "-inf", "+1.5", "-0", "1e-45", "1e39"][(r // 13) % 7]
        lines.append(",".join(vals))
    body = "\n".join(lines)
    with np.errstate(over="ignore"):
        ref = _matrix(encoder.csv_to_dmatrix(body, dtype=float))
        assert _same(T._reference_route(body), ref)
        assert _same(_matrix(serving.csv_to_dmatrix(body, dtype=float)), ref)          # CPU engine: the mirror's host route
    for delim in (";", "\t", " "):
        b2 = "\n".join(delim.join("%g" % v for v in row) for row in rng.standard_normal((20, 4)))
        assert _same(_matrix(serving.csv_to_dmatrix(b2, dtype=float)), _matrix(encoder.csv_to_dmatrix(b2, dtype=float)))
