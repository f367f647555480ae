"""Training loaders that reach the device without a dense float32 host matrix (csrc/ingest.cu):
column buffers (Parquet / pandas) converted + transposed on the device, CSR (libsvm / scipy) densified on the device.  Both must
give bit for bit the float32 matrix the container's host route builds (data_utils.py:348-390: to_pandas().to_numpy() /
scipy -> DMatrix, i.e. numpy's astype(float32) and NaN for absent entries)."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "abalone")
DTYPES = [np.float32, np.float64, np.int32, np.int64, np.uint8, np.int8, np.int16, np.uint16, np.uint32, np.uint64, np.bool_]


def _raw(xgb, d):
    return xgb.get_backend().dmatrix_get_raw(d.handle).reshape(d.num_row(), d.num_col())


def _random_columns(rng, n, ncols):
    cols = []
    for c in range(ncols):
        dt = DTYPES[(c * 7 + ncols) % len(DTYPES)]
        if dt in (np.float32, np.float64):
            v = (rng.standard_normal(n) * np.exp(rng.uniform(-30, 30, n))).astype(dt)
            v[rng.random(n) < 0.03] = np.nan
            if n > 3:
                v[:3] = [np.inf, -np.inf, -0.0]
        elif dt is np.bool_:
            v = rng.random(n) < 0.5
        else:
            info = np.iinfo(dt)
            v = rng.integers(info.min, info.max, n, dtype=dt, endpoint=True)       # full range: int64 / uint64 values that round in float32
        cols.append(v)
    return cols


@pytest.mark.parametrize("n,ncols,label,weight", [(1, 1, -1, -1), (31, 2, 0, -1), (33, 6, 2, 0), (1000, 32, -1, -1), (1000, 33, 0, -1), (1000, 34, 0, 1),
                                                  (70001, 71, 5, -1), (5000, 101, 0, -1), (4097, 130, 64, 129)])
def test_columns_match_numpy_astype_bit_for_bit(xgb, n, ncols, label, weight):
    rng = np.random.default_rng(n * 131 + ncols)
    cols = _random_columns(rng, n, ncols)
    if weight >= 0:                                                      # weights must be non-negative numbers (the setter checks)
        cols[weight] = rng.integers(0, 30000, n).astype(np.uint16)
    be = xgb.get_backend()
    d = xgb.DMatrix._from_handle(be.dmatrix_from_columns(cols, label_column=label, weight_column=weight))
    feats = [c for i, c in enumerate(cols) if i not in (label, weight)]
    assert (d.num_row(), d.num_col()) == (n, len(feats))
    if feats:
        ref = np.column_stack([c.astype(np.float32) for c in feats])
        got = _raw(xgb, d)
        assert got.view(np.uint32).tobytes() == ref.view(np.uint32).tobytes()
    if label >= 0:
        assert d.get_label().view(np.uint32).tobytes() == cols[label].astype(np.float32).view(np.uint32).tobytes()
    if weight >= 0:
        assert d.get_weight().view(np.uint32).tobytes() == cols[weight].astype(np.float32).view(np.uint32).tobytes()


def test_columns_longer_than_one_staging_chunk(xgb):
    n = (1 << 22) + 77                                                   # ingest.cu stages 2^22 rows at a time
    rng = np.random.default_rng(5)
    cols = [rng.standard_normal(n), rng.integers(-1000, 1000, n, dtype=np.int32), rng.standard_normal(n).astype(np.float32)]
    d = xgb.DMatrix._from_handle(xgb.get_backend().dmatrix_from_columns(cols, label_column=0))
    got = _raw(xgb, d)
    assert np.array_equal(got[:, 0], cols[1].astype(np.float32)) and np.array_equal(got[:, 1], cols[2])
    assert np.array_equal(d.get_label(), cols[0].astype(np.float32))


def test_parquet_channel_equals_the_container_route(xgb, tmp_path):
    import pyarrow as pa
    import pyarrow.parquet as pq
    from sagemaker_xgboost_container_b200 import data
    rng = np.random.default_rng(9)
    n, F = 30011, 28
    y = rng.integers(0, 2, n).astype(np.int64)
    X = rng.standard_normal((n, F))
    tabs = {"label": y}
    for j in range(F):
        col = X[:, j].astype(np.float32) if j % 3 == 0 else X[:, j]
        tabs["f%d" % j] = pa.array(col, mask=(rng.random(n) < 0.02) if j % 5 == 0 else None)     # nulls -> NaN
    ch = tmp_path / "train"
    ch.mkdir()
    t = pa.table(tabs)
    pq.write_table(t.slice(0, 12000), ch / "part-0.parquet", row_group_size=5000)
    pq.write_table(t.slice(12000), ch / "part-1.parquet", row_group_size=7000)
    ref = pq.read_table(str(ch)).to_pandas().to_numpy()                  # data_utils.py:375-385
    d = data.parquet_to_dmatrix(str(ch))
    got = _raw(xgb, d)
    want = ref[:, 1:].astype(np.float32)
    assert got.shape == want.shape and got.view(np.uint32).tobytes() == want.view(np.uint32).tobytes()
    np.testing.assert_array_equal(d.get_label(), ref[:, 0].astype(np.float32))
    # and a model trained on it equals the model trained on the container-route matrix
    d2 = xgb.DMatrix(ref[:, 1:], label=ref[:, 0])
    p = dict(objective="binary:logistic", max_depth=4, eta=0.3)
    be = xgb.get_backend()
    b1, b2 = xgb.train(p, d, num_boost_round=3, verbose_eval=False), xgb.train(p, d2, num_boost_round=3, verbose_eval=False)
    m1, m2 = be.booster_export_model(b1.handle), be.booster_export_model(b2.handle)
    for k in ("left", "split_index", "split_cond"):
        np.testing.assert_array_equal(m1[k], m2[k])


def test_pandas_frame_goes_through_the_column_path(xgb):
    import pandas as pd
    rng = np.random.default_rng(2)
    df = pd.DataFrame({"a": rng.standard_normal(999), "b": rng.integers(0, 9, 999), "c": rng.standard_normal(999).astype(np.float32), "d": rng.random(999) < 0.3})
    d = xgb.DMatrix(df, label=np.zeros(999, np.float32))
    want = df.to_numpy(dtype=np.float32)
    assert _raw(xgb, d).view(np.uint32).tobytes() == want.view(np.uint32).tobytes()
    assert d.feature_names == ["a", "b", "c", "d"]


def test_csr_is_densified_on_the_device(xgb):
    import scipy.sparse as sp
    rng = np.random.default_rng(4)
    n, F = 20011, 57
    dense = rng.standard_normal((n, F)).astype(np.float32)
    mask = rng.random((n, F)) < 0.2
    mask[::7] = False                                                    # empty rows
    csr = sp.csr_matrix(np.where(mask, dense, 0))
    csr.eliminate_zeros()
    d = xgb.DMatrix(csr)
    want = np.full((n, F), np.nan, np.float32)
    rows = np.repeat(np.arange(n), np.diff(csr.indptr))
    want[rows, csr.indices] = csr.data
    got = _raw(xgb, d)
    assert (d.num_row(), d.num_col()) == (n, F)
    assert got.view(np.uint32).tobytes() == want.view(np.uint32).tobytes()
    d0 = xgb.DMatrix(sp.csr_matrix((3, 4), dtype=np.float32))           # no entries at all
    assert np.isnan(_raw(xgb, d0)).all() and _raw(xgb, d0).shape == (3, 4)


def test_libsvm_channel_keeps_indices_and_values(xgb):
    path = os.path.join(G, "abalone.train_0")
    d = xgb.DMatrix(path + "?format=libsvm")
    assert (d.num_row(), d.num_col()) == (1461, 9)                       # 1-based indices kept as they are (test_data_utils.py:119-127)
    want = np.full((1461, 9), np.nan, np.float32)
    labels = []
    for r, line in enumerate(open(path)):
        p = line.split()
        labels.append(float(p[0]))
        for kv in p[1:]:
            k, v = kv.split(":")
            want[r, int(k)] = np.float32(float(v))
    got = _raw(xgb, d)
    assert got.view(np.uint32).tobytes() == want.view(np.uint32).tobytes()
    np.testing.assert_array_equal(d.get_label(), np.asarray(labels, np.float32))
