"""Host-side logic of the package on CPU: URI/CSV/libsvm loaders, parameter flattening, train() + callbacks, checkpoint
resume, model IO, tracker/collective -- driven through an oracle-backed engine (tests/ may use the oracle; the product
never does)."""
import json
import multiprocessing as mp
import os
import pickle

import numpy as np
import pytest

from util import synth


@pytest.fixture()
def cpu_xgb(monkeypatch):
    import sagemaker_xgboost_container_b200 as xgb
    from sagemaker_xgboost_container_b200 import backend
    from oracle.engine import OracleBackend
    monkeypatch.setattr(backend, "_BACKEND", OracleBackend(error_cls=xgb.XGBoostError))
    return xgb


def test_csv_uri_directory_with_weights(cpu_xgb, tmp_path):
    d = tmp_path / "train"
    d.mkdir()
    (d / "part-b.csv").write_text("1,0.5,3,4\n0,2.0,5,6\n")
    (d / "part-a.csv").write_text("1,1.0,7,8\n")
    dm = cpu_xgb.DMatrix("%s?format=csv&label_column=0&delimiter=,&weight_column=1" % d)
    assert (dm.num_row(), dm.num_col()) == (3, 2)
    np.testing.assert_array_equal(dm.get_label(), [1, 1, 0])           # files in sorted order: part-a first
    np.testing.assert_array_equal(dm.get_weight(), [1.0, 0.5, 2.0])
    assert dm                                                            # truthy (train.py:243,271 use `if train_dmatrix:`)
    dm2 = cpu_xgb.DMatrix("%s?format=csv&label_column=0&delimiter=," % (d / "part-a.csv"))
    assert (dm2.num_row(), dm2.num_col()) == (1, 3)


def test_libsvm_uri_keeps_indices(cpu_xgb):
    path = os.path.join(os.path.dirname(__file__), "golden", "abalone", "abalone.train_0")
    dm = cpu_xgb.DMatrix(path + "?format=libsvm")
    assert (dm.num_row(), dm.num_col()) == (1461, 9)                   # 1-based indices kept as-is (test_data_utils.py:119-127)
    assert abs(float(dm.get_label().mean()) - 10.026694) < 1e-5


def test_dmatrix_inputs_and_slice(cpu_xgb):
    import scipy.sparse as sp
    X, y = synth(50, 4, 1)
    d = cpu_xgb.DMatrix(X, label=y)
    s = d.slice([3, 1, 7])
    np.testing.assert_array_equal(s.get_label(), y[[3, 1, 7]])
    assert s.num_row() == 3
    dsp = cpu_xgb.DMatrix(sp.csr_matrix(np.array([[0, 1.5], [2.0, 0]], np.float32)))
    assert (dsp.num_row(), dsp.num_col()) == (2, 2)
    with pytest.raises(ValueError):
        cpu_xgb.DMatrix(np.zeros((2, 2, 2)))


def test_train_callbacks_monitor_format_and_early_stopping(cpu_xgb, capsys):
    X, y = synth(600, 5, 2)
    Xv, yv = synth(200, 5, 3)
    dtr, dva = cpu_xgb.DMatrix(X, label=y), cpu_xgb.DMatrix(Xv, label=np.random.default_rng(0).permutation(yv))
    es = cpu_xgb.callback.EarlyStopping(rounds=2, data_name="validation", metric_name="rmse", maximize=False, save_best=True)
    res = {}
    bst = cpu_xgb.train({"objective": "reg:squarederror", "max_depth": 3, "eval_metric": ["rmse"], "verbosity": 1, "nthread": 2}, dtr,
                        num_boost_round=30, evals=[(dtr, "train"), (dva, "validation")], callbacks=[cpu_xgb.callback.EvaluationMonitor(), es],
                        evals_result=res, verbose_eval=False)
    out = capsys.readouterr().out.splitlines()
    import re
    assert re.match(r"^\[0\]\ttrain-rmse:\d+\.\d{5}\tvalidation-rmse:\d+\.\d{5}$", out[0])      # CloudWatch regex shape (metrics.py:21-42)
    assert len(res["validation"]["rmse"]) < 30                                                # stopped early on the shuffled labels
    assert bst.num_boosted_rounds() == bst.best_iteration + 1                                  # save_best slices the model
    assert isinstance(bst.attr("best_score"), str)


def test_checkpoint_resume_matches_uninterrupted_run(cpu_xgb, tmp_path):
    X, y = synth(500, 6, 4, "bin")
    d = cpu_xgb.DMatrix(X, label=y)
    p = {"objective": "binary:logistic", "max_depth": 3}
    full = cpu_xgb.train(p, d, num_boost_round=6, verbose_eval=False)
    ck = cpu_xgb.callback.TrainingCheckPoint(directory=str(tmp_path), interval=2, name="xgboost-checkpoint")
    cpu_xgb.train(p, d, num_boost_round=4, callbacks=[ck], verbose_eval=False)
    files = sorted(os.listdir(tmp_path))
    assert files and all(f.startswith("xgboost-checkpoint_") and f.endswith(".ubj") for f in files)
    part = cpu_xgb.train(p, d, num_boost_round=4, verbose_eval=False)
    f = str(tmp_path / "xgboost-checkpoint.3")
    part.save_model(f)
    resumed = cpu_xgb.train(p, d, num_boost_round=2, xgb_model=f, verbose_eval=False)
    assert resumed.num_boosted_rounds() == 6
    np.testing.assert_allclose(resumed.predict(d), full.predict(d), atol=1e-6)
    zero = cpu_xgb.train(p, d, num_boost_round=0, xgb_model=f, verbose_eval=False)            # test_checkpointing.py:222-244
    assert isinstance(zero, cpu_xgb.Booster) and zero.num_boosted_rounds() == 4


def test_booster_api_surface(cpu_xgb, tmp_path):
    import inspect
    X, y = synth(300, 4, 5, "multi", K=3)
    d = cpu_xgb.DMatrix(X, label=y)
    bst = cpu_xgb.train({"objective": "multi:softprob", "num_class": 3, "max_depth": 2}, d, num_boost_round=3, verbose_eval=False)
    sig = inspect.signature(bst.predict)
    assert "ntree_limit" not in sig.parameters and "iteration_range" in sig.parameters and "validate_features" in sig.parameters   # serve_utils.py:238-250
    p = bst.predict(d, validate_features=False)
    assert p.shape == (300, 3) and np.allclose(p.sum(1), 1, atol=1e-5)
    assert bst.predict(d, iteration_range=(0, 1)).shape == (300, 3)
    assert bst.predict(d, pred_leaf=True).shape == (300, 9)
    cfg = json.loads(bst.save_config())
    assert cfg["learner"]["objective"]["name"] == "multi:softprob" and cfg["learner"]["learner_model_param"]["num_class"] == "3"   # serve.py:85-88
    f = str(tmp_path / "xgboost-model")
    bst.save_model(f)
    b2 = cpu_xgb.Booster()
    b2.load_model(f)
    b2.set_param("nthread", 1)                                                                 # serve_utils.py:193
    np.testing.assert_allclose(b2.predict(d), p, atol=1e-6)
    b3 = pickle.loads(pickle.dumps(bst))
    np.testing.assert_allclose(b3.predict(d), p, atol=1e-6)
    assert bst.copy().num_boosted_rounds() == 3 and bst[:2].num_boosted_rounds() == 2
    assert bst.num_features() == 4
    with pytest.raises(cpu_xgb.XGBoostError):
        cpu_xgb.Booster(model_file=str(tmp_path / "does-not-exist"))


def test_label_errors_carry_the_customer_error_substrings(cpu_xgb):
    X, y = synth(100, 3, 6)
    with pytest.raises(cpu_xgb.XGBoostError, match=r"label must be in \[0,1\] for logistic regression"):
        cpu_xgb.train({"objective": "binary:logistic"}, cpu_xgb.DMatrix(X, label=y * 10), num_boost_round=1, verbose_eval=False)


def test_param_flattening():
    from sagemaker_xgboost_container_b200.core import _param_items
    items = _param_items({"eval_metric": ["rmse", "mae"], "max_depth": 3, "monotone_constraints": (1, -1), "flag": True, "none": None})
    assert ("eval_metric", "rmse") in items and ("eval_metric", "mae") in items and ("max_depth", 3) in items
    assert ("monotone_constraints", "(1,-1)") in items and ("flag", "1") in items and all(k != "none" for k, _ in items)


def _tracker_worker(port, task_id, q):
    from sagemaker_xgboost_container_b200.tracker import TrackerClient
    c = TrackerClient("127.0.0.1", port, task_id, timeout=20)
    c.connect()
    got = c.broadcast({"host": task_id} if c.rank == 1 else None, 1)
    c.barrier()
    q.put((task_id, c.rank, c.world, got))
    c.close()


def test_tracker_ranks_by_task_and_broadcasts():
    """Multi-node without a cluster, like test/unit/test_distributed.py: processes on one box, 127.0.0.1."""
    from sagemaker_xgboost_container_b200.tracker import RabitTracker
    tr = RabitTracker(n_workers=3, host_ip="127.0.0.1", port=0, sortby="task")
    tr.start()
    args = tr.worker_args()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_tracker_worker, args=(args["dmlc_tracker_port"], t, q)) for t in ("algo-3", "algo-1", "algo-2")]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=60) for _ in procs)
    for p in procs:
        p.join(timeout=30)
    tr.wait_for(timeout=30)
    tr.free()
    assert [(t, r, w) for t, r, w, _ in res] == [("algo-1", 0, 3), ("algo-2", 1, 3), ("algo-3", 2, 3)]
    assert all(g == {"host": "algo-2"} for *_, g in res)


def _gloo_worker(rank, world, port, q):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), LOCAL_RANK=str(rank))
    from sagemaker_xgboost_container_b200 import backend, collective
    from oracle.engine import OracleBackend
    backend._BACKEND = OracleBackend()
    collective.init_from_env(backend="gloo")
    obj = collective.broadcast({"who": rank} if rank == 1 else None, 1)
    s = collective.allreduce_sum(np.array([rank + 1.0, 10.0 * (rank + 1)]))
    q.put((collective.get_rank(), collective.get_world_size(), obj, s.tolist()))
    collective.finalize()


def test_collective_world_size_2_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29650 + os.getpid() % 200
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    assert [(r, w) for r, w, *_ in res] == [(0, 2), (1, 2)]
    assert all(o == {"who": 1} and s == [3.0, 30.0] for *_, o, s in res)


def test_sklearn_wrappers_and_cv(cpu_xgb, tmp_path):
    """Script-mode surface used by the container's customer-script resource (test/resources/boston/...:54,94)."""
    X, y = synth(400, 5, 9)
    reg = cpu_xgb.XGBRegressor(objective="reg:squarederror", colsample_bytree=1.0, learning_rate=0.3, max_depth=3, n_estimators=8)
    reg.fit(X, y, eval_set=[(X, y)], verbose=False)
    p = reg.predict(X)
    assert p.shape == (400,) and np.sqrt(np.mean((p - y) ** 2)) < np.std(y)
    assert len(reg.evals_result()["validation_0"]["rmse"]) == 8
    assert reg.feature_importances_.shape == (5,) and abs(reg.feature_importances_.sum() - 1) < 1e-5
    f = str(tmp_path / "m.json")
    reg.save_model(f)
    reg2 = cpu_xgb.XGBRegressor()
    reg2.load_model(f)
    np.testing.assert_allclose(reg2.predict(X), p, atol=1e-6)
    Xc, yc = synth(300, 4, 10, "multi", K=3)
    clf = cpu_xgb.XGBClassifier(max_depth=2, n_estimators=4).fit(Xc, yc, verbose=False)
    assert clf.predict_proba(Xc).shape == (300, 3) and set(np.unique(clf.predict(Xc))) <= {0, 1, 2}
    d = cpu_xgb.DMatrix(X, label=y)
    res = cpu_xgb.cv({"objective": "reg:squarederror", "max_depth": 3}, d, num_boost_round=5, nfold=3, metrics="rmse", as_pandas=False, seed=123)
    assert len(res["test-rmse-mean"]) == 5 and res["train-rmse-mean"][-1] < res["train-rmse-mean"][0]


def test_tracker_frames_are_json_and_bad_handshakes_are_dropped():
    """no pickle on the tracker's sockets; a port probe / oversized hello must not take a worker slot."""
    import socket, struct, inspect
    from sagemaker_xgboost_container_b200 import tracker as T
    assert "pickle" not in inspect.getsource(T).replace("No pickle", "").replace("never pickle", "")
    for obj in ({"a": [1, 2.5, None, True], "t": (1, "x"), "b": b"\x00\xff"}, "s", 3, None):
        assert T._decode(__import__("json").loads(__import__("json").dumps(T._encode(obj)))) == obj
    with pytest.raises(TypeError):
        T._encode({"f": object()})
    tr = T.RabitTracker(n_workers=1, host_ip="127.0.0.1", port=0, sortby="task")
    tr.start()
    port = tr.worker_args()["dmlc_tracker_port"]
    probe = socket.create_connection(("127.0.0.1", port))           # connect and say nothing useful
    probe.sendall(struct.pack("!Q", 1 << 40))                       # absurd length: dropped at the cap
    junk = socket.create_connection(("127.0.0.1", port))
    junk.sendall(struct.pack("!Q", 4) + b"\x80\x04N.")              # a pickle frame: not JSON, dropped
    c = T.TrackerClient("127.0.0.1", port, "algo-1", timeout=20)
    c.connect()
    assert (c.rank, c.world) == (0, 1)
    assert c.broadcast({"k": (1, 2)}, 0) == {"k": (1, 2)}
    c.close()
    tr.wait_for(timeout=20)
    tr.free()
    probe.close(); junk.close()


def test_libsvm_fast_and_plain_loaders_agree(tmp_path):
    """data._load_libsvm takes scikit-learn's C parser when it can and a plain Python loop otherwise (weights in the label
    token, qid); on files both accept they must give the same CSR matrix and labels."""
    from sagemaker_xgboost_container_b200 import data
    rng = np.random.default_rng(3)
    files = []
    for k in range(3):
        lines = []
        for r in range(int(rng.integers(1, 40))):
            idx = np.sort(rng.choice(30, size=int(rng.integers(0, 8)), replace=False))
            toks = ["%g" % rng.standard_normal()] + ["%d:%.7g" % (i, v) for i, v in zip(idx, rng.standard_normal(len(idx)) * 10 ** rng.uniform(-5, 5, len(idx)))]
            lines.append(" ".join(toks) + ("  # trailing comment" if r % 9 == 0 else ""))
        p = tmp_path / ("part-%d.libsvm" % k)
        p.write_text("\n".join(lines) + "\n")
        files.append(str(p))
    Xf, yf, wf = data._load_libsvm_fast(files)
    import unittest.mock as um
    with um.patch.object(data, "_load_libsvm_fast", return_value=None):
        Xs, ys, ws = data._load_libsvm(files, {})
    assert wf is None and ws is None
    np.testing.assert_array_equal(yf, ys)
    assert Xf.shape == Xs.shape and (Xf != Xs).nnz == 0
    np.testing.assert_array_equal(Xf.toarray().view(np.uint32), Xs.toarray().astype(np.float32).view(np.uint32))
