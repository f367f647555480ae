"""`use_dask_gpu_training=true` on the native one-process-per-GPU path (multi_gpu.py; SURVEY.md section 8(f) row 3, reference
algorithm_mode/train.py:183-214 + distributed_gpu/distributed_gpu_training.py:93-222).

CPU: world_size-2 runs of the launcher on the oracle-backed test engine (tracker rendezvous, row shards read per rank,
master-only model / checkpoints, failures surfaced) -- the sharded job must write the model the single-process job writes.
GPU (needs 2 GPUs): the same through the CUDA engine, model bit-identical to the 1-GPU model."""
import functools
import os
import sys

import numpy as np
import pytest

import reference_stubs

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "helpers"))
import oracle_worker_init  # noqa: E402

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "abalone")


def _libsvm_to_csv(src, dst):
    with open(dst, "w") as out:
        for line in open(src):
            p = line.split()
            vals = {int(k): v for k, v in (kv.split(":") for kv in p[1:])}
            out.write(",".join([p[0]] + [vals.get(i, "0") for i in range(1, 9)]) + "\n")


def _channels(tmp_path):
    tr, va = tmp_path / "train", tmp_path / "validation"
    tr.mkdir(); va.mkdir()
    _libsvm_to_csv(os.path.join(G, "abalone.train_0"), tr / "abalone.train_0.csv")
    _libsvm_to_csv(os.path.join(G, "abalone.train_1"), tr / "abalone.train_1.csv")
    _libsvm_to_csv(os.path.join(G, "abalone.validation"), va / "abalone.validation.csv")
    return tr, va


def test_shards_cover_the_channel_exactly_once(tmp_path):
    from sagemaker_xgboost_container_b200 import multi_gpu
    tr, _ = _channels(tmp_path)
    files = multi_gpu._channel_files(str(tr), "csv")
    whole = b"\n".join(open(f, "rb").read().strip(b"\n") for f in files)
    for world in (1, 2, 3, 8):
        parts = [multi_gpu._csv_shard_text(files, r, world) for r in range(world)]
        assert all(n == 2922 for _, n in parts)
        assert b"\n".join(t for t, _ in parts) == whole
        sizes = [t.count(b"\n") + 1 for t, _ in parts]
        assert max(sizes) - min(sizes) <= 1
    assert multi_gpu._csv_shard_text(files, 0, 4000)[0] == b""           # more workers than lines: an empty shard, reported by load_shard
    import pyarrow as pa
    import pyarrow.parquet as pq
    pqd = tmp_path / "pq"
    pqd.mkdir()
    arr = np.genfromtxt(whole.decode().splitlines(), delimiter=",", dtype=np.float32)
    for i, (a, b) in enumerate([(0, 1000), (1000, 2922)]):
        pq.write_table(pa.table({"c%d" % j: arr[a:b, j] for j in range(arr.shape[1])}), pqd / ("part-%d.parquet" % i), row_group_size=300)
    # (the DMatrix itself needs an engine: only the row bookkeeping is checked here)
    metas = [pq.ParquetFile(f) for f in multi_gpu._channel_files(str(pqd), "parquet")]
    assert sum(m.metadata.num_rows for m in metas) == 2922


def test_validation_mirrors_the_reference_checks():
    from sagemaker_xgboost_container_b200 import multi_gpu as mg
    rep = {"train": {"S3DistributionType": "FullyReplicated"}}
    assert mg.validate_gpu_train_configuration("hist", 1, 8, "File", "csv", rep) == []
    assert mg.validate_gpu_train_configuration("gpu_hist", 2, 8, "File", "parquet", rep) == []
    assert mg.validate_gpu_train_configuration("approx", 1, 8, "File", "csv", rep) == [mg.NON_GPU_ERROR_MSG]
    assert mg.validate_gpu_train_configuration("hist", 1, 0, "Pipe", "libsvm", rep) == [mg.NON_GPU_ERROR_MSG, mg.PIPE_MODE_ERROR_MSG, mg.INPUT_FORMAT_ERROR_MSG]
    sharded = {"train": {"S3DistributionType": "ShardedByS3Key"}}
    assert mg.validate_gpu_train_configuration("hist", 1, 8, "File", "csv", sharded) == []
    assert mg.validate_gpu_train_configuration("hist", 2, 8, "File", "csv", sharded) == [mg.NOT_REPLICATED_ERROR_MSG]


@pytest.mark.skipif(not reference_stubs.reference_available(), reason="/root/reference is not mounted here")
def test_sagemaker_train_with_use_dask_gpu_training_runs_the_native_launcher(tmp_path, monkeypatch, capfd):
    """The reference's entry point, unchanged, with the HP set: its call at train.py:203 lands on multi_gpu (the one-line
    binding of INTEGRATION.md); two worker processes; the model equals the one the ordinary single-process route writes."""
    import sagemaker_xgboost_container_b200 as xgb
    from sagemaker_xgboost_container_b200 import backend, multi_gpu
    from oracle.engine import OracleBackend
    monkeypatch.setattr(backend, "_BACKEND", OracleBackend(error_cls=xgb.XGBoostError))
    reference_stubs.install(xgb)
    from sagemaker_xgboost_container.algorithm_mode import train as ref_train
    from sagemaker_xgboost_container.distributed_gpu import distributed_gpu_training as dgt
    monkeypatch.setattr(dgt, "run_training_with_dask", functools.partial(multi_gpu.run_training_with_dask, worker_init=oracle_worker_init.use_oracle_engine))
    monkeypatch.setenv("SM_NUM_GPUS", "2")
    tr, va = _channels(tmp_path)
    ck = tmp_path / "ck"
    ck.mkdir()
    hp = {"objective": "reg:squarederror", "tree_method": "hist", "num_round": "8", "max_depth": "4", "eta": "0.3", "eval_metric": "rmse"}
    dc = {"train": {"ContentType": "text/csv", "TrainingInputMode": "File", "S3DistributionType": "FullyReplicated"},
          "validation": {"ContentType": "text/csv", "TrainingInputMode": "File", "S3DistributionType": "FullyReplicated"}}
    m2 = tmp_path / "model2"
    ref_train.sagemaker_train(train_config=dict(hp, use_dask_gpu_training="true"), data_config=dc, train_path=str(tr), val_path=str(va), model_dir=str(m2),
                              sm_hosts=["algo-1"], sm_current_host="algo-1", checkpoint_config={"LocalPath": str(ck)})
    out = capfd.readouterr().out
    lines = [l for l in out.splitlines() if l.startswith("[") and "train-rmse:" in l]
    assert len(lines) == 8 and "validation-rmse:" in lines[-1]          # the master's monitor only: one line per round, not two
    assert sorted(os.listdir(m2)) == ["xgboost-model"]
    assert len(os.listdir(ck)) >= 1                                       # checkpoints written once (by the master)
    m1 = tmp_path / "model1"
    ref_train.sagemaker_train(train_config=dict(hp), data_config=dc, train_path=str(tr), val_path=str(va), model_dir=str(m1),
                              sm_hosts=["algo-1"], sm_current_host="algo-1", checkpoint_config={})
    from oracle import ubjson
    a = ubjson.model_from_xgb_json(ubjson.load(str(m1 / "xgboost-model")))
    b = ubjson.model_from_xgb_json(ubjson.load(str(m2 / "xgboost-model")))
    for k in ("left", "right", "split_index", "split_cond", "tree_offset"):
        np.testing.assert_array_equal(a[k], b[k])


def _boom():
    raise RuntimeError("engine refused to start")


def test_worker_failures_reach_the_caller(tmp_path):
    from sagemaker_xgboost_container_b200 import multi_gpu
    tr, _ = _channels(tmp_path)
    with pytest.raises(Exception, match="engine refused to start"):
        multi_gpu.run_training_with_dask({"num_round": 2}, str(tr), None, str(tmp_path / "m"), "csv", ["algo-1"], "algo-1", None, 2, worker_init=_boom)


def _ngpu():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


@pytest.mark.gpu
def test_two_gpu_launcher_equals_single_gpu(xgb, tmp_path):
    if _ngpu() < 2:
        pytest.skip("needs 2 GPUs")
    from sagemaker_xgboost_container_b200 import multi_gpu
    rng = np.random.default_rng(11)
    n, F = 60000, 12
    X = (np.round(np.clip(rng.standard_normal((n, F)), -4, 4 - 1 / 32) * 32) / 32).astype(np.float32)
    y = (X @ (rng.standard_normal(F) / np.sqrt(F)) + 0.1 * rng.standard_normal(n)).astype(np.float32)
    tr = tmp_path / "train"
    tr.mkdir()
    for i, (a, b) in enumerate([(0, 25000), (25000, n)]):
        np.savetxt(tr / ("part-%d.csv" % i), np.column_stack([y[a:b], X[a:b]]), delimiter=",", fmt="%.9g")
    hp = {"objective": "reg:squarederror", "tree_method": "hist", "num_round": 6, "max_depth": 5, "eta": 0.3, "eval_metric": ["rmse"]}
    multi_gpu.run_training_with_dask(dict(hp), str(tr), None, str(tmp_path / "m2"), "csv", ["algo-1"], "algo-1", None, 2, worker_init=oracle_worker_init.bind_package)
    multi = xgb.Booster(model_file=str(tmp_path / "m2" / "xgboost-model"))
    d = xgb.DMatrix(str(tr) + "?format=csv&label_column=0&delimiter=,")
    single = xgb.train({k: v for k, v in hp.items() if k != "num_round"}, d, num_boost_round=6, verbose_eval=False)
    be = xgb.get_backend()
    m1, m2 = be.booster_export_model(single.handle), be.booster_export_model(multi.handle)
    for k in ("left", "right", "split_index", "split_cond"):
        np.testing.assert_array_equal(m1[k], m2[k])


def test_two_hosts_with_one_gpu_each_rendezvous_through_the_master_tracker(tmp_path):
    """Several hosts (train.py:236-269 analogue for the GPU route): every host calls run_training_with_dask with the same host
    list; the first host runs the tracker on the fixed port, ranks follow (host, gpu), rank 0 alone writes the model.  Two
    'hosts' on this machine (two names of the loopback interface), CPU test engine."""
    import threading
    from sagemaker_xgboost_container_b200 import multi_gpu
    tr, va = _channels(tmp_path)
    hosts = ["localhost", "127.0.0.1"]
    hp = {"objective": "reg:squarederror", "tree_method": "hist", "num_round": 4, "max_depth": 3, "eta": 0.3}
    dirs = [tmp_path / "model-host0", tmp_path / "model-host1"]
    errors = []

    def host(i):
        try:
            multi_gpu.run_training_with_dask(dict(hp), str(tr), str(va), str(dirs[i]), "csv", hosts, hosts[i], None, 1,
                                             worker_init=oracle_worker_init.use_oracle_engine)
        except BaseException as e:      # noqa: BLE001
            errors.append((i, e))
    threads = [threading.Thread(target=host, args=(i,)) for i in (1, 0)]          # the non-master host may well come up first
    for t in threads:
        t.start()
    for t in threads:
        t.join(300)
    assert not errors, errors
    assert os.path.exists(dirs[0] / "xgboost-model") and not os.path.exists(dirs[1])      # the master (rank 0 lives on host 0) saves, nobody else
    # same model as one process on all rows
    import sagemaker_xgboost_container_b200 as xgb
    from sagemaker_xgboost_container_b200 import backend
    from oracle.engine import OracleBackend
    from oracle import ubjson
    old = backend._BACKEND
    backend._BACKEND = OracleBackend(error_cls=xgb.XGBoostError)
    try:
        arr = np.concatenate([np.loadtxt(f, delimiter=",", dtype=np.float32) for f in sorted(str(p) for p in tr.iterdir())])
        single = xgb.train({k: v for k, v in hp.items() if k != "num_round"}, xgb.DMatrix(arr[:, 1:], label=arr[:, 0]), num_boost_round=4, verbose_eval=False)
        a = ubjson.model_from_xgb_json(ubjson.loads(bytes(single.save_raw("ubj"))))
    finally:
        backend._BACKEND = old
    b = ubjson.model_from_xgb_json(ubjson.load(str(dirs[0] / "xgboost-model")))
    for k in ("left", "split_index", "split_cond"):
        np.testing.assert_array_equal(a[k], b[k])


def test_memory_mapped_shards_equal_the_whole_file_split(tmp_path, monkeypatch):
    """_csv_shard_text never loads the channel: it maps the files and scans them in chunks.  Against the obvious whole-file
    implementation on random channels: several files, CRLF and LF endings, leading / trailing blank lines, empty files, scan
    chunks far smaller than a line run."""
    from sagemaker_xgboost_container_b200 import multi_gpu as mg

    def whole_file(files, rank, world):
        chunks = [b for b in (open(f, "rb").read().replace(b"\r\n", b"\n").strip(b"\n") for f in files) if b]
        buf = np.frombuffer(b"\n".join(chunks), np.uint8)
        ends = np.flatnonzero(buf == 10)
        n = len(ends) + (1 if len(buf) else 0)
        lo, hi = mg.shard_bounds(n, rank, world)
        if hi <= lo:
            return b"", n
        return buf[(0 if lo == 0 else int(ends[lo - 1]) + 1):(len(buf) if hi == n else int(ends[hi - 1]))].tobytes(), n
    monkeypatch.setattr(mg, "_SCAN_CHUNK", 37)
    rng = np.random.default_rng(0)
    for trial in range(120):
        files = []
        for k in range(int(rng.integers(1, 5))):
            eol = b"\r\n" if rng.random() < 0.3 else b"\n"
            body = eol.join(b",".join(b"%d" % int(x) for x in rng.integers(0, 1000, int(rng.integers(1, 6)))) for _ in range(int(rng.integers(0, 12))))
            p = tmp_path / ("t%d_%d.csv" % (trial, k))
            p.write_bytes(b"\n" * int(rng.integers(0, 3)) + body + eol * int(rng.integers(0, 3)))
            files.append(str(p))
        for world in (1, 2, 3, 7):
            for r in range(world):
                assert mg._csv_shard_text(files, r, world) == whole_file(files, r, world)


@pytest.mark.skipif(not reference_stubs.reference_available(), reason="/root/reference is not mounted here")
def test_the_references_own_validation_tests_pass_on_the_native_module(monkeypatch):
    """test/unit/distributed_gpu/test_distributed_gpu_training.py, unmodified, with the module it imports
    (distributed_gpu_training: validate_gpu_train_configuration + its four message constants) replaced by multi_gpu"""
    import importlib.util
    import unittest
    import sagemaker_xgboost_container_b200 as xgb
    from sagemaker_xgboost_container_b200 import multi_gpu
    reference_stubs.install(xgb)
    import sagemaker_xgboost_container.distributed_gpu  # noqa: F401  (the parent package must be importable)
    monkeypatch.setitem(sys.modules, "sagemaker_xgboost_container.distributed_gpu.distributed_gpu_training", multi_gpu)
    spec = importlib.util.spec_from_file_location("ref_distributed_gpu_training_tests", "/root/reference/test/unit/distributed_gpu/test_distributed_gpu_training.py")
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    assert mod.validate_gpu_train_configuration is multi_gpu.validate_gpu_train_configuration
    result = unittest.TextTestRunner(verbosity=0).run(unittest.defaultTestLoader.loadTestsFromModule(mod))
    assert result.wasSuccessful() and result.testsRun >= 9, (result.failures, result.errors)


@pytest.mark.skipif(not reference_stubs.reference_available(), reason="/root/reference is not mounted here")
def test_unbound_reference_dask_route_still_trains_through_xgboost_dask(tmp_path, monkeypatch, capfd):
    """Without the multi_gpu binding the container's own run_training_with_dask (distributed_gpu_training.py:93-222) reaches
    `xgboost.dask.DaskDMatrix` / `xgboost.dask.train` of this package: they must exist (train.py:46 imports the module at
    start-up) and train.  The Dask cluster itself is stubbed (no dask in this image): a Client that is a context manager, and
    dask.array.from_array returning the array."""
    import types
    import sagemaker_xgboost_container_b200 as xgb
    from sagemaker_xgboost_container_b200 import backend
    from oracle.engine import OracleBackend
    monkeypatch.setattr(backend, "_BACKEND", OracleBackend(error_cls=xgb.XGBoostError))
    reference_stubs.install(xgb)
    import xgboost.dask as dxgb
    assert dxgb is xgb.dask and sys.modules["xgboost"].dask is xgb.dask
    from sagemaker_xgboost_container.distributed_gpu import distributed_gpu_training as dgt, dask_data_utils as ddu

    class Client:
        def __init__(self, address): self.address = address
        def __enter__(self): return self
        def __exit__(self, *a): return False
        def wait_for_workers(self, n, timeout): pass
        def scheduler_info(self): return {"workers": {"w0": {}, "w1": {}}}
    monkeypatch.setattr(dgt, "Client", Client)
    monkeypatch.setattr(dgt, "dxgb", xgb.dask)
    monkeypatch.setattr(ddu, "dxgb", xgb.dask)
    monkeypatch.setattr(ddu, "da", types.SimpleNamespace(from_array=lambda a, chunks=None: a))
    monkeypatch.setattr(dgt, "start_daemons_in_current_instance", lambda *a, **k: None)
    monkeypatch.setattr(dgt, "get_host_ip", lambda h: "127.0.0.1")
    tr, va = _channels(tmp_path)
    hp = {"objective": "reg:squarederror", "tree_method": "hist", "num_round": 5, "max_depth": 3, "eval_metric": ["rmse"]}
    (tmp_path / "m").mkdir()                                             # /opt/ml/model exists in the container
    with pytest.warns(UserWarning, match="multi_gpu.run_training_with_dask"):
        dgt.run_training_with_dask(hyperparameters=dict(hp), train_path=str(tr), validation_path=str(va), model_dir=str(tmp_path / "m"),
                                   content_type="csv", sm_hosts=["algo-1"], current_host="algo-1", checkpoint_dir=None, num_gpus=2)
    out = capfd.readouterr().out
    # every round twice, as with the real library: get_callbacks adds an EvaluationMonitor and the reference leaves dask.train's
    # verbose_eval at its default True (distributed_gpu_training.py:184-192), which adds another
    assert len([l for l in out.splitlines() if l.startswith("[") and "validation-rmse:" in l]) == 10
    b = xgb.Booster(model_file=str(tmp_path / "m" / "xgboost-model"))
    assert b.num_boosted_rounds() == 5 and b.num_features() == 8
