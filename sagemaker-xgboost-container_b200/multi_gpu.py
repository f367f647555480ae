"""All GPUs of the instance(s) for one training job: the native replacement of the container's Dask path.

The reference reaches multi-GPU training only through `use_dask_gpu_training=true` (algorithm_mode/train.py:183-214), which
starts a Dask scheduler + one dask-cuda worker per GPU and calls `xgboost.dask.train`
(distributed_gpu/distributed_gpu_training.py:93-222, dask_data_utils.py:27-95).  Here the same hyperparameter lands on the
engine's own multi-GPU mode: ONE PROCESS PER GPU, rows sharded by global rank, per-level int64 histogram
all-reduce over NVLink / NCCL inside libb200xgb.so, bootstrap through this package's RabitTracker (the NCCL unique id travels
over its TCP links, collective.init).  `run_training_with_dask` keeps the reference's name, arguments and side effects
(model written to `<model_dir>/xgboost-model` by the master only, checkpoints by the master only, evaluation lines on the
master's stdout), so the binding in the container is one assignment:

    from sagemaker_xgboost_container_b200 import multi_gpu
    distributed_gpu_training.run_training_with_dask = multi_gpu.run_training_with_dask

Differences that are deliberate: no Dask cluster is started (nothing listens on 8786); every worker reads only ITS row range
of the channel (the reference reads everything on the scheduler host and scatters); a checkpoint found in `checkpoint_dir` is
resumed (the reference's Dask path loads it and then ignores it, distributed_gpu_training.py:166-185).
"""
import logging
import multiprocessing as mp
import os
import socket
import traceback

import numpy as np

CSV, PARQUET = "csv", "parquet"                       # data_utils.CSV / PARQUET
MODEL_NAME = "xgboost-model"                          # constants/xgb_constants.py MODEL_NAME
TRACKER_PORT = 9099
SUPPORTED_TRAINING_CONTENT_TYPES = {CSV, PARQUET}
NON_GPU_ERROR_MSG = "Multi-GPU training is only available for `hist` (or `gpu_hist`) training on GPU instances."
PIPE_MODE_ERROR_MSG = "Multi-GPU training is not supported for pipe mode input. Please use File mode."
INPUT_FORMAT_ERROR_MSG = "Multi-GPU training is only supported for CSV and Parquet input."
NOT_REPLICATED_ERROR_MSG = "Multi-GPU distributed training requires FullyReplicated data."

logger = logging.getLogger(__name__)


def validate_gpu_train_configuration(tree_method_hp, num_hosts, num_gpus, input_mode, input_format, data_config):
    """Same checks, same order, as distributed_gpu_training.validate_gpu_train_configuration (lines 61-90)."""
    errors = []
    if tree_method_hp not in ("gpu_hist", "hist") or num_gpus == 0:
        errors.append(NON_GPU_ERROR_MSG)
    if input_mode == "Pipe":
        errors.append(PIPE_MODE_ERROR_MSG)
    if input_format not in SUPPORTED_TRAINING_CONTENT_TYPES:
        errors.append(INPUT_FORMAT_ERROR_MSG)
    not_replicated = any(ch.get("S3DistributionType", None) != "FullyReplicated" for ch in data_config.values())
    if not_replicated and num_hosts > 1:                  # on one host replicated and sharded mean the same thing
        errors.append(NOT_REPLICATED_ERROR_MSG)
    return errors


# ------------------------------------------------------------------------------------------------ row shards
def _channel_files(path, content_type):
    """The files the reference's reader takes (dask_data_utils.read_data: sorted glob of *.csv / *.parquet); a channel whose
    files carry no extension (the symlink farm of data_utils.py:520-545) is taken whole."""
    if os.path.isfile(path):
        return [path]
    names = sorted(f for f in os.listdir(path) if not f.startswith(".") and os.path.isfile(os.path.join(path, f)))
    ext = [f for f in names if f.lower().endswith("." + content_type)]
    return [os.path.join(path, f) for f in (ext or names)]


def shard_bounds(n, rank, world):
    return rank * n // world, (rank + 1) * n // world


_SCAN_CHUNK = 64 << 20                                # bytes looked at per step when counting / locating newlines


def _file_extent(mm):
    """[s, e) of a file's bytes without leading / trailing line terminators"""
    s, e = 0, len(mm)
    while e > s and mm[e - 1] in (10, 13):
        e -= 1
    while s < e and mm[s] in (10, 13):
        s += 1
    return s, e


def _count_newlines(mm, s, e):
    n = 0
    for a in range(s, e, _SCAN_CHUNK):
        n += int(np.count_nonzero(mm[a:min(e, a + _SCAN_CHUNK)] == 10))
    return n


def _offset_after_newline(mm, s, e, k):
    """byte offset right after the k-th newline (1-based) inside [s, e)"""
    seen = 0
    for a in range(s, e, _SCAN_CHUNK):
        chunk = mm[a:min(e, a + _SCAN_CHUNK)]
        c = int(np.count_nonzero(chunk == 10))
        if seen + c >= k:
            return a + int(np.flatnonzero(chunk == 10)[k - seen - 1]) + 1
        seen += c
    raise ValueError("line index out of range")


def _csv_shard_text(files, rank, world):
    """Lines [lo, hi) of the concatenated files as one text block, and the total line count.  The files are memory-mapped and
    scanned in chunks: a worker holds its own shard in memory, never the channel (8 workers x a 50 GB channel otherwise)."""
    maps, extents, counts = [], [], []
    for f in files:
        if os.path.getsize(f) == 0:
            continue
        mm = np.memmap(f, dtype=np.uint8, mode="r")
        s, e = _file_extent(mm)
        if e > s:
            maps.append(mm); extents.append((s, e)); counts.append(_count_newlines(mm, s, e) + 1)
    n = int(sum(counts))
    lo, hi = shard_bounds(n, rank, world)
    if hi <= lo:
        return b"", n
    parts, first = [], 0
    for mm, (s, e), c in zip(maps, extents, counts):
        a, z = max(lo, first), min(hi, first + c)             # this file's lines [a, z) belong to the shard
        if a < z:
            ba = s if a == first else _offset_after_newline(mm, s, e, a - first)
            bz = e if z == first + c else _offset_after_newline(mm, s, e, z - first) - 1
            part = bytes(mm[ba:bz])
            if part.endswith(b"\r") and bz < e:                # the cut fell inside a CRLF: the '\r' belongs to the terminator
                part = part[:-1]
            parts.append(part.replace(b"\r\n", b"\n"))
        first += c
    return b"\n".join(parts), n


def load_shard(path, content_type, rank, world):
    """This rank's rows of a channel as a DMatrix (column 0 = label, as dask_data_utils.read_data:48-52).  CSV text goes to
    the device parser (csv.cu) when the backend has it; Parquet is read row-group-wise through pyarrow."""
    from . import DMatrix
    from .backend import get_backend
    files = _channel_files(path, content_type)
    if not files:
        raise ValueError("No %s files found under %s" % (content_type, path))
    if content_type == CSV:
        text, n = _csv_shard_text(files, rank, world)
        if not text:
            raise ValueError("worker %d of %d has no rows: %s holds %d lines" % (rank, world, path, n))
        first = text[:text.find(b"\n")] if b"\n" in text else text
        delim = "," if b"," in first else (";" if b";" in first else ("\t" if b"\t" in first else (" " if b" " in first.strip() else ",")))
        be = get_backend()
        if hasattr(be, "dmatrix_from_csv_labeled"):
            handle, status = be.dmatrix_from_csv_labeled(text, delim, 0, -1)
            if status == 0:
                return DMatrix._from_handle(handle), n
        import io
        import pandas as pd
        arr = pd.read_csv(io.BytesIO(text), header=None, sep=delim, dtype=np.float32).to_numpy(np.float32)
        return DMatrix(arr[:, 1:], label=arr[:, 0]), n
    import pyarrow as pa
    import pyarrow.parquet as pq
    metas = [pq.ParquetFile(f) for f in files]
    n = sum(m.metadata.num_rows for m in metas)
    lo, hi = shard_bounds(n, rank, world)
    parts, base = [], 0
    for m in metas:                                       # only the row groups that intersect [lo, hi) are read
        for g in range(m.metadata.num_row_groups):
            rows = m.metadata.row_group(g).num_rows
            a, z = max(lo, base), min(hi, base + rows)
            if a < z:
                parts.append(m.read_row_group(g).slice(a - base, z - a))
            base += rows
    if not parts:
        raise ValueError("worker %d of %d has no rows: %s holds %d rows" % (rank, world, path, n))
    t = pa.concat_tables(parts)
    from .data import _arrow_columns
    be = get_backend()
    if hasattr(be, "dmatrix_from_columns"):               # arrow column buffers straight to the device (csrc/ingest.cu)
        return DMatrix._from_handle(be.dmatrix_from_columns(_arrow_columns(t), label_column=0)), n
    cols = [np.asarray(c, np.float32) for c in _arrow_columns(t)]
    X = np.empty((len(cols[0]), len(cols) - 1), np.float32)
    for j, c in enumerate(cols[1:]):
        X[:, j] = c
    return DMatrix(X, label=cols[0]), n


# ------------------------------------------------------------------------------------------------ one worker = one GPU
def _train_on_shard(hyperparameters, train_path, validation_path, model_dir, content_type, checkpoint_dir, is_master, rank, world):
    """The body of the reference's scheduler-side block (distributed_gpu_training.py:107-213) on this rank's rows."""
    import sagemaker_xgboost_container_b200 as xgb
    hp = dict(hyperparameters)
    dtrain, n_train = load_shard(train_path, content_type, rank, world)
    if is_master:
        logging.info("Train features matrix has %d rows and %d columns (%d on this GPU)", n_train, dtrain.num_col(), dtrain.num_row())
    watchlist = [(dtrain, "train")]
    dvalid = None
    if validation_path:
        dvalid, _ = load_shard(validation_path, content_type, rank, world)
        watchlist.append((dvalid, "validation"))
    num_round = int(hp.pop("num_round"))
    save_model_on_termination = hp.pop("save_model_on_termination", "false")
    tuning_metric_param = hp.pop("_tuning_objective_metric", None)
    eval_metric = hp.pop("eval_metric", None)
    early_stopping_rounds = hp.pop("early_stopping_rounds", None)
    for k in ("use_dask_gpu_training", "_kfold", "_num_cv_round"):
        hp.pop(k, None)
    try:                                                  # the container's own helpers when it is installed next to this package
        from sagemaker_xgboost_container.algorithm_mode import train_utils
        from sagemaker_xgboost_container.callback import get_callbacks
    except ImportError:
        train_utils = get_callbacks = None
    feval, tuning_metric = None, None
    if train_utils is not None:
        cleaned, feval, tuning_metric = train_utils.get_eval_metrics_and_feval(tuning_metric_param, eval_metric)
        if cleaned:
            hp["eval_metric"] = cleaned
    elif eval_metric:
        hp["eval_metric"] = eval_metric
    es_metric = None
    if early_stopping_rounds:
        es_metric = tuning_metric[-1] if tuning_metric else (eval_metric[-1] if eval_metric else None)
    xgb_model, iteration, callbacks = None, 0, []
    if get_callbacks is not None:
        xgb_model, iteration, callbacks = get_callbacks(
            model_dir=model_dir, checkpoint_dir=checkpoint_dir, early_stopping_data_name="validation" if dvalid else None,
            early_stopping_metric=es_metric, early_stopping_rounds=early_stopping_rounds,
            save_model_on_termination=save_model_on_termination, is_master=is_master)
    bst = xgb.train(hp, dtrain, num_boost_round=num_round - iteration, evals=watchlist, custom_metric=feval, callbacks=callbacks,
                    xgb_model=xgb_model, verbose_eval=False if callbacks else is_master)
    if is_master:
        os.makedirs(model_dir, exist_ok=True)
        bst.save_model(os.path.join(model_dir, MODEL_NAME))
        logging.info("Training complete. Model saved.")


def _worker(local_rank, host_index, num_gpus, world, tracker_uri, tracker_port, kwargs, worker_init, errq):
    try:
        os.environ["LOCAL_RANK"] = str(local_rank)        # the engine binds to this GPU (csrc/booster.cu engine_stream)
        os.environ.pop("RANK", None)
        os.environ.pop("WORLD_SIZE", None)
        if worker_init is not None:
            worker_init()
        from sagemaker_xgboost_container_b200 import collective
        rank_hint = host_index * num_gpus + local_rank
        args = {"dmlc_tracker_uri": tracker_uri, "dmlc_tracker_port": tracker_port, "dmlc_task_id": "%06d" % rank_hint, "dmlc_timeout": 300}
        with collective.CommunicatorContext(**args):
            rank = collective.get_rank()
            assert collective.get_world_size() == world
            _train_on_shard(is_master=rank == 0, rank=rank, world=world, **kwargs)
        errq.put((local_rank, None))
    except BaseException as e:                            # noqa: BLE001 -- everything goes back to the parent
        errq.put((local_rank, "%s: %s\n%s" % (type(e).__name__, e, traceback.format_exc())))


def run_training_with_dask(hyperparameters, train_path, validation_path, model_dir, content_type, sm_hosts, current_host,
                           checkpoint_dir, num_gpus, worker_init=None):
    """Signature of distributed_gpu_training.run_training_with_dask (lines 93-103).  Blocks until the job is done on this
    host; raises AlgorithmError (the container's, when importable) with the first worker's traceback on failure.
    `worker_init`: optional picklable callable run first in every worker process (tests select their engine with it)."""
    hosts = list(sm_hosts)
    host_index = hosts.index(current_host)
    num_gpus = int(num_gpus)
    world = len(hosts) * num_gpus
    if num_gpus < 1:
        raise ValueError(NON_GPU_ERROR_MSG)
    master_ip = "127.0.0.1" if len(hosts) == 1 else socket.gethostbyname(hosts[0])
    tracker = None
    if host_index == 0:
        from .tracker import RabitTracker
        tracker = RabitTracker(n_workers=world, host_ip=master_ip, port=0 if len(hosts) == 1 else TRACKER_PORT, sortby="task")
        tracker.start()
    port = tracker.port if tracker is not None else TRACKER_PORT
    kwargs = dict(hyperparameters=dict(hyperparameters), train_path=train_path, validation_path=validation_path, model_dir=model_dir,
                  content_type=content_type, checkpoint_dir=checkpoint_dir)
    ctx = mp.get_context("spawn")                         # a forked CUDA context is unusable; spawn gives every GPU a fresh process
    errq = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(g, host_index, num_gpus, world, master_ip, port, kwargs, worker_init, errq), daemon=False)
             for g in range(num_gpus)]
    logger.info("Starting %d GPU worker process(es) on %s (world size %d).", num_gpus, current_host, world)
    for p in procs:
        p.start()
    failures, reported = [], 0
    try:
        while reported < num_gpus:
            try:
                g, err = errq.get(timeout=1.0)
                reported += 1
                if err is not None:
                    failures.append((g, err))
                    break                                 # the peers of a failed rank would wait in a collective forever
            except Exception:                             # queue.Empty: look for workers that died without reporting
                dead = [i for i, p in enumerate(procs) if p.exitcode not in (None, 0)]
                if dead:
                    failures.append((dead[0], "worker process exited with code %s" % procs[dead[0]].exitcode))
                    break
    finally:
        for p in procs:
            p.join(timeout=None if not failures else 5)
            if p.is_alive():
                p.terminate()
                p.join(5)
        if tracker is not None:
            tracker.free()
    if failures:
        g, err = failures[0]
        try:
            from sagemaker_algorithm_toolkit import exceptions as exc
            raise exc.AlgorithmError("XGB train call failed with exception (GPU worker %d):\n %s" % (g, err))
        except ImportError:
            raise RuntimeError("XGB train call failed with exception (GPU worker %d):\n %s" % (g, err))
    # the hyperparameter dict of the caller is consumed like the reference consumes it (num_round etc. are popped there)
    for k in ("num_round", "save_model_on_termination", "_tuning_objective_metric", "eval_metric", "early_stopping_rounds"):
        if isinstance(hyperparameters, dict):
            hyperparameters.pop(k, None)


run_training = run_training_with_dask
