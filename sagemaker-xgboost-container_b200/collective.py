"""`xgboost.collective` surface used by the container (distributed.py:119-136,219-220,238-243).

Two layers:
  * in-engine NCCL communicator (histogram / statistics all-reduce inside libb200xgb.so) -- initialised here from
    a CommunicatorContext or from the torchrun environment (RANK / WORLD_SIZE / LOCAL_RANK);
  * a small host-side object broadcast for the container's membership sync (RabitHelper.synchronize), carried by
    torch.distributed (gloo) when available, or by the tracker's TCP links.
"""
import json
import os
import sys

from .backend import XGBoostError, get_backend

_state = {"rank": 0, "world": 1, "engine": False, "pg": None, "tracker_client": None}


def get_rank():
    return _state["rank"]


def get_world_size():
    return _state["world"]


def is_distributed():
    return _state["world"] > 1


def communicator_print(msg):
    msg = str(msg)
    sys.stdout.write(msg if msg.endswith("\n") else msg + "\n")
    sys.stdout.flush()


def get_processor_name():
    import socket
    return socket.gethostname()


def _torch_dist():
    import torch.distributed as dist
    return dist


def broadcast(data, root):
    """Broadcast an object from `root` (contract of xgboost.collective.broadcast).  Over the tracker's TCP links the
    payload is framed as JSON, never pickle (see tracker.py): dict / list / tuple / str / number / bool / None / bytes."""
    if _state["world"] <= 1:
        return data
    if _state["tracker_client"] is not None:
        return _state["tracker_client"].broadcast(data, root)
    dist = _torch_dist()
    box = [data if _state["rank"] == root else None]
    dist.broadcast_object_list(box, src=root, group=_state["pg"])
    return box[0]


def allreduce_sum(arr):
    import numpy as np
    if _state["world"] <= 1:
        return arr
    import torch
    dist = _torch_dist()
    t = torch.from_numpy(np.ascontiguousarray(arr).copy())
    dist.all_reduce(t, group=_state["pg"])
    return t.numpy()


def init_engine(unique_id_hex, rank, world):
    """Create the NCCL communicator inside the CUDA engine."""
    get_backend().comm_init({"nccl_unique_id": unique_id_hex, "rank": rank, "world_size": world})
    _state["engine"] = True


def init_from_env(backend="gloo"):
    """torchrun-style bootstrap: host-side process group over `backend`, NCCL unique id shipped through it."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    _state["rank"], _state["world"] = rank, world
    if world <= 1:
        return
    dist = _torch_dist()
    if not dist.is_initialized():
        dist.init_process_group(backend=backend, rank=rank, world_size=world)
    _state["pg"] = None
    be = get_backend()
    if getattr(be, "name", "") == "cuda":
        uid = be.comm_unique_id() if rank == 0 else None
        uid = broadcast(uid, 0)
        init_engine(uid, rank, world)


def init(**args):
    """xgboost.collective.init(**{dmlc_tracker_uri, dmlc_tracker_port, dmlc_task_id, dmlc_timeout, ...})."""
    if "dmlc_tracker_uri" in args:
        from .tracker import TrackerClient
        client = TrackerClient(args["dmlc_tracker_uri"], int(args["dmlc_tracker_port"]), str(args.get("dmlc_task_id", "")),
                               timeout=float(args.get("dmlc_timeout", 300)))
        client.connect()
        _state["tracker_client"] = client
        _state["rank"], _state["world"] = client.rank, client.world
        be = get_backend()
        if getattr(be, "name", "") == "cuda" and client.world > 1:
            uid = be.comm_unique_id() if client.rank == 0 else None
            uid = client.broadcast(uid, 0)
            init_engine(uid, client.rank, client.world)
    else:
        init_from_env()


def finalize():
    if _state["engine"]:
        try:
            get_backend().comm_finalize()
        except XGBoostError:
            pass
        _state["engine"] = False
    if _state["tracker_client"] is not None:
        _state["tracker_client"].close()
        _state["tracker_client"] = None
    _state["rank"], _state["world"] = 0, 1


class CommunicatorContext:
    """Context manager controlling the communicator lifetime (distributed.py:219-220)."""

    def __init__(self, **args):
        self.args = args

    def __enter__(self):
        init(**self.args)
        assert is_distributed() or True
        return self.args

    def __exit__(self, *exc):
        finalize()
