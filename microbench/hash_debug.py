import os, sys, hashlib, time, argparse
import numpy as np
sys.path.insert(0, '/root/repo')
import torch, bench
rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); lr = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(lr)
import sagemaker_xgboost_container_b200 as xgb
from sagemaker_xgboost_container_b200 import collective
be = xgb.get_backend()
if world > 1: collective.init_from_env(backend="gloo")
rows = int(sys.argv[1]); rounds = int(sys.argv[2])
a = argparse.Namespace(rows=rows, cols=100, seed=43, objective="reg:squarederror", num_class=0)
r0, r1 = rank * rows // world, (rank + 1) * rows // world
X, y = bench.gen_shard(a, r0, r1, torch.device("cuda", lr))
d = xgb.DMatrix(X, label=y.cpu().numpy())
bst = xgb.Booster({"objective": "reg:squarederror", "tree_method": "hist", "max_depth": 6, "max_bin": 256, "eta": 0.3}, [d])
t0 = time.perf_counter()
for i in range(rounds): bst.update(d, i)
host_s = time.perf_counter() - t0
be.synchronize(); tot = time.perf_counter() - t0
if rank == 0:
    m = be.booster_export_model(bst.handle)
    print("world", world, "host enqueue s/round", host_s / rounds, "total s/round", tot / rounds, "base_score %.9g" % m["base_score"])
    for k in ("tree_offset", "left", "split_index", "split_bin", "default_left", "split_cond", "base_weight", "loss_chg", "sum_hess"):
        print(k, hashlib.sha256(np.ascontiguousarray(m[k]).tobytes()).hexdigest()[:12], len(m[k]))
    np.save("/root/repo/gpurun_out/model_w%d.npy" % world, {k: m[k] for k in m if hasattr(m[k], "shape")}, allow_pickle=True)
if world > 1:
    collective.finalize(); torch.distributed.destroy_process_group()
