"""Worker for the multi-GPU parity test: every rank trains on its row shard (NCCL histogram all-reduce inside the
engine) and rank 0 writes the model; launched with torchrun."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    out, n, F, rounds, objective = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]), sys.argv[5]
    extra = eval(sys.argv[6]) if len(sys.argv) > 6 else {}          # test-supplied dict literal of extra hyperparameters
    import sagemaker_xgboost_container_b200 as xgb
    from sagemaker_xgboost_container_b200 import collective
    from util import synth
    collective.init_from_env(backend="gloo")
    rank, world = collective.get_rank(), collective.get_world_size()
    kind = "bin" if objective.startswith("binary") else "reg"
    X, y = synth(n, F, 7, kind)
    lo, hi = rank * n // world, (rank + 1) * n // world
    d = xgb.DMatrix(X[lo:hi], label=y[lo:hi])
    res = {}
    bst = xgb.train(dict(dict(objective=objective, max_depth=5, eta=0.3, max_bin=256), **extra), d, num_boost_round=rounds, evals=[(d, "train")],
                    evals_result=res, verbose_eval=False)
    if rank == 0:
        with open(out + ".path", "w") as f:
            f.write("nvlink-peer-kernel" if xgb.get_backend().comm_peer_reduce_active() else "nccl")
        bst.save_model(out)
        with open(out + ".metric", "w") as f:
            f.write(repr(list(res["train"].values())[0][-1]))
    collective.finalize()


if __name__ == "__main__":
    main()
