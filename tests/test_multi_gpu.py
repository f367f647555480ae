"""Multi-GPU parity (needs >= 2 GPUs; `gpurun --gpus 2 -- pytest tests/test_multi_gpu.py -m gpu`): rows sharded over the
ranks + NCCL all-reduce of the int64 histograms must give the SAME model as one GPU on all rows (fixed-point sums are
order independent), and the all-reduced evaluation metric must equal the single-GPU one."""
import os
import subprocess
import sys

import numpy as np
import pytest

from util import assert_same_structure, synth

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ngpu():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


@pytest.mark.parametrize("collective_path", ["nvlink-peer-kernel", "nccl"])
@pytest.mark.parametrize("objective,extra", [("reg:squarederror", {}), ("binary:logistic", {}),
                                             ("reg:squarederror", dict(grow_policy="lossguide", max_leaves=12, max_depth=0))])
def test_two_rank_training_equals_single_gpu(xgb, tmp_path, objective, extra, collective_path):
    """Both histogram all-reduce paths (the NVLink peer-memory kernel inside the tree graph, and NCCL between graph segments);
    depth-wise and loss-guided growth (the latter all-reduces through the fixed staging slot)."""
    if _ngpu() < 2:
        pytest.skip("needs 2 GPUs")
    env = dict(os.environ)
    if collective_path == "nccl":
        env["B200XGB_NO_PEER_REDUCE"] = "1"
    n, F, rounds = 40000, 20, 6
    out = str(tmp_path / "model.ubj")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1", "--master-port",
           "29611", os.path.join(ROOT, "tests", "helpers", "train_shard_worker.py"), out, str(n), str(F), str(rounds), objective, repr(extra)]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    kind = "bin" if objective.startswith("binary") else "reg"
    X, y = synth(n, F, 7, kind)
    d = xgb.DMatrix(X, label=y)
    res = {}
    single = xgb.train(dict(dict(objective=objective, max_depth=5, eta=0.3, max_bin=256), **extra), d, num_boost_round=rounds, evals=[(d, "train")],
                       evals_result=res, verbose_eval=False)
    multi = xgb.Booster(model_file=out)
    be = xgb.get_backend()
    m1, m2 = be.booster_export_model(single.handle), be.booster_export_model(multi.handle)
    assert_same_structure(m2, m1)
    np.testing.assert_array_equal(m2["split_cond"], m1["split_cond"])          # bit-identical leaves: exact integer histograms
    assert open(out + ".path").read() == collective_path
    metric = float(open(out + ".metric").read())
    assert abs(metric - list(res["train"].values())[0][-1]) < 1e-9
