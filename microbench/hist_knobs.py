#!/usr/bin/env python
"""Experiment matrix over the root kernel's knobs (B200XGB_ROOT_FLAGS / _R / _S), one process, env flipped between calls."""
import argparse, json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=20_000_000)
    ap.add_argument("--cols", type=int, default=100)
    ap.add_argument("--only", default="")
    a = ap.parse_args()
    import torch, bench
    import sagemaker_xgboost_container_b200 as xgb
    be = xgb.get_backend()
    ba = argparse.Namespace(rows=a.rows, cols=a.cols, seed=43, objective="reg:squarederror", num_class=0)
    X, y = bench.gen_shard(ba, 0, a.rows, torch.device("cuda", 0))
    d = xgb.DMatrix(X, label=y.cpu().numpy()); del X; torch.cuda.empty_cache()
    b = xgb.Booster({"max_bin": 256}, [d])
    rng = np.random.default_rng(1)
    gpair = np.empty((a.rows, 2), np.float32); gpair[:, 0] = rng.standard_normal(a.rows, dtype=np.float32); gpair[:, 1] = 1.0
    combos = []
    for flags in (0, 1, 4, 5, 8, 9, 2, 3, 10, 11):
        combos.append((2, flags, None, None))
    for S in (6, 8, 10):
        combos.append((2, 1, None, S))
    if a.only:
        combos = [c for i, c in enumerate(combos) if str(i) in a.only.split(",")]
    for mode, flags, R, S in combos:
        os.environ["B200XGB_ROOT_FLAGS"] = str(flags)
        for k, v in (("B200XGB_ROOT_R", R), ("B200XGB_ROOT_S", S)):
            if v is None: os.environ.pop(k, None)
            else: os.environ[k] = str(v)
        try:
            be.build_histogram_ex(b.handle, d.handle, gpair, mode=mode, repeats=1)
            hist, scales, ms, kernel = be.build_histogram_ex(b.handle, d.handle, gpair, mode=mode, repeats=3)
            print(json.dumps({"mode": mode, "flags": flags, "R": R, "S": S, "kernel": kernel, "ms": ms, "gbs": a.rows * (a.cols + 8) / ms / 1e6, "cs": int(hist[:, :, 0].sum())}), flush=True)
        except Exception as e:
            print(json.dumps({"mode": mode, "flags": flags, "R": R, "S": S, "error": str(e)[:200]}), flush=True)


if __name__ == "__main__":
    main()
