#!/usr/bin/env python
"""One launch of each histogram kernel variant for an `ncu --set full -k regex:hist_` capture (20 M x 100 by default)."""
import argparse, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
ap = argparse.ArgumentParser()
ap.add_argument("--rows", type=int, default=20_000_000)
ap.add_argument("--cols", type=int, default=100)
ap.add_argument("--modes", default="1,s,2,0")
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
for m in a.modes.split(","):
    if m == "s":
        sub = np.sort(rng.choice(a.rows, size=a.rows // 4, replace=False).astype(np.uint32))
        r = be.build_histogram_ex(b.handle, d.handle, gpair[:len(sub)], mode=4, row_ids=sub, repeats=1)     # 4: tail words by position, like the training path
    else:
        r = be.build_histogram_ex(b.handle, d.handle, gpair, mode=int(m), repeats=1)
    print(m, r[3], r[2], flush=True)
