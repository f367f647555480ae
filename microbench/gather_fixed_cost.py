#!/usr/bin/env python
"""Fixed cost of a gather-kernel launch: time vs number of gathered rows (row subsets of a 20M x 100 matrix)."""
import argparse, json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch, bench
import sagemaker_xgboost_container_b200 as xgb
be = xgb.get_backend()
rows = 20_000_000
ba = argparse.Namespace(rows=rows, cols=100, seed=43, objective="reg:squarederror", num_class=0)
X, y = bench.gen_shard(ba, 0, rows, torch.device("cuda", 0))
d = xgb.DMatrix(X, label=y.cpu().numpy()); del X; torch.cuda.empty_cache()
b = xgb.Booster({"max_bin": 256}, [d])
rng = np.random.default_rng(1)
gpair = np.empty((rows, 2), np.float32); gpair[:, 0] = rng.standard_normal(rows, dtype=np.float32); gpair[:, 1] = 1.0
for m in (150_000, 600_000, 1_200_000, 2_400_000, 4_800_000):
    sub = np.sort(rng.choice(rows, size=m, replace=False).astype(np.uint32))
    be.build_histogram_ex(b.handle, d.handle, gpair[:m], mode=0, row_ids=sub, repeats=2)
    r = be.build_histogram_ex(b.handle, d.handle, gpair[:m], mode=0, row_ids=sub, repeats=10)
    print(json.dumps({"rows": m, "ms": r[2], "ns_per_row": r[2] * 1e6 / m}), flush=True)
