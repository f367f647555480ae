#!/usr/bin/env python
"""Where the time of the device CSV path goes (config 5: 1M x 28 request body): stage times of parse_csv_device (stderr, via
B200XGB_CSV_PROFILE=1) next to the wall time of serving.csv_to_dmatrix and its Python-side steps."""
import os, sys, time
os.environ["B200XGB_CSV_PROFILE"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import bench
import sagemaker_xgboost_container_b200 as xgb
from sagemaker_xgboost_container_b200 import serving
rows = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
rng = np.random.default_rng(45)
X = rng.standard_normal((rows, 28)).astype(np.float32)
import io
import pandas as pd
t0 = time.time(); buf = io.StringIO(); pd.DataFrame(X).to_csv(buf, header=False, index=False, float_format="%.6g"); body = buf.getvalue().strip("\n"); print("body %.1f MB built in %.1f s" % (len(body) / 1e6, time.time() - t0))
for kind, payload in (("str", body), ("bytes", body.encode("utf-8"))):
    for rep in range(3):
        t0 = time.perf_counter(); d = serving.csv_to_dmatrix(payload); t1 = time.perf_counter()
        print("csv_to_dmatrix(%s) wall %.1f ms (%d x %d)" % (kind, (t1 - t0) * 1e3, d.num_row(), d.num_col()), flush=True)
