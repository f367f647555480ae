"""GPU-backend conformance with the container's own entry points.

The reference tree does not exist on the GPU box, so the container's calls are REPLAYED: tests/golden/make_container_goldens.py
ran `sagemaker_train` and `serve_utils.parse_content_data / predict` unchanged (build container, oracle-backed engine) and
recorded the exact `xgb.train` keyword arguments, the DMatrix URIs' shape, the saved model, the last evaluation line and the
served predictions.  Here the same calls run on the CUDA backend through the C-ABI and must reproduce those files:
tree structure identical, leaf values within 1e-5 (BASELINE.json north_star), evaluation lines to 5 decimals."""
import json
import os
import shutil

import numpy as np
import pytest

from util import assert_same_structure, max_leaf_diff

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(__file__), "golden")
INDEX = json.load(open(os.path.join(G, "container", "index.json")))
LEAF_TOL = 1e-5


def _libsvm_to_csv(src, dst):
    with open(dst, "w") as out:
        for line in open(src):
            p = line.split()
            vals = {int(k): v for k, v in (kv.split(":") for kv in p[1:])}
            out.write(",".join([p[0]] + [vals.get(i, "") for i in range(1, 9)]) + "\n")


def _channels(tmp_path, fmt):
    tr, va = tmp_path / "train", tmp_path / "validation"
    tr.mkdir(); va.mkdir()
    A = os.path.join(G, "abalone")
    if fmt == "csv":
        _libsvm_to_csv(os.path.join(A, "abalone.train_0"), tr / "abalone.train_0.csv")
        _libsvm_to_csv(os.path.join(A, "abalone.train_1"), tr / "abalone.train_1.csv")
        _libsvm_to_csv(os.path.join(A, "abalone.validation"), va / "abalone.validation.csv")
        uri = "{}?format=csv&label_column=0&delimiter=,"            # data_utils.py:309-313
    else:
        shutil.copy(os.path.join(A, "abalone.train_0"), tr)
        shutil.copy(os.path.join(A, "abalone.validation"), va)
        uri = "{}?format=libsvm"                                    # data_utils.py:361
    return uri.format(tr), uri.format(va)


@pytest.mark.parametrize("case", sorted(INDEX))
def test_sagemaker_train_call_replayed_on_the_cuda_backend(xgb, case, tmp_path):
    from oracle import ubjson
    rec = INDEX[case]
    call = rec["train_call"]
    utrain, uval = _channels(tmp_path, rec["format"])
    dtrain, dval = xgb.DMatrix(utrain), xgb.DMatrix(uval)          # the loaders the container uses, straight onto the device
    assert xgb.get_backend().name == "cuda"
    res = {}
    bst = xgb.train(call["params"], dtrain, num_boost_round=call["num_boost_round"], evals=[(dtrain, "train"), (dval, "validation")],
                    evals_result=res, verbose_eval=False)
    out = str(tmp_path / "xgboost-model")
    bst.save_model(out)                                            # train.py:479-480: extension-less => UBJSON
    got = ubjson.model_from_xgb_json(ubjson.load(out))
    ref = ubjson.model_from_xgb_json(ubjson.load(os.path.join(G, "container", case + "_model.ubj")))
    assert abs(got["base_score"] - ref["base_score"]) <= 1e-6 * max(1.0, abs(ref["base_score"]))
    assert_same_structure(got, ref)
    assert max_leaf_diff(got, ref) <= LEAF_TOL
    line = "[%d]\ttrain-rmse:%.5f\tvalidation-rmse:%.5f" % (call["num_boost_round"] - 1, res["train"]["rmse"][-1], res["validation"]["rmse"][-1])
    assert line == rec["last_eval_line"]                           # what the container prints for CloudWatch (metrics.py:27,36)


def test_serve_utils_predict_replayed_on_the_cuda_backend(xgb):
    """serve_utils.parse_content_data (CSV payload -> DMatrix) + serve_utils.predict (Booster.predict on the loaded model)."""
    rec = INDEX["cfg1_csv"]["serve"]
    bst = xgb.Booster()
    bst.load_model(os.path.join(G, "container", "cfg1_csv_model.ubj"))           # serve_utils.get_loaded_booster: Booster.load_model
    rows = [r.split(",") for r in rec["payload"].strip().split("\n")]           # encoder.csv_to_dmatrix: split on ',' -> float32 matrix
    X = np.array([[np.nan if v == "" else float(v) for v in r] for r in rows], dtype=np.float32)
    d = xgb.DMatrix(X)
    pred = bst.predict(d, validate_features=False)                               # serve_utils.py:244-250
    np.testing.assert_allclose(pred, np.array(rec["predictions"], np.float32), rtol=0, atol=1e-5)
    leaves = bst.predict(d, pred_leaf=True)
    assert leaves.shape == (len(rows), 50)
