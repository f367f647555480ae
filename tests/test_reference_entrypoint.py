"""BASELINE config 1: the reference container's own entry point `algorithm_mode.train.sagemaker_train` run UNCHANGED on
top of this package bound as `xgboost` -- on CPU through the oracle-backed engine (the reference tree only exists in the
build container; skipped elsewhere).  Asserts what test/integration/local/test_abalone.py asserts (model file present, no
failure) plus numbers the reference's tests never pin: the eval lines, and a model file the oracle can read back."""
import json
import os
import re

import numpy as np
import pytest

import reference_stubs

pytestmark = pytest.mark.skipif(not reference_stubs.reference_available(), reason="/root/reference is not mounted here")
G = os.path.join(os.path.dirname(__file__), "golden", "abalone")


@pytest.fixture()
def container(monkeypatch):
    import sagemaker_xgboost_container_b200 as xgb
    from sagemaker_xgboost_container_b200 import backend
    from oracle.engine import OracleBackend
    monkeypatch.setattr(backend, "_BACKEND", OracleBackend(error_cls=xgb.XGBoostError))
    reference_stubs.install(xgb)
    from sagemaker_xgboost_container.algorithm_mode import train as ref_train
    return xgb, ref_train


def _libsvm_to_csv(src, dst):
    with open(dst, "w") as out:
        for line in open(src):
            p = line.split()
            vals = {int(k): v for k, v in (kv.split(":") for kv in p[1:])}
            out.write(",".join([p[0]] + [vals.get(i, "") for i in range(1, 9)]) + "\n")


def test_sagemaker_train_abalone_csv_50_rounds(container, tmp_path, capsys):
    xgb, ref_train = container
    tr, va, model_dir = tmp_path / "train", tmp_path / "validation", tmp_path / "model"
    tr.mkdir(); va.mkdir()
    _libsvm_to_csv(os.path.join(G, "abalone.train_0"), tr / "abalone.train_0.csv")
    _libsvm_to_csv(os.path.join(G, "abalone.train_1"), tr / "abalone.train_1.csv")
    _libsvm_to_csv(os.path.join(G, "abalone.validation"), va / "abalone.validation.csv")
    hp = {"objective": "reg:squarederror", "tree_method": "hist", "num_round": "50", "max_depth": "5", "eta": "0.2", "gamma": "4", "min_child_weight": "6"}
    data_config = {"train": {"ContentType": "text/csv", "TrainingInputMode": "File", "S3DistributionType": "FullyReplicated"},
                   "validation": {"ContentType": "text/csv", "TrainingInputMode": "File", "S3DistributionType": "FullyReplicated"}}
    ref_train.sagemaker_train(train_config=hp, data_config=data_config, train_path=str(tr), val_path=str(va), model_dir=str(model_dir),
                              sm_hosts=["algo-1"], sm_current_host="algo-1", checkpoint_config={})
    out = capsys.readouterr().out
    lines = [l for l in out.splitlines() if re.match(r"^\[\d+\]\ttrain-rmse:", l)]
    assert len(lines) == 50 and "validation-rmse:" in lines[-1]
    last_val = float(lines[-1].split("validation-rmse:")[1])
    assert last_val < 2.6
    model_file = model_dir / "xgboost-model"
    assert model_file.exists()
    # the saved model is UBJSON in the reference's schema: the independent reader + the oracle reproduce the eval line
    from oracle import gbt_oracle as O, ubjson
    m = ubjson.model_from_xgb_json(ubjson.load(str(model_file)))
    assert len(m["tree_info"]) == 50 and m["num_feature"] == 8
    Xv = np.genfromtxt(va / "abalone.validation.csv", delimiter=",", dtype=np.float32)
    rm = float(np.sqrt(np.mean((O.predict_margin(m, Xv[:, 1:])[:, 0] - Xv[:, 0]) ** 2)))
    assert abs(rm - last_val) < 1e-4
    # and serving loads it the way serve_utils.get_loaded_booster does
    b = xgb.Booster()
    b.load_model(str(model_file))
    assert json.loads(b.save_config())["learner"]["objective"]["name"] == "reg:squarederror"


def test_sagemaker_train_libsvm_with_checkpoints_and_early_stopping(container, tmp_path, capsys):
    xgb, ref_train = container
    import shutil
    tr, va, model_dir, ck = tmp_path / "train", tmp_path / "validation", tmp_path / "model", tmp_path / "ck"
    tr.mkdir(); va.mkdir(); ck.mkdir()
    shutil.copy(os.path.join(G, "abalone.train_0"), tr / "abalone.train_0")
    shutil.copy(os.path.join(G, "abalone.validation"), va / "abalone.validation")
    hp = {"objective": "reg:linear", "num_round": "12", "max_depth": "4", "eta": "0.3", "early_stopping_rounds": "3", "eval_metric": "rmse",
          "save_model_on_termination": "true"}
    dc = {"train": {"ContentType": "libsvm", "TrainingInputMode": "File", "S3DistributionType": "FullyReplicated"},
          "validation": {"ContentType": "libsvm", "TrainingInputMode": "File", "S3DistributionType": "FullyReplicated"}}
    ref_train.sagemaker_train(train_config=hp, data_config=dc, train_path=str(tr), val_path=str(va), model_dir=str(model_dir),
                              sm_hosts=["algo-1"], sm_current_host="algo-1", checkpoint_config={"LocalPath": str(ck)})
    assert (model_dir / "xgboost-model").exists()
    b = xgb.Booster(model_file=str(model_dir / "xgboost-model"))
    assert 1 <= b.num_boosted_rounds() <= 12 and b.num_features() == 9          # libsvm indices kept: 9 columns


def test_bad_labels_become_user_error(container, tmp_path):
    xgb, ref_train = container
    from sagemaker_algorithm_toolkit import exceptions as exc
    tr = tmp_path / "train"
    tr.mkdir()
    (tr / "d.csv").write_text("5,1,2\n7,3,4\n")
    hp = {"objective": "binary:logistic", "num_round": "2"}
    dc = {"train": {"ContentType": "text/csv", "TrainingInputMode": "File", "S3DistributionType": "FullyReplicated"}}
    with pytest.raises(exc.UserError, match="label must be in"):
        ref_train.sagemaker_train(train_config=hp, data_config=dc, train_path=str(tr), val_path=None, model_dir=str(tmp_path / "m"),
                                  sm_hosts=["algo-1"], sm_current_host="algo-1", checkpoint_config={})


def test_sagemaker_train_kfold_branch(container, tmp_path, monkeypatch):
    """train.py:378-459: RepeatedKFold -> DMatrix.slice(idx) -> xgb.train per fold -> booster.predict(fold) -> N model files
    (what test/integration/local/test_kfold.py asserts), plus predictions.csv from the ValidationPredictionRecorder."""
    xgb, ref_train = container
    import shutil
    tr, va, model_dir, out = tmp_path / "train", tmp_path / "validation", tmp_path / "model", tmp_path / "output"
    tr.mkdir(); va.mkdir(); out.mkdir()
    shutil.copy(os.path.join(G, "abalone.train_0"), tr / "abalone.train_0")
    shutil.copy(os.path.join(G, "abalone.validation"), va / "abalone.validation")
    monkeypatch.setenv("SM_OUTPUT_DATA_DIR", str(out))
    hp = {"objective": "reg:squarederror", "num_round": "5", "max_depth": "3", "_kfold": "3", "eval_metric": "rmse"}
    dc = {"train": {"ContentType": "libsvm", "TrainingInputMode": "File", "S3DistributionType": "FullyReplicated"},
          "validation": {"ContentType": "libsvm", "TrainingInputMode": "File", "S3DistributionType": "FullyReplicated"}}
    ref_train.sagemaker_train(train_config=hp, data_config=dc, train_path=str(tr), val_path=str(va), model_dir=str(model_dir),
                              sm_hosts=["algo-1"], sm_current_host="algo-1", checkpoint_config={})
    assert sorted(os.listdir(model_dir)) == ["xgboost-model-0", "xgboost-model-1", "xgboost-model-2"]
    assert (out / "predictions.csv").exists()
    rows = open(out / "predictions.csv").read().strip().splitlines()
    assert len(rows) == 1461 + 626          # one out-of-fold prediction per row of train + validation
