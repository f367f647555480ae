#!/usr/bin/env python
"""Golden vectors of the container's OWN entry points, generated in the build container where /root/reference is mounted:

    python tests/golden/make_container_goldens.py

Runs the reference's `algorithm_mode.train.sagemaker_train` and `algorithm_mode.serve_utils.{parse_content_data, predict}`
UNCHANGED on top of this package bound as `xgboost`, with the oracle-backed engine (CPU), and records
  * the exact keyword arguments the container hands to `xgb.train` (so the GPU tests can replay the call),
  * the model file it saves, its last evaluation line,
  * the predictions serve_utils returns for a CSV payload.
tests/test_gpu_container_conformance.py (-m gpu, no reference tree on the GPU box) replays the same calls on the CUDA
backend and compares against these files (structure identical, leaves <= 1e-5)."""
import io
import json
import os
import shutil
import sys
import tempfile
import contextlib

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
OUT = os.path.join(HERE, "container")
G = os.path.join(HERE, "abalone")


def libsvm_to_csv(src, dst):
    with open(dst, "w") as out:
        for line in open(src):
            p = line.split()
            vals = {int(k): v for k, v in (kv.split(":") for kv in p[1:])}
            out.write(",".join([p[0]] + [vals.get(i, "") for i in range(1, 9)]) + "\n")


def main():
    import reference_stubs
    import sagemaker_xgboost_container_b200 as xgb
    from sagemaker_xgboost_container_b200 import backend, training
    from oracle.engine import OracleBackend
    backend._BACKEND = OracleBackend(error_cls=xgb.XGBoostError)
    reference_stubs.install(xgb)
    from sagemaker_xgboost_container.algorithm_mode import train as ref_train
    from sagemaker_xgboost_container.algorithm_mode import serve_utils
    os.makedirs(OUT, exist_ok=True)

    calls = []
    real_train = training.train

    def spy(params, dtrain, **kw):
        calls.append({"params": {k: (list(v) if isinstance(v, (list, tuple)) else v) for k, v in dict(params).items()},
                      "num_boost_round": kw.get("num_boost_round"), "evals": [n for _, n in kw.get("evals") or []],
                      "has_custom_metric": kw.get("custom_metric") is not None})
        return real_train(params, dtrain, **kw)
    xgb.train = spy
    sys.modules["xgboost"].train = spy

    cases = {
        "cfg1_csv": dict(hp={"objective": "reg:squarederror", "tree_method": "hist", "num_round": "50"}, fmt="csv"),           # BASELINE config 1
        "fixture_hp_libsvm": dict(hp={"objective": "reg:linear", "max_depth": "5", "eta": "0.2", "gamma": "4", "min_child_weight": "6",
                                      "subsample": "0.7", "num_round": "50"}, fmt="libsvm"),                                  # test_abalone.py:36-47
    }
    index = {}
    for name, case in cases.items():
        tmp = tempfile.mkdtemp()
        tr, va, md = os.path.join(tmp, "train"), os.path.join(tmp, "validation"), os.path.join(tmp, "model")
        os.makedirs(tr); os.makedirs(va)
        if case["fmt"] == "csv":
            libsvm_to_csv(os.path.join(G, "abalone.train_0"), os.path.join(tr, "abalone.train_0.csv"))
            libsvm_to_csv(os.path.join(G, "abalone.train_1"), os.path.join(tr, "abalone.train_1.csv"))
            libsvm_to_csv(os.path.join(G, "abalone.validation"), os.path.join(va, "abalone.validation.csv"))
            ct = "text/csv"
        else:
            shutil.copy(os.path.join(G, "abalone.train_0"), tr); shutil.copy(os.path.join(G, "abalone.validation"), va)
            ct = "libsvm"
        dc = {"train": {"ContentType": ct, "TrainingInputMode": "File", "S3DistributionType": "FullyReplicated"},
              "validation": {"ContentType": ct, "TrainingInputMode": "File", "S3DistributionType": "FullyReplicated"}}
        calls.clear()
        buf = io.StringIO()
        with contextlib.redirect_stdout(buf):
            ref_train.sagemaker_train(train_config=dict(case["hp"]), data_config=dc, train_path=tr, val_path=va, model_dir=md,
                                      sm_hosts=["algo-1"], sm_current_host="algo-1", checkpoint_config={})
        lines = [l for l in buf.getvalue().splitlines() if l.startswith("[")]
        shutil.copy(os.path.join(md, "xgboost-model"), os.path.join(OUT, name + "_model.ubj"))
        index[name] = {"format": case["fmt"], "hyperparameters": case["hp"], "train_call": calls[0], "last_eval_line": lines[-1], "eval_lines": len(lines)}
        # serving: the container's own parse + predict on a CSV payload (first 40 validation rows, label column dropped)
        if name == "cfg1_csv":
            rows = open(os.path.join(va, "abalone.validation.csv")).read().splitlines()[:40]
            payload = "\n".join(",".join(r.split(",")[1:]) for r in rows).encode("utf-8")
            dtest, ctype = serve_utils.parse_content_data(payload, "text/csv")
            boosters, formats = serve_utils.get_loaded_booster(md)
            preds = serve_utils.predict(boosters, formats, dtest, ctype, objective="reg:squarederror")
            index[name]["serve"] = {"payload": payload.decode("utf-8"), "content_type": "text/csv", "model_format": formats[0],
                                    "predictions": [float(p) for p in preds]}
        shutil.rmtree(tmp)
    json.dump(index, open(os.path.join(OUT, "index.json"), "w"), indent=1)
    print("wrote", sorted(os.listdir(OUT)))


if __name__ == "__main__":
    main()
