"""`train()` with the semantics of xgboost.training.train @ v3.0.5.

Call sites in the container: algorithm_mode/train.py:367-376,432-442 and checkpointing.py:74.
"""
import os

from .callback import CallbackContainer, EarlyStopping, EvaluationMonitor
from .core import Booster, DMatrix


def train(params, dtrain, num_boost_round=10, *, evals=None, obj=None, maximize=None, early_stopping_rounds=None,
          evals_result=None, verbose_eval=True, xgb_model=None, callbacks=None, custom_metric=None, feval=None):
    # upstream: `custom_metric` sees TRANSFORMED predictions when a built-in objective is used, the legacy `feval` sees raw
    # margins (output_margin = callable(obj) or metric_fn is feval).  The container passes custom_metric= (train.py:372,437).
    if feval is not None and custom_metric is not None:
        raise ValueError("Both `feval` and `custom_metric` are supplied.  Use `custom_metric` instead.")
    metric_fn = custom_metric if custom_metric is not None else feval
    output_margin = callable(obj) or (feval is not None and custom_metric is None)
    callbacks = [] if callbacks is None else list(callbacks)
    evals = list(evals) if evals else []
    for va, _ in evals:
        if not isinstance(va, DMatrix):
            raise TypeError("Invalid type for the `evals`.")
    bst = Booster(params, [dtrain] + [d[0] for d in evals], model_file=xgb_model)
    if verbose_eval:
        period = 1 if isinstance(verbose_eval, bool) else int(verbose_eval)
        callbacks.append(EvaluationMonitor(period=period))
    if early_stopping_rounds:
        callbacks.append(EarlyStopping(rounds=early_stopping_rounds, maximize=maximize))
    cb_container = CallbackContainer(callbacks, metric=metric_fn, output_margin=output_margin)
    bst = cb_container.before_training(bst)
    start = 0
    for i in range(start, num_boost_round):
        if cb_container.before_iteration(bst, i, dtrain, evals):
            break
        bst.update(dtrain, iteration=i, fobj=obj)
        if cb_container.after_iteration(bst, i, dtrain, evals):
            break
    bst = cb_container.after_training(bst)
    if evals_result is not None:
        evals_result.update(cb_container.history)
    return bst.copy()


def cv(params, dtrain, num_boost_round=10, nfold=3, stratified=False, folds=None, metrics=(), obj=None, maximize=None, early_stopping_rounds=None,
       as_pandas=True, verbose_eval=None, show_stdv=True, seed=0, callbacks=None, shuffle=True, custom_metric=None, feval=None):
    """k-fold cross validation with the semantics of xgboost.cv (used by script-mode customer code,
    test/resources/boston/single_machine_customer_script.py:94): returns mean/std of every metric per round."""
    import numpy as np
    output_margin = callable(obj) or (feval is not None and custom_metric is None)
    if feval is not None and custom_metric is None:
        custom_metric = feval
    n = dtrain.num_row()
    rng = np.random.RandomState(seed)
    idx = rng.permutation(n) if shuffle else np.arange(n)
    if folds is not None:
        splits = list(folds.split(np.zeros(n), dtrain.get_label())) if hasattr(folds, "split") else list(folds)
    elif stratified:
        y = dtrain.get_label()
        order = np.argsort(y[idx], kind="stable")
        parts = [idx[order][k::nfold] for k in range(nfold)]
        splits = [(np.concatenate([parts[j] for j in range(nfold) if j != k]), parts[k]) for k in range(nfold)]
    else:
        parts = np.array_split(idx, nfold)
        splits = [(np.concatenate([parts[j] for j in range(nfold) if j != k]), parts[k]) for k in range(nfold)]
    params = dict(params) if isinstance(params, dict) else dict(params or [])
    if metrics:
        params["eval_metric"] = list(metrics) if not isinstance(metrics, str) else [metrics]
    packs = []
    for tr, te in splits:
        dtr, dte = dtrain.slice(np.sort(tr)), dtrain.slice(np.sort(te))
        packs.append((Booster(params, [dtr, dte]), dtr, dte))
    history = {}
    best, best_round, wait = None, 0, 0
    for i in range(num_boost_round):
        per_fold = []
        for bst, dtr, dte in packs:
            bst.update(dtr, i, obj)
            msg = bst.eval_set([(dtr, "train"), (dte, "test")], i, custom_metric, output_margin)
            per_fold.append([tuple(s.split(":")) for s in msg.split()[1:]])
        keys = [k for k, _ in per_fold[0]]
        for j, key in enumerate(keys):
            vals = np.array([float(f[j][1]) for f in per_fold])
            history.setdefault(key + "-mean", []).append(float(vals.mean()))
            history.setdefault(key + "-std", []).append(float(vals.std()))
        if verbose_eval:
            print("[%d]" % i + "".join("\t%s:%.5f%s" % (k, history[k + "-mean"][-1], ("+%.5f" % history[k + "-std"][-1]) if show_stdv else "") for k in keys))
        if early_stopping_rounds:
            key = [k for k in keys if k.startswith("test-")][-1]
            cur = history[key + "-mean"][-1]
            mx = maximize if maximize is not None else any(key.split("-", 1)[1].startswith(m) for m in ("auc", "aucpr", "map", "ndcg"))
            if best is None or (cur > best if mx else cur < best):
                best, best_round, wait = cur, i, 0
            else:
                wait += 1
                if wait >= early_stopping_rounds:
                    history = {k: v[:best_round + 1] for k, v in history.items()}
                    break
    if as_pandas:
        try:
            import pandas as pd
            return pd.DataFrame.from_dict(history)
        except ImportError:
            pass
    return history
