"""`train()` with the semantics of xgboost.training.train @ v3.0.5.

Call sites in the container: algorithm_mode/train.py:367-376,432-442 and checkpointing.py:74.
"""
import os

from .callback import CallbackContainer, EarlyStopping, EvaluationMonitor
from .core import Booster, DMatrix


def train(params, dtrain, num_boost_round=10, *, evals=None, obj=None, maximize=None, early_stopping_rounds=None,
          evals_result=None, verbose_eval=True, xgb_model=None, callbacks=None, custom_metric=None, feval=None):
    if feval is not None and custom_metric is None:
        custom_metric = feval
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
    cb_container = CallbackContainer(callbacks, metric=custom_metric, output_margin=callable(obj) or custom_metric is not None)
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
