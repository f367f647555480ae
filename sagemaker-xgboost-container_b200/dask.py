"""`xgboost.dask` as far as the container touches it: `DaskDMatrix(client, data, label)` and `train(client, params, dtrain, ...)`
(distributed_gpu/dask_data_utils.py:78-83, distributed_gpu_training.py:184-195).  The container imports this module at start-up
(`algorithm_mode/train.py:46` -> `from xgboost import dask as dxgb`), so it has to exist for the package to be bound as `xgboost`.

This engine does not train through Dask: its multi-GPU mode is one process per GPU with an in-engine all-reduce (DESIGN.md
section 5), reached from the container through `multi_gpu.run_training_with_dask`.  What is here keeps the
UNBOUND reference path alive: the collections handed over (dask arrays / frames, or what the container's own reader produced:
pandas / numpy) are materialised on the calling process and trained on ITS GPU, with a warning that says so and names the
binding that uses every GPU.  Same call signatures and return value (`{"booster": Booster, "history": {...}}`) as upstream.
"""
import logging
import warnings

import numpy as np

from .core import Booster, DMatrix
from .training import train as _train

_HINT = ("xgboost.dask on the B200 engine gathers the collection onto the calling process and trains on one GPU; bind "
         "sagemaker_xgboost_container_b200.multi_gpu.run_training_with_dask (INTEGRATION.md) to use every GPU of the job")


def _materialise(x):
    if x is None:
        return None
    if hasattr(x, "compute"):                            # dask array / dataframe / series
        x = x.compute()
    if hasattr(x, "to_numpy"):
        x = x.to_numpy()
    return np.asarray(x)


class DaskDMatrix:
    """Holds the collections until `train` / `predict` needs them (upstream keeps per-worker partitions)."""

    def __init__(self, client, data, label=None, *, weight=None, base_margin=None, missing=None, silent=False, feature_names=None,
                 feature_types=None, group=None, qid=None, label_lower_bound=None, label_upper_bound=None, feature_weights=None,
                 enable_categorical=False):
        if group is not None or qid is not None:
            raise ValueError("ranking (group / qid) data is not supported on the B200 hist path")
        self.client = client
        self._data, self._label, self._weight, self._base_margin = data, label, weight, base_margin
        self.missing, self.feature_names, self.feature_types = missing, feature_names, feature_types
        self._local = None

    def num_col(self):
        return int(self._data.shape[1])

    def _dmatrix(self):
        if self._local is None:
            self._local = DMatrix(_materialise(self._data), label=_materialise(self._label), weight=_materialise(self._weight),
                                  base_margin=_materialise(self._base_margin), missing=self.missing, feature_names=self.feature_names,
                                  feature_types=self.feature_types)
        return self._local


def _local(d):
    return d._dmatrix() if isinstance(d, DaskDMatrix) else d


def train(client, params, dtrain, num_boost_round=10, *, evals=None, obj=None, early_stopping_rounds=None, xgb_model=None,
          verbose_eval=True, callbacks=None, custom_metric=None, feval=None, maximize=None):
    warnings.warn(_HINT)
    logging.getLogger(__name__).warning(_HINT)
    history = {}
    bst = _train(dict(params), _local(dtrain), num_boost_round=num_boost_round, evals=[(_local(d), name) for d, name in (evals or [])],
                 obj=obj, early_stopping_rounds=early_stopping_rounds, evals_result=history, verbose_eval=verbose_eval, xgb_model=xgb_model,
                 callbacks=callbacks, custom_metric=custom_metric, feval=feval, maximize=maximize)
    return {"booster": bst, "history": history}


def predict(client, model, data, output_margin=False, missing=np.nan, pred_leaf=False, pred_contribs=False, validate_features=True,
            iteration_range=(0, 0), strict_shape=False, **kwargs):
    bst = model["booster"] if isinstance(model, dict) else model
    if not isinstance(bst, Booster):
        raise TypeError("model must be a Booster or the dictionary returned by xgboost.dask.train")
    d = _local(data) if isinstance(data, (DaskDMatrix, DMatrix)) else DMatrix(_materialise(data), missing=missing)
    return bst.predict(d, output_margin=output_margin, pred_leaf=pred_leaf, pred_contribs=pred_contribs, validate_features=validate_features,
                       iteration_range=iteration_range, strict_shape=strict_shape)


__all__ = ["DaskDMatrix", "train", "predict"]
