"""Callback protocol of xgboost.callback @ v3.0.5 as consumed by the container
(callback.py:85-121, checkpointing.py:12,79,115-125,193,443): TrainingCallback, EvaluationMonitor,
TrainingCheckPoint, EarlyStopping, LearningRateScheduler, plus the CallbackContainer used by train()."""
import collections
import os
import pickle

import numpy as np

from . import collective
from .core import Booster, XGBoostError


class TrainingCallback:
    def before_training(self, model):
        return model

    def after_training(self, model):
        return model

    def before_iteration(self, model, epoch, evals_log):
        return False

    def after_iteration(self, model, epoch, evals_log):
        return False


def _aggcv(rlist):  # pragma: no cover - cv() is not part of the container path
    raise NotImplementedError


class CallbackContainer:
    def __init__(self, callbacks, metric=None, output_margin=True, is_cv=False):
        self.callbacks = set(callbacks)
        self._ordered = list(callbacks)
        for cb in callbacks:
            if not isinstance(cb, TrainingCallback):
                raise TypeError("callback must be an instance of `TrainingCallback`.")
        if metric is not None and not callable(metric):
            raise TypeError("metric must be callable object for monitoring.")
        self.metric = metric
        self.history = collections.OrderedDict()
        self._output_margin = output_margin
        self.is_cv = is_cv

    def before_training(self, model):
        for c in self._ordered:
            model = c.before_training(model=model)
            if not isinstance(model, Booster):
                raise TypeError("before_training should return the model")
        return model

    def after_training(self, model):
        for c in self._ordered:
            model = c.after_training(model=model)
            if not isinstance(model, Booster):
                raise TypeError("after_training should return the model")
        return model

    def before_iteration(self, model, epoch, dtrain, evals):
        return any(c.before_iteration(model, epoch, self.history) for c in self._ordered)

    def _update_history(self, score, epoch):
        for d in score:
            name, s = d[0], float(d[1])
            data_name, _, metric_name = name.partition("-")
            self.history.setdefault(data_name, collections.OrderedDict()).setdefault(metric_name, []).append(s)

    def after_iteration(self, model, epoch, dtrain, evals):
        evals = evals or []
        for _, name in evals:
            if name.find("-") != -1:
                raise ValueError("Dataset name should not contain `-`")
        score = model.eval_set(evals, epoch, self.metric, self._output_margin)
        metric_score = [tuple(s.split(":")) for s in score.split()[1:]]      # into datasets
        self._update_history(metric_score, epoch)
        ret = any(c.after_iteration(model, epoch, self.history) for c in self._ordered)
        return ret


class LearningRateScheduler(TrainingCallback):
    def __init__(self, learning_rates):
        if callable(learning_rates):
            self.learning_rates = learning_rates
        else:
            rates = list(learning_rates)
            self.learning_rates = lambda epoch: rates[epoch]
        super().__init__()

    def after_iteration(self, model, epoch, evals_log):
        model.set_param("learning_rate", self.learning_rates(epoch))
        return False


class EarlyStopping(TrainingCallback):
    def __init__(self, rounds, metric_name=None, data_name=None, maximize=None, save_best=False, min_delta=0.0):
        self.data = data_name
        self.metric_name = metric_name
        self.rounds = rounds
        self.save_best = save_best
        self.maximize = maximize
        self.stopping_history = {}
        self._min_delta = min_delta
        if self._min_delta < 0:
            raise ValueError("min_delta must be greater or equal to 0.")
        self.current_rounds = 0
        self.best_scores = {}
        self.starting_round = 0
        super().__init__()

    def before_training(self, model):
        self.starting_round = model.num_boosted_rounds()
        return model

    def _update_rounds(self, score, name, metric, model, epoch):
        def get_s(value):
            return value[0] if isinstance(value, tuple) else value

        def maximize(new, best):
            return np.greater(get_s(new) - self._min_delta, get_s(best))

        def minimize(new, best):
            return np.greater(get_s(best) - self._min_delta, get_s(new))

        if self.maximize is None:
            maximize_metrics = ("auc", "aucpr", "pre", "pre@", "map", "ndcg", "auc@", "aucpr@", "map@", "ndcg@")
            if metric != "mape" and any(metric.startswith(x) for x in maximize_metrics):
                self.maximize = True
            else:
                self.maximize = False
        improve_op = maximize if self.maximize else minimize
        if not self.stopping_history:
            self.current_rounds = 0
            self.stopping_history[name] = {metric: [score]}
            self.best_scores[name] = {metric: [score]}
            model.set_attr(best_score=str(score), best_iteration=str(epoch))
        elif not improve_op(score, self.best_scores[name][metric][-1]):
            self.stopping_history[name][metric].append(score)
            self.current_rounds += 1
        else:
            self.stopping_history[name][metric].append(score)
            self.best_scores[name][metric].append(score)
            record = self.stopping_history[name][metric][-1]
            model.set_attr(best_score=str(record), best_iteration=str(epoch))
            self.current_rounds = 0
        if self.current_rounds >= self.rounds:
            return True
        return False

    def after_iteration(self, model, epoch, evals_log):
        epoch += self.starting_round
        msg = "Must have at least 1 validation dataset for early stopping."
        if len(evals_log.keys()) < 1:
            raise ValueError(msg)
        if self.data:
            data_name = self.data
        else:
            data_name = list(evals_log.keys())[-1]
        if data_name not in evals_log:
            raise ValueError("No dataset named: %s" % data_name)
        data_log = evals_log[data_name]
        if self.metric_name:
            metric_name = self.metric_name
        else:
            metric_name = list(data_log.keys())[-1]
        if metric_name not in data_log:
            raise ValueError("No metric named: %s" % metric_name)
        score = data_log[metric_name][-1]
        return self._update_rounds(score, data_name, metric_name, model, epoch)

    def after_training(self, model):
        if not self.save_best:
            return model
        try:
            best_iteration = model.best_iteration
            best_score = model.best_score
            model = model[: best_iteration + 1]
            model.best_iteration = best_iteration
            model.best_score = best_score
        except XGBoostError as e:
            raise XGBoostError("`save_best` is not applicable to the current booster") from e
        return model


class EvaluationMonitor(TrainingCallback):
    """Prints "[epoch]\\t<data>-<metric>:<score>" -- the CloudWatch regexes of the container
    (algorithm_mode/metrics.py:21-42) depend on this exact shape."""

    def __init__(self, rank=0, period=1, show_stdv=False):
        self.printer_rank = rank
        self.show_stdv = show_stdv
        self.period = period
        assert period > 0
        self._latest = None
        super().__init__()

    def _fmt_metric(self, data, metric, score, std):
        if std is not None and self.show_stdv:
            return "\t%s:%.5f+%.5f" % (data + "-" + metric, score, std)
        return "\t%s:%.5f" % (data + "-" + metric, score)

    def after_iteration(self, model, epoch, evals_log):
        if not evals_log:
            return False
        msg = "[%d]" % epoch
        if collective.get_rank() == self.printer_rank:
            for data, metric in evals_log.items():
                for metric_name, log in metric.items():
                    stdv = None
                    if isinstance(log[-1], tuple):
                        score, stdv = log[-1][0], log[-1][1]
                    else:
                        score = log[-1]
                    msg += self._fmt_metric(data, metric_name, score, stdv)
            msg += "\n"
            if (epoch % self.period) == 0 or self.period == 1:
                collective.communicator_print(msg)
                self._latest = None
            else:
                self._latest = msg
        return False

    def after_training(self, model):
        if collective.get_rank() == self.printer_rank and self._latest is not None:
            collective.communicator_print(self._latest)
        return model


class TrainingCheckPoint(TrainingCallback):
    """Every `interval` iterations rank 0 writes <directory>/<name>_<iter>.ubj (or .pkl)."""

    default_format = "ubj"

    def __init__(self, directory, name="model", as_pickle=False, interval=100):
        self._path = os.fspath(directory)
        self._name = name
        self._as_pickle = as_pickle
        self._iterations = interval
        self._epoch = 0
        self._start = 0
        super().__init__()

    def before_training(self, model):
        self._start = model.num_boosted_rounds()
        return model

    def after_iteration(self, model, epoch, evals_log):
        if self._epoch == self._iterations:
            path = os.path.join(self._path, self._name + "_" + str(epoch + self._start) +
                                (".pkl" if self._as_pickle else "." + self.default_format))
            self._epoch = 0
            if collective.get_rank() == 0:
                if self._as_pickle:
                    with open(path, "wb") as fd:
                        pickle.dump(model, fd)
                else:
                    model.save_model(path)
        self._epoch += 1
        return False
