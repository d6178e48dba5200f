"""Callbacks that collect training curves into pandas DataFrames (parity: python/mxnet/notebook/callback.py ``PandasLogger``; the live
bokeh charts of the reference need the ``bokeh`` package, which is not part of this environment — ``LiveBokehChart`` and its subclasses say
so when constructed)."""
from __future__ import annotations

import datetime
import time

__all__ = ["PandasLogger", "LiveBokehChart", "LiveTimeSeries", "LiveLearningCurve", "args_wrapper"]


def _add_new_columns(df, valueset):
    for c in valueset:
        if c not in df.columns:
            df[c] = None


def _extend(df, row):
    import pandas as pd
    _add_new_columns(df, row)
    df.loc[len(df)] = [row.get(c) for c in df.columns]
    return df


class PandasLogger:
    """``train_df`` / ``eval_df`` / ``epoch_df`` DataFrames filled by the ``train_cb`` / ``eval_cb`` / ``epoch_cb`` callbacks::

        log = PandasLogger(batch_size=32, frequent=10)
        mod.fit(it, eval_data=val, num_epoch=3, **log.callback_args())
    """

    def __init__(self, batch_size, frequent=50):
        import pandas as pd
        self.batch_size, self.frequent = batch_size, frequent
        self._dataframes = {"train": pd.DataFrame(), "eval": pd.DataFrame(), "epoch": pd.DataFrame()}
        self.last_time = time.time()
        self.start_time = datetime.datetime.now()
        self.last_epoch_time = datetime.datetime.now()

    train_df = property(lambda self: self._dataframes["train"])
    eval_df = property(lambda self: self._dataframes["eval"])
    epoch_df = property(lambda self: self._dataframes["epoch"])
    all_dataframes = property(lambda self: self._dataframes)

    def elapsed(self):
        return datetime.datetime.now() - self.start_time

    def append_metrics(self, metrics, df_name):
        self._dataframes[df_name] = _extend(self._dataframes[df_name], dict(metrics))

    def train_cb(self, param):
        if param.nbatch % self.frequent == 0:
            self._process_batch(param, "train")

    def eval_cb(self, param):
        self._process_batch(param, "eval")

    def _process_batch(self, param, dataframe):
        now = time.time()
        metrics = dict(param.eval_metric.get_name_value()) if param.eval_metric is not None else {}
        speed = self.frequent / (now - self.last_time) if now > self.last_time else float("inf")
        metrics.update({"batches_per_sec": speed, "records_per_sec": speed * self.batch_size, "elapsed": self.elapsed(),
                        "minibatch_count": param.nbatch, "epoch": param.epoch})
        self.append_metrics(metrics, dataframe)
        self.last_time = now

    def epoch_cb(self, epoch=None, symbol=None, arg_params=None, aux_params=None):
        now = datetime.datetime.now()
        self.append_metrics({"elapsed": self.elapsed(), "epoch_time": now - self.last_epoch_time}, "epoch")
        self.last_epoch_time = now

    def callback_args(self):
        return {"batch_end_callback": self.train_cb, "epoch_end_callback": self.epoch_cb}


class LiveBokehChart:
    def __init__(self, *args, **kwargs):
        raise ImportError("mx.notebook live charts need the 'bokeh' package, which is not installed in this environment; use PandasLogger")


LiveTimeSeries = LiveLearningCurve = LiveBokehChart


def args_wrapper(*args):
    """Merge several ``callback_args()`` dicts: callbacks for the same hook are chained into a list."""
    out = {}
    for a in args:
        for k, v in a.callback_args().items():
            out.setdefault(k, []).append(v)
    return out
