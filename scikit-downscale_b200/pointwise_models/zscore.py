"""ZScoreRegressor on the B200 path (skdownscale/pointwise_models/zscore.py:11-353).

Z-score bias correction: ``fit`` compares, for every day of the year, the mean and standard deviation of the
historical model series X and the observations y pooled over all years and a centred window of ``window_width`` day
columns (``shift_ = mean_y - mean_X``, ``scale_ = std_y / std_X``); ``predict`` standardises a new series with its own
centred rolling mean / standard deviation and rescales it with the corrected statistics.

The calendar bookkeeping (which row of the record is which (year, day-of-year), which day columns a window pools)
is built here on the host as three small integer tables; the arithmetic runs in ``csrc/zscore_kernels.cu`` for all
cells at once.  The reference builds the same layout with xarray (``_reshape`` / ``_calc_stats``); xarray is not
installed where this package is developed, so the FIT statistics follow the reference's code and xarray's documented
window semantics and are checked against the reference's known-answer tests only — predict is checked against the
live reference (``tests/golden/zscore_*.npz``).
"""

from __future__ import annotations

import numpy as np
import pandas as pd
import torch
from sklearn.base import RegressorMixin
from sklearn.exceptions import NotFittedError

from .. import engine
from .base import TimeSynchronousDownscaler, cuda_device, series_to_device


def day_tables(index, window_width: int):
    """Host tables of ``sdb_zscore_fit`` for a record with time axis ``index``:

    ``day_rows [n_years, n_days]``  row of the record per (year, day column), -1 = absent — the outer join of the yearly
    pieces that ``groupby('time.year').map(split)`` builds (zscore.py:145-148): the day axis is the sorted union of the
    days of year that occur; ``pos_col [n_days + w]`` the day column behind each position of the bookended year
    ``[last ceil(w/2) columns | all | first w//2]`` (zscore.py:150-157; ``-window_width // 2`` is ``-ceil(w/2)``);
    ``col_count [n_days]`` years holding each column; ``n_kept`` retained windows: positions ``[n, len - n)`` of the
    centred rolling window, ``n = w//2 + 1`` (zscore.py:185-190) — window k pools positions k+1 .. k+w."""
    index = pd.DatetimeIndex(index)
    w = int(window_width)
    years = np.unique(index.year)
    days = np.unique(index.dayofyear)
    rows = np.full((len(years), len(days)), -1, dtype=np.int32)
    rows[np.searchsorted(years, index.year), np.searchsorted(days, index.dayofyear)] = np.arange(len(index), dtype=np.int32)
    n_days = len(days)
    late, early = -((-w) // 2), w // 2
    if late > n_days or early > n_days:
        raise ValueError(f'window_width={w} is wider than the {n_days} days of year in the record')
    pos_col = np.concatenate([np.arange(n_days - late, n_days), np.arange(n_days), np.arange(early)]).astype(np.int32)
    n_kept = len(pos_col) - 2 * (w // 2 + 1)
    col_count = (rows >= 0).sum(axis=0).astype(np.int32)
    return rows, pos_col, col_count, n_kept, days


class ZScoreRegressor(RegressorMixin, TimeSynchronousDownscaler):
    """Z Score Regressor bias correction (zscore.py:11-122).

    Parameters
    ----------
    window_width : int
        Size of the moving window in days (default 31)."""

    _fit_attributes = ['shift_', 'scale_']
    _timestep = 'M'

    def __init__(self, window_width: int = 31) -> None:
        if window_width <= 0:
            raise ValueError(f'window_width must be positive, got {window_width}')
        self.window_width = window_width

    # ---- batched (all cells): X, y [T, C] CUDA tensors sharing the time axis ``index``
    def fit_batched(self, X: torch.Tensor, y: torch.Tensor, index, valid=None, want_stats: bool = False):
        if index is None or len(index) != X.shape[0]:
            raise ValueError('ZScoreRegressor needs the time axis labels of the record: pass a DatetimeIndex of len(X)')
        rows, pos_col, col_count, n_kept, days = day_tables(index, self.window_width)
        if n_kept <= 0:
            raise ValueError(f'the record has {len(days)} days of year: no complete window of {self.window_width}')
        self._flag = torch.zeros(1, dtype=torch.int32, device=X.device)
        self._shift, self._scale, self._stats = engine.zscore_fit(X, y.to(X.dtype), rows, pos_col, col_count, self.window_width,
                                                                  n_kept, valid, self._flag, want_stats)
        self._valid = valid
        n = self.window_width // 2 + 1
        self._days = days[pos_col[n:n + n_kept]]     # day-of-year label of each retained window's centre position
        self.n_features_in_ = 1
        return self

    def check_fit(self):
        if int(self._flag.item()) != 0:
            self._flag.zero_()
            raise ValueError('Input contains NaN or infinity.')

    def predict_batched(self, X: torch.Tensor, out_dtype=None, out=None) -> torch.Tensor:
        if not hasattr(self, '_shift'):
            raise NotFittedError(f"This {type(self).__name__} instance is not fitted yet. Call 'fit' with "
                                 "appropriate arguments before using this estimator.")
        need = min(X.shape[0], 364)                                 # zscore.py:300
        if self._shift.shape[0] < need:
            raise IndexError('positional indexers are out-of-bounds')   # shift.iloc[inds], zscore.py:314
        if X.dtype != self._shift.dtype:
            X = X.to(self._shift.dtype)
        return engine.zscore_predict(X, self._shift, self._scale, self.window_width, out_dtype, self._valid, self._flag, out)

    # ---- per-cell API of the reference: DataFrames / Series with a DatetimeIndex
    def fit(self, X, y):
        X, y = self._frames(X, y)
        dev = cuda_device()
        x_t, index, _ = series_to_device(X, dev)
        y_t, _, _ = series_to_device(y, dev)
        if x_t.shape[1] != 1:
            raise ValueError(f'Zscore only supports 1 feature, found {x_t.shape[1]}')
        if y_t.shape[1] != 1:
            raise ValueError('y must have exactly one column')
        if x_t.dtype != y_t.dtype:
            x_t, y_t = x_t.to(torch.float64), y_t.to(torch.float64)
        self.fit_batched(x_t, y_t, index, want_stats=True)
        self.check_fit()
        day = pd.Index(self._days, name='day')
        stats = self._stats.cpu().numpy()[:, :, 0]
        self.fit_stats_dict_ = {k: pd.Series(stats[i], index=day) for i, k in enumerate(('X_mean', 'X_std', 'y_mean', 'y_std'))}
        self.shift_ = pd.Series(self._shift.cpu().numpy()[:, 0], index=day)
        self.scale_ = pd.Series(self._scale.cpu().numpy()[:, 0], index=day)
        return self

    def predict(self, X):
        if not hasattr(self, 'shift_') or not hasattr(self, 'scale_'):
            raise NotFittedError(f"This {type(self).__name__} instance is not fitted yet. Call 'fit' with "
                                 "appropriate arguments before using this estimator.")
        X = self._frames(X)
        if X.shape[1] != 1:
            raise ValueError(f'X must have exactly 1 feature, got {X.shape[1]}')
        dev = cuda_device()
        x_t, index, columns = series_to_device(X, dev)
        # like the reference, predict only needs shift_ / scale_ (its own test assigns them by hand)
        dt = x_t.dtype
        self._shift = torch.as_tensor(np.asarray(self.shift_, dtype=np.float64), device=dev).to(dt).reshape(-1, 1)
        self._scale = torch.as_tensor(np.asarray(self.scale_, dtype=np.float64), device=dev).to(dt).reshape(-1, 1)
        self._valid = None
        self._flag = torch.zeros(1, dtype=torch.int32, device=dev)
        out = self.predict_batched(x_t, out_dtype=torch.float64)
        self.check_fit()
        return pd.DataFrame(out.cpu().numpy(), index=index, columns=columns)
