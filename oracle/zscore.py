"""ZScoreRegressor restated in numpy (skdownscale/pointwise_models/zscore.py:11-353) — TEST INFRASTRUCTURE ONLY.

Pin status (see tests/test_oracle_golden.py):

* ``zscore_predict`` (zscore.py:68-110, 242-353: pandas rolling mean / std, ``_expand_params``,
  ``_correct_fut_stats``) is pinned to the LIVE reference: ``tests/golden/zscore_*.npz`` hold what the reference's own
  ``ZScoreRegressor.predict`` returns (imported in the build container with a placeholder ``xarray`` module — predict
  never touches xarray).
* ``zscore_calc_stats`` (zscore.py:124-193) is built on xarray (``groupby('time.year').map``, ``concat``,
  ``rolling(...).construct``), which is not installed anywhere this repository runs: its window mapping is restated
  from the reference's code and xarray's documented semantics and is pinned ONLY by the reference's three known-answer
  tests (test_pointwise_models.py:236-299: 364 values for a leap-day-free record, scale 2, shift 1) —
  **fit parity otherwise unpinned**.
"""

from __future__ import annotations

import numpy as np
import pandas as pd


def zscore_day_table(index: pd.DatetimeIndex) -> np.ndarray:
    """zscore.py:124-158 (``_reshape``): ``groupby('time.year').map(split)`` lays the record out as
    [year, day-of-year]; the day axis is the sorted union of the days of year that occur (outer join of the yearly
    pieces), entries a year does not have are NaN.  Returns the row of the record per (year, day column), -1 = absent."""
    index = pd.DatetimeIndex(index)
    years = np.unique(index.year)
    days = np.unique(index.dayofyear)
    table = np.full((len(years), len(days)), -1, dtype=np.int64)
    table[np.searchsorted(years, index.year), np.searchsorted(days, index.dayofyear)] = np.arange(len(index))
    return table


def zscore_window_columns(n_days: int, window_width: int) -> np.ndarray:
    """Columns of the [year, day] layout that enter each retained rolling window (zscore.py:150-157, 185-190).

    ``da_rsh`` = [last ceil(w/2) day columns | all D columns | first w//2 columns] (``-window_width // 2`` is
    ``-ceil(w/2)``); a centred window of w positions is built at every position and positions
    ``[n, len - n)``, ``n = w//2 + 1``, are kept: retained window k covers ``da_rsh`` positions k+1 .. k+w.
    Returns int array [n_kept, w] of source day columns."""
    w = int(window_width)
    late = -((-w) // 2)                       # ceil(w / 2) columns taken from the end of the year
    early = w // 2
    src = np.concatenate([np.arange(n_days - late, n_days), np.arange(n_days), np.arange(early)])
    n = w // 2 + 1
    kept = len(src) - 2 * n
    if kept <= 0:
        return np.zeros((0, w), dtype=np.int64)
    first = np.arange(kept) + n - w // 2      # centre p = n + k, window p - w//2 .. p - w//2 + w - 1
    return src[first[:, None] + np.arange(w)[None, :]]


def zscore_calc_stats(values: np.ndarray, index, window_width: int):
    """zscore.py:161-193: mean and population standard deviation (ddof = 0, NaN skipped) over all years and the w
    window columns, in the input dtype (xarray reduces float32 data in float32)."""
    v = np.asarray(values).reshape(-1)
    table = zscore_day_table(index)
    grid = np.full(table.shape, np.nan, dtype=v.dtype if v.dtype.kind == 'f' else np.float64)
    grid[table >= 0] = v[table[table >= 0]]
    cols = zscore_window_columns(table.shape[1], window_width)
    mean = np.empty(len(cols), dtype=grid.dtype)
    std = np.empty(len(cols), dtype=grid.dtype)
    for k in range(len(cols)):
        block = grid[:, cols[k]]
        mean[k] = np.nanmean(block)
        std[k] = np.nanstd(block)
    return mean, std


def zscore_fit(X: np.ndarray, y: np.ndarray, index, window_width: int = 31) -> dict:
    """zscore.py:32-66, 196-239: shift = mean(y) - mean(X), scale = std(y) / std(X) per retained day."""
    x_mean, x_std = zscore_calc_stats(X, index, window_width)
    y_mean, y_std = zscore_calc_stats(y, index, window_width)
    with np.errstate(divide='ignore', invalid='ignore'):
        return {'shift': y_mean - x_mean, 'scale': y_std / x_std, 'window_width': int(window_width),
                'X_mean': x_mean, 'X_std': x_std, 'y_mean': y_mean, 'y_std': y_std}


def zscore_expand_index(n_samples: int) -> np.ndarray:
    """zscore.py:278-318: the fitted values are repeated every ``min(n, 364)`` steps from the first step on."""
    len_avgyr = min(n_samples, 364)
    return np.arange(n_samples) % len_avgyr


def zscore_predict(st: dict, X: np.ndarray) -> np.ndarray:
    """zscore.py:68-110: centred rolling mean / sample standard deviation (pandas: float64, NaN where the window is
    incomplete), z-score, corrected by the expanded shift / scale.  float64 [n]."""
    s = pd.Series(np.asarray(X).reshape(-1))
    w = st['window_width']
    fut_mean = s.rolling(w, center=True).mean()
    fut_std = s.rolling(w, center=True).std()
    z = (s - fut_mean) / fut_std
    inds = zscore_expand_index(len(s))
    shift = np.asarray(st['shift'])[inds]
    scale = np.asarray(st['scale'])[inds]
    return (z * (fut_std * scale) + (fut_mean + shift)).to_numpy(dtype=np.float64)
