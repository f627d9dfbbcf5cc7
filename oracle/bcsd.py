"""Oracle: BCSD temperature / precipitation for ONE cell (TEST INFRASTRUCTURE — see oracle/__init__.py).

Follows skdownscale/pointwise_models/bcsd.py:115-185 (BcsdPrecipitation) and
:197-281 (BcsdTemperature).  Group structures are passed in explicitly as
lists of ``(key, rows)`` (see oracle/groupers.py) so the same functions cover
the monthly mode (bcsd.py:48-49) and the 'daily_nasa-nex' mode
(bcsd.py:36-38,50-55), whose predict side is keyed by day-of-month
(bcsd.py:53,250,275).
"""

from __future__ import annotations

import numpy as np

from .quantile import (quantile_mapper_fit, quantile_mapper_fit_detrend, quantile_mapper_transform,
                       quantile_mapper_transform_detrend)


def _group_mean_like_pandas(v: np.ndarray, how: str = 'groupby') -> np.generic:
    """Mean of one group in the arithmetic pandas (a third-party dependency of the
    reference; 2.3.3 pinned in uv.lock:2181-2182, 3.0.2 installed) actually uses.

    * ``how='groupby'`` — ``df.groupby(...).mean()`` (bcsd.py:138,222-223): the cython
      ``group_mean`` kernel: Kahan-compensated running sum IN THE INPUT DTYPE, rows in
      time order, then one division by the count in that dtype.  (float32 input ⇒
      float32 arithmetic throughout — verified against the live reference.)
    * ``how='frame'`` — ``DataFrame.mean()`` used by PaddedDOYGrouper.mean()
      (groupers.py:84-89): ``nanops.nanmean`` = numpy's pairwise ``sum`` in the input
      dtype divided by the count in that dtype.
    """
    v = np.asarray(v)
    ft = v.dtype.type if v.dtype.kind == 'f' else np.float64
    if how == 'frame':
        return ft(v.sum(dtype=ft) / ft(len(v)))
    s = ft(0)
    comp = ft(0)
    with np.errstate(all='ignore'):
        for x in v.astype(ft, copy=False):
            y = ft(x - comp)
            t = ft(s + y)
            comp = ft(ft(t - s) - y)
            if comp != comp:
                comp = ft(0)
            s = t
    return ft(s / ft(len(v)))


def _map_group(x, fitted, qt):
    """One group through its fitted QuantileMapper (bcsd.py:69-79): plain, or — when the group was
    fitted with detrend=True (a dict) — the detrending transform."""
    if isinstance(fitted, dict):
        return quantile_mapper_transform_detrend(x, fitted, return_rank=True, **(qt or {}))
    return quantile_mapper_transform(x, fitted, return_rank=True, **(qt or {}))


def rolling9_centered(x: np.ndarray) -> np.ndarray:
    """``x.rolling(9, center=True, min_periods=1).mean()`` (bcsd.py:247-248) on one
    time-ordered group subsequence; float64 result."""
    x = np.asarray(x, dtype=np.float64)
    n = len(x)
    tot = np.zeros(n, dtype=np.float64)
    cnt = np.zeros(n, dtype=np.float64)
    for d in range(-4, 5):
        lo = max(0, -d)
        hi = min(n, n - d)
        if hi > lo:
            tot[lo:hi] += x[lo + d:hi + d]
            cnt[lo:hi] += 1.0
    return tot / cnt


def bcsd_temperature_fit(X: np.ndarray, y: np.ndarray, fit_groups, mean_how: str = 'groupby',
                         detrend: bool = False) -> dict:
    """BcsdTemperature.fit (bcsd.py:197-228): per group x/y climatology + sorted y."""
    X = np.asarray(X).reshape(-1)
    y = np.asarray(y).reshape(-1)
    st = {'x_climo': {}, 'y_climo': {}, 'sorted': {}}
    for key, rows in fit_groups:
        st['x_climo'][key] = _group_mean_like_pandas(X[rows], mean_how)       # bcsd.py:222
        st['y_climo'][key] = _group_mean_like_pandas(y[rows], mean_how)       # bcsd.py:223
        # bcsd.py:226 → quantile.py:462 (qm_kwargs detrend=True: the mapper of each group detrends the
        # group's own subsequence, positions 0..len-1, quantile.py:94-98)
        st['sorted'][key] = quantile_mapper_fit_detrend(y[rows]) if detrend else quantile_mapper_fit(y[rows])
    return st


def bcsd_temperature_predict(st: dict, X: np.ndarray, roll_groups, qm_groups,
                             return_anoms: bool = True, return_rank: bool = False, qt: dict | None = None):
    """BcsdTemperature.predict (bcsd.py:230-281).

    ``roll_groups``: groups of ``climate_trend`` (bcsd.py:250); ``qm_groups``: groups
    used for climatology removal AND quantile mapping (bcsd.py:253,260 — the
    time_grouper groups in monthly mode, day-of-month groups in daily mode).
    Returns float64 [T] (and the per-group 1-based ranks when asked).
    """
    X = np.asarray(X).reshape(-1)
    T = len(X)
    roll = np.empty(T, dtype=np.float64)
    for _, rows in roll_groups:                                      # bcsd.py:247-250
        roll[rows] = rolling9_centered(X[rows])
    shift = np.empty(T, dtype=np.float64)
    for key, rows in qm_groups:                                      # bcsd.py:253, 271-281
        shift[rows] = roll[rows] - np.float64(st['x_climo'][key])
    no_shift = X.astype(np.float64) - shift                          # bcsd.py:256
    xqm = np.empty(T, dtype=np.float64)
    ranks = np.empty(T, dtype=np.int64)
    for key, rows in qm_groups:                                      # bcsd.py:260, 69-79
        xqm[rows], ranks[rows] = _map_group(no_shift[rows], st['sorted'][key], qt)
    out = shift + xqm                                                # bcsd.py:263
    if return_anoms:                                                 # bcsd.py:266-267
        for key, rows in qm_groups:
            out[rows] = out[rows] - np.float64(st['y_climo'][key])
    return (out, ranks) if return_rank else out


def bcsd_precipitation_fit(y: np.ndarray, fit_groups, return_anoms: bool = True,
                           mean_how: str = 'groupby', detrend: bool = False) -> dict:
    """BcsdPrecipitation.fit (bcsd.py:115-147).  X_train is validated only, never used."""
    y = np.asarray(y).reshape(-1)
    st = {'y_climo': {}, 'sorted': {}}
    for key, rows in fit_groups:
        st['y_climo'][key] = _group_mean_like_pandas(y[rows], mean_how)       # bcsd.py:138
    if return_anoms and min(float(v) for v in st['y_climo'].values()) <= 0:   # bcsd.py:140-141
        raise ValueError('Invalid value in target climatology')
    for key, rows in fit_groups:
        st['sorted'][key] = quantile_mapper_fit_detrend(y[rows]) if detrend else quantile_mapper_fit(y[rows])  # bcsd.py:145
    return st


def bcsd_precipitation_predict(st: dict, X: np.ndarray, qm_groups, return_anoms: bool = True,
                               return_rank: bool = False, qt: dict | None = None):
    """BcsdPrecipitation.predict (bcsd.py:149-185): QM per group, then ratio anomalies."""
    X = np.asarray(X).reshape(-1)
    out = np.empty(len(X), dtype=np.float64)
    ranks = np.empty(len(X), dtype=np.int64)
    for key, rows in qm_groups:                                      # bcsd.py:167
        out[rows], ranks[rows] = _map_group(X[rows], st['sorted'][key], qt)
    if return_anoms:                                                 # bcsd.py:170-185
        for key, rows in qm_groups:
            out[rows] = out[rows] / np.float64(st['y_climo'][key])
    return (out, ranks) if return_rank else out
