"""Oracle: Cunnane empirical-CDF quantile mapping (TEST INFRASTRUCTURE — see oracle/__init__.py).

Follows skdownscale/pointwise_models/quantile.py:23-43 (plotting positions),
:81-147 (QuantileMapper, detrend=False) and :438-545 (CunnaneTransformer).
All functions work on ONE 1-D series, like the reference.
"""

from __future__ import annotations

import numpy as np


def plotting_positions(n: int, alpha: float = 0.4, beta: float = 0.4) -> np.ndarray:
    """quantile.py:23-43 — ``(arange(1, n+1) - alpha) / (n + 1.0 - alpha - beta)`` (float64)."""
    return (np.arange(1, n + 1) - alpha) / (n + 1.0 - alpha - beta)


def rank_max_ties(x: np.ndarray) -> np.ndarray:
    """1-based rank of every element among the series itself, ties take the HIGHEST rank.

    This is what ``CunnaneTransformer().fit_transform(x)`` reduces to
    (quantile.py:462,488): ``np.interp(x, np.sort(x), pp)`` lands exactly on a
    knot and numpy's binary search picks the last duplicate.
    """
    s = np.sort(x)
    return np.searchsorted(s, x, side='right').astype(np.int64)


def quantile_mapper_fit(v: np.ndarray) -> np.ndarray:
    """QuantileMapper.fit → CunnaneTransformer.fit (quantile.py:81-107, 438-463):
    the fitted state is ``np.sort(v)`` (dtype preserved); pp is implied by its length."""
    return np.sort(np.asarray(v).reshape(-1))


def _ols_line(x: np.ndarray, y: np.ndarray) -> tuple[float, float]:
    """Closed form of sklearn LinearRegression on one feature (used at quantile.py:532-543):
    centred least squares, returns (slope, intercept) in float64."""
    x = np.asarray(x, dtype=np.float64)
    y = np.asarray(y, dtype=np.float64)
    xm = x.mean()
    ym = y.mean()
    dx = x - xm
    sxx = np.dot(dx, dx)
    slope = np.dot(dx, y - ym) / sxx if sxx > 0 else 0.0     # single point: lstsq's minimum-norm answer
    return float(slope), float(ym - slope * xm)


def cunnane_inverse(q: np.ndarray, sorted_fit: np.ndarray, n_endpoints: int = 10, alpha: float = 0.4,
                    beta: float = 0.4, extrapolate='both') -> np.ndarray:
    """CunnaneTransformer.inverse_transform (quantile.py:523-545).

    ``q`` float64 quantiles, ``sorted_fit`` the fitted sorted values.  Interior:
    ``np.interp(q, pp_fit, vals_fit)``; outside ``[pp_fit[0], pp_fit[-1]]``: OLS
    line through the first / last ``n_endpoints`` (pp, val) pairs on the tails
    ``extrapolate`` names ('min', 'max', 'both'), np.interp's end-value clamp on the others
    (None and '1to1' clamp both, quantile.py:526-527).
    """
    q = np.asarray(q, dtype=np.float64)
    # quantile.py:462: ``Cdf(plotting_positions(len(X)), np.sort(X))`` — the constructor's alpha / beta
    # are never handed to plotting_positions, so the reference always uses 0.4 / 0.4 (replicated)
    del alpha, beta
    pp = plotting_positions(len(sorted_fit))
    left = -np.inf if extrapolate in ('min', 'both') else None
    right = np.inf if extrapolate in ('max', 'both') else None
    vals = np.interp(q, pp, sorted_fit, left=left, right=right)
    if np.isinf(vals).any():
        lower = np.nonzero(-np.inf == vals)[0]
        upper = np.nonzero(np.inf == vals)[0]
        if len(lower):
            s = slice(None, n_endpoints)
            a, b = _ols_line(pp[s], sorted_fit[s])
            vals[lower] = a * q[lower] + b
        if len(upper):
            s = slice(-n_endpoints, None)
            a, b = _ols_line(pp[s], sorted_fit[s])
            vals[upper] = a * q[upper] + b
    return vals


def quantile_mapper_transform(x: np.ndarray, sorted_fit: np.ndarray, n_endpoints: int = 10,
                              return_rank: bool = False, alpha: float = 0.4, beta: float = 0.4,
                              extrapolate='both'):
    """QuantileMapper.transform (quantile.py:109-147, detrend=False).

    ``x`` is ranked against ITSELF (quantile.py:138), the resulting quantiles are
    pushed through the fitted inverse CDF (quantile.py:139).  Returns float64.
    The keyword arguments are the ``qt_kwargs`` of the reference (CunnaneTransformer settings).
    """
    x = np.asarray(x).reshape(-1)
    r = rank_max_ties(x)
    q = plotting_positions(len(x))[r - 1]                # alpha / beta are ignored by the reference (quantile.py:462)
    out = cunnane_inverse(q, sorted_fit, n_endpoints=n_endpoints, alpha=alpha, beta=beta, extrapolate=extrapolate)
    return (out, r) if return_rank else out
