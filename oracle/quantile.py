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


def linear_trend_fit(v: np.ndarray) -> tuple[float, float]:
    """LinearTrendTransformer.fit (trend.py:40-52): sklearn LinearRegression of the series on
    ``arange(n)`` — centred least squares in float64.  Returns (slope, intercept)."""
    y = np.asarray(v, dtype=np.float64).reshape(-1)
    t = np.arange(len(y), dtype=np.float64)
    tm, ym = t.mean(), y.mean()
    dt = t - tm
    stt = np.dot(dt, dt)
    slope = float(np.dot(dt, y - ym) / stt) if stt > 0 else 0.0
    return slope, float(ym - tm * slope)


def quantile_mapper_fit_detrend(v: np.ndarray) -> dict:
    """QuantileMapper(detrend=True).fit (quantile.py:94-105): remove the series' own linear trend
    (float64: ``X - trendline``, trend.py:54-64,79-83), keep the sorted residuals and the fitted
    intercept (needed by transform, quantile.py:145)."""
    v = np.asarray(v).reshape(-1)
    slope, icpt = linear_trend_fit(v)
    resid = v - (np.arange(len(v)) * slope + icpt)
    return {'sorted': np.sort(resid), 'intercept': icpt}


def quantile_mapper_transform_detrend(x: np.ndarray, st: dict, return_rank: bool = False, **qt):
    """QuantileMapper(detrend=True).transform (quantile.py:127-147): detrend the NEW data with its own
    trend, map the residuals, add that trend back and move the baseline to the fitted intercept."""
    x = np.asarray(x).reshape(-1)
    slope, icpt = linear_trend_fit(x)
    trend = np.arange(len(x)) * slope + icpt
    mapped, r = quantile_mapper_transform(x - trend, st['sorted'], return_rank=True, **qt)
    out = mapped + trend                                         # quantile.py:143  (inverse_transform)
    out -= icpt - st['intercept']                                # quantile.py:145
    return (out, r) if return_rank else out


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


# ---------------------------------------------------------------------------------------------
# QuantileMappingReressor / EquidistantCdfMatcher (quantile.py:160-395, 556-636) — SURVEY §8(f) row 1
# ---------------------------------------------------------------------------------------------
SYNTHETIC_PP_MIN = -1e20          # quantile.py:17-18
SYNTHETIC_PP_MAX = 1e20


def extended_cdf(sorted_data: np.ndarray, extrapolate=None, n_endpoints: int = 10):
    """``_calc_extrapolated_cdf`` (quantile.py:311-388) for already sorted data.

    Returns (pp, vals), both float64 of length n + 2: the Cunnane positions / sorted values framed
    by one synthetic point on either side.  On a tail that extrapolates ('min' → lower, 'max' →
    upper, 'both') the synthetic position is -1e20 / +1e20 and its value the OLS line through the
    ``n_endpoints`` nearest (pp, value) pairs evaluated there; otherwise the end point is repeated."""
    if extrapolate not in (None, '1to1', 'both', 'max', 'min'):
        raise ValueError(f'unknown value for extrapolate: {extrapolate}')
    d = np.asarray(sorted_data, dtype=np.float64).reshape(-1)
    n = len(d)
    pp = np.empty(n + 2)
    pp[1:-1] = plotting_positions(n)
    vals = np.empty(n + 2)
    vals[1:-1] = d
    vals[0], vals[-1] = d[0], d[-1]
    lower = extrapolate in ('min', 'both')
    upper = extrapolate in ('max', 'both')
    pp[0] = SYNTHETIC_PP_MIN if lower else pp[1]
    pp[-1] = SYNTHETIC_PP_MAX if upper else pp[-2]
    if lower:
        a, b = _ols_line(pp[1:n_endpoints + 1], vals[1:n_endpoints + 1])
        vals[0] = a * pp[0] + b
    if upper:
        a, b = _ols_line(pp[-n_endpoints - 1:-1], vals[-n_endpoints - 1:-1])
        vals[-1] = a * pp[-1] + b
    return pp, vals


def qm_regressor_fit(X: np.ndarray, y: np.ndarray, extrapolate=None, n_endpoints: int = 10) -> dict:
    """QuantileMappingReressor.fit (quantile.py:193-222): the two framed CDFs."""
    if n_endpoints < 2:
        raise ValueError('Invalid number of n_endpoints, must be >= 2')
    xs = np.sort(np.asarray(X).reshape(-1))
    ys = np.sort(np.asarray(y).reshape(-1))
    xpp, xv = extended_cdf(xs, extrapolate, n_endpoints)
    ypp, yv = extended_cdf(ys, extrapolate, n_endpoints)
    return {'x_pp': xpp, 'x_vals': xv, 'y_pp': ypp, 'y_vals': yv, 'extrapolate': extrapolate,
            'n_endpoints': n_endpoints}


def _one_to_one_tails(st: dict, X: np.ndarray, y_hat: np.ndarray) -> np.ndarray:
    """``_extrapolate_1to1`` (quantile.py:268-309): values outside the fitted X range keep their
    distance to the range end.  X and y were fitted on the same number of samples in every
    reachable call (fit(X, y)), the unequal-length branches are restated for completeness."""
    xv, yv, xpp, ypp = st['x_vals'], st['y_vals'], st['x_pp'], st['y_pp']
    nx, ny = len(xv), len(yv)
    x_min, x_max, y_min, y_max = xv[0], xv[-1], yv[0], yv[-1]
    over = X > x_max
    if over.any():
        if nx == ny:
            y_hat[over] = y_max + (X[over] - x_max)
        elif nx > ny:
            y_hat[over] = y_max + (X[over] - np.interp(ypp[-1], xpp, xv))
        else:
            y_hat[over] = np.interp(xpp[-1], ypp, yv) + (X[over] - x_max)
    under = X < x_min
    if under.any():
        if nx == ny:
            y_hat[under] = y_min + (X[under] - x_min)
        elif nx > ny:
            y_hat[under] = x_min + (X[under] - np.interp(ypp[0], xpp, xv))       # sic: quantile.py:304
        else:
            y_hat[under] = np.interp(xpp[0], ypp, yv) + (X[under] - x_min)
    return y_hat


def qm_regressor_predict(st: dict, X: np.ndarray) -> np.ndarray:
    """QuantileMappingReressor.predict (quantile.py:224-266).  Result in X's dtype (``full_like``)."""
    X = np.asarray(X).reshape(-1)
    ex, ne = st['extrapolate'], st['n_endpoints']
    order = np.argsort(X)
    pp, vals = extended_cdf(X[order], ex, ne)
    left = -np.inf if ex in ('min', 'both') else None
    right = np.inf if ex in ('max', 'both') else None
    pp[:] = np.interp(vals, st['x_vals'], st['x_pp'], left=left, right=right)        # quantile.py:246-248
    if np.isinf(pp).any():                                                             # quantile.py:252-263
        lower = np.nonzero(-np.inf == pp)[0]
        upper = np.nonzero(np.inf == pp)[0]
        if len(lower):
            s = slice(lower[-1] + 1, lower[-1] + 1 + ne)
            a, b = _ols_line(pp[s], vals[s])
            pp[lower] = a * vals[lower] + b                  # sic: the model maps pp → value, fed with values
        if len(upper):
            s = slice(upper[0] - ne, upper[0])
            a, b = _ols_line(pp[s], vals[s])
            pp[upper] = a * vals[upper] + b
    y_hat = np.full_like(X, np.nan)
    y_hat[order] = np.interp(pp, st['y_pp'], st['y_vals'])[1:-1]                      # quantile.py:266
    if ex == '1to1':
        y_hat = _one_to_one_tails(st, X, y_hat)
    return y_hat


def edcdf_predict(st: dict, X: np.ndarray, kind: str = 'difference') -> np.ndarray:
    """EquidistantCdfMatcher.predict (quantile.py:594-636), ``max_ratio=None`` (with a value the
    reference calls ``np.min(ratio, max_ratio)`` — an axis argument — and raises, quantile.py:624)."""
    if kind not in ('difference', 'ratio'):
        raise NotImplementedError('kind must be either difference or ratio')
    X = np.asarray(X).reshape(-1)
    ex, ne = st['extrapolate'], st['n_endpoints']
    order = np.argsort(X)
    pp, vals = extended_cdf(X[order], ex, ne)
    x_train = np.interp(pp, st['x_pp'], st['x_vals'])                                  # quantile.py:612-613
    y_at_pp = np.interp(pp, st['y_pp'], st['y_vals'])
    sorted_hat = y_at_pp + (vals - x_train) if kind == 'difference' else y_at_pp * (vals / x_train)
    y_hat = np.full_like(X, np.nan)
    y_hat[order] = sorted_hat[1:-1]
    if ex == '1to1':
        y_hat = _one_to_one_tails(st, X, y_hat)
    return y_hat


def qmr_well_conditioned(st: dict, X: np.ndarray, estimator: str) -> np.ndarray:
    """Mask of the prediction steps whose reference result is numerically meaningful.

    On a tail that extrapolates ('min' / 'max' / 'both') the reference interpolates through a
    synthetic CDF point at pp = -+1e20 whose value is ~1e21 (quantile.py:17-18, 375-386):
    ``slope * (x - xp[j]) + fp[j]`` there subtracts two ~1e21 numbers, so the result depends on the
    last bit of sklearn's LinearRegression output and carries no significant digits.  Those steps
    are the ones that reach the synthetic segment: for the regressor, values outside the fitted X
    range on an extrapolating side; for EDCDFm, plotting positions outside the fitted ones (and, for
    EDCDFm, steps whose value is exactly tied with another one: their order is np.argsort's choice)."""
    X = np.asarray(X).reshape(-1)
    ex = st['extrapolate']
    ok = np.ones(len(X), dtype=bool)
    lower, upper = ex in ('min', 'both'), ex in ('max', 'both')
    if estimator == 'regressor':
        if lower:
            ok &= X >= st['x_vals'][1]
        if upper:
            ok &= X <= st['x_vals'][-2]
    else:
        # exact ties: which of the tied steps takes which plotting position is decided by the internals
        # of np.argsort's unstable default sort (quantile.py:607) — not defined by the algorithm
        _, inv, cnt = np.unique(X, return_inverse=True, return_counts=True)
        ok &= cnt[inv] == 1
        q = plotting_positions(len(X))[np.argsort(np.argsort(X, kind='stable'), kind='stable')]
        if lower:
            ok &= q >= st['x_pp'][1]
        if upper:
            ok &= q <= st['x_pp'][-2]
    return ok


def trend_aware_qm_fit_predict(X_train, y_train, X_pred, extrapolate=None, n_endpoints: int = 10) -> np.ndarray:
    """TrendAwareQuantileMappingRegressor(QuantileMappingReressor(extrapolate, n_endpoints)) for ONE cell
    (quantile.py:639-716; the default LinearTrendTransformer — passing another one leaves the attribute
    unset in the reference, quantile.py:655-656).  NOT yet in the product: pinned here for the next round.

    fit: remove the linear trends of X and y (each its own), fit the CDF-to-CDF regressor on the
    residuals; predict: remove the new X's trend, map the residuals, add the new trend line centred at
    zero plus ``(mean(X_new) - mean(X_fit)) + mean(y_fit)`` (column means summed in the input dtype, like
    DataFrame.mean).  Returns float64 [n, 1] like the reference."""
    Xf = np.asarray(X_train).reshape(-1)
    yf = np.asarray(y_train).reshape(-1)
    Xp = np.asarray(X_pred).reshape(-1)

    def detrended(v):
        slope, icpt = linear_trend_fit(v)
        line = np.arange(len(v)) * slope + icpt
        return v - line, line

    x_res, _ = detrended(Xf)
    y_res, _ = detrended(yf)
    st = qm_regressor_fit(x_res, y_res, extrapolate, n_endpoints)
    xp_res, line = detrended(Xp)
    y_hat = qm_regressor_predict(st, xp_res).reshape(-1, 1).astype(np.float64)
    # DataFrame.mean() of a one-column float32 frame: pandas' nanmean sums in the frame's own dtype (numpy
    # pairwise summation) and divides by the count — the same value as ndarray.mean() in that dtype
    col_mean = lambda a: a.sum(dtype=a.dtype) / a.dtype.type(len(a))     # noqa: E731
    delta = (col_mean(Xp) - col_mean(Xf)) + col_mean(yf)
    centred = line - line.mean()
    return y_hat + (centred.reshape(-1, 1) + delta)
