"""QuantileMapper — drop-in for skdownscale.pointwise_models.QuantileMapper
(skdownscale/pointwise_models/quantile.py:46-157), executed on the GPU.

``qt_kwargs`` (alpha, beta, extrapolate, n_endpoints of the reference's CunnaneTransformer,
quantile.py:420-432) travel to the kernels as ``sdb_cunnane_opts``.
"""

from __future__ import annotations

import copy

import numpy as np
import torch
from sklearn.base import BaseEstimator, RegressorMixin, TransformerMixin
from sklearn.exceptions import NotFittedError

from .. import _lib, engine
from .base import cuda_device, series_to_device
from .utils import default_none_kwargs

_QT_DEFAULTS = {'alpha': 0.4, 'beta': 0.4, 'extrapolate': 'both', 'n_endpoints': 10}


def cunnane_opts(qt_kwargs):
    """CunnaneTransformer(**qt_kwargs) (quantile.py:420-432) as the C-ABI option block, or None for
    the defaults.  Unknown keywords raise like the reference's constructor would."""
    kw = dict(_QT_DEFAULTS)
    for k, v in default_none_kwargs(qt_kwargs).items():
        if k not in _QT_DEFAULTS:
            raise TypeError(f"CunnaneTransformer.__init__() got an unexpected keyword argument '{k}'")
        kw[k] = v
    if kw == _QT_DEFAULTS:
        return None
    if kw['extrapolate'] not in _lib.EXTRAPOLATE:
        raise ValueError(f"unknown value for extrapolate: {kw['extrapolate']}")
    if int(kw['n_endpoints']) < 1:
        raise ValueError('n_endpoints must be >= 1')
    # quantile.py:462 builds the CDF with plotting_positions(len(X)) — the transformer's own alpha / beta
    # never reach it, so the reference always maps with 0.4 / 0.4; accepted and (like there) without effect
    float(kw['alpha']), float(kw['beta'])
    return _lib.CunnaneOpts(0.4, 0.4, int(kw['n_endpoints']), _lib.EXTRAPOLATE[kw['extrapolate']])


def check_qt_kwargs(qt_kwargs):
    cunnane_opts(qt_kwargs)


def check_lt_kwargs(lt_kwargs):
    """lt_kwargs → LinearTrendTransformer(**lt_kwargs) (quantile.py:96): only its own ``lr_kwargs``
    argument exists (trend.py:31-32), and only settings that keep the plain least-squares line."""
    for k, v in default_none_kwargs(lt_kwargs).items():
        if k != 'lr_kwargs':
            raise TypeError(f"LinearTrendTransformer.__init__() got an unexpected keyword argument '{k}'")
        for kk, vv in default_none_kwargs(v).items():
            if not ((kk == 'fit_intercept' and vv) or kk in ('copy_X', 'n_jobs', 'tol') or (kk == 'positive' and not vv)):
                raise NotImplementedError(f'lr_kwargs {kk}={vv!r} is not supported on the B200 path')


def fit_detrended(v: torch.Tensor, table, valid):
    """Fit of detrending mappers for every (cell, group) of ``table`` (quantile.py:94-105): returns the
    fitted state of the float64 residuals and the trend intercepts ``[group, C]`` (quantile.py:145 needs them)."""
    flag = torch.zeros(1, dtype=torch.int32, device=v.device)
    slope, icpt = engine.group_trend(v, table, valid, flag)
    resid = engine.trend_apply(_lib.TREND_REMOVE, v, table, slope, icpt, valid=valid)
    st = engine.qm_fit(resid, table, valid=valid, want_y_climo=False)
    st.extra['raw_dtype'] = v.dtype
    if int(flag.item()) != 0:
        st.nonfinite.fill_(1)
    return st, icpt


def whole_series_table(n_rows: int) -> engine.GroupTable:
    return engine.GroupTable([(0, np.arange(n_rows))])


class QuantileMapper(TransformerMixin, BaseEstimator):
    """Transform features using quantile mapping (quantile.py:46-157)."""

    _fit_attributes = ['x_cdf_fit_']

    def __init__(self, detrend=False, lt_kwargs=None, qt_kwargs=None):
        self.detrend = detrend
        self.lt_kwargs = lt_kwargs
        self.qt_kwargs = qt_kwargs

    # ---- batched (all cells) API used by PointWiseDownscaler
    def fit_batched(self, X: torch.Tensor, valid=None):
        check_qt_kwargs(self.qt_kwargs)
        check_lt_kwargs(self.lt_kwargs)
        table = whole_series_table(X.shape[0])
        if self.detrend:
            # quantile.py:94-98: the CDF is fitted on X minus its own linear trend (float64 residuals)
            self._state, self._icpt_fit = fit_detrended(X, table, valid)
        else:
            self._state = engine.qm_fit(X, table, valid=valid, want_y_climo=False)
        return self

    def transform_batched(self, X: torch.Tensor, out_dtype=None, want_rank=False):
        if not hasattr(self, '_state'):
            raise NotFittedError(f"This {type(self).__name__} instance is not fitted yet.")
        fit_dtype = self._state.extra.get('raw_dtype', self._state.dtype)
        if X.dtype != fit_dtype:
            X = X.to(fit_dtype)                    # the reference accepts a transform dtype that differs from fit's
        if self.detrend:
            if want_rank:
                raise NotImplementedError('rank instrumentation is not available with detrend=True')
            return engine.qm_predict_detrended(self._state, self._state, self._icpt_fit, X, whole_series_table(X.shape[0]),
                                               _lib.MODE_QM, out_dtype=out_dtype, cunnane=cunnane_opts(self.qt_kwargs))
        return engine.qm_predict(self._state, X, whole_series_table(X.shape[0]), _lib.MODE_QM,
                                 out_dtype=out_dtype, want_rank=want_rank, cunnane=cunnane_opts(self.qt_kwargs))

    # ---- per-cell API of the reference (one series)
    def fit(self, X, y=None):
        x, _, _ = series_to_device(X, cuda_device())
        if x.shape[1] != 1:
            raise ValueError('CunnaneTransformer.fit() only supports a single feature')
        self.fit_batched(x)
        self._state.check_finite()
        self.n_features_in_ = 1
        self.x_cdf_fit_ = True
        return self

    def transform(self, X):
        x, _, _ = series_to_device(X, cuda_device())
        if x.shape[1] != 1:
            raise ValueError('CunnaneTransformer.transform() only supports a single feature')
        if not hasattr(self, '_state'):
            raise NotFittedError(f"This {type(self).__name__} instance is not fitted yet.")
        if x.dtype != self._state.dtype:
            x = x.to(self._state.dtype)
        out = self.transform_batched(x, out_dtype=torch.float64)
        self._state.check_finite()
        return out.cpu().numpy().reshape(-1, 1)


class QuantileMappingReressor(RegressorMixin, BaseEstimator):
    """Transform features using quantile mapping — CDF of X onto the CDF of y
    (quantile.py:160-395; the class name keeps the reference's spelling).

    ``extrapolate`` in {None, 'min', 'max', 'both', '1to1'}, ``n_endpoints >= 2``.  Note for
    'min' / 'max' / 'both': outside the fitted X range the reference interpolates through a
    synthetic CDF point at pp = -1e20 / +1e20 (quantile.py:17-18), a cancellation of ~1e21-sized
    numbers — those outputs are computed the same way here but carry no significant digits in
    either implementation."""

    _fit_attributes = ['_X_cdf', '_y_cdf']
    _kind = _lib.QMR_REGRESSOR

    def __init__(self, extrapolate=None, n_endpoints=10):
        self.extrapolate = extrapolate
        self.n_endpoints = n_endpoints
        if self.n_endpoints < 2:
            raise ValueError('Invalid number of n_endpoints, must be >= 2')

    # ---- batched (all cells) API used by PointWiseDownscaler
    def fit_batched(self, X: torch.Tensor, y: torch.Tensor, valid=None):
        if self.extrapolate not in _lib.EXTRAPOLATE:
            raise ValueError(f'unknown value for extrapolate: {self.extrapolate}')
        n = X.shape[0]
        need = 2 * self.n_endpoints + 1                               # check_array(ensure_min_samples=...)  quantile.py:205-210
        if n < need or y.shape[0] < need:
            raise ValueError(f'Found array with {min(n, y.shape[0])} sample(s) while a minimum of {need} is required.')
        if y.shape != X.shape:
            raise ValueError(f'X {tuple(X.shape)} and y {tuple(y.shape)} must have the same shape')
        table = whole_series_table(n)
        self._sx = engine.qm_fit(X, table, valid=valid, want_y_climo=False)
        self._sy = engine.qm_fit(y, table, valid=valid, want_y_climo=False)
        self._sy.nonfinite = self._sx.nonfinite                       # one flag for the model
        self._frame = engine.qmr_frame(self._sx, self._sy, _lib.EXTRAPOLATE[self.extrapolate], int(self.n_endpoints))
        return self

    def check_fit(self):
        self._sx.check_finite()

    def _ranks(self, X):
        return None

    def predict_batched(self, X: torch.Tensor):
        if not hasattr(self, '_sx'):
            raise NotFittedError(f"This {type(self).__name__} instance is not fitted yet. Call 'fit' with "
                                 "appropriate arguments before using this estimator.")
        return engine.qmr_predict(self._kind, X, self._sx, self._sy, self._frame, _lib.EXTRAPOLATE[self.extrapolate],
                                  self.extrapolate == '1to1', rank=self._ranks(X))

    # ---- per-cell API of the reference (one series)
    def fit(self, X, y, **kwargs):
        dev = cuda_device()
        x, _, _ = series_to_device(X, dev)
        yt, _, _ = series_to_device(y, dev)
        if x.shape[1] != 1:
            raise ValueError(f'X should have up to 1 features, found {x.shape[1]}')     # utils.check_max_features
        if yt.shape[1] != 1:
            raise ValueError('y must be 1-dimensional')
        if x.dtype != yt.dtype:
            x, yt = x.to(torch.float64), yt.to(torch.float64)
        self.fit_batched(x, yt)
        self.check_fit()
        self._X_cdf = self._y_cdf = True
        return self

    def predict(self, X, **kwargs):
        x, _, _ = series_to_device(X, cuda_device())
        if not hasattr(self, '_sx'):
            raise NotFittedError(f"This {type(self).__name__} instance is not fitted yet. Call 'fit' with "
                                 "appropriate arguments before using this estimator.")
        in_dtype = x.dtype
        if x.dtype != self._sx.dtype:
            x = x.to(self._sx.dtype)
        out = self.predict_batched(x[:, :1].contiguous())
        self._sx.check_finite()
        return out[:, 0].to(in_dtype).cpu().numpy()                   # y_hat = np.full_like(X)  quantile.py:265


class EquidistantCdfMatcher(QuantileMappingReressor):
    """Equidistant CDF matching (quantile.py:556-636): quantile mapping that preserves the difference
    or the ratio between the new and the training X at equal plotting positions.  ``max_ratio`` other
    than None makes the reference raise (``np.min(ratio, max_ratio)``, quantile.py:624) and is rejected."""

    def __init__(self, kind='difference', extrapolate=None, n_endpoints=10, max_ratio=None):
        if kind not in ['difference', 'ratio']:
            raise NotImplementedError('kind must be either difference or ratio')
        self.kind = kind
        self.extrapolate = extrapolate
        self.n_endpoints = n_endpoints
        self.max_ratio = max_ratio
        if self.n_endpoints < 2:
            raise ValueError('Invalid number of n_endpoints, must be >= 2')

    @property
    def _kind(self):
        return _lib.QMR_EDCDF_DIFFERENCE if self.kind == 'difference' else _lib.QMR_EDCDF_RATIO

    def _ranks(self, X):
        if self.kind == 'ratio' and self.max_ratio is not None:
            raise TypeError("'float' object cannot be interpreted as an integer")   # np.min(ratio, max_ratio)
        # np.argsort position of every step inside its cell's series (quantile.py:607-609)
        return engine.series_rank(X, whole_series_table(X.shape[0]), ordinal=True, valid=self._sx.valid,
                                  nonfinite=self._sx.nonfinite)



class LinearTrendTransformer(TransformerMixin, BaseEstimator):
    """Transform features by removing linear trends (trend.py:14-91): the least-squares line of every series on
    ``arange(n)`` (``sdb_group_trend``), subtracted / added back by ``sdb_trend_apply``.  ``transform`` /
    ``inverse_transform`` / ``trendline`` evaluate the FITTED line at positions ``0 .. len(X) - 1`` of the array
    they are given, like the reference.  Results are float64 (the reference's ``X - lr_model_.predict(...)``)."""

    def __init__(self, lr_kwargs=None):
        self.lr_kwargs = lr_kwargs

    # ---- batched (all cells): X [T, C] CUDA tensor
    def fit_batched(self, X: torch.Tensor, valid=None):
        check_lt_kwargs({'lr_kwargs': self.lr_kwargs})
        self._flag = torch.zeros(1, dtype=torch.int32, device=X.device)
        self._slope, self._icpt = engine.group_trend(X, whole_series_table(X.shape[0]), valid, self._flag)
        self._valid = valid
        return self

    def _require_fit(self):
        if not hasattr(self, '_slope'):
            raise NotFittedError(f"This {type(self).__name__} instance is not fitted yet. Call 'fit' with "
                                 "appropriate arguments before using this estimator.")

    def transform_batched(self, X: torch.Tensor) -> torch.Tensor:
        self._require_fit()
        return engine.trend_apply(_lib.TREND_REMOVE, X, whole_series_table(X.shape[0]), self._slope, self._icpt, valid=self._valid)

    def inverse_transform_batched(self, X: torch.Tensor) -> torch.Tensor:
        self._require_fit()
        # RESTORE = (v + line) - (intercept - intercept_ref): intercept_ref = intercept leaves v + line
        return engine.trend_apply(_lib.TREND_RESTORE, X, whole_series_table(X.shape[0]), self._slope, self._icpt, self._icpt,
                                  valid=self._valid)

    def trendline_batched(self, n: int) -> torch.Tensor:
        self._require_fit()
        zeros = torch.zeros((n, self._slope.shape[1]), dtype=torch.float64, device=self._slope.device)
        return self.inverse_transform_batched(zeros)

    def check_fit(self):
        if int(self._flag.item()) != 0:
            self._flag.zero_()
            raise ValueError('Input contains NaN or infinity.')

    # ---- per-cell API of the reference: X [n, n_features] array-like
    def fit(self, X, y=None):
        x, _, _ = series_to_device(X, cuda_device())
        self.fit_batched(x)               # every column is a series of its own (multi-output LinearRegression)
        self.check_fit()
        self.lr_model_ = True
        self.n_features_in_ = x.shape[1]
        return self

    def _columns(self, X):
        x, _, _ = series_to_device(X, cuda_device())
        self._require_fit()
        if x.shape[1] != self._slope.shape[1]:
            raise ValueError(f'X has {x.shape[1]} features, but LinearTrendTransformer is expecting {self._slope.shape[1]} features as input.')
        return x

    def transform(self, X):
        return self.transform_batched(self._columns(X)).cpu().numpy()

    def inverse_transform(self, X):
        return self.inverse_transform_batched(self._columns(X)).cpu().numpy()

    def trendline(self, X):
        return self.trendline_batched(self._columns(X).shape[0]).cpu().numpy()


class TrendAwareQuantileMappingRegressor(RegressorMixin, BaseEstimator):
    """Experimental meta estimator for trend-aware quantile mapping (quantile.py:639-716): the CDF-to-CDF regressor
    is fitted on linearly detrended X and y; predict maps the detrended new X and adds back the new trend line
    (centred at zero) plus ``(mean(X_new) - mean(X_fit)) + mean(y_fit)``.  Like the reference, only the default
    ``LinearTrendTransformer()`` exists (passing another one leaves the attribute unset there, quantile.py:655-656)."""

    def __init__(self, qm_estimator=None, trend_transformer=None):
        self.qm_estimator = qm_estimator
        if trend_transformer is None:
            self.trend_transformer = LinearTrendTransformer()

    def _check(self):
        if not isinstance(self.qm_estimator, QuantileMappingReressor):
            raise TypeError('qm_estimator must be a QuantileMappingReressor / EquidistantCdfMatcher of this package; '
                            'there is no per-cell Python fallback on the B200 path')
        if not hasattr(self, 'trend_transformer'):
            raise AttributeError("'TrendAwareQuantileMappingRegressor' object has no attribute 'trend_transformer'")

    # ---- batched (all cells): X, y [T, C] CUDA tensors
    def fit_batched(self, X: torch.Tensor, y: torch.Tensor, valid=None):
        self._check()
        table = whole_series_table(X.shape[0])
        flag = torch.zeros(1, dtype=torch.int32, device=X.device)
        self._x_mean_fit = engine.group_mean(X, table, _lib.MEAN_NUMPY, valid, flag)            # [1, C], input dtype
        self._y_mean_fit = engine.group_mean(y, whole_series_table(y.shape[0]), _lib.MEAN_NUMPY, valid, flag)
        x_res = copy.deepcopy(self.trend_transformer).fit_batched(X, valid).transform_batched(X)
        y_res = copy.deepcopy(self.trend_transformer).fit_batched(y, valid).transform_batched(y)
        self.qm_estimator.fit_batched(x_res, y_res, valid=valid)
        self._valid, self._flag = valid, flag
        return self

    def check_fit(self):
        self.qm_estimator.check_fit()
        if int(self._flag.item()) != 0:
            self._flag.zero_()
            raise ValueError('Input contains NaN or infinity.')

    def predict_batched(self, X: torch.Tensor) -> torch.Tensor:
        if not hasattr(self, '_x_mean_fit'):
            raise NotFittedError(f"This {type(self).__name__} instance is not fitted yet. Call 'fit' with "
                                 "appropriate arguments before using this estimator.")
        n = X.shape[0]
        table = whole_series_table(n)
        lt = copy.deepcopy(self.trend_transformer).fit_batched(X, self._valid)
        y_hat = self.qm_estimator.predict_batched(lt.transform_batched(X))                      # float64 [n, C]
        x_mean = engine.group_mean(X, table, _lib.MEAN_NUMPY, self._valid, self._flag)
        delta = ((x_mean - self._x_mean_fit.to(x_mean.dtype)) + self._y_mean_fit.to(x_mean.dtype)).to(torch.float64)
        line_mean = lt._slope * ((n - 1) / 2.0) + lt._icpt                                       # mean of the new trend line
        # y_hat + (line - mean(line)) + delta, in one pass: RESTORE adds the line and moves its intercept
        return engine.trend_apply(_lib.TREND_RESTORE, y_hat, table, lt._slope, lt._icpt, lt._icpt - line_mean + delta,
                                  valid=self._valid)

    # ---- per-cell API of the reference (one series)
    def fit(self, X, y):
        dev = cuda_device()
        x, _, _ = series_to_device(X, dev)
        yt, _, _ = series_to_device(y, dev)
        if x.shape[1] != 1 or yt.shape[1] != 1:
            raise ValueError('X and y must be single-column')
        if x.dtype != yt.dtype:
            x, yt = x.to(torch.float64), yt.to(torch.float64)
        self.fit_batched(x, yt)
        self.check_fit()
        return self

    def predict(self, X):
        x, _, _ = series_to_device(X, cuda_device())
        if hasattr(self, '_x_mean_fit') and x.dtype != self._x_mean_fit.dtype:
            x = x.to(self._x_mean_fit.dtype)
        out = self.predict_batched(x[:, :1].contiguous())
        self.check_fit()
        return out.cpu().numpy().reshape(-1, 1)
