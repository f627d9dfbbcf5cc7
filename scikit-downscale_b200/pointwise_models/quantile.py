"""QuantileMapper — drop-in for skdownscale.pointwise_models.QuantileMapper
(skdownscale/pointwise_models/quantile.py:46-157), executed on the GPU.

``qt_kwargs`` (alpha, beta, extrapolate, n_endpoints of the reference's CunnaneTransformer,
quantile.py:420-432) travel to the kernels as ``sdb_cunnane_opts``.
"""

from __future__ import annotations

import numpy as np
import torch
from sklearn.base import BaseEstimator, TransformerMixin
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
        if self.detrend:
            raise NotImplementedError('QuantileMapper(detrend=True) is not on the B200 path yet')
        check_qt_kwargs(self.qt_kwargs)
        self._state = engine.qm_fit(X, whole_series_table(X.shape[0]), valid=valid, want_y_climo=False)
        return self

    def transform_batched(self, X: torch.Tensor, out_dtype=None, want_rank=False):
        if not hasattr(self, '_state'):
            raise NotFittedError(f"This {type(self).__name__} instance is not fitted yet.")
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
