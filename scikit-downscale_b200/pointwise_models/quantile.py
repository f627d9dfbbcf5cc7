"""QuantileMapper — drop-in for skdownscale.pointwise_models.QuantileMapper
(skdownscale/pointwise_models/quantile.py:46-157), executed on the GPU.

Only the default configuration of the reference's CunnaneTransformer is on the hot path
(alpha = beta = 0.4, extrapolate='both', n_endpoints=10; quantile.py:420-432); other
settings and ``detrend=True`` raise NotImplementedError (SURVEY.md §8(f) "next").
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


def check_qt_kwargs(qt_kwargs):
    for k, v in default_none_kwargs(qt_kwargs).items():
        if k not in _QT_DEFAULTS:
            raise TypeError(f"CunnaneTransformer.__init__() got an unexpected keyword argument '{k}'")
        if v != _QT_DEFAULTS[k]:
            raise NotImplementedError(f'qt_kwargs {k}={v!r}: only the default Cunnane settings '
                                      f'{_QT_DEFAULTS} run on the B200 path')


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
                                 out_dtype=out_dtype, want_rank=want_rank)

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
