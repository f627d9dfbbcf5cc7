"""PureAnalog / AnalogRegression — drop-ins for skdownscale/pointwise_models/gard.py:55-364,
executed for all cells at once on the GPU (exact float64 k-nearest-neighbour search +
fused statistic / OLS epilogue, csrc/analog_kernels.cu).
"""

from __future__ import annotations

import os
import warnings

import numpy as np
import pandas as pd
import torch
from sklearn.base import BaseEstimator, RegressorMixin
from sklearn.exceptions import NotFittedError

from .. import _lib, engine
from .base import cuda_device, series_to_device
from .utils import default_none_kwargs

_KINDS = {'best_analog': _lib.ANALOG_BEST, 'sample_analogs': _lib.ANALOG_SAMPLE,
          'weight_analogs': _lib.ANALOG_WEIGHT, 'mean_analogs': _lib.ANALOG_MEAN}
_MAX_ANALOGS = 256


def _check_tree_kwargs(kdtree_kwargs, query_kwargs):
    """The search is exact brute force: tree-shape options are accepted and ignored, options that
    would change the answer are refused."""
    for k, v in default_none_kwargs(kdtree_kwargs).items():
        if k == 'leaf_size' or (k == 'metric' and v in ('euclidean', 'minkowski', 'l2')):
            continue
        raise NotImplementedError(f'kdtree_kwargs {k}={v!r} is not supported on the B200 path')
    for k, v in default_none_kwargs(query_kwargs).items():
        if k in ('dualtree', 'breadth_first') or (k == 'sort_results' and v) or (k == 'return_distance' and v):
            continue
        raise NotImplementedError(f'query_kwargs {k}={v!r} is not supported on the B200 path')


class AnalogBase(RegressorMixin, BaseEstimator):
    """gard.py:55-98.  ``fit`` keeps the training window on the device (the reference builds a
    KDTree per cell); the neighbour search happens inside ``predict``."""

    _fit_attributes = ['kdtree_', 'X_', 'y_', 'k_']
    n_outputs = 3
    output_names = ['pred', 'exceedance_prob', 'prediction_error']

    # ---- batched API
    def fit_batched(self, X: torch.Tensor, y: torch.Tensor, valid=None):
        """X ``[T, p, C]``, y ``[T, C]`` CUDA tensors."""
        _check_tree_kwargs(self.kdtree_kwargs, self.query_kwargs)
        T = X.shape[0]
        if T >= self.n_analogs:                               # gard.py:75-79
            self.k_ = self.n_analogs
        else:
            warnings.warn('length of X is less than n_analogs, setting n_analogs = len(X)')
            self.k_ = T
        if self.k_ > _MAX_ANALOGS:
            raise NotImplementedError(f'n_analogs > {_MAX_ANALOGS} is not supported on the B200 path')
        self._Xtr, self._ytr, self._valid = X.contiguous(), y.contiguous(), valid
        self._nonfinite = torch.zeros(1, dtype=torch.int32, device=X.device)
        # the counterpart of the reference's KDTree build (gard.py:82): every cell's training rows grouped by the
        # boxes of a quantile grid over the first predictors, which is what the search prunes on
        self._order_train = engine.analog_grid_fit(self._Xtr, valid) if engine.analog_grid_supported(self._Xtr, self.k_) else None
        self.n_features_in_ = X.shape[1]
        return self

    def check_fit(self):
        pass

    def _check_finite(self):
        flags = int(self._nonfinite.item())
        if flags != 0:
            self._nonfinite.zero_()
            if flags & 1:
                raise ValueError('Input contains NaN or infinity.')
            # gard.py:208-209: LogisticRegression.fit on a timestep whose analogs are all at or below thresh
            raise ValueError('This solver needs samples of at least 2 classes in the data, but the data '
                             'contains only one class: 0')

    def _run(self, kind, k, X, out_dtype, thresh=None, rand_idx=None, want_idx=False, logistic_C=1.0):
        if not hasattr(self, '_Xtr'):
            raise NotFittedError(f"This {type(self).__name__} instance is not fitted yet. Call 'fit' with "
                                 "appropriate arguments before using this estimator.")
        if X.shape[1] != self.n_features_in_:
            raise ValueError(f'X has {X.shape[1]} features, but {self.__class__.__name__} is expecting '
                             f'{self.n_features_in_} features as input.')
        if X.dtype != self._Xtr.dtype:
            X = X.to(self._Xtr.dtype)
        prune = self._order_train is not None and engine.analog_grid_supported(self._Xtr, k)
        return engine.analog_predict(kind, self._Xtr, self._ytr, X, k, thresh=thresh, rand_idx=rand_idx,
                                     out_dtype=out_dtype, want_idx=want_idx, valid=self._valid,
                                     nonfinite=self._nonfinite, logistic_C=logistic_C, prune=prune,
                                     grid=self._order_train if prune else None)

    # ---- per-cell API
    def fit(self, X, y):
        dev = cuda_device()
        x_t, _, _ = series_to_device(X, dev)
        y_t, _, _ = series_to_device(y, dev)
        if y_t.shape[1] != 1:
            raise ValueError('y should be a 1d array or a column vector')
        if x_t.shape[0] != y_t.shape[0]:
            raise ValueError(f'Found input variables with inconsistent numbers of samples: [{x_t.shape[0]}, {y_t.shape[0]}]')
        if x_t.dtype != y_t.dtype:
            x_t, y_t = x_t.to(torch.float64), y_t.to(torch.float64)
        self.fit_batched(x_t.unsqueeze(-1), y_t)
        return self

    def predict(self, X):
        return_df = isinstance(X, pd.DataFrame)
        x_t, _, _ = series_to_device(X, cuda_device())
        out = self.predict_batched(x_t.unsqueeze(-1), out_dtype=torch.float64)
        self._check_finite()
        out = out[:, :, 0].cpu().numpy()
        return pd.DataFrame(out, columns=self.output_names) if return_df else out


class AnalogRegression(AnalogBase):
    """AnalogRegression (gard.py:101-224): k analogs → OLS → prediction, exceedance
    probability and in-sample RMSE.  With ``thresh`` the exceedance probability is P(class 0) of
    the per-timestep logistic regression (gard.py:205-212; the reference's lbfgs solution to its
    tol = 1e-4, here the exact optimum) and the OLS uses the analogs above ``thresh`` only."""

    def __init__(self, n_analogs=200, thresh=None, kdtree_kwargs=None, query_kwargs=None,
                 logistic_kwargs=None, lr_kwargs=None):
        self.n_analogs = n_analogs
        self.thresh = thresh
        self.kdtree_kwargs = kdtree_kwargs
        self.query_kwargs = query_kwargs
        self.logistic_kwargs = logistic_kwargs
        self.lr_kwargs = lr_kwargs

    def predict_batched(self, X: torch.Tensor, out_dtype=None, want_idx=False, **_):
        # logistic_kwargs (gard.py:171-172): the exceedance model is solved to its optimum on the device,
        # so only what defines the objective matters — C; solver tolerances are accepted and ignored
        C_reg = 1.0
        for k, v in default_none_kwargs(self.logistic_kwargs).items():
            if k == 'C':
                C_reg = float(v)
            elif k in ('tol', 'max_iter', 'n_jobs', 'verbose', 'random_state', 'warm_start'):
                pass
            elif (k, v) in (('penalty', 'l2'), ('fit_intercept', True), ('solver', 'lbfgs'), ('dual', False),
                            ('class_weight', None), ('intercept_scaling', 1), ('l1_ratio', None)):
                pass
            else:
                raise NotImplementedError(f'logistic_kwargs {k}={v!r} is not supported on the B200 path')
        for k, v in default_none_kwargs(self.lr_kwargs).items():
            if not (k == 'fit_intercept' and v) and not (k in ('copy_X', 'n_jobs', 'tol')) and not (k == 'positive' and not v):
                raise NotImplementedError(f'lr_kwargs {k}={v!r} is not supported on the B200 path')
        return self._run(_lib.ANALOG_REGRESSION, self.k_, X, out_dtype, thresh=self.thresh, want_idx=want_idx,
                         logistic_C=C_reg)


class PureAnalog(AnalogBase):
    """PureAnalog (gard.py:227-364): best / sampled / distance-weighted / mean analog."""

    def __init__(self, n_analogs=200, kind='best_analog', thresh=None, kdtree_kwargs=None, query_kwargs=None):
        self.n_analogs = n_analogs
        self.kind = kind
        self.thresh = thresh
        self.kdtree_kwargs = kdtree_kwargs
        self.query_kwargs = query_kwargs

    def predict_batched(self, X: torch.Tensor, out_dtype=None, want_idx=False, rand_idx=None, **_):
        if self.kind == 'best_analog' or self.n_analogs == 1:      # gard.py:290-296
            k, kind = 1, 'best_analog'
        else:
            k, kind = getattr(self, 'k_', None), self.kind
        if kind not in _KINDS:
            raise ValueError(f'got unexpected kind {kind}')
        if kind == 'sample_analogs' and rand_idx is None:
            # gard.py:315 draws from numpy's GLOBAL generator, one call per cell in cell order
            # (masked cells never reach predict in the reference, so they draw nothing)
            Tq, C = X.shape[0], X.shape[-1]
            ok = np.ones(C, dtype=bool) if self._valid is None else self._valid.cpu().numpy().astype(bool)
            rand_idx = np.zeros((Tq, C), dtype=np.int32)
            for c in range(C):
                if ok[c]:
                    rand_idx[:, c] = np.random.randint(low=0, high=k, size=Tq)
        return self._run(_KINDS[kind], k, X, out_dtype, thresh=self.thresh, rand_idx=rand_idx, want_idx=want_idx)


class PureRegression(RegressorMixin, BaseEstimator):
    """PureRegression (gard.py:367-504): one linear regression per cell over the whole training window
    (the rows above ``thresh`` when given), ``exceedance_prob`` = P(exceed) of a logistic regression on
    all rows, ``prediction_error`` = the in-sample RMSE."""

    _fit_attributes = ['logistic_model_', 'linear_model_', 'fit_error_']
    n_outputs = 3
    output_names = ['pred', 'exceedance_prob', 'prediction_error']

    def __init__(self, thresh=None, logistic_kwargs=None, linear_kwargs=None):
        self.thresh = thresh
        self.logistic_kwargs = logistic_kwargs
        self.linear_kwargs = linear_kwargs

    def fit_batched(self, X: torch.Tensor, y: torch.Tensor, valid=None):
        """X ``[T, p, C]``, y ``[T, C]`` CUDA tensors."""
        C_reg = 1.0
        for k, v in default_none_kwargs(self.logistic_kwargs).items():
            if k == 'C':
                C_reg = float(v)
            elif k not in ('tol', 'max_iter', 'n_jobs', 'verbose', 'random_state', 'warm_start'):
                raise NotImplementedError(f'logistic_kwargs {k}={v!r} is not supported on the B200 path')
        for k, v in default_none_kwargs(self.linear_kwargs).items():
            if not ((k == 'fit_intercept' and v) or k in ('copy_X', 'n_jobs', 'tol') or (k == 'positive' and not v)):
                raise NotImplementedError(f'linear_kwargs {k}={v!r} is not supported on the B200 path')
        self._valid = valid
        self._nonfinite = torch.zeros(1, dtype=torch.int32, device=X.device)
        self._dtype = X.dtype
        self._model = engine.pure_regression_fit(X, y, thresh=self.thresh, logistic_C=C_reg, valid=valid,
                                                 nonfinite=self._nonfinite)
        self.n_features_in_ = X.shape[1]
        return self

    def check_fit(self):
        self._check_finite()

    def _check_finite(self):
        flags = int(self._nonfinite.item())
        if flags:
            self._nonfinite.zero_()
            if flags & 1:
                raise ValueError('Input contains NaN or infinity.')
            # gard.py:435: no row above thresh → LinearRegression.fit on an empty selection
            raise ValueError(f'Found array with 0 sample(s) (shape=(0, {self.n_features_in_})) while a minimum of 1 '
                             'is required by LinearRegression.')

    def predict_batched(self, X: torch.Tensor, out_dtype=None, **_):
        if not hasattr(self, '_model'):
            raise NotFittedError(f"This {type(self).__name__} instance is not fitted yet. Call 'fit' with "
                                 "appropriate arguments before using this estimator.")
        if X.shape[1] != self.n_features_in_:
            raise ValueError(f'X has {X.shape[1]} features, but {self.__class__.__name__} is expecting '
                             f'{self.n_features_in_} features as input.')
        if X.dtype != self._dtype:
            X = X.to(self._dtype)
        return engine.pure_regression_predict(self._model, X, out_dtype=out_dtype, valid=self._valid,
                                              nonfinite=self._nonfinite)

    # ---- per-cell API of the reference
    def fit(self, X, y):
        dev = cuda_device()
        x_t, _, _ = series_to_device(X, dev)
        y_t, _, _ = series_to_device(y, dev)
        if y_t.shape[1] != 1:
            raise ValueError('y should be a 1d array or a column vector')
        if x_t.dtype != y_t.dtype:
            x_t, y_t = x_t.to(torch.float64), y_t.to(torch.float64)
        self.fit_batched(x_t.unsqueeze(-1), y_t)
        self._check_finite()
        self.fit_error_ = float(self._model[0, self.n_features_in_ + 1 if self.n_features_in_ <= 4 else 9].item())
        return self

    def predict(self, X):
        return_df = isinstance(X, pd.DataFrame)
        x_t, _, _ = series_to_device(X, cuda_device())
        out = self.predict_batched(x_t.unsqueeze(-1), out_dtype=torch.float64)
        self._check_finite()
        out = out[:, :, 0].cpu().numpy()
        return pd.DataFrame(out, columns=self.output_names) if return_df else out
