"""Input conventions shared by the estimators — the error behaviour of
skdownscale/pointwise_models/base.py:13-57 at the Python boundary, and the
host → device hand-off of the per-cell (single series) API.
"""

from __future__ import annotations

import warnings

import numpy as np
import pandas as pd
import torch
from sklearn.base import BaseEstimator


def cuda_device(device=None) -> torch.device:
    """The device the kernels run on.  There is no CPU path: fail loudly without CUDA."""
    if not torch.cuda.is_available():
        raise RuntimeError('skdownscale_b200 needs a CUDA device (sm_100a); there is no CPU fallback')
    if device is None:
        return torch.device('cuda', torch.cuda.current_device())
    return torch.device(device)


def _float_dtype(a: np.ndarray) -> np.dtype:
    return a.dtype if a.dtype in (np.float32, np.float64) else np.dtype(np.float64)


def series_to_device(obj, device) -> tuple[torch.Tensor, pd.Index | None, object]:
    """One cell's samples → ``[T, p]`` CUDA tensor (+ its index and column labels)."""
    if isinstance(obj, pd.Series):
        obj = obj.to_frame()
    if isinstance(obj, pd.DataFrame):
        index, columns = obj.index, obj.columns
        a = obj.to_numpy()
    else:
        index, columns = None, None
        a = np.asarray(obj)
    if a.ndim == 1:
        a = a.reshape(-1, 1)
    if a.ndim != 2:
        raise ValueError(f'Found array with dim {a.ndim}. Expected 2 (n_samples, n_features).')
    if a.shape[0] == 0:
        raise ValueError('Found array with 0 sample(s) while a minimum of 1 is required.')
    a = np.ascontiguousarray(a, dtype=_float_dtype(a))
    return torch.from_numpy(a).to(device), index, columns


class TimeSynchronousDownscaler(BaseEstimator):
    """Base of the BCSD estimators (base.py:12-136): X and y must share one time index."""

    _timestep = 'M'

    def _frames(self, X, y=None):
        """base.py:13-36: DataFrames must have equal indexes (AssertionError otherwise);
        anything else gets a made-up monthly index and a warning."""
        if y is not None:
            if isinstance(X, pd.DataFrame) and isinstance(y, pd.DataFrame):
                pd.testing.assert_index_equal(X.index, y.index)
                return X, y
            Xa, ya = np.asarray(X), np.asarray(y)
            if len(Xa) != len(ya):
                raise ValueError(f'Found input variables with inconsistent numbers of samples: [{len(Xa)}, {len(ya)}]')
            warnings.warn('X and y do not have pandas DateTimeIndexes, making one up...')
            index = pd.date_range(periods=len(Xa), start='1950', freq='MS')
            return (pd.DataFrame(Xa.reshape(len(Xa), -1), index=index),
                    pd.DataFrame(ya.reshape(len(ya), -1), index=index))
        if isinstance(X, pd.DataFrame):
            return X
        Xa = np.asarray(X)
        warnings.warn('array does not have a pandas DateTimeIndex, making one up...')
        index = pd.date_range(periods=len(Xa), start='1950', freq='MS')
        return pd.DataFrame(Xa.reshape(len(Xa), -1), index=index)
