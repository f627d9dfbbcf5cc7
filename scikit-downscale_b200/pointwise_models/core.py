"""PointWiseDownscaler — drop-in for skdownscale.pointwise_models.PointWiseDownscaler
(skdownscale/pointwise_models/core.py:200-448).

The reference applies a copy of the model to every grid cell in a Python loop
(core.py:86-96, 137-141).  Here the B200-native estimators are applied to ALL cells of the
block with one batched device call per fit / predict.  Inputs may be

* ``xarray.DataArray`` / ``Dataset`` (when xarray is installed) with a ``time`` dimension —
  same contract as the reference, the result is a DataArray with the reference's dims/coords;
* plain ``numpy`` arrays or ``torch`` tensors shaped ``(time, *cell_dims)`` — or
  ``(time, feature, *cell_dims)`` for the multi-feature GARD models — with the time axis
  labels passed as ``time=<DatetimeIndex>`` (needed by the BCSD models).

Only estimators of this package run (there is no generic per-cell Python fallback).
"""

from __future__ import annotations

import copy

import numpy as np
import pandas as pd
import torch

from .. import engine
from .base import cuda_device
from .bcsd import BcsdBase
from .gard import AnalogBase, PureRegression
from .quantile import (LinearTrendTransformer, QuantileMapper, QuantileMappingReressor,
                       TrendAwareQuantileMappingRegressor)
from .zscore import ZScoreRegressor

try:  # xarray is optional (absent in the build container)
    import xarray as xr
except Exception:  # pragma: no cover
    xr = None

DEFAULT_FEATURE_DIM = 'variable'


class _Block:
    """A (time, feature, cells) view of one input plus what is needed to rebuild the output."""

    def __init__(self, data, index, cell_shape, kind, template=None, feature_dim=None, pinned=False):
        self.data = data            # numpy or torch, shape [T, p, C]
        self.index = index          # pandas index of the time axis or None
        self.cell_shape = cell_shape
        self.kind = kind            # 'xarray' | 'numpy' | 'torch'
        self.template = template
        self.feature_dim = feature_dim


class PointWiseDownscaler:
    """Pointwise downscaling model wrapper (core.py:200-448), batched over cells on the GPU.

    Parameters
    ----------
    model : estimator of this package (BcsdTemperature, BcsdPrecipitation, QuantileMapper,
        PureAnalog, AnalogRegression)
    dim : str, optional
        Dimension to apply the model along. Default is ``time``.
    device : torch device, optional (default: current CUDA device)
    """

    def __init__(self, model, dim: str = 'time', device=None, chunk_cells: int = 16384) -> None:
        self._dim = dim
        self._model = model
        self._models = None
        self._device = device
        # host blocks with more cells than this are streamed through the GPU in cell chunks
        # (H2D / kernels / D2H overlapped) instead of being copied whole
        self._chunk_cells = chunk_cells
        if not hasattr(model, 'fit'):
            raise TypeError(f'Type {type(model)} does not have the fit method required by PointWiseDownscaler')

    # ------------------------------------------------------------------ input handling
    def _multi_feature(self):
        return isinstance(self._model, (AnalogBase, PureRegression))

    def _to_block(self, X, feature_dim, time=None) -> _Block:
        """core.py:427-440 (_to_feature_x) + core.py:40-66 (_da_to_df) for a whole block."""
        if xr is not None and isinstance(X, (xr.DataArray, xr.Dataset)):
            if isinstance(X, xr.Dataset):
                X = X.to_array(feature_dim)
            if feature_dim not in X.dims:
                X = X.expand_dims(**{feature_dim: [f'{feature_dim}_0']}, axis=1)
            X = X.transpose(self._dim, feature_dim, ...)
            a = np.asarray(X.data)
            try:
                index = X.indexes[self._dim]
            except (KeyError, AttributeError):
                index = pd.RangeIndex(X.sizes[self._dim])
            return _Block(a.reshape(a.shape[0], a.shape[1], -1), index, a.shape[2:], 'xarray', template=X,
                          feature_dim=feature_dim)
        is_torch = isinstance(X, torch.Tensor)
        a = X if is_torch else np.asarray(X)
        if a.ndim < 1:
            raise ValueError('X must have a time axis')
        if self._multi_feature():
            if a.ndim < 2:
                raise ValueError('multi-feature models take arrays shaped (time, feature, *cells)')
            cell_shape = tuple(a.shape[2:])
            a = a.reshape(a.shape[0], a.shape[1], -1)
        else:
            cell_shape = tuple(a.shape[1:])
            a = a.reshape(a.shape[0], 1, -1)
        index = None if time is None else pd.Index(time)
        if index is not None and len(index) != a.shape[0]:
            raise ValueError(f'time has {len(index)} labels, the array has {a.shape[0]} rows')
        return _Block(a, index, cell_shape, 'torch' if is_torch else 'numpy')

    def _dev(self):
        return cuda_device(self._device)

    def _wrap(self, out: torch.Tensor, blk: _Block, n_outputs=1, output_names=None):
        """core.py:119-135: dims = X dims minus the feature dim, or feature dim := output names."""
        T = out.shape[0]
        shape = (T,) + ((n_outputs,) if n_outputs > 1 else ()) + tuple(blk.cell_shape)
        if blk.kind == 'torch':
            return out.reshape(shape)
        res = out.cpu().numpy().reshape(shape)
        if blk.kind == 'numpy':
            return res
        X = blk.template
        fd = blk.feature_dim
        dims = [d for d in X.dims if d != fd]
        coords = {k: v for k, v in X.coords.items() if fd not in v.dims}
        if n_outputs > 1:
            dims = [dims[0], fd] + dims[1:]
            coords[fd] = output_names
        return xr.DataArray(res, dims=dims, coords=coords)

    # ------------------------------------------------------------------ fit / predict
    def fit(self, X, *args, **kwargs):
        """Fit the model for every cell (core.py:225-264).  ``fit(X, y)`` or ``fit(X)``."""
        kws = {'feature_dim': DEFAULT_FEATURE_DIM} | kwargs
        if len(args) > 1:
            raise ValueError(f'Expected at most 1 positional argument, got {len(args)}')
        time = kws.pop('time', None)
        fd = kws.pop('feature_dim')
        kws.pop('along_dim', None)
        if kws:
            raise TypeError(f'unsupported fit parameters on the B200 path: {sorted(kws)}')
        dev = self._dev()
        bx = self._to_block(X, fd, time)
        # core.py:87: the reference fits deep copies, the estimator handed to the wrapper is never fitted itself —
        # two wrappers sharing one estimator (or a refit on another block) must not overwrite each other's state
        model = copy.deepcopy(self._model)
        if self._streamed(bx) and args:
            by = self._to_block_y(args[0], fd, time)
            if bx.index is None:
                raise ValueError('BCSD models need the time axis labels: pass time=<DatetimeIndex>')
            if bx.data.shape[1] != 1:
                raise ValueError(f'BCSD only supports 1 feature, found {bx.data.shape[1]}')
            model.fit_host(bx.data[:, 0], by.data[:, 0], bx.index, device=dev, chunk_cells=self._chunk_cells)
            model.check_fit()
            self._models = model
            self._valid = model._state.valid
            return
        x = engine.as_device(bx.data, dev)
        x = self._float(x)
        valid = engine.cell_mask(x[0, 0])                      # core.py:35-37,77-78
        if isinstance(model, QuantileMapper):
            if x.shape[1] != 1:
                raise ValueError('CunnaneTransformer.fit() only supports a single feature')
            model.fit_batched(x[:, 0], valid=valid)
        elif isinstance(model, LinearTrendTransformer):
            if x.shape[1] != 1:
                raise ValueError('LinearTrendTransformer behind PointWiseDownscaler takes one feature per cell')
            model.fit_batched(x[:, 0], valid=valid)
        else:
            if not args:
                raise TypeError(f'{type(model).__name__}.fit() missing 1 required positional argument: y')
            by = self._to_block_y(args[0], fd, time)
            y = self._float(engine.as_device(by.data, dev)).to(x.dtype)
            if y.shape[0] != x.shape[0] or y.shape[-1] != x.shape[-1]:
                raise ValueError(f'X {tuple(x.shape)} and y {tuple(y.shape)} do not describe the same block')
            if isinstance(model, BcsdBase):
                if x.shape[1] != 1:
                    raise ValueError(f'BCSD only supports 1 feature, found {x.shape[1]}')
                if bx.index is None:
                    raise ValueError('BCSD models need the time axis labels: pass time=<DatetimeIndex>')
                if by.index is not None:
                    pd.testing.assert_index_equal(pd.Index(bx.index), pd.Index(by.index))   # base.py:17
                model.fit_batched(x[:, 0], y[:, 0], bx.index, valid=valid)
            elif isinstance(model, ZScoreRegressor):
                if x.shape[1] != 1:
                    raise ValueError(f'Zscore only supports 1 feature, found {x.shape[1]}')
                if bx.index is None:
                    raise ValueError('ZScoreRegressor needs the time axis labels: pass time=<DatetimeIndex>')
                model.fit_batched(x[:, 0], y[:, 0], bx.index, valid=valid)
            elif isinstance(model, (AnalogBase, PureRegression)):
                model.fit_batched(x, y[:, 0], valid=valid)
            elif isinstance(model, (QuantileMappingReressor, TrendAwareQuantileMappingRegressor)):
                if x.shape[1] != 1:
                    raise ValueError(f'X should have up to 1 features, found {x.shape[1]}')
                model.fit_batched(x[:, 0], y[:, 0], valid=valid)
            else:
                raise TypeError(f'{type(model).__name__} is not a B200-native estimator; PointWiseDownscaler '
                                'has no per-cell Python fallback')
        model.check_fit() if hasattr(model, 'check_fit') else None
        self._models = model
        self._valid = valid

    def _streamed(self, blk) -> bool:
        """Host-resident BCSD block big enough to be worth the chunked copy/compute pipeline."""
        if not isinstance(self._model, BcsdBase) or blk.kind == 'xarray':
            return False
        if (self._model.qm_kwargs or {}).get('detrend', False):
            return False                                   # composed float64 path: whole block on the device
        d = blk.data
        on_host = (not isinstance(d, torch.Tensor)) or (not d.is_cuda)
        ok_dtype = (d.dtype in (torch.float32, torch.float64)) if isinstance(d, torch.Tensor) else (d.dtype in (np.float32, np.float64))
        return on_host and ok_dtype and d.shape[-1] > self._chunk_cells

    def _to_block_y(self, y, fd, time):
        multi = self._multi_feature()
        if not (xr is not None and isinstance(y, (xr.DataArray, xr.Dataset))) and multi:
            # y of a multi-feature model is still (time, *cells)
            is_torch = isinstance(y, torch.Tensor)
            a = y if is_torch else np.asarray(y)
            return _Block(a.reshape(a.shape[0], 1, -1), None if time is None else pd.Index(time),
                          tuple(a.shape[1:]), 'torch' if is_torch else 'numpy')
        return self._to_block(y, fd, time) if not multi else self._to_block_single(y, fd)

    def _to_block_single(self, y, fd):
        if isinstance(y, xr.Dataset):
            y = y.to_array(fd)
        if fd not in y.dims:
            y = y.expand_dims(**{fd: [f'{fd}_0']}, axis=1)
        y = y.transpose(self._dim, fd, ...)
        a = np.asarray(y.data)
        return _Block(a.reshape(a.shape[0], a.shape[1], -1), None, a.shape[2:], 'xarray', template=y, feature_dim=fd)

    @staticmethod
    def _float(t: torch.Tensor) -> torch.Tensor:
        return t if t.dtype in (torch.float32, torch.float64) else t.to(torch.float64)

    def _require_fit(self):
        if self._models is None:
            raise RuntimeError('PointWiseDownscaler is not fitted yet')

    def predict(self, X, **kwargs):
        """Predict for every cell (core.py:266-338).  Output dtype = input dtype, NaN where the cell
        was masked at fit time (core.py:129)."""
        self._require_fit()
        kws = {'feature_dim': DEFAULT_FEATURE_DIM} | kwargs
        time = kws.pop('time', None)
        fd = kws.pop('feature_dim')
        kws.pop('along_dim', None)
        kws_out = kws.pop('out', None)      # optional (pinned) host array [time, cells] receiving a streamed result
        if kws:
            raise TypeError(f'unsupported predict parameters on the B200 path: {sorted(kws)}')
        model = self._models
        if isinstance(model, QuantileMapper):
            raise AttributeError("'QuantileMapper' object has no attribute 'predict'")
        dev = self._dev()
        out_host = kws_out
        bx = self._to_block(X, fd, time)
        if self._streamed(bx):
            if bx.index is None:
                raise ValueError('BCSD models need the time axis labels: pass time=<DatetimeIndex>')
            res = model.predict_host(bx.data[:, 0], bx.index, out=out_host, chunk_cells=self._chunk_cells)
            res = res.reshape((res.shape[0],) + tuple(bx.cell_shape))
            return res if bx.kind == 'torch' else res.numpy()
        x = self._float(engine.as_device(bx.data, dev))
        if isinstance(model, BcsdBase):
            if bx.index is None:
                raise ValueError('BCSD models need the time axis labels: pass time=<DatetimeIndex>')
            out = model.predict_batched(x[:, 0], bx.index)
            model._state.check_finite()
            return self._wrap(out, bx)
        if isinstance(model, LinearTrendTransformer):
            raise AttributeError("'LinearTrendTransformer' object has no attribute 'predict'")
        if isinstance(model, ZScoreRegressor):
            if x.shape[1] != 1:
                raise ValueError(f'X must have exactly 1 feature, got {x.shape[1]}')
            out = model.predict_batched(x[:, 0])
            model.check_fit()
            return self._wrap(out, bx)
        if isinstance(model, (QuantileMappingReressor, TrendAwareQuantileMappingRegressor)):
            out = model.predict_batched(x[:, 0]).to(x.dtype)          # core.py:129: the wrapper's array has X.dtype
            model.check_fit()
            return self._wrap(out, bx)
        out = model.predict_batched(x)
        model._check_finite()
        return self._wrap(out, bx, model.n_outputs, model.output_names)

    def _apply_transformer(self, X, direction, **kwargs):
        self._require_fit()
        kws = {'feature_dim': DEFAULT_FEATURE_DIM} | kwargs
        time = kws.pop('time', None)
        fd = kws.pop('feature_dim')
        if kws:
            raise TypeError(f'unsupported {direction} parameters on the B200 path: {sorted(kws)}')
        model = self._models
        fn = getattr(model, direction + '_batched', None)
        if fn is None:
            raise AttributeError(f"'{type(model).__name__}' object has no attribute '{direction}'")
        bx = self._to_block(X, fd, time)
        x = self._float(engine.as_device(bx.data, self._dev()))
        if x.shape[1] != 1:
            raise ValueError(f'{type(model).__name__}.{direction} takes one feature per cell, found {x.shape[1]}')
        out = fn(x[:, 0])
        model._state.check_finite() if isinstance(model, QuantileMapper) else model.check_fit()
        out = out.to(x.dtype)                            # core.py:129-131: the wrapper stores into an X.dtype array
        T = out.shape[0]
        if bx.kind == 'torch':
            return out.reshape((T,) + tuple(bx.cell_shape))
        res = out.cpu().numpy()
        if bx.kind == 'numpy':
            return res.reshape((T,) + tuple(bx.cell_shape))
        return xr.DataArray(res.reshape(bx.template.shape), dims=bx.template.dims, coords=bx.template.coords)

    def transform(self, X, **kwargs):
        """core.py:340-370: ``QuantileMapper`` / ``LinearTrendTransformer`` fitted for every cell, applied to X."""
        return self._apply_transformer(X, 'transform', **kwargs)

    def inverse_transform(self, X, **kwargs):
        """core.py:372-403 (``LinearTrendTransformer``: add the fitted trend line back)."""
        return self._apply_transformer(X, 'inverse_transform', **kwargs)

    def get_attr(self, key, dtype, template_output=None):
        """core.py:405-425 for the fitted climatologies (``y_climo_``, ``_x_climo``): returns an array
        ``[group, *cells]`` instead of one object per cell."""
        self._require_fit()
        st = getattr(self._models, '_state', None)
        table = {'y_climo_': 'y_climo', '_x_climo': 'x_climo'}
        if st is None or key not in table or getattr(st, table[key]) is None:
            raise AttributeError(f'{type(self._models).__name__} has no batched attribute {key!r}')
        return getattr(st, table[key]).cpu().numpy().astype(dtype)

    def __repr__(self):
        summary = [f'<skdownscale_b200.{self.__class__.__name__}>',
                   f'  Fit Status: {self._models is not None}',
                   f'  Model:\n    {self._model}']
        return '\n'.join(summary)
