"""BcsdTemperature / BcsdPrecipitation — drop-ins for
skdownscale/pointwise_models/bcsd.py:14-289, executed for all cells at once on the GPU.
"""

from __future__ import annotations

import numpy as np
import pandas as pd
import torch
from sklearn.exceptions import NotFittedError

from .. import _lib, engine
from .base import TimeSynchronousDownscaler, cuda_device, series_to_device
from .groupers import (DAY_GROUPER, MONTH_GROUPER, PaddedDOYGrouper, grouper_keys, groups_from_keys,
                       padded_doy_groups, rolling_neighbours)
from .quantile import check_lt_kwargs, check_qt_kwargs, cunnane_opts, fit_detrended
from .utils import default_none_kwargs


class BcsdBase(TimeSynchronousDownscaler):
    """Base class for BCSD model (bcsd.py:14-93)."""

    _fit_attributes = ['y_climo_', 'quantile_mappers_']
    _timestep = 'M'
    _mode = None
    _needs_x_climo = False

    def __init__(self, time_grouper=MONTH_GROUPER, climate_trend_grouper=DAY_GROUPER,
                 climate_trend=MONTH_GROUPER, return_anoms=True, qm_kwargs=None):
        self.time_grouper = time_grouper
        self.climate_trend_grouper = climate_trend_grouper
        self.climate_trend = climate_trend
        self.return_anoms = return_anoms
        self.qm_kwargs = qm_kwargs

    def _pre_fit(self):
        """bcsd.py:34-44, including the in-place replacement of the constructor argument by the
        grouper class that the reference's own test asserts (test_pointwise_models.py:315-320)."""
        if isinstance(self.time_grouper, str):
            if self.time_grouper == 'daily_nasa-nex':
                self.time_grouper = PaddedDOYGrouper
                self.timestep = 'daily'
            else:
                raise NotImplementedError(f"time_grouper={self.time_grouper!r}: only callables and "
                                          "'daily_nasa-nex' run on the B200 path")
        elif self.time_grouper is PaddedDOYGrouper:
            self.timestep = 'daily'
        else:
            self.time_grouper_ = self.time_grouper
            self.timestep = 'monthly'
        qm = default_none_kwargs(self.qm_kwargs)
        for k in qm:
            if k not in ('detrend', 'lt_kwargs', 'qt_kwargs'):
                raise TypeError(f"QuantileMapper.__init__() got an unexpected keyword argument '{k}'")
        self._detrend = bool(qm.get('detrend', False))
        if self._detrend and self.timestep == 'daily':
            raise NotImplementedError("qm_kwargs detrend=True with 'daily_nasa-nex' is not on the B200 path")
        check_qt_kwargs(qm.get('qt_kwargs'))
        check_lt_kwargs(qm.get('lt_kwargs'))

    def _cunnane(self):
        """qm_kwargs['qt_kwargs'] → CunnaneTransformer settings of every group's mapper (bcsd.py:65-67)."""
        return cunnane_opts(default_none_kwargs(self.qm_kwargs).get('qt_kwargs'))

    # ------------------------------------------------------------------ group tables (host, exact)
    def _fit_tables(self, index):
        if self.timestep == 'monthly':
            t = engine.GroupTable(groups_from_keys(grouper_keys(self.time_grouper, index)))
            return t, t, _lib.MEAN_GROUPBY
        full = engine.GroupTable(padded_doy_groups(index))
        # predict only ever looks up day-of-month keys 1..31 (bcsd.py:53,275): sort those groups,
        # keep the climatology of all 366
        return full.subset(range(31)), full, _lib.MEAN_FRAME

    def _predict_tables(self, index):
        """(mapping groups, neighbour table or None)."""
        if self.timestep == 'monthly':
            qm_keys = grouper_keys(self.time_grouper, index)
        else:
            qm_keys = grouper_keys(self.climate_trend_grouper, index)
        table = engine.GroupTable(groups_from_keys(qm_keys))
        nbr = None
        if self._mode == _lib.MODE_BCSD_T:
            roll_keys = grouper_keys(self.climate_trend, index)
            if not np.array_equal(np.asarray(roll_keys), np.asarray(qm_keys)):
                nbr = rolling_neighbours(groups_from_keys(roll_keys), len(index))
        return table, nbr

    # ------------------------------------------------------------------ batched API
    def fit_batched(self, X: torch.Tensor, y: torch.Tensor, index, valid=None):
        """fit for all cells: X, y ``[T, C]`` CUDA tensors sharing the time ``index``."""
        self._pre_fit()
        sort_t, mean_t, how = self._fit_tables(index)
        self._state = engine.qm_fit(y, sort_t, valid=valid, X=X if self._needs_x_climo else None,
                                    mean_table=mean_t, mean_how=how)
        if self._detrend:
            # bcsd.py:65-67 with QuantileMapper(detrend=True): every group's mapper is fitted on the group's
            # residuals about its own trend line (positions 0..len-1 of the group's subsequence)
            self._state_res, self._icpt_fit = fit_detrended(y, sort_t, valid)
        self.n_features_in_ = 1
        return self

    def fit_predict_batched(self, X: torch.Tensor, y: torch.Tensor, X_pred: torch.Tensor, index, valid=None, out=None,
                            keep_state: bool = True, stats=None, fused: bool | None = None):
        """``fit(X, y)`` followed by ``predict(X_pred)`` on the SAME time index.

        ``fused=True`` runs both in one pass (``sdb_bcsd_fit_predict``, the counting-rank kernel of
        csrc/qm_fused.cuh): bit-identical to the two calls and without the fitted state's round trip through
        HBM (dram traffic = the algorithmic bytes), but on B200 it is the SLOWER path (shared-memory bound,
        one CTA per SM — profiles/README.md, round 2), so the default (``fused=None`` → environment variable
        ``SDB_FUSED``, off) issues the two calls.  The fused kernel does not cover float64, 'daily_nasa-nex',
        detrending mappers or groups longer than 1024 steps; those always take the two calls.  The model is
        fitted afterwards (``keep_state=False`` on the fused path keeps the climatologies only)."""
        import os
        self._pre_fit()
        if fused is None:
            fused = os.environ.get('SDB_FUSED', '0') not in ('', '0')
        if fused and self.timestep == 'monthly' and not self._detrend:
            table = engine.GroupTable(groups_from_keys(grouper_keys(self.time_grouper, index)))
            if engine.fused_supported(y.dtype, table) and X_pred.dtype == y.dtype and X_pred.shape == y.shape:
                res, st = engine.qm_fit_predict(y, X_pred, table, self._mode, X_train=X if self._needs_x_climo else None,
                                                return_anoms=self.return_anoms, valid=valid, out=out,
                                                keep_state=keep_state, stats=stats)
                if keep_state:
                    self._state = st
                else:
                    self._climo_state = st
                    self.__dict__.pop('_state', None)
                self.n_features_in_ = 1
                return res
        self.fit_batched(X, y, index, valid=valid)
        return self.predict_batched(X_pred, index, out=out)

    def check_fit(self):
        """Deferred (synchronising) checks of fit: NaN/inf inputs (base.py:18-20) and the
        precipitation climatology (bcsd.py:140-141)."""
        self._state.check_finite()

    def predict_batched(self, X: torch.Tensor, index, out_dtype=None, want_rank=False, out=None):
        if not hasattr(self, '_state'):
            raise NotFittedError(f"This {type(self).__name__} instance is not fitted yet. Call 'fit' with "
                                 "appropriate arguments before using this estimator.")
        if self.timestep == 'daily' and self.return_anoms:
            # bcsd.py:267,271-281 / 170-185: the padded groups overlap, the regrouped frame has a
            # different shape and the reference raises
            raise ValueError('shape of climo is not equal to input array')
        table, nbr = self._predict_tables(index)
        if X.dtype != self._state.dtype:
            X = X.to(self._state.dtype)          # fit float32 / predict float64 works like the reference (core.py:254)
        if getattr(self, '_detrend', False):
            if want_rank:
                raise NotImplementedError('rank instrumentation is not available with detrend=True')
            res = engine.qm_predict_detrended(self._state, self._state_res, self._icpt_fit, X, table, self._mode,
                                              return_anoms=self.return_anoms, roll_nbr=nbr, out_dtype=out_dtype,
                                              cunnane=self._cunnane())
            if out is not None:
                out.copy_(res)
                return out
            return res
        return engine.qm_predict(self._state, X, table, self._mode, return_anoms=self.return_anoms,
                                 roll_nbr=nbr, out_dtype=out_dtype, want_rank=want_rank, out=out,
                                 cunnane=self._cunnane())

    def predict_gathered(self, X: torch.Tensor, index, gather, chunk_cells: int = 16200):
        """predict this rank's shard in cell chunks straight into its column block of ``gather.full``
        (:class:`skdownscale_b200.distributed.PeerGather`) and push every finished chunk to the peers while the
        next one is computed.  Returns the complete ``[T, n_cells]`` field (valid on every rank)."""
        if not hasattr(self, '_state'):
            raise NotFittedError(f"This {type(self).__name__} instance is not fitted yet.")
        if getattr(self, '_detrend', False):
            raise NotImplementedError('predict_gathered does not cover qm_kwargs detrend=True')
        if self.timestep == 'daily' and self.return_anoms:
            raise ValueError('shape of climo is not equal to input array')
        table, nbr = self._predict_tables(index)
        st = self._state
        for c0, c1 in gather.chunks(chunk_cells):
            engine.qm_predict(st.cells(c0, c1), X[:, c0:c1], table, self._mode, return_anoms=self.return_anoms,
                              roll_nbr=nbr, out=gather.local[:, c0:c1], cunnane=self._cunnane())
            gather.push(c0, c1)
        return gather.finish()

    # ------------------------------------------------------------------ host arrays: chunked H2D → kernels → D2H
    @staticmethod
    def _host2d(a, name):
        t = a if isinstance(a, torch.Tensor) else torch.from_numpy(np.asarray(a))
        if t.is_cuda or t.dim() != 2:
            raise ValueError(f'{name} must be a host [time, cell] array')
        if t.shape[1] > 1 and t.stride(1) != 1:
            t = t.contiguous()
        return t

    @staticmethod
    def _spans(n_cells, chunk):
        return [(c0, min(c0 + chunk, n_cells)) for c0 in range(0, n_cells, chunk)]

    def _no_detrend_streaming(self):
        if bool(default_none_kwargs(self.qm_kwargs).get('detrend', False)):
            raise NotImplementedError('the host-streaming path does not cover qm_kwargs detrend=True; '
                                      'pass device tensors (or fewer cells than chunk_cells)')

    def fit_host(self, X, y, index, device=None, chunk_cells: int = 16384):
        """fit from HOST arrays ``[T, C]``: cell chunks travel host → device on a copy stream
        (strided 2-D DMA, pinned memory recommended) while the previous chunk is being fitted;
        the fitted state of all cells stays on the device."""
        self._no_detrend_streaming()
        dev = cuda_device(device)
        Xh, yh = self._host2d(X, 'X'), self._host2d(y, 'y')
        if Xh.shape != yh.shape:
            raise ValueError(f'X {tuple(Xh.shape)} and y {tuple(yh.shape)} must have the same shape')
        if Xh.dtype != yh.dtype or yh.dtype not in (torch.float32, torch.float64):
            Xh, yh = Xh.to(torch.float64), yh.to(torch.float64)
        T, C = yh.shape
        self._pre_fit()
        sort_t, mean_t, how = self._fit_tables(index)
        st = engine.alloc_state(yh.dtype, C, dev, sort_t, mean_t, need_x_climo=self._needs_x_climo,
                                need_y_climo=True, with_valid=True)
        chunk = max(8, min(chunk_cells, C))
        with torch.cuda.device(dev):
            comp = torch.cuda.current_stream(dev)
            s_in = torch.cuda.Stream(dev)
            buf_y = [torch.empty((T, chunk), dtype=yh.dtype, device=dev) for _ in range(2)]
            buf_x = [torch.empty((T if self._needs_x_climo else 1, chunk), dtype=yh.dtype, device=dev) for _ in range(2)]
            free = [None, None]
            for k, (c0, c1) in enumerate(self._spans(C, chunk)):
                b, w = k & 1, c1 - c0
                if free[b] is not None:
                    s_in.wait_event(free[b])
                engine.copy2d(buf_y[b][:, :w], yh[:, c0:c1], True, stream=s_in.cuda_stream)
                xrows = T if self._needs_x_climo else 1          # precipitation only needs X for the cell mask
                engine.copy2d(buf_x[b][:xrows, :w], Xh[:xrows, c0:c1], True, stream=s_in.cuda_stream)
                landed = torch.cuda.Event()
                landed.record(s_in)
                comp.wait_event(landed)
                view = st.cells(c0, c1)
                view.valid.copy_(engine.cell_mask(buf_x[b][0, :w]))          # core.py:35-37
                engine.qm_fit_into(view, buf_y[b][:, :w], buf_x[b][:, :w] if self._needs_x_climo else None, how)
                free[b] = torch.cuda.Event()
                free[b].record(comp)
        self._state = st
        self.n_features_in_ = 1
        return self

    def predict_host(self, X, index, out=None, device=None, chunk_cells: int = 16384):
        """predict from a HOST array ``[T, C]`` into a host array (``out`` or a new pinned tensor):
        H2D of chunk k+1, the kernels of chunk k and D2H of chunk k-1 overlap on three streams."""
        if not hasattr(self, '_state'):
            raise NotFittedError(f"This {type(self).__name__} instance is not fitted yet. Call 'fit' with "
                                 "appropriate arguments before using this estimator.")
        if self.timestep == 'daily' and self.return_anoms:
            raise ValueError('shape of climo is not equal to input array')
        self._no_detrend_streaming()
        st = self._state
        dev = st.sorted_state.device
        Xh = self._host2d(X, 'X')
        if Xh.dtype != st.dtype:
            Xh = Xh.to(st.dtype)
        T, C = Xh.shape
        if C != st.n_cells:
            raise ValueError(f'X has {C} cells, the model was fitted on {st.n_cells}')
        if out is None:
            out_h = torch.empty((T, C), dtype=st.dtype, pin_memory=True)
        else:
            out_h = self._host2d(out, 'out')
            if out_h.shape != (T, C) or out_h.dtype != st.dtype:
                raise ValueError('out must be a host array shaped and typed like the result')
        table, nbr = self._predict_tables(index)
        chunk = max(8, min(chunk_cells, C))
        with torch.cuda.device(dev):
            comp = torch.cuda.current_stream(dev)
            s_in, s_out = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
            buf_x = [torch.empty((T, chunk), dtype=st.dtype, device=dev) for _ in range(2)]
            buf_o = [torch.empty((T, chunk), dtype=st.dtype, device=dev) for _ in range(2)]
            x_free, o_free = [None, None], [None, None]
            for k, (c0, c1) in enumerate(self._spans(C, chunk)):
                b, w = k & 1, c1 - c0
                if x_free[b] is not None:
                    s_in.wait_event(x_free[b])
                engine.copy2d(buf_x[b][:, :w], Xh[:, c0:c1], True, stream=s_in.cuda_stream)
                landed = torch.cuda.Event()
                landed.record(s_in)
                comp.wait_event(landed)
                if o_free[b] is not None:
                    comp.wait_event(o_free[b])
                engine.qm_predict(st.cells(c0, c1), buf_x[b][:, :w], table, self._mode,
                                  return_anoms=self.return_anoms, roll_nbr=nbr, out=buf_o[b][:, :w],
                                  cunnane=self._cunnane())
                done = torch.cuda.Event()
                done.record(comp)
                x_free[b] = done
                s_out.wait_event(done)
                engine.copy2d(out_h[:, c0:c1], buf_o[b][:, :w], False, stream=s_out.cuda_stream)
                o_free[b] = torch.cuda.Event()
                o_free[b].record(s_out)
            s_out.synchronize()
        st.check_finite()
        return out_h

    # ------------------------------------------------------------------ per-cell API of the reference
    def fit(self, X, y):
        X, y = self._frames(X, y)
        dev = cuda_device()
        x_t, index, _ = series_to_device(X, dev)
        y_t, _, _ = series_to_device(y, dev)
        if x_t.shape[1] != 1:
            raise ValueError(f'BCSD only supports 1 feature, found {x_t.shape[1]}')
        if y_t.shape[1] != 1:
            raise ValueError('y must have exactly one column')
        if x_t.dtype != y_t.dtype:
            x_t, y_t = x_t.to(torch.float64), y_t.to(torch.float64)
        self.fit_batched(x_t, y_t, index)
        self.check_fit()
        self.y_climo_ = pd.DataFrame(self._state.y_climo.cpu().numpy(), index=self._state.mean_table.keys)
        if self._state.x_climo is not None:
            self._x_climo = pd.DataFrame(self._state.x_climo.cpu().numpy(), index=self._state.mean_table.keys)
        return self

    def predict(self, X):
        if not hasattr(self, '_state'):
            raise NotFittedError(f"This {type(self).__name__} instance is not fitted yet. Call 'fit' with "
                                 "appropriate arguments before using this estimator.")
        X = self._frames(X)
        x_t, index, columns = series_to_device(X, cuda_device())
        if x_t.shape[1] != self.n_features_in_:
            raise ValueError(f'X has {x_t.shape[1]} features, but {self.__class__.__name__} was fitted with '
                             f'{self.n_features_in_} features.')
        if x_t.dtype != self._state.dtype:
            x_t = x_t.to(self._state.dtype)
        out = self.predict_batched(x_t, index, out_dtype=torch.float64)
        self._state.check_finite()
        return pd.DataFrame(out.cpu().numpy(), index=index, columns=columns)


class BcsdPrecipitation(BcsdBase):
    """Classic BCSD model for precipitation (bcsd.py:96-193)."""

    _mode = _lib.MODE_BCSD_P
    _needs_x_climo = False

    def check_fit(self):
        super().check_fit()
        st = self._state
        if self.return_anoms:                                   # bcsd.py:140-141
            yc = st.y_climo if st.valid is None else st.y_climo[:, st.valid.bool()]
            if yc.numel() and bool((yc <= 0).any().item()):
                raise ValueError('Invalid value in target climatology')


class BcsdTemperature(BcsdBase):
    """Classic BCSD model for temperature (bcsd.py:196-289)."""

    _mode = _lib.MODE_BCSD_T
    _needs_x_climo = True
