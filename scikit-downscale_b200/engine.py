"""Batched host-side dispatcher: builds the (tiny, exact, integer) group tables on the host,
owns the device-resident fitted state, and calls the CUDA kernels through the C ABI
(``include/sdb.h``) once per (fit | predict) for ALL cells of a device shard.

This replaces the reference's two Python loops over cells
(skdownscale/pointwise_models/core.py:86-96 and :137-141) and everything they call per
cell.  torch is used only for device memory and streams.
"""

from __future__ import annotations

from dataclasses import dataclass, field

import ctypes
import os

import numpy as np
import torch

from . import _lib

_TORCH_CODE = {torch.float32: _lib.SDB_F32, torch.float64: _lib.SDB_F64}


def _code(t: torch.Tensor) -> int:
    try:
        return _TORCH_CODE[t.dtype]
    except KeyError:
        raise TypeError(f'only float32 / float64 are supported on the device path, got {t.dtype}') from None


def _ptr(t):
    return None if t is None else t.data_ptr()


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _on_device_of_first_tensor(fn):
    """Run an engine entry point with the device of its first CUDA tensor argument current (kernel launches and the
    stream handed to the C ABI then belong to that device — not to whatever device the caller had selected) and
    check that every CUDA tensor argument lives on that one device."""
    import functools

    @functools.wraps(fn)
    def wrapped(*args, **kwargs):
        dev = None
        stack = list(args) + list(kwargs.values())
        while stack:
            v = stack.pop()
            if isinstance(v, torch.Tensor):
                if v.is_cuda:
                    if dev is None:
                        dev = v.device
                    elif v.device != dev:
                        raise RuntimeError(f'{fn.__name__}: operands on different devices ({dev} and {v.device})')
            elif isinstance(v, QMFitted):
                stack.extend(t for t in (v.sorted_state, v.x_climo, v.y_climo, v.valid) if t is not None)
            elif isinstance(v, (tuple, list)):
                stack.extend(v)
        if dev is None:
            return fn(*args, **kwargs)
        with torch.cuda.device(dev):
            return fn(*args, **kwargs)
    return wrapped


def _check_2d(t: torch.Tensor, name: str) -> int:
    if not t.is_cuda:
        raise RuntimeError(f'{name} must be a CUDA tensor (skdownscale_b200 has no CPU path)')
    if t.dim() != 2 or (t.stride(1) != 1 and t.shape[1] != 1):
        raise ValueError(f'{name} must be [time, cell] with the cell axis contiguous, got shape '
                         f'{tuple(t.shape)} strides {t.stride()}')
    ld = t.stride(0) if t.shape[0] > 1 else max(t.shape[1], 1)
    if ld < t.shape[1]:
        raise ValueError(f'{name}: rows overlap (shape {tuple(t.shape)}, strides {t.stride()})')
    return ld


def as_device(a, device, dtype=None) -> torch.Tensor:
    """numpy / torch → CUDA tensor (the H2D copy of the public API)."""
    if isinstance(a, torch.Tensor):
        t = a
    else:
        a = np.asarray(a)
        t = torch.from_numpy(np.ascontiguousarray(a))
    if dtype is not None and t.dtype != dtype:
        t = t.to(dtype)
    return t.to(device, non_blocking=True)


class GroupTable:
    """A set of time groups: ``rows[g, j]`` = row number of the j-th member of group g in the
    order the reference would hand them to numpy, ``-1`` beyond ``len[g]``.  Built on the host
    (exact integer work, done once per call — the reference redoes it per cell)."""

    def __init__(self, groups):
        groups = [(k, np.asarray(r, dtype=np.int64)) for k, r in groups]
        self.keys = [k for k, _ in groups]
        self.key_to_gid = {k: i for i, k in enumerate(self.keys)}
        self.len = np.array([len(r) for _, r in groups], dtype=np.int32)
        self.max_len = int(self.len.max()) if len(groups) else 0
        self.n_groups = len(groups)
        self.rows = np.full((self.n_groups, max(self.max_len, 1)), -1, dtype=np.int32)
        for i, (_, r) in enumerate(groups):
            self.rows[i, :len(r)] = r
        self._dev = {}

    def subset(self, gids) -> 'GroupTable':
        return GroupTable([(self.keys[g], self.rows[g, :self.len[g]]) for g in gids])

    def device(self, device):
        key = str(device)
        if key not in self._dev:
            self._dev[key] = (torch.from_numpy(self.rows).to(device), torch.from_numpy(self.len).to(device))
        return self._dev[key]


@dataclass
class QMFitted:
    """Fitted empirical CDFs + climatologies of every cell of a device shard."""
    dtype: torch.dtype
    n_cells: int
    sort_table: GroupTable                 # groups whose sorted values are held
    state_off: np.ndarray                  # int64 [n sorted groups], offset inside a cell record
    state_ld: int
    sorted_state: torch.Tensor             # [C, state_ld]
    fit_len_dev: torch.Tensor
    state_off_dev: torch.Tensor
    mean_table: GroupTable | None = None   # groups of the climatologies (superset of sort_table keys)
    x_climo: torch.Tensor | None = None    # [n mean groups, C]
    y_climo: torch.Tensor | None = None
    valid: torch.Tensor | None = None      # uint8 [C]
    nonfinite: torch.Tensor | None = None  # int32 [1]
    extra: dict = field(default_factory=dict)

    def cells(self, c0: int, c1: int) -> 'QMFitted':
        """View of the state of cells [c0, c1) (shares storage and the non-finite flag)."""
        return QMFitted(dtype=self.dtype, n_cells=c1 - c0, sort_table=self.sort_table, state_off=self.state_off,
                        state_ld=self.state_ld, sorted_state=self.sorted_state[c0:c1],
                        fit_len_dev=self.fit_len_dev, state_off_dev=self.state_off_dev, mean_table=self.mean_table,
                        x_climo=None if self.x_climo is None else self.x_climo[:, c0:c1],
                        y_climo=None if self.y_climo is None else self.y_climo[:, c0:c1],
                        valid=None if self.valid is None else self.valid[c0:c1], nonfinite=self.nonfinite,
                        extra={})

    def check_finite(self):
        """Raise like the reference's sklearn validation (base.py:18-20) if a kernel met NaN/inf."""
        if self.nonfinite is not None and int(self.nonfinite.item()) != 0:
            self.nonfinite.zero_()
            raise ValueError('Input contains NaN or infinity.')


def cell_mask(first_row: torch.Tensor) -> torch.Tensor:
    """core.py:35-37: a cell takes part iff its first timestep (first feature) is not NaN."""
    return (~torch.isnan(first_row)).to(torch.uint8).contiguous()


@_on_device_of_first_tensor
def group_mean(v: torch.Tensor, table: GroupTable, how: int, valid=None, nonfinite=None, out=None) -> torch.Tensor:
    lib = _lib.load()
    ld = _check_2d(v, 'v')
    C = v.shape[1]
    rows, length = table.device(v.device)
    if out is None:
        out = torch.empty((table.n_groups, C), dtype=v.dtype, device=v.device)
    ld_out = _check_2d(out, 'climo')
    _lib.check(lib.sdb_group_mean(_ptr(v), _code(v), ld, C, _ptr(rows), _ptr(length), table.n_groups,
                                  table.rows.shape[1], how, _ptr(out), ld_out, _ptr(valid), _ptr(nonfinite),
                                  _stream()), 'sdb_group_mean')
    return out


def copy2d(dst: torch.Tensor, src: torch.Tensor, to_device, stream=None) -> None:
    """Asynchronous strided copy of a [rows, cols] block between host and device (both tensors may
    be column slices of wider arrays) — cudaMemcpy2DAsync through the C ABI.  ``to_device``: True = host →
    device, False = device → host, ``'peer'`` = device → device (possibly another GPU's memory mapped by IPC)."""
    lib = _lib.load()
    if dst.shape != src.shape or dst.dim() != 2 or dst.dtype != src.dtype:
        raise ValueError('copy2d needs two [rows, cols] tensors of the same shape and dtype')
    if (dst.stride(1) != 1 and dst.shape[1] != 1) or (src.stride(1) != 1 and src.shape[1] != 1):
        raise ValueError('copy2d needs unit stride along the columns')
    es = dst.element_size()
    rows, cols = dst.shape
    dp = (dst.stride(0) if rows > 1 else cols) * es
    sp = (src.stride(0) if rows > 1 else cols) * es
    _lib.check(lib.sdb_memcpy2d_async(_ptr(dst), dp, _ptr(src), sp, cols * es, rows, 2 if to_device == 'peer' else (0 if to_device else 1),
                                      stream if stream is not None else _stream()), 'sdb_memcpy2d_async')


@_on_device_of_first_tensor
def peer_copy2d(dst: torch.Tensor, src: torch.Tensor, stream=None, n_ctas: int = 32, method: str = 'kernel') -> None:
    """Push a [rows, cols] block into (possibly peer) device memory: ``method='kernel'`` = ``sdb_peer_copy2d``
    (SM loads / stores over NVLink; needs 16-byte aligned rows, else falls back), ``'ce'`` = ``cudaMemcpy2DAsync``."""
    lib = _lib.load()
    if dst.shape != src.shape or dst.dim() != 2 or dst.dtype != src.dtype:
        raise ValueError('peer_copy2d needs two [rows, cols] tensors of the same shape and dtype')
    es = dst.element_size()
    rows, cols = dst.shape
    dp = (dst.stride(0) if rows > 1 else cols) * es
    sp = (src.stride(0) if rows > 1 else cols) * es
    st = stream if stream is not None else _stream()
    aligned = not ((cols * es | dp | sp | dst.data_ptr() | src.data_ptr()) & 15)
    if method == 'kernel' and aligned:
        _lib.check(lib.sdb_peer_copy2d(_ptr(dst), dp, _ptr(src), sp, cols * es, rows, n_ctas, st), 'sdb_peer_copy2d')
    else:
        _lib.check(lib.sdb_memcpy2d_async(_ptr(dst), dp, _ptr(src), sp, cols * es, rows, 2, st), 'sdb_memcpy2d_async')


def alloc_state(dtype, n_cells: int, device, sort_table: GroupTable, mean_table: GroupTable | None = None,
                need_x_climo: bool = False, need_y_climo: bool = True, with_valid: bool = False) -> QMFitted:
    """Device storage of the fitted state of ``n_cells`` cells (filled by :func:`qm_fit_into`)."""
    lib = _lib.load()
    if sort_table.max_len > lib.sdb_max_group_len():
        raise NotImplementedError(f'time groups longer than {lib.sdb_max_group_len()} steps are not supported yet '
                                  f'(got {sort_table.max_len})')
    lens = sort_table.len.astype(np.int64)
    align = int(os.environ.get('SDB_STATE_ALIGN', '4'))    # elements; experiment knob (sector-aligned groups: 8)
    padded = (lens + align - 1) // align * align       # keep every group 16-byte aligned inside a record
    off = np.concatenate(([0], np.cumsum(padded)[:-1])).astype(np.int64)
    state_ld = int(padded.sum())
    mt = mean_table if mean_table is not None else sort_table
    _, length = sort_table.device(device)
    return QMFitted(dtype=dtype, n_cells=n_cells, sort_table=sort_table, state_off=off, state_ld=state_ld,
                    sorted_state=torch.empty((n_cells, state_ld), dtype=dtype, device=device),
                    fit_len_dev=length, state_off_dev=torch.from_numpy(off).to(device), mean_table=mt,
                    x_climo=torch.empty((mt.n_groups, n_cells), dtype=dtype, device=device) if need_x_climo else None,
                    y_climo=torch.empty((mt.n_groups, n_cells), dtype=dtype, device=device) if need_y_climo else None,
                    valid=torch.ones(n_cells, dtype=torch.uint8, device=device) if with_valid else None,
                    nonfinite=torch.zeros(1, dtype=torch.int32, device=device))


@_on_device_of_first_tensor
def qm_fit_into(st: QMFitted, y: torch.Tensor, X: torch.Tensor | None = None,
                mean_how: int = _lib.MEAN_GROUPBY) -> QMFitted:
    """Run the fit kernels for the cells of ``st`` (a whole state or a :meth:`QMFitted.cells` view)."""
    lib = _lib.load()
    ld = _check_2d(y, 'y')
    T, C = y.shape
    if C != st.n_cells or y.dtype != st.dtype:
        raise ValueError('y does not match the state block')
    rows, length = st.sort_table.device(y.device)
    _lib.check(lib.sdb_qm_fit(_ptr(y), _code(y), ld, C, _ptr(rows), _ptr(length), _ptr(st.state_off_dev),
                              st.sort_table.n_groups, st.sort_table.rows.shape[1], _ptr(st.sorted_state),
                              st.state_ld, _ptr(st.valid), _ptr(st.nonfinite), _stream()), 'sdb_qm_fit')
    if st.y_climo is not None:
        group_mean(y, st.mean_table, mean_how, st.valid, st.nonfinite, out=st.y_climo)
    if st.x_climo is not None:
        if X is None or X.shape != y.shape:
            raise ValueError('X and y must have the same shape')
        group_mean(X, st.mean_table, mean_how, st.valid, st.nonfinite, out=st.x_climo)
    return st


@_on_device_of_first_tensor
def qm_fit(y: torch.Tensor, sort_table: GroupTable, *, valid=None, X=None, mean_table=None,
           mean_how: int = _lib.MEAN_GROUPBY, want_y_climo: bool = True) -> QMFitted:
    """fit: sort every (cell, group) of ``y`` and compute the climatologies.

    BcsdTemperature.fit / BcsdPrecipitation.fit / QuantileMapper.fit for all cells
    (bcsd.py:115-147, 197-228; quantile.py:81-107)."""
    _check_2d(y, 'y')
    st = alloc_state(y.dtype, y.shape[1], y.device, sort_table, mean_table, need_x_climo=X is not None,
                     need_y_climo=want_y_climo)
    st.valid = valid
    return qm_fit_into(st, y, X, mean_how)


FUSED_MAX_LEN = 1024


def fused_supported(dtype, table: GroupTable) -> bool:
    """Whether :func:`qm_fit_predict` covers this case (float32, groups of up to 1024 steps)."""
    return dtype == torch.float32 and 0 < table.max_len <= FUSED_MAX_LEN


@_on_device_of_first_tensor
def qm_fit_predict(y: torch.Tensor, X_pred: torch.Tensor, table: GroupTable, mode: int, *, X_train=None,
                   return_anoms: bool = False, valid=None, out=None, keep_state: bool = True, stats=None):
    """fit + predict in one pass when both share the time index (``sdb_bcsd_fit_predict``): the sorted
    training values stay in shared memory.  Returns ``(out, fitted state)``; with ``keep_state=False`` the
    state holds the climatologies only (no sorted values are written).  bcsd.py:115-185, 197-269;
    quantile.py:81-147."""
    lib = _lib.load()
    ld_y = _check_2d(y, 'y')
    ld_p = _check_2d(X_pred, 'X_pred')
    T, C = y.shape
    if X_pred.shape != (T, C):
        raise ValueError('the fused path needs X_pred shaped like y (same time index)')
    if y.dtype != torch.float32 or X_pred.dtype != torch.float32:
        raise TypeError('the fused path is float32 only')
    if X_train is not None:
        if X_train.shape != (T, C) or X_train.dtype != y.dtype or _check_2d(X_train, 'X_train') != ld_y:
            raise ValueError('X_train must be laid out like y')
    dev = y.device
    need_x = mode == _lib.MODE_BCSD_T
    need_y = mode != _lib.MODE_QM
    if need_x and X_train is None:
        raise ValueError('BcsdTemperature needs X_train')
    if keep_state:
        st = alloc_state(y.dtype, C, dev, table, None, need_x_climo=need_x, need_y_climo=need_y)
    else:
        _, length = table.device(dev)
        st = QMFitted(dtype=y.dtype, n_cells=C, sort_table=table, state_off=np.zeros(0, np.int64), state_ld=0,
                      sorted_state=None, fit_len_dev=length, state_off_dev=None, mean_table=table,
                      x_climo=torch.empty((table.n_groups, C), dtype=y.dtype, device=dev) if need_x else None,
                      y_climo=torch.empty((table.n_groups, C), dtype=y.dtype, device=dev) if need_y else None,
                      nonfinite=torch.zeros(1, dtype=torch.int32, device=dev))
    st.valid = valid
    rows, length = table.device(dev)
    if out is None:
        out = torch.empty((T, C), dtype=y.dtype, device=dev)
    ld_out = _check_2d(out, 'out')
    _lib.check(lib.sdb_bcsd_fit_predict(mode, _ptr(X_train) if need_x else None, _ptr(y), _ptr(X_pred), _code(y),
                                        ld_y, ld_p, C, _ptr(rows), _ptr(length), table.n_groups, table.rows.shape[1],
                                        _ptr(st.x_climo), _ptr(st.y_climo), C, int(bool(return_anoms)),
                                        _ptr(st.sorted_state) if keep_state else None, st.state_ld,
                                        _ptr(st.state_off_dev) if keep_state else None,
                                        _ptr(out), ld_out, _ptr(valid), _ptr(st.nonfinite), _ptr(stats), _stream()),
               'sdb_bcsd_fit_predict')
    return out, st


@_on_device_of_first_tensor
def qm_predict(st: QMFitted, X: torch.Tensor, table: GroupTable, mode: int, *, return_anoms: bool = False,
               roll_nbr: np.ndarray | None = None, out_dtype=None, want_rank: bool = False, out=None,
               climo_gid: np.ndarray | None = None, cunnane=None):
    """predict: quantile-map every (cell, group) of ``X`` through the fitted CDFs.

    ``table`` = mapping groups of the prediction index; group g uses the fitted group with
    the same key (``KeyError`` like the reference's dict / ``.loc`` lookup otherwise)."""
    lib = _lib.load()
    ld = _check_2d(X, 'X')
    T, C = X.shape
    if C != st.n_cells:
        raise ValueError(f'X has {C} cells, the model was fitted on {st.n_cells}')
    if X.dtype != st.dtype:
        raise TypeError(f'X is {X.dtype}, the model was fitted on {st.dtype}')
    if table.max_len > lib.sdb_max_group_len():
        raise NotImplementedError(f'time groups longer than {lib.sdb_max_group_len()} steps are not supported yet')
    dev = X.device
    try:
        gid = np.array([st.sort_table.key_to_gid[k] for k in table.keys], dtype=np.int32)
    except KeyError as e:
        raise KeyError(e.args[0]) from None
    # climatologies are indexed by the MEAN table; re-index them by sorted group when they differ
    x_climo, y_climo = st.x_climo, st.y_climo
    if st.mean_table is not st.sort_table and (x_climo is not None or y_climo is not None):
        sel = torch.as_tensor([st.mean_table.key_to_gid[k] for k in st.sort_table.keys], device=dev)
        key = 'climo_by_sort'
        if key not in st.extra:
            st.extra[key] = (None if x_climo is None else x_climo.index_select(0, sel).contiguous(),
                             None if y_climo is None else y_climo.index_select(0, sel).contiguous())
        x_climo, y_climo = st.extra[key]
    ld_climo = C
    for cl in (x_climo, y_climo):
        if cl is not None:
            ld_climo = _check_2d(cl, 'climo')
    gid_dev = torch.from_numpy(gid).to(dev)
    rows, length = table.device(dev)
    od = out_dtype or X.dtype
    if out is None:
        out = torch.empty((T, C), dtype=od, device=dev)
    ld_out = _check_2d(out, 'out')
    rank = torch.zeros((T, C), dtype=torch.int32, device=dev) if want_rank else None
    if rank is not None and ld_out != C:
        raise ValueError('want_rank needs a contiguous output')
    nbr_dev = None
    if roll_nbr is not None:
        nbr_dev = torch.from_numpy(np.ascontiguousarray(roll_nbr, dtype=np.int32)).to(dev)
    _lib.check(lib.sdb_qm_predict(mode, _ptr(X), _code(X), ld, C,
                                  _ptr(rows), _ptr(length), _ptr(gid_dev), table.n_groups, table.rows.shape[1],
                                  _ptr(st.fit_len_dev), _ptr(st.state_off_dev), st.sort_table.max_len,
                                  _ptr(st.sorted_state), st.state_ld,
                                  _ptr(x_climo), _ptr(y_climo), ld_climo,
                                  int(bool(return_anoms)), _ptr(nbr_dev),
                                  ctypes.byref(cunnane) if cunnane is not None else None,
                                  _ptr(out), _TORCH_CODE[out.dtype], ld_out, _ptr(rank),
                                  _ptr(st.valid), _ptr(st.nonfinite), _stream()), 'sdb_qm_predict')
    return (out, rank) if want_rank else out


@_on_device_of_first_tensor
def series_argsort(x: torch.Tensor, row_stride: int, n_steps: int, n_cells: int, valid=None) -> torch.Tensor:
    """Per cell, the time steps in ascending order of ``x[t * row_stride + c]`` → int32 ``[n_steps, n_cells]``."""
    lib = _lib.load()
    order = torch.empty((n_steps, n_cells), dtype=torch.int32, device=x.device)
    _lib.check(lib.sdb_series_argsort(_ptr(x), _code(x), row_stride, n_cells, n_steps, _ptr(order), n_cells, _ptr(valid),
                                      _stream()), 'sdb_series_argsort')
    return order


ANALOG_GRID_MIN_STEPS = 2048        # below this the whole window is cheaper to scan than to index


def analog_grid_supported(X_train: torch.Tensor, k: int) -> bool:
    lib = _lib.load()
    T, p = X_train.shape[0], X_train.shape[1]
    return (X_train.dtype == torch.float32 and T >= ANALOG_GRID_MIN_STEPS and os.environ.get('SDB_ANALOG_PRUNE', '1') != '0'
            and bool(lib.sdb_analog_pruned_supported(_lib.SDB_F32, T, p, k)))


@_on_device_of_first_tensor
def analog_grid_fit(X_train: torch.Tensor, valid=None):
    """The spatial index of every cell's training window (the reference's KDTree build, gard.py:82) →
    ``(perm_train [T, C], box_start [65, C], bounds [63, C])``."""
    lib = _lib.load()
    T, p, C = X_train.shape
    dev = X_train.device
    work = torch.empty((T, C), dtype=torch.int32, device=dev)
    perm = torch.empty((T, C), dtype=torch.int32, device=dev)
    start = torch.zeros((lib.sdb_analog_grid_boxes() + 1, C), dtype=torch.int32, device=dev)
    bounds = torch.zeros((lib.sdb_analog_grid_planes(), C), dtype=torch.float32, device=dev)
    _lib.check(lib.sdb_analog_grid_fit(_ptr(X_train), _code(X_train), C, C, T, p, _ptr(work), _ptr(bounds), _ptr(perm),
                                       _ptr(start), C, _ptr(valid), _stream()), 'sdb_analog_grid_fit')
    return perm, start, bounds


@_on_device_of_first_tensor
def analog_predict(kind: int, X_train: torch.Tensor, y_train: torch.Tensor, X_query: torch.Tensor, k: int, *,
                   thresh=None, rand_idx=None, out_dtype=None, want_idx: bool = False, valid=None,
                   nonfinite=None, logistic_C: float = 1.0, prune: bool | None = None, grid=None):
    """PureAnalog / AnalogRegression fit+predict for all cells (gard.py:58-87, 152-224, 273-364).

    X_train [T, p, C], y_train [T, C], X_query [Tq, p, C] → out [Tq, 3, C]."""
    lib = _lib.load()
    for name, t in (('X_train', X_train), ('y_train', y_train), ('X_query', X_query)):
        if not t.is_cuda:
            raise RuntimeError(f'{name} must be a CUDA tensor (skdownscale_b200 has no CPU path)')
    X_train = X_train.contiguous()
    X_query = X_query.contiguous()
    y_train = y_train.contiguous()
    T, p, C = X_train.shape
    Tq = X_query.shape[0]
    if X_query.shape[1:] != (p, C) or y_train.shape != (T, C):
        raise ValueError('shape mismatch between X_train, y_train and X_query')
    if X_query.dtype != X_train.dtype or y_train.dtype != X_train.dtype:
        raise TypeError('X_train, y_train and X_query must share one dtype')
    dev = X_train.device
    od = out_dtype or X_train.dtype
    out = torch.empty((Tq, 3, C), dtype=od, device=dev)
    idx = torch.empty((Tq, k, C), dtype=torch.int32, device=dev) if want_idx else None
    if rand_idx is not None:
        rand_idx = torch.as_tensor(np.ascontiguousarray(rand_idx, dtype=np.int32)).to(dev)
    # exact pruning: the training rows and the query steps grouped by the boxes of the quantile grid
    can_prune = analog_grid_supported(X_train, k)
    if prune is None:
        prune = can_prune
    if prune and not can_prune:
        raise ValueError('the pruned analog search covers float32 windows of at least %d steps, 1..3 predictors, k <= 16'
                         % ANALOG_GRID_MIN_STEPS)
    args = (kind, _ptr(X_train), _ptr(y_train), _ptr(X_query), _code(X_train), C, C, T, Tq, p, k, int(thresh is not None),
            float(thresh) if thresh is not None else 0.0, float(logistic_C), _ptr(rand_idx),
            _ptr(out), _TORCH_CODE[od], C, _ptr(idx), _ptr(valid), _ptr(nonfinite))
    if prune:
        perm_t, start, bounds = grid if grid is not None else analog_grid_fit(X_train, valid)
        perm_q = torch.empty((Tq, C), dtype=torch.int32, device=dev)
        _lib.check(lib.sdb_analog_grid_assign(_ptr(X_query), _code(X_query), C, C, Tq, p, _ptr(bounds), _ptr(perm_q), C,
                                              _ptr(valid), _stream()), 'sdb_analog_grid_assign')
        _lib.check(lib.sdb_analog_predict_pruned(*args, _ptr(perm_t), _ptr(perm_q), _ptr(start), _ptr(bounds), C, _stream()),
                   'sdb_analog_predict_pruned')
    else:
        _lib.check(lib.sdb_analog_predict(*args, _stream()), 'sdb_analog_predict')
    return (out, idx) if want_idx else out


# ---------------------------------------------------------------------------------------------
# CDF-to-CDF regressors (QuantileMappingReressor / EquidistantCdfMatcher, quantile.py:160-395, 556-636)
# ---------------------------------------------------------------------------------------------
@_on_device_of_first_tensor
def series_rank(X: torch.Tensor, table: GroupTable, *, ordinal: bool, valid=None, nonfinite=None) -> torch.Tensor:
    """1-based rank of every value inside its (cell, group) series → int32 ``[T, C]``."""
    lib = _lib.load()
    ld = _check_2d(X, 'X')
    T, C = X.shape
    if table.max_len > lib.sdb_max_group_len():
        raise NotImplementedError(f'series longer than {lib.sdb_max_group_len()} steps are not supported yet')
    rows, length = table.device(X.device)
    rank = torch.zeros((T, C), dtype=torch.int32, device=X.device)
    _lib.check(lib.sdb_series_rank(_ptr(X), _code(X), ld, C, _ptr(rows), _ptr(length), table.n_groups,
                                   table.rows.shape[1], int(ordinal), _ptr(rank), C, _ptr(valid), _ptr(nonfinite),
                                   _stream()), 'sdb_series_rank')
    return rank


@_on_device_of_first_tensor
def qmr_frame(sx: QMFitted, sy: QMFitted, extrapolate: int, n_endpoints: int) -> torch.Tensor:
    """Synthetic frame points of the fitted X / y CDFs of every cell → float64 ``[C, 4]``."""
    lib = _lib.load()
    n_fit = int(sx.sort_table.max_len)
    frame = torch.empty((sx.n_cells, 4), dtype=torch.float64, device=sx.sorted_state.device)
    _lib.check(lib.sdb_qmr_frame(_ptr(sx.sorted_state), _ptr(sy.sorted_state), _code(sx.sorted_state), sx.state_ld,
                                 sx.n_cells, n_fit, extrapolate, n_endpoints, _ptr(frame), _ptr(sx.valid), _stream()),
               'sdb_qmr_frame')
    return frame


@_on_device_of_first_tensor
def qmr_predict(kind: int, X: torch.Tensor, sx: QMFitted, sy: QMFitted, frame: torch.Tensor, extrapolate: int,
                one_to_one: bool, rank: torch.Tensor | None = None) -> torch.Tensor:
    lib = _lib.load()
    ld = _check_2d(X, 'X')
    T, C = X.shape
    if C != sx.n_cells:
        raise ValueError(f'X has {C} cells, the model was fitted on {sx.n_cells}')
    if X.dtype != sx.dtype:
        raise TypeError(f'X is {X.dtype}, the model was fitted on {sx.dtype}')
    out = torch.empty((T, C), dtype=X.dtype, device=X.device)
    _lib.check(lib.sdb_qmr_predict(kind, _ptr(X), _code(X), ld, C, T, _ptr(sx.sorted_state), _ptr(sy.sorted_state),
                                   sx.state_ld, int(sx.sort_table.max_len), _ptr(frame), extrapolate, int(one_to_one),
                                   _ptr(rank), C, _ptr(out), _code(X), C, _ptr(sx.valid), _ptr(sx.nonfinite), _stream()),
               'sdb_qmr_predict')
    return out


# ---------------------------------------------------------------------------------------------
# Detrending quantile map (QuantileMapper(detrend=True), quantile.py:94-98, 127-145)
# ---------------------------------------------------------------------------------------------
@_on_device_of_first_tensor
def group_trend(v: torch.Tensor, table: GroupTable, valid=None, nonfinite=None):
    """LinearTrendTransformer.fit of every (cell, group) → (slope, intercept) float64 ``[G, C]``."""
    lib = _lib.load()
    ld = _check_2d(v, 'v')
    T, C = v.shape
    rows, length = table.device(v.device)
    slope = torch.empty((table.n_groups, C), dtype=torch.float64, device=v.device)
    icpt = torch.empty_like(slope)
    _lib.check(lib.sdb_group_trend(_ptr(v), _code(v), ld, C, _ptr(rows), _ptr(length), table.n_groups,
                                   table.rows.shape[1], _ptr(slope), _ptr(icpt), C, _ptr(valid), _ptr(nonfinite),
                                   _stream()), 'sdb_group_trend')
    return slope, icpt


@_on_device_of_first_tensor
def trend_apply(mode: int, v: torch.Tensor, table: GroupTable, slope, icpt, icpt_ref=None, valid=None) -> torch.Tensor:
    """Remove (``_lib.TREND_REMOVE``) or restore (``_lib.TREND_RESTORE``) the group trend lines → float64."""
    lib = _lib.load()
    ld = _check_2d(v, 'v')
    T, C = v.shape
    rows, length = table.device(v.device)
    out = torch.empty((T, C), dtype=torch.float64, device=v.device)
    _lib.check(lib.sdb_trend_apply(mode, _ptr(v), _code(v), ld, C, _ptr(rows), _ptr(length), table.n_groups,
                                   table.rows.shape[1], _ptr(slope), _ptr(icpt), _ptr(icpt_ref), C, _ptr(out), C,
                                   _ptr(valid), _stream()), 'sdb_trend_apply')
    return out


@_on_device_of_first_tensor
def zscore_fit(X: torch.Tensor, y: torch.Tensor, day_rows: np.ndarray, pos_col: np.ndarray, col_count: np.ndarray,
               window: int, n_kept: int, valid=None, flag=None, want_stats: bool = False):
    """ZScoreRegressor.fit for every cell (``sdb_zscore_fit``): ``shift``, ``scale`` ``[n_kept, C]`` in X's dtype
    (+ the four fitted statistics ``[4, n_kept, C]``)."""
    lib = _lib.load()
    ld = _check_2d(X, 'X')
    if y.shape != X.shape or y.dtype != X.dtype or _check_2d(y, 'y') != ld:
        raise ValueError('zscore_fit needs X and y of one shape, dtype and row stride')
    T, C = X.shape
    n_years, n_days = day_rows.shape
    dev = X.device
    tabs = [torch.from_numpy(np.ascontiguousarray(a, dtype=np.int32)).to(dev) for a in (day_rows, pos_col, col_count)]
    ws = torch.empty(int(lib.sdb_zscore_workspace_bytes(C, n_days)) // 8, dtype=torch.float64, device=dev)
    shift = torch.empty((n_kept, C), dtype=X.dtype, device=dev)
    scale = torch.empty((n_kept, C), dtype=X.dtype, device=dev)
    stats = torch.empty((4, n_kept, C), dtype=X.dtype, device=dev) if want_stats else None
    _lib.check(lib.sdb_zscore_fit(_ptr(X), _ptr(y), _code(X), ld, C, _ptr(tabs[0]), n_years, n_days, _ptr(tabs[1]), _ptr(tabs[2]),
                                  int(window), int(n_kept), _ptr(ws), _ptr(shift), _ptr(scale), _ptr(stats), C,
                                  _ptr(valid), _ptr(flag), _stream()), 'sdb_zscore_fit')
    return shift, scale, stats


@_on_device_of_first_tensor
def zscore_predict(X: torch.Tensor, shift: torch.Tensor, scale: torch.Tensor, window: int, out_dtype=None, valid=None,
                   flag=None, out=None) -> torch.Tensor:
    """ZScoreRegressor.predict for every cell (``sdb_zscore_predict``)."""
    lib = _lib.load()
    ld = _check_2d(X, 'X')
    T, C = X.shape
    if shift.shape != scale.shape or shift.shape[1] != C or shift.dtype != X.dtype or scale.dtype != X.dtype:
        raise ValueError('zscore_predict: shift / scale must be [n, cells] arrays of X\'s dtype')
    out_dtype = out_dtype or X.dtype
    if out is None:
        out = torch.empty((T, C), dtype=out_dtype, device=X.device)
    ld_out = _check_2d(out, 'out')
    _lib.check(lib.sdb_zscore_predict(_ptr(X), _code(X), ld, C, T, int(window), _ptr(shift), _ptr(scale), C, shift.shape[0],
                                      _ptr(out), _code(out), ld_out, _ptr(valid), _ptr(flag), _stream()), 'sdb_zscore_predict')
    return out


def _fit_gids(st: QMFitted, table: GroupTable, device) -> torch.Tensor:
    try:
        gid = np.array([st.sort_table.key_to_gid[k] for k in table.keys], dtype=np.int32)
    except KeyError as e:
        raise KeyError(e.args[0]) from None
    return torch.from_numpy(gid).to(device)


def _climo_by_sort(st: QMFitted, device):
    x_climo, y_climo = st.x_climo, st.y_climo
    if st.mean_table is not st.sort_table and (x_climo is not None or y_climo is not None):
        sel = torch.as_tensor([st.mean_table.key_to_gid[k] for k in st.sort_table.keys], device=device)
        x_climo = None if x_climo is None else x_climo.index_select(0, sel).contiguous()
        y_climo = None if y_climo is None else y_climo.index_select(0, sel).contiguous()
    return x_climo, y_climo


@_on_device_of_first_tensor
def qm_predict_detrended(st_raw: QMFitted, st_res: QMFitted, icpt_fit: torch.Tensor, X: torch.Tensor, table: GroupTable,
                         mode: int, *, return_anoms: bool = False, roll_nbr=None, out_dtype=None, cunnane=None):
    """predict with detrending mappers: (shift →) remove the new data's trend → map the residuals through
    the fitted residual CDFs → restore the trend at the fitted baseline (→ combine).  ``st_raw`` holds the
    climatologies (and the mask), ``st_res`` the sorted float64 residuals, ``icpt_fit`` the fitted
    intercepts ``[fitted group, C]``."""
    lib = _lib.load()
    ld = _check_2d(X, 'X')
    T, C = X.shape
    if C != st_raw.n_cells:
        raise ValueError(f'X has {C} cells, the model was fitted on {st_raw.n_cells}')
    raw_dtype = st_res.extra.get('raw_dtype', st_raw.dtype)
    if X.dtype != raw_dtype:
        # sdb_bcsd_shift reads x_climo in X's element type: a float64 X against float32 climatologies would read
        # out of bounds.  The reference accepts a predict dtype that differs from fit's — cast like predict() does.
        X = X.to(raw_dtype)
        ld = _check_2d(X, 'X')
    dev = X.device
    rows, length = table.device(dev)
    gid = _fit_gids(st_res, table, dev)
    x_climo, y_climo = _climo_by_sort(st_raw, dev)
    valid, flag = st_raw.valid, st_raw.nonfinite
    shift = None
    key = X
    if mode == _lib.MODE_BCSD_T:
        shift = torch.empty((T, C), dtype=torch.float64, device=dev)
        key = torch.empty((T, C), dtype=torch.float64, device=dev)
        nbr_dev = None if roll_nbr is None else torch.from_numpy(np.ascontiguousarray(roll_nbr, dtype=np.int32)).to(dev)
        _lib.check(lib.sdb_bcsd_shift(_ptr(X), _code(X), ld, C, _ptr(rows), _ptr(length), _ptr(gid), table.n_groups,
                                      table.rows.shape[1], _ptr(nbr_dev), _ptr(x_climo), C, _ptr(shift), _ptr(key), C,
                                      _ptr(valid), _ptr(flag), _stream()), 'sdb_bcsd_shift')
    slope, icpt = group_trend(key, table, valid, flag)
    resid = trend_apply(_lib.TREND_REMOVE, key, table, slope, icpt, valid=valid)
    mapped = qm_predict(st_res, resid, table, _lib.MODE_QM, out_dtype=torch.float64, cunnane=cunnane)
    icpt_ref = icpt_fit.index_select(0, gid.long()).contiguous()
    restored = trend_apply(_lib.TREND_RESTORE, mapped, table, slope, icpt, icpt_ref, valid=valid)
    od = out_dtype or X.dtype
    out = torch.empty((T, C), dtype=od, device=dev)
    climo = y_climo if (return_anoms and mode != _lib.MODE_QM) else None
    _lib.check(lib.sdb_bcsd_combine(mode, _ptr(restored), _ptr(shift), C, C, _ptr(rows), _ptr(length), _ptr(gid),
                                    table.n_groups, table.rows.shape[1], _ptr(climo),
                                    _code(climo) if climo is not None else _lib.SDB_F32, C, int(bool(return_anoms)),
                                    _ptr(out), _TORCH_CODE[od], C, _ptr(valid), _stream()), 'sdb_bcsd_combine')
    return out


# ---------------------------------------------------------------------------------------------
# PureRegression (gard.py:367-504)
# ---------------------------------------------------------------------------------------------
@_on_device_of_first_tensor
def pure_regression_fit(X_train: torch.Tensor, y_train: torch.Tensor, *, thresh=None, logistic_C: float = 1.0,
                        valid=None, nonfinite=None) -> torch.Tensor:
    """One OLS (+ logistic exceedance model) per cell → model ``[C, model_ld]`` float64 (library-private layout)."""
    lib = _lib.load()
    X_train, y_train = X_train.contiguous(), y_train.contiguous()
    T, p, C = X_train.shape
    if y_train.shape != (T, C) or y_train.dtype != X_train.dtype:
        raise ValueError('X_train [T, p, C] and y_train [T, C] must agree in shape and dtype')
    model = torch.empty((C, lib.sdb_pure_regression_model_ld()), dtype=torch.float64, device=X_train.device)
    _lib.check(lib.sdb_pure_regression_fit(_ptr(X_train), _ptr(y_train), _code(X_train), C, C, T, p,
                                           int(thresh is not None), float(thresh) if thresh is not None else 0.0,
                                           float(logistic_C), _ptr(model), _ptr(valid), _ptr(nonfinite), _stream()),
               'sdb_pure_regression_fit')
    return model


@_on_device_of_first_tensor
def pure_regression_predict(model: torch.Tensor, X_query: torch.Tensor, *, out_dtype=None, valid=None,
                            nonfinite=None) -> torch.Tensor:
    lib = _lib.load()
    X_query = X_query.contiguous()
    Tq, p, C = X_query.shape
    od = out_dtype or X_query.dtype
    out = torch.empty((Tq, 3, C), dtype=od, device=X_query.device)
    step = 65535                                     # query steps per launch (grid.y)
    for q0 in range(0, Tq, step):
        q1 = min(Tq, q0 + step)
        _lib.check(lib.sdb_pure_regression_predict(_ptr(X_query[q0:q1]), _code(X_query), C, C, q1 - q0, p, _ptr(model),
                                                   _ptr(out[q0:q1]), _TORCH_CODE[od], C, _ptr(valid), _ptr(nonfinite),
                                                   _stream()), 'sdb_pure_regression_predict')
    return out
