"""ctypes binding of libsdb.so (the C ABI declared in include/sdb.h).

The library is built in-tree by ``__graft_entry__.build()`` / ``csrc/Makefile``.  Loading
fails loudly when it is missing — there is no CPU fallback behind this module.
"""

from __future__ import annotations

import ctypes
import os
from ctypes import c_char_p, c_double, c_int, c_int64, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get('SDB_LIBRARY') or os.path.join(_HERE, 'csrc', 'libsdb.so')   # SDB_LIBRARY: experiment builds

SDB_F32, SDB_F64 = 0, 1
MODE_QM, MODE_BCSD_P, MODE_BCSD_T = 0, 1, 2
MEAN_GROUPBY, MEAN_FRAME, MEAN_NUMPY = 0, 1, 2
EXTRAPOLATE = {None: 0, '1to1': 0, 'min': 1, 'max': 2, 'both': 3}    # SDB_EXTRAPOLATE_* (include/sdb.h)


class CunnaneOpts(ctypes.Structure):
    """``sdb_cunnane_opts`` of include/sdb.h: CunnaneTransformer settings (quantile.py:420-432)."""
    _fields_ = [('alpha', c_double), ('beta', c_double), ('n_endpoints', ctypes.c_int32), ('extrapolate', ctypes.c_int32)]


TREND_REMOVE, TREND_RESTORE = 0, 1
QMR_REGRESSOR, QMR_EDCDF_DIFFERENCE, QMR_EDCDF_RATIO = 0, 1, 2
ANALOG_BEST, ANALOG_SAMPLE, ANALOG_WEIGHT, ANALOG_MEAN, ANALOG_REGRESSION = 0, 1, 2, 3, 4

# every symbol include/sdb.h declares: (restype, argtypes)
SIGNATURES = {
    'sdb_version': (c_int, []),
    'sdb_last_error': (c_char_p, []),
    'sdb_max_group_len': (c_int, []),
    'sdb_set_debug_flags': (c_int, [c_int]),
    'sdb_memcpy2d_async': (c_int, [c_void_p, c_int64, c_void_p, c_int64, c_int64, c_int64, c_int, c_void_p]),
    'sdb_enable_peer_access': (c_int, [c_int, c_int]),
    'sdb_peer_alloc': (c_int, [c_int64, c_void_p, c_void_p]),
    'sdb_peer_open': (c_int, [c_void_p, c_void_p]),
    'sdb_peer_close': (c_int, [c_void_p]),
    'sdb_peer_free': (c_int, [c_void_p]),
    'sdb_peer_bcast2d': (c_int, [c_void_p, c_int, c_int64, c_void_p, c_int64, c_int64, c_int64, c_int, c_void_p]),
    'sdb_peer_copy2d': (c_int, [c_void_p, c_int64, c_void_p, c_int64, c_int64, c_int64, c_int, c_void_p]),
    'sdb_group_mean': (c_int, [c_void_p, c_int, c_int64, c_int64, c_void_p, c_void_p, c_int, c_int,
                               c_int, c_void_p, c_int64, c_void_p, c_void_p, c_void_p]),
    'sdb_qm_fit': (c_int, [c_void_p, c_int, c_int64, c_int64, c_void_p, c_void_p, c_void_p, c_int, c_int,
                           c_void_p, c_int64, c_void_p, c_void_p, c_void_p]),
    'sdb_qm_predict': (c_int, [c_int, c_void_p, c_int, c_int64, c_int64,
                               c_void_p, c_void_p, c_void_p, c_int, c_int,
                               c_void_p, c_void_p, c_int,
                               c_void_p, c_int64,
                               c_void_p, c_void_p, c_int64,
                               c_int, c_void_p, c_void_p,
                               c_void_p, c_int, c_int64, c_void_p,
                               c_void_p, c_void_p, c_void_p]),
    'sdb_bcsd_fit_predict': (c_int, [c_int, c_void_p, c_void_p, c_void_p, c_int,
                                     c_int64, c_int64, c_int64,
                                     c_void_p, c_void_p, c_int, c_int,
                                     c_void_p, c_void_p, c_int64, c_int,
                                     c_void_p, c_int64, c_void_p,
                                     c_void_p, c_int64,
                                     c_void_p, c_void_p, c_void_p, c_void_p]),
    'sdb_analog_predict': (c_int, [c_int, c_void_p, c_void_p, c_void_p,
                                   c_int, c_int64, c_int64,
                                   c_int, c_int, c_int, c_int,
                                   c_int, c_double, c_double, c_void_p,
                                   c_void_p, c_int, c_int64, c_void_p,
                                   c_void_p, c_void_p, c_void_p]),
    'sdb_analog_predict_pruned': (c_int, [c_int, c_void_p, c_void_p, c_void_p,
                                          c_int, c_int64, c_int64,
                                          c_int, c_int, c_int, c_int,
                                          c_int, c_double, c_double, c_void_p,
                                          c_void_p, c_int, c_int64, c_void_p,
                                          c_void_p, c_void_p,
                                          c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_void_p]),
    'sdb_analog_grid_fit': (c_int, [c_void_p, c_int, c_int64, c_int64, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p,
                                    c_int64, c_void_p, c_void_p]),
    'sdb_analog_grid_assign': (c_int, [c_void_p, c_int, c_int64, c_int64, c_int, c_int, c_void_p, c_void_p, c_int64,
                                       c_void_p, c_void_p]),
    'sdb_analog_pruned_supported': (c_int, [c_int, c_int, c_int, c_int]),
    'sdb_analog_grid_boxes': (c_int, []),
    'sdb_analog_grid_planes': (c_int, []),
    'sdb_series_argsort': (c_int, [c_void_p, c_int, c_int64, c_int64, c_int, c_void_p, c_int64, c_void_p, c_void_p]),
    'sdb_series_argsort_max_steps': (c_int, []),
    'sdb_series_rank': (c_int, [c_void_p, c_int, c_int64, c_int64, c_void_p, c_void_p, c_int, c_int,
                                c_int, c_void_p, c_int64, c_void_p, c_void_p, c_void_p]),
    'sdb_qmr_frame': (c_int, [c_void_p, c_void_p, c_int, c_int64, c_int64, c_int, c_int, c_int,
                              c_void_p, c_void_p, c_void_p]),
    'sdb_qmr_predict': (c_int, [c_int, c_void_p, c_int, c_int64, c_int64, c_int,
                                c_void_p, c_void_p, c_int64, c_int,
                                c_void_p, c_int, c_int,
                                c_void_p, c_int64,
                                c_void_p, c_int, c_int64,
                                c_void_p, c_void_p, c_void_p]),
    'sdb_group_trend': (c_int, [c_void_p, c_int, c_int64, c_int64, c_void_p, c_void_p, c_int, c_int,
                                c_void_p, c_void_p, c_int64, c_void_p, c_void_p, c_void_p]),
    'sdb_zscore_workspace_bytes': (c_int64, [c_int64, c_int]),
    'sdb_zscore_fit': (c_int, [c_void_p, c_void_p, c_int, c_int64, c_int64, c_void_p, c_int, c_int, c_void_p, c_void_p, c_int, c_int,
                               c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_void_p, c_void_p, c_void_p]),
    'sdb_zscore_predict': (c_int, [c_void_p, c_int, c_int64, c_int64, c_int, c_int, c_void_p, c_void_p, c_int64, c_int,
                                   c_void_p, c_int, c_int64, c_void_p, c_void_p, c_void_p]),
    'sdb_trend_apply': (c_int, [c_int, c_void_p, c_int, c_int64, c_int64, c_void_p, c_void_p, c_int, c_int,
                                c_void_p, c_void_p, c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_void_p]),
    'sdb_bcsd_shift': (c_int, [c_void_p, c_int, c_int64, c_int64, c_void_p, c_void_p, c_void_p, c_int, c_int,
                               c_void_p, c_void_p, c_int64, c_void_p, c_void_p, c_int64, c_void_p, c_void_p, c_void_p]),
    'sdb_bcsd_combine': (c_int, [c_int, c_void_p, c_void_p, c_int64, c_int64, c_void_p, c_void_p, c_void_p, c_int, c_int,
                                 c_void_p, c_int, c_int64, c_int, c_void_p, c_int, c_int64, c_void_p, c_void_p]),
    'sdb_pure_regression_model_ld': (c_int, []),
    'sdb_pure_regression_fit': (c_int, [c_void_p, c_void_p, c_int, c_int64, c_int64, c_int, c_int, c_int, c_double,
                                        c_double, c_void_p, c_void_p, c_void_p, c_void_p]),
    'sdb_pure_regression_predict': (c_int, [c_void_p, c_int, c_int64, c_int64, c_int, c_int, c_void_p, c_void_p, c_int,
                                            c_int64, c_void_p, c_void_p, c_void_p]),
}

_lib = None


class SdbError(RuntimeError):
    pass


def load() -> ctypes.CDLL:
    """Load libsdb.so once and declare the argument types of every entry point."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f'{LIB_PATH} is missing: build the CUDA library first '
            '(python -c "import __graft_entry__ as g; g.build()" or make -C scikit-downscale_b200/csrc). '
            'skdownscale_b200 has no CPU fallback.')
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)      # AttributeError here = the library does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(code: int, what: str) -> None:
    if code != 0:
        msg = load().sdb_last_error()
        raise SdbError(f'{what} failed ({code}): {msg.decode() if msg else "?"}')
