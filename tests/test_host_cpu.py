"""CPU-only checks of the host side: the C-ABI library loads and exports every symbol that
include/sdb.h declares, the group tables match the oracle, and the product refuses to run
without CUDA (no CPU fallback)."""

import os
import re

import numpy as np
import pandas as pd
import pytest

import oracle
import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope='module')
def lib():
    import __graft_entry__ as ge
    from skdownscale_b200 import _lib
    if not os.path.exists(_lib.LIB_PATH):
        ge.build()
    return _lib.load()


def test_header_symbols_exported(lib):
    from skdownscale_b200 import _lib
    hdr = open(os.path.join(ROOT, 'include', 'sdb.h')).read()
    declared = set(re.findall(r'^\s*(?:int|int64_t|const char\*)\s+(sdb_\w+)\s*\(', hdr, flags=re.M))
    assert declared == set(_lib.SIGNATURES), (declared, set(_lib.SIGNATURES))
    for name in declared:
        assert getattr(lib, name) is not None
    assert lib.sdb_version() >= 100
    assert lib.sdb_max_group_len() == 16384


def test_abi_rejects_bad_arguments(lib):
    from skdownscale_b200 import _lib
    rc = lib.sdb_qm_fit(None, 0, 1, 1, None, None, None, 1, 1, None, 1, None, None, None)
    assert rc == -1 and b'NULL' in lib.sdb_last_error()
    with pytest.raises(_lib.SdbError):
        _lib.check(rc, 'sdb_qm_fit')


def test_round2_entries_reject_bad_arguments(lib):
    """The round-2 entry points validate before touching the device (no GPU needed)."""
    rc = lib.sdb_bcsd_fit_predict(2, None, None, None, 0, 1, 1, 1, None, None, 1, 1, None, None, 1, 1, None, 0, None, None, 1,
                                  None, None, None, None)
    assert rc == -1 and b'NULL' in lib.sdb_last_error()
    assert lib.sdb_series_argsort(None, 0, 1, 1, 1, None, 1, None, None) == -1
    assert lib.sdb_analog_grid_fit(None, 0, 1, 1, 100, 3, None, None, None, None, 1, None, None) == -1
    assert lib.sdb_analog_grid_assign(None, 0, 1, 1, 100, 3, None, None, 1, None, None) == -1
    assert lib.sdb_peer_copy2d(None, 16, None, 16, 16, 1, 0, None) == -1
    assert lib.sdb_peer_alloc(0, None, None) == -1
    # what the pruned analog search covers: float32, 1..3 predictors, k <= 16, the window must fit in shared memory
    assert lib.sdb_analog_pruned_supported(0, 18250, 3, 10) == 1
    assert lib.sdb_analog_pruned_supported(0, 10950, 3, 200) == 0
    assert lib.sdb_analog_pruned_supported(1, 10950, 3, 10) == 0
    assert lib.sdb_analog_pruned_supported(0, 10950, 4, 10) == 0
    assert lib.sdb_analog_pruned_supported(0, 30000, 3, 10) == 0
    assert lib.sdb_analog_grid_boxes() == 512 and lib.sdb_analog_grid_planes() == 511
    assert lib.sdb_series_argsort_max_steps() == 32768
    assert lib.sdb_peer_bcast2d(None, 1, 16, None, 16, 16, 1, 0, None) == -1
    import ctypes
    one = (ctypes.c_void_p * 1)(16)
    assert lib.sdb_peer_bcast2d(one, 9, 16, ctypes.c_void_p(16), 16, 16, 1, 0, None) == -1        # 1..8 destinations
    assert lib.sdb_peer_bcast2d(one, 1, 16, ctypes.c_void_p(16), 16, 24, 1, 0, None) == -1        # pitch < width
    assert lib.sdb_peer_bcast2d(one, 1, 24, ctypes.c_void_p(16), 24, 24, 1, 0, None) == -3        # SDB_E_UNSUPPORTED: 16-byte rows only
    assert lib.sdb_peer_bcast2d(one, 1, 16, ctypes.c_void_p(16), 16, 0, 1, 0, None) == 0          # empty block
    assert lib.sdb_zscore_fit(None, None, 0, 1, 1, None, 1, 365, None, None, 31, 364, None, None, None, None, 1, None, None, None) == -1
    assert b'NULL' in lib.sdb_last_error()
    assert lib.sdb_zscore_predict(None, 0, 1, 1, 10, 31, None, None, 1, 364, None, 0, 1, None, None, None) == -1
    assert lib.sdb_zscore_workspace_bytes(129600, 366) == 4 * 366 * 129600 * 8
    p = ctypes.c_void_p(64)
    # fewer fitted values than min(n_steps, 364): the reference's positional IndexError (zscore.py:314)
    assert lib.sdb_zscore_predict(p, 0, 1, 1, 1000, 30, p, p, 1, 363, p, 0, 1, None, None, None) == -1
    assert b'out-of-bounds' in lib.sdb_last_error()


def test_zscore_host_tables():
    """The calendar tables of ZScoreRegressor.fit (pointwise_models/zscore.py::day_tables) against the oracle's
    restatement of zscore.py:124-193, and the constructor check of zscore.py:27-30."""
    import oracle
    from skdownscale_b200.pointwise_models import ZScoreRegressor
    from skdownscale_b200.pointwise_models.zscore import day_tables
    for start, T, w in (('1999-03-01', 1461, 31), ('2018-01-01', 731, 31), ('1999-06-01', 1100, 30), ('1999-03-01', 1461, 11),
                        ('2001-01-01', 1000, 30), ('1981-01-01', 10950, 31), ('2000-02-10', 500, 1), ('2000-02-10', 500, 2)):
        idx = pd.date_range(start, periods=T)
        rows, pos_col, col_count, n_kept, days = day_tables(idx, w)
        cols = oracle.zscore_window_columns(rows.shape[1], w)
        assert n_kept == len(cols) and len(pos_col) == rows.shape[1] + w
        assert np.array_equal(np.stack([pos_col[k + 1:k + 1 + w] for k in range(n_kept)]), cols)
        assert np.array_equal(rows, oracle.zscore_day_table(idx))
        assert np.array_equal(col_count, (rows >= 0).sum(0)) and col_count.sum() == T
    assert day_tables(pd.date_range('1981-01-01', periods=10950), 31)[3] == 365
    assert day_tables(pd.date_range('2018-01-01', '2020-01-01'), 31)[3] == 364          # the reference's own test record
    with pytest.raises(ValueError, match='window_width must be positive'):
        ZScoreRegressor(window_width=-3)
    from sklearn.base import clone
    assert clone(ZScoreRegressor(window_width=15)).get_params() == {'window_width': 15}


def test_fused_path_is_opt_in_and_scoped():
    import torch
    from skdownscale_b200 import engine
    t = engine.GroupTable([(1, range(900)), (2, range(900, 1800))])
    assert engine.fused_supported(torch.float32, t)
    assert not engine.fused_supported(torch.float64, t)
    assert not engine.fused_supported(torch.float32, engine.GroupTable([(0, range(1025))]))


def test_group_tables_match_oracle():
    from skdownscale_b200.pointwise_models import groupers as g
    idx = synth.daily_index(10950)
    a = g.groups_from_keys(g.grouper_keys(g.MONTH_GROUPER, idx))
    b = oracle.groups_from_keys(oracle.month_keys(idx))
    assert [k for k, _ in a] == [k for k, _ in b]
    for (_, ra), (_, rb) in zip(a, b):
        np.testing.assert_array_equal(ra, rb)
    assert max(len(r) for _, r in a) == 930 and min(len(r) for _, r in a) == 847
    # arbitrary callable grouper goes through Index.map like df.groupby(callable)
    c = g.groups_from_keys(g.grouper_keys(lambda ts: ts.quarter, idx))
    assert [k for k, _ in c] == [1, 2, 3, 4]
    d = g.padded_doy_groups(idx)
    e = oracle.padded_doy_groups(idx)
    for (ka, ra), (kb, rb) in zip(d, e):
        assert ka == kb
        np.testing.assert_array_equal(ra, rb)


def test_padded_doy_grouper_class(golden):
    from skdownscale_b200.pointwise_models import PaddedDOYGrouper
    index = pd.date_range(start='1980-01-01', end='1982-12-31')
    X = pd.DataFrame({'foo': np.arange(len(index), dtype=np.float64)}, index=index)
    groups = dict(list(PaddedDOYGrouper(X)))          # skdownscale/test/test_pointwise_models.py:302-312
    np.testing.assert_array_equal(np.unique(groups[123].index.dayofyear), np.arange(108, 139))
    gref = golden('padded_doy_1980_1982')
    for d in (1, 60, 123, 365, 366):
        np.testing.assert_array_equal(groups[d]['foo'].values.astype(int), gref['rows'][d - 1][:gref['lens'][d - 1]])
    assert PaddedDOYGrouper(X).mean().shape == (366, 1)


def test_rolling_neighbours():
    from skdownscale_b200.pointwise_models import groupers as g
    idx = synth.daily_index(800)
    groups = g.groups_from_keys(g.grouper_keys(g.MONTH_GROUPER, idx))
    nbr = g.rolling_neighbours(groups, 800)
    x = np.random.default_rng(0).standard_normal(800)
    for _, rows in groups:
        want = oracle.bcsd.rolling9_centered(x[rows])
        got = np.array([x[nbr[r][nbr[r] >= 0]].mean() for r in rows])
        np.testing.assert_allclose(got, want, rtol=1e-13)


def test_estimator_surface_and_no_cpu_fallback():
    import torch
    from skdownscale_b200.pointwise_models import (AnalogRegression, BcsdPrecipitation, BcsdTemperature,
                                                   PointWiseDownscaler, PureAnalog, QuantileMapper)
    from sklearn.base import clone
    m = BcsdTemperature(return_anoms=False)
    assert clone(m).get_params()['return_anoms'] is False
    assert PureAnalog.n_outputs == 3 and AnalogRegression.output_names == ['pred', 'exceedance_prob', 'prediction_error']
    with pytest.raises(TypeError):
        PointWiseDownscaler(object())                  # core.py:220-223
    pw = PointWiseDownscaler(BcsdPrecipitation())
    with pytest.raises(ValueError):
        pw.fit(np.zeros((4, 2)), np.zeros((4, 2)), np.zeros((4, 2)))   # core.py:249-250
    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError, match='no CPU fallback'):
            pw.fit(np.zeros((40, 2), np.float32), np.zeros((40, 2), np.float32), time=synth.daily_index(40))
        with pytest.raises(RuntimeError, match='no CPU fallback'):
            QuantileMapper().fit(np.zeros((10, 1)))
    # _pre_fit replaces the constructor argument by the grouper class (test_pointwise_models.py:315-320)
    from skdownscale_b200.pointwise_models import PaddedDOYGrouper
    m = BcsdTemperature(time_grouper='daily_nasa-nex', return_anoms=False)
    m._pre_fit()
    assert issubclass(m.time_grouper, PaddedDOYGrouper)
    with pytest.raises(NotImplementedError):
        BcsdTemperature(time_grouper='daily_nasa-nex', qm_kwargs={'detrend': True})._pre_fit()
    with pytest.raises(TypeError):
        QuantileMapper(qt_kwargs={'gamma': 1}).fit_batched(None)       # CunnaneTransformer has no such argument


def test_next_row_estimators_host_side():
    """Constructor / option validation of the "next" estimators runs on the host (no GPU needed) and follows
    the reference: quantile.py:183-191, 577-592 (n_endpoints, kind), 420-432 (Cunnane keywords), trend.py:31-32."""
    from skdownscale_b200 import _lib
    from skdownscale_b200.pointwise_models import EquidistantCdfMatcher, QuantileMappingReressor
    from skdownscale_b200.pointwise_models.quantile import check_lt_kwargs, cunnane_opts
    from sklearn.base import clone
    with pytest.raises(ValueError, match='n_endpoints'):
        QuantileMappingReressor(n_endpoints=1)
    with pytest.raises(NotImplementedError):
        EquidistantCdfMatcher(kind='product')
    m = EquidistantCdfMatcher(kind='ratio', extrapolate='1to1', n_endpoints=4)
    assert clone(m).get_params() == {'kind': 'ratio', 'extrapolate': '1to1', 'n_endpoints': 4, 'max_ratio': None}
    assert m._kind == _lib.QMR_EDCDF_RATIO and QuantileMappingReressor()._kind == _lib.QMR_REGRESSOR
    assert cunnane_opts(None) is None and cunnane_opts({'alpha': 0.4, 'extrapolate': 'both'}) is None
    o = cunnane_opts({'extrapolate': 'min', 'n_endpoints': 3, 'alpha': 0.1})
    # the reference never hands alpha / beta to plotting_positions (quantile.py:462): always 0.4
    assert (o.alpha, o.beta, o.n_endpoints, o.extrapolate) == (0.4, 0.4, 3, 1)
    assert cunnane_opts({'extrapolate': None}).extrapolate == 0 and cunnane_opts({'extrapolate': '1to1'}).extrapolate == 0
    with pytest.raises(ValueError):
        cunnane_opts({'extrapolate': 'upwards'})
    check_lt_kwargs(None)
    check_lt_kwargs({'lr_kwargs': {'fit_intercept': True}})
    with pytest.raises(TypeError):
        check_lt_kwargs({'degree': 2})
    with pytest.raises(NotImplementedError):
        check_lt_kwargs({'lr_kwargs': {'positive': True}})
