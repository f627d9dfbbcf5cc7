"""The numpy oracle (oracle/) against the golden vectors produced by the LIVE reference
(tests/golden/make_golden.py).  CPU only.  Tolerances: the oracle restates float64
arithmetic, so it must agree with the reference far tighter than the 1e-5 product bar."""

import numpy as np
import pandas as pd
import pytest

import oracle
import synth


def _close(a, b, rtol=1e-12, atol=1e-12):
    np.testing.assert_allclose(a, b, rtol=rtol, atol=atol, equal_nan=True)


def test_qm_known_answer(golden):
    g = golden('qm_known_answer')            # skdownscale/test/test_pointwise_models.py:81-90
    st = oracle.quantile_mapper_fit(g['fit'][:, 0])
    out = oracle.quantile_mapper_transform(g['x'][:, 0], st)
    np.testing.assert_almost_equal(out, g['fit'][:, 0])
    _close(out, g['out'][:, 0])


QT_VARIANTS = {          # the settings tests/golden/make_golden.py used for the qm_qt_* files
    'ab': dict(alpha=0.3, beta=0.5, n_endpoints=5),
    'none': dict(extrapolate=None),
    'min': dict(extrapolate='min', n_endpoints=4),
    'max': dict(extrapolate='max', alpha=0.0, beta=1.0),
    '1to1': dict(extrapolate='1to1'),
}


@pytest.mark.parametrize('tag', sorted(QT_VARIANTS))
def test_qm_cunnane_settings(golden, tag):
    """QuantileMapper(qt_kwargs=...) — non-default CunnaneTransformer settings (quantile.py:420-432)."""
    g = golden(f'qm_qt_{tag}')
    for c in range(g['Xp'].shape[1]):
        st = oracle.quantile_mapper_fit(g['ytr'][:, c])
        _close(oracle.quantile_mapper_transform(g['Xp'][:, c], st, **QT_VARIANTS[tag]), g['out'][:, c],
               rtol=1e-13, atol=1e-13)


@pytest.mark.parametrize('name', ['qm_equal_len', 'qm_pred_longer', 'qm_pred_shorter', 'qm_ties',
                                  'qm_f64', 'qm_tiny'])
def test_qm_cases(golden, name):
    g = golden(name)
    out = oracle.pointwise_fit_predict({'name': 'QuantileMapper'}, None, g['ytr'], g['Xp'])
    _close(out, g['out'].astype(g['Xp'].dtype))
    for c in range(g['Xp'].shape[1]):
        st = oracle.quantile_mapper_fit(g['ytr'][:, c])
        _close(oracle.quantile_mapper_transform(g['Xp'][:, c], st), g['out'][:, c], rtol=1e-13, atol=1e-13)


def test_plotting_positions_and_rank():
    pp = oracle.plotting_positions(5)
    _close(pp, (np.arange(1, 6) - 0.4) / (((5 + 1.0) - 0.4) - 0.4), rtol=0, atol=0)   # NOT 5.2: evaluation order matters
    r = oracle.rank_max_ties(np.array([3.0, 1.0, 3.0, 2.0, 1.0]))
    assert r.tolist() == [5, 2, 5, 3, 2]


def test_padded_doy_grouper(golden):
    g = golden('padded_doy_1980_1982')       # test_pointwise_models.py:302-312 + full table
    index = pd.date_range(start='1980-01-01', end='1982-12-31')
    groups = oracle.padded_doy_groups(index)
    assert len(groups) == 366
    for (key, rows), ref_rows, n in zip(groups, g['rows'], g['lens']):
        assert len(rows) == n
        np.testing.assert_array_equal(rows, ref_rows[:n])
    doy = np.asarray(index.dayofyear)
    np.testing.assert_array_equal(np.unique(doy[groups[122][1]]), np.arange(108, 139))
    # the irregular DOY-366 group: non-leap rows are DOYs 351..365 and 2..15
    rows366 = groups[365][1]
    nl = rows366[~np.asarray(index.is_leap_year)[rows366]]
    assert sorted(set(doy[nl].tolist())) == list(range(2, 16)) + list(range(351, 366))


@pytest.mark.parametrize('name,kw', [
    ('bcsd_t_month_anoms', {}),
    ('bcsd_t_month_abs', {'return_anoms': False}),
    ('bcsd_t_month_future', {}),
    ('bcsd_t_month_future_qt', {'qt_kwargs': dict(alpha=0.3, beta=0.5, n_endpoints=5, extrapolate='max')}),
    ('bcsd_t_month_detrend', {'detrend': True}),
    ('bcsd_t_month_detrend_future', {'detrend': True, 'return_anoms': False}),
    ('bcsd_t_month_f64', {}),
    ('bcsd_t_month_30yr', {}),
    ('bcsd_t_nasanex', {'time_grouper': 'daily_nasa-nex', 'return_anoms': False}),
])
def test_bcsd_temperature(golden, name, kw):
    g = golden(name)
    idx_f = synth.daily_index(len(g['Xtr']), str(g['start_fit']))
    idx_p = synth.daily_index(len(g['Xp']), str(g['start_pred']))
    spec = {'name': 'BcsdTemperature', 'time_grouper': 'month', **kw}
    out = oracle.pointwise_fit_predict(spec, g['Xtr'], g['ytr'], g['Xp'], idx_f, idx_p)
    assert out.dtype == g['Xp'].dtype
    ref = g['out64']
    scale = np.nanstd(g['ytr'])
    assert np.nanmax(np.abs(out.astype(np.float64) - ref.astype(g['Xp'].dtype))) <= 1e-6 * scale
    # float64 agreement of the unrounded result, cell by cell
    fit_groups, roll_groups, qm_groups = oracle.wrapper._bcsd_groups(spec['time_grouper'], idx_f, idx_p)
    for c in range(g['Xp'].shape[1]):
        if np.isnan(g['Xtr'][0, c]):
            assert np.isnan(out[:, c]).all() and np.isnan(g['out'][:, c]).all()
            continue
        how = 'frame' if spec['time_grouper'] == 'daily_nasa-nex' else 'groupby'
        st = oracle.bcsd_temperature_fit(g['Xtr'][:, c], g['ytr'][:, c], fit_groups, how,
                                         detrend=kw.get('detrend', False))
        o = oracle.bcsd_temperature_predict(st, g['Xp'][:, c], roll_groups, qm_groups,
                                            kw.get('return_anoms', True), qt=kw.get('qt_kwargs'))
        _close(o, ref[:, c], rtol=0, atol=2e-12 * max(1.0, scale))


@pytest.mark.parametrize('name,kw', [
    ('bcsd_p_month_anoms', {}),
    ('bcsd_p_month_30yr', {}),
    ('bcsd_p_month_detrend', {'detrend': True}),
    ('bcsd_p_month_abs_future', {'return_anoms': False}),
    ('bcsd_p_nasanex', {'time_grouper': 'daily_nasa-nex', 'return_anoms': False}),
])
def test_bcsd_precipitation(golden, name, kw):
    g = golden(name)
    idx_f = synth.daily_index(len(g['Xtr']), str(g['start_fit']))
    idx_p = synth.daily_index(len(g['Xp']), str(g['start_pred']))
    spec = {'name': 'BcsdPrecipitation', 'time_grouper': 'month', **kw}
    out = oracle.pointwise_fit_predict(spec, g['Xtr'], g['ytr'], g['Xp'], idx_f, idx_p)
    _close(out, g['out'], rtol=1e-6, atol=0)
    fit_groups, _, qm_groups = oracle.wrapper._bcsd_groups(spec['time_grouper'], idx_f, idx_p)
    for c in range(g['Xp'].shape[1]):
        how = 'frame' if spec['time_grouper'] == 'daily_nasa-nex' else 'groupby'
        st = oracle.bcsd_precipitation_fit(g['ytr'][:, c], fit_groups, kw.get('return_anoms', True), how,
                                           detrend=kw.get('detrend', False))
        o = oracle.bcsd_precipitation_predict(st, g['Xp'][:, c], qm_groups, kw.get('return_anoms', True))
        _close(o, g['out64'][:, c], rtol=1e-12, atol=1e-12)


def test_bcsd_precipitation_bad_climatology():
    idx = synth.daily_index(400)
    y = np.zeros(400, dtype=np.float32)
    groups = oracle.groups_from_keys(oracle.month_keys(idx))
    with pytest.raises(ValueError, match='Invalid value in target climatology'):
        oracle.bcsd_precipitation_fit(y, groups, True)


@pytest.mark.parametrize('kind', ['best_analog', 'mean_analogs', 'weight_analogs', 'sample_analogs'])
@pytest.mark.parametrize('suffix,thresh', [('', None), ('_thresh', 0.0)])
def test_pure_analog(golden, kind, suffix, thresh):
    g = golden(f'pure_{kind}{suffix}')
    spec = {'name': 'PureAnalog', 'n_analogs': 10, 'kind': kind, 'thresh': thresh}
    C = g['Xq'].shape[-1]
    for c in range(C):
        rand = g['rand'][:, c] if 'rand' in g else None
        o = oracle.pure_analog_predict(g['Xtr'][..., c], g['ytr'][:, c], g['Xq'][..., c], 10, kind, thresh, rand)
        _close(o.astype(np.float32), g['out'][:, :, c], rtol=2e-6, atol=1e-7)
    if 'rand' not in g:
        out = oracle.pointwise_fit_predict(spec, g['Xtr'], g['ytr'], g['Xq'])
        _close(out, g['out'], rtol=2e-6, atol=1e-7)


def test_pure_analog_k200(golden):
    g = golden('pure_mean_analogs_k200')
    o = oracle.pure_analog_predict(g['Xtr'][..., 0], g['ytr'][:, 0], g['Xq'][..., 0], 200, 'mean_analogs')
    _close(o.astype(np.float32), g['out'][:, :, 0], rtol=2e-6, atol=1e-7)


@pytest.mark.parametrize('name,k', [('analogreg_k10', 10), ('analogreg_k200', 200), ('analogreg_k10_30yr', 10)])
def test_analog_regression(golden, name, k):
    g = golden(name)
    for c in range(g['Xq'].shape[-1]):
        o = oracle.analog_regression_predict(g['Xtr'][..., c], g['ytr'][:, c], g['Xq'][..., c], k)
        _close(o, g['out64'][:, :, c], rtol=1e-9, atol=1e-10)
    out = oracle.pointwise_fit_predict({'name': 'AnalogRegression', 'n_analogs': k}, g['Xtr'], g['ytr'], g['Xq'])
    _close(out, g['out'], rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize('name,k', [('analogreg_thresh_k20', 20), ('analogreg_thresh_k10_C', 10),
                                    ('analogreg_thresh_k200', 200)])
def test_analog_regression_thresh(golden, name, k):
    """AnalogRegression(thresh=...) (gard.py:201-215): prediction and RMSE (least squares on the analogs
    above the threshold, minimum norm when only a few exceed) to 1e-9; the exceedance probability to
    1e-7 against the reference run with a tight lbfgs tolerance, and to 2e-3 against the reference's
    DEFAULT run — whose own solver stops at a gradient of 1e-4, i.e. is only defined to that level."""
    g = golden(name)
    o = oracle.analog_regression_predict(g['Xtr'][..., 0], g['ytr'][:, 0], g['Xq'][..., 0], k,
                                         thresh=float(g['thresh']), logistic_C=float(g['C_reg']))
    ref, tight = g['out64'][:, :, 0], g['out64_tight'][:, :, 0]
    _close(o[:, [0, 2]], ref[:, [0, 2]], rtol=1e-9, atol=1e-10)
    _close(o[:, 1], tight[:, 1], rtol=0, atol=1e-7)
    _close(o[:, 1], ref[:, 1], rtol=0, atol=2e-3)
    assert ((ref[:, 1] > 0) & (ref[:, 1] < 1)).sum() > 20        # the logistic branch is exercised
    if k < 200:
        assert (ref[:, 1] == 1.0).sum() > 0                      # ... and the all-exceed shortcut


def test_analog_regression_thresh_one_class():
    Xtr, ytr, Xq = synth.analog(200, 20, 1, 3, seed=3)
    with pytest.raises(ValueError, match='at least 2 classes'):
        oracle.analog_regression_predict(Xtr[..., 0], ytr[:, 0], Xq[..., 0], 5, thresh=1e6)


# ------------------------------------------------------------------ QuantileMappingReressor / EquidistantCdfMatcher
EX_MODES = [None, 'min', 'max', 'both', '1to1']


@pytest.mark.parametrize('name', ['qmr_equal_len', 'qmr_pred_longer_shifted', 'qmr_f64_shorter'])
def test_qm_regressor_and_edcdf(golden, name):
    """quantile.py:160-395, 556-636 against the live-reference vectors, every extrapolate mode.
    Steps that reach the reference's synthetic +-1e20 CDF points are excluded (no significant digits
    in the reference itself, see oracle.qmr_well_conditioned); everything else is bit-exact."""
    g = golden(name)
    ne = int(g['n_endpoints'])
    for ex in EX_MODES:
        tag = 'none' if ex is None else ex
        n_checked = 0
        for c in range(g['Xp'].shape[1]):
            st = oracle.qm_regressor_fit(g['Xtr'][:, c], g['ytr'][:, c], ex, ne)
            x = g['Xp'][:, c]
            o = oracle.qm_regressor_predict(st, x)
            assert o.dtype == x.dtype
            ok = oracle.qmr_well_conditioned(st, x, 'regressor')
            np.testing.assert_array_equal(o[ok], g[f'qmr_{tag}'][ok, c])
            n_checked += ok.sum()
            ok = oracle.qmr_well_conditioned(st, x, 'edcdf')
            for kind in ('difference', 'ratio'):
                o = oracle.edcdf_predict(st, x, kind)
                np.testing.assert_array_equal(o[ok], g[f"{'diff' if kind == 'difference' else 'ratio'}_{tag}"][ok, c])
        assert n_checked > 0.5 * g['Xp'].size


def test_edcdf_known_answer(golden):
    """The reference's own test (test_pointwise_models.py:323-344): exact equality."""
    g = golden('edcdf_known_answer')
    x = g['x']
    st = oracle.qm_regressor_fit(x, x + 3)
    assert (oracle.edcdf_predict(st, x + 2, 'difference') == (x + 3) + 2).all()
    assert (oracle.edcdf_predict(st, x * 2, 'ratio') == (x + 3) * 2).all()
    assert (g['difference'] == (x + 3) + 2).all() and (g['ratio'] == (x + 3) * 2).all()


@pytest.mark.parametrize('name', ['qm_detrend_equal', 'qm_detrend_longer', 'qm_detrend_f64'])
def test_qm_detrend(golden, name):
    """QuantileMapper(detrend=True) (quantile.py:94-98, 127-145; trend.py:40-83) against the live reference."""
    g = golden(name)
    for c in range(g['Xp'].shape[1]):
        st = oracle.quantile_mapper_fit_detrend(g['ytr'][:, c])
        _close(oracle.quantile_mapper_transform_detrend(g['Xp'][:, c], st), g['out'][:, c], rtol=1e-12, atol=1e-12)
    out = oracle.pointwise_fit_predict({'name': 'QuantileMapper', 'detrend': True}, None, g['ytr'], g['Xp'])
    _close(out, g['out'].astype(g['Xp'].dtype), rtol=1e-6, atol=1e-6)


@pytest.mark.parametrize('name', ['pure_regression', 'pure_regression_thresh', 'pure_regression_f64_thresh'])
def test_pure_regression(golden, name):
    """PureRegression (gard.py:367-504) against the live reference.  float32 inputs are fitted in float32 by
    the reference (LAPACK sgelsd): agreement to 1e-5 of the target spread; float64 inputs: 1e-9.  The
    exceedance probability: 1e-7 against the tightly converged reference, 2e-3 against its default lbfgs run."""
    g = golden(name)
    th = float(g['thresh']) if 'thresh' in g else None
    f64 = g['Xtr'].dtype == np.float64
    for c in range(g['Xq'].shape[-1]):
        o = oracle.pure_regression_fit_predict(g['Xtr'][..., c], g['ytr'][:, c], g['Xq'][..., c], th)
        ref = g['out'][:, :, c]
        tol = 1e-9 if f64 else 1e-5 * np.std(g['ytr'][:, c])
        _close(o[:, [0, 2]], ref[:, [0, 2]], rtol=0, atol=tol)
        if th is None:
            assert (o[:, 1] == 1.0).all() and (ref[:, 1] == 1.0).all()
        else:
            _close(o[:, 1], g['prob_tight'][:, c], rtol=0, atol=1e-7)
            _close(o[:, 1], ref[:, 1], rtol=0, atol=2e-3)


@pytest.mark.parametrize('name', ['trend_aware_qmr', 'trend_aware_qmr_f64'])
def test_trend_aware_qm_regressor_oracle(golden, name):
    """TrendAwareQuantileMappingRegressor (quantile.py:639-716) — oracle only (the estimator is a "next"
    item, not in the product yet): pinned against the live reference so the next round starts from a checker."""
    g = golden(name)
    for ex, key in ((None, 'out_none'), ('1to1', 'out_1to1')):
        for c in range(g['Xp'].shape[1]):
            o = oracle.trend_aware_qm_fit_predict(g['Xtr'][:, c], g['ytr'][:, c], g['Xp'][:, c], ex, 6)
            _close(o[:, 0], g[key][:, c], rtol=1e-12, atol=1e-11)


# ------------------------------------------------------------------ ZScoreRegressor (zscore.py)
def test_zscore_reference_known_answers():
    """The reference's own three tests (test_pointwise_models.py:236-299): a record without a 29 February gives 364
    fitted values; y = 2 X → scale 2; X = 0, y = 1 → shift 1; unit scale / zero shift → predict is the identity
    except for the NaN half-windows at both ends."""
    time = pd.date_range(start='2018-01-01', end='2020-01-01')
    x = np.linspace(0, 1, len(time))
    st = oracle.zscore_fit(x, x * 2, time, 31)
    assert st['scale'].shape == (364,) and st['shift'].shape == (364,)
    np.testing.assert_allclose(st['scale'], np.full(364, 2.0))
    st = oracle.zscore_fit(np.zeros(len(time)), np.ones(len(time)), time, 31)
    np.testing.assert_allclose(st['shift'], np.ones(364))
    out = oracle.zscore_predict({'shift': np.zeros(364), 'scale': np.ones(364), 'window_width': 31}, x)
    want = x.copy()
    want[:15] = np.nan
    want[-15:] = np.nan
    np.testing.assert_allclose(out, want)


def test_zscore_window_columns():
    """The retained windows of zscore.py:150-157,185-190 for w = 31 on a 366-column year: 365 windows; window 0 =
    the last 15 days of the year + days 1-16; the last one = days 351-366 + days 1-15 (the wrap stays inside one
    year's row)."""
    cols = oracle.zscore_window_columns(366, 31)
    assert cols.shape == (365, 31)
    assert cols[0].tolist() == list(range(351, 366)) + list(range(0, 16))
    assert cols[15].tolist() == list(range(0, 31))
    assert cols[-1].tolist() == list(range(349, 366)) + list(range(0, 14))
    assert oracle.zscore_window_columns(365, 30).shape == (363, 30)      # even window, no leap day: 363 < 364
    tab = oracle.zscore_day_table(pd.date_range('1999-12-30', periods=400, freq='D'))
    assert tab.shape == (3, 366) and tab[0, 363] == 0 and tab[0, 0] == -1 and tab[1, 0] == 2 and tab[1, 365] == 367


@pytest.mark.parametrize('name', ['zscore_4yr', 'zscore_pred_longer', 'zscore_w30_f64', 'zscore_short_pred'])
def test_zscore_predict_against_live_reference(golden, name):
    """oracle.zscore_predict against the reference's own ZScoreRegressor.predict (zscore.py:68-110) given the same
    shift_ / scale_."""
    g = golden(name)
    idx = pd.date_range(str(g['start']), periods=len(g['Xtr']), freq='D')
    for c in range(g['Xp'].shape[1]):
        st = oracle.zscore_fit(g['Xtr'][:, c], g['ytr'][:, c], idx, int(g['window']))
        assert np.array_equal(st['shift'], g['shift'][:, c]) and np.array_equal(st['scale'], g['scale'][:, c])
        _close(oracle.zscore_predict(st, g['Xp'][:, c]), g['out'][:, c], rtol=1e-12, atol=1e-12)


@pytest.mark.parametrize('kind', ['difference', 'ratio'])
def test_edcdf_tied_inputs_against_live_reference(golden, kind):
    """EquidistantCdfMatcher on exactly tied inputs (quantile.py:594-636): which tied step takes which plotting
    position is np.argsort's unstable choice (:607), so the oracle must agree with the live reference as a MULTISET
    inside every tie run — and therefore exactly on every untied step.  (The GPU path orders ties by time index and
    is held to the oracle in the same way: test_gpu_parity.py::test_edcdf_tied_inputs_multiset, same inputs.)"""
    g = golden('edcdf_tied')
    for c in range(g['Xp'].shape[1]):
        st = oracle.qm_regressor_fit(g['Xtr'][:, c], g['ytr'][:, c], None, 10)
        got = oracle.edcdf_predict(st, g['Xp'][:, c], kind).astype(np.float32)
        vals, inv, cnt = np.unique(g['Xp'][:, c], return_inverse=True, return_counts=True)
        assert (cnt > 1).sum() > 5 and (cnt == 1).sum() > 0
        for v in range(len(vals)):
            sel = inv == v
            np.testing.assert_allclose(np.sort(got[sel]), np.sort(g[kind][sel, c]), rtol=1e-6, atol=1e-6)
