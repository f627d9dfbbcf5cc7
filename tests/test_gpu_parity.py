"""GPU parity tests: the CUDA path (through the C ABI / ctypes) against

* the golden vectors produced by the LIVE reference (tests/golden/), and
* the numpy oracle on seeded inputs (in-group ranks and analog indices BIT-EXACT, values
  within the north-star tolerance 1e-5 relative: |a-b| <= 1e-5 * max(|b|, sigma_y)),
* size-independent properties at larger sizes.

Run on the B200 box with ``pytest -m gpu``.  Nothing here reads /root/reference.
"""

import numpy as np
import pandas as pd
import pytest
import torch

import oracle
import synth

pytestmark = pytest.mark.gpu

RTOL = 1e-5      # north_star: mapped floating values within 1e-5 relative


@pytest.fixture(scope='module')
def dev():
    assert torch.cuda.is_available(), 'GPU tests need a CUDA device'
    import skdownscale_b200  # noqa: F401
    return torch.device('cuda:0')


def pm():
    import skdownscale_b200.pointwise_models as m
    return m


def eng():
    from skdownscale_b200 import engine
    return engine


class force_generic:
    """Kernel family selection for a block of code: 'tile' (default float32 tile kernels),
    'generic' (any dtype / any length kernels), 'scalar' (tile kernels without 16-byte row accesses).
    ``True`` / ``False`` are accepted for generic / tile."""

    FLAGS = {'tile': 0, 'generic': 1, 'scalar': 2, True: 1, False: 0}

    def __init__(self, on=True):
        self.on = on

    def __enter__(self):
        from skdownscale_b200 import _lib
        self.old = _lib.load().sdb_set_debug_flags(self.FLAGS[self.on])

    def __exit__(self, *a):
        from skdownscale_b200 import _lib
        _lib.load().sdb_set_debug_flags(self.old)


def assert_close(got, ref, scale=1.0, rtol=RTOL):
    got = np.asarray(got, dtype=np.float64)
    ref = np.asarray(ref, dtype=np.float64)
    assert got.shape == ref.shape, (got.shape, ref.shape)
    nan_g, nan_r = np.isnan(got), np.isnan(ref)
    assert np.array_equal(nan_g, nan_r), 'NaN pattern differs'
    tol = rtol * np.maximum(np.abs(ref), scale)
    bad = np.abs(got - ref) > tol
    bad &= ~nan_r
    assert not bad.any(), f'{bad.sum()} / {bad.size} outside tolerance, max abs err {np.nanmax(np.abs(got - ref))}'


# ------------------------------------------------------------------ QuantileMapper
def test_qm_known_answer(dev, golden):
    g = golden('qm_known_answer')      # reference test_pointwise_models.py:81-90
    mapper = pm().QuantileMapper().fit(g['fit'])
    actual = mapper.transform(g['x'])
    assert actual.shape == (100, 1) and actual.dtype == np.float64
    np.testing.assert_almost_equal(actual, g['fit'])
    np.testing.assert_allclose(actual, g['out'], rtol=1e-12, atol=1e-12)


@pytest.mark.parametrize('name', ['qm_equal_len', 'qm_pred_longer', 'qm_pred_shorter', 'qm_ties', 'qm_f64', 'qm_tiny'])
def test_qm_golden(dev, golden, name):
    g = golden(name)
    pw = pm().PointWiseDownscaler(pm().QuantileMapper())
    pw.fit(g['ytr'])
    got = pw.transform(g['Xp'])
    assert got.dtype == g['Xp'].dtype
    assert_close(got, g['out'].astype(g['Xp'].dtype), scale=np.std(g['ytr']))
    # float64 result of the per-cell API: the only difference left is summation order in the tails
    for c in range(g['Xp'].shape[1]):
        m = pm().QuantileMapper().fit(g['ytr'][:, c:c + 1])
        o = m.transform(g['Xp'][:, c:c + 1])[:, 0]
        np.testing.assert_allclose(o, g['out'][:, c], rtol=1e-11, atol=1e-11)


QT_VARIANTS = {          # the settings tests/golden/make_golden.py used for the qm_qt_* files
    'ab': dict(alpha=0.3, beta=0.5, n_endpoints=5),
    'none': dict(extrapolate=None),
    'min': dict(extrapolate='min', n_endpoints=4),
    'max': dict(extrapolate='max', alpha=0.0, beta=1.0),
    '1to1': dict(extrapolate='1to1'),
}


@pytest.mark.parametrize('tag', sorted(QT_VARIANTS))
def test_qm_cunnane_settings_golden(dev, golden, tag):
    """QuantileMapper(qt_kwargs=...): plotting-position parameters, tail selection and end-point count
    of the reference's CunnaneTransformer (quantile.py:420-432) against the live-reference vectors."""
    g = golden(f'qm_qt_{tag}')
    qt = QT_VARIANTS[tag]
    pw = pm().PointWiseDownscaler(pm().QuantileMapper(qt_kwargs=qt))
    pw.fit(g['ytr'])
    got = pw.transform(g['Xp'])                                   # float32 tile kernels
    assert_close(got, g['out'].astype(g['Xp'].dtype), scale=np.std(g['ytr']))
    for c in range(g['Xp'].shape[1]):                             # per-cell API, float64 out
        o = pm().QuantileMapper(qt_kwargs=qt).fit(g['ytr'][:, c:c + 1]).transform(g['Xp'][:, c:c + 1])[:, 0]
        np.testing.assert_allclose(o, g['out'][:, c], rtol=1e-11, atol=1e-11)


def test_qm_cunnane_bad_settings(dev):
    with pytest.raises(TypeError, match='unexpected keyword'):
        pm().QuantileMapper(qt_kwargs={'gamma': 1}).fit(np.arange(10.0).reshape(-1, 1))
    with pytest.raises(ValueError, match='extrapolate'):
        pm().QuantileMapper(qt_kwargs={'extrapolate': 'sideways'}).fit(np.arange(10.0).reshape(-1, 1))


@pytest.mark.parametrize('n_fit,n_pred', [(1, 1), (2, 5), (255, 256), (256, 256), (257, 257), (1024, 1024),
                                          (1025, 1000), (4096, 4096), (4097, 5000), (10950, 10950), (16384, 16384)])
def test_qm_sizes_vs_oracle(dev, n_fit, n_pred):
    """every padded-size boundary of the sorting network; ranks bit-exact, values 1e-5."""
    C = 5
    rng = np.random.default_rng(n_fit * 31 + n_pred)
    y = rng.standard_normal((n_fit, C)).astype(np.float32) * 5 + 3
    x = rng.standard_normal((n_pred, C)).astype(np.float32) * 6 + 4
    x[: n_pred // 3, 0] = x[0, 0]                  # a long run of exact ties in cell 0
    m = pm().QuantileMapper()
    m.fit_batched(eng().as_device(y, dev))
    out, rank = m.transform_batched(eng().as_device(x, dev), out_dtype=torch.float64, want_rank=True)
    out, rank = out.cpu().numpy(), rank.cpu().numpy()
    for c in range(C):
        st = oracle.quantile_mapper_fit(y[:, c])
        o, r = oracle.quantile_mapper_transform(x[:, c], st, return_rank=True)
        assert np.array_equal(rank[:, c], r), f'rank mismatch in cell {c}'
        np.testing.assert_allclose(out[:, c], o, rtol=1e-10, atol=1e-10)


def test_qm_group_too_long(dev):
    y = torch.zeros((16385, 2), device=dev)
    with pytest.raises(NotImplementedError):
        pm().QuantileMapper().fit_batched(y)


def test_qm_sorted_multiset_property(dev):
    """n == m: the mapped value of every element is the fitted order statistic at its tie-max rank
    (size-independent property, larger shape; exact equality, computed independently with torch)."""
    T, C = 10950, 512
    gen = torch.Generator(device=dev).manual_seed(0)
    y = torch.randn((T, C), device=dev, generator=gen)
    x = torch.randn((T, C), device=dev, generator=gen) * 2 + 1
    m = pm().QuantileMapper()
    m.fit_batched(y)
    out = m.transform_batched(x)
    torch.testing.assert_close(out, _expected_rank_map(x, y), rtol=0, atol=0)


def _expected_rank_map(x, y):
    """sorted(y)[rank(x) - 1] per column, rank = number of elements <= x (ties take the highest)."""
    ys = torch.sort(y, dim=0).values
    xs = torch.sort(x, dim=0).values
    r = torch.searchsorted(xs.T.contiguous(), x.T.contiguous(), right=True).T
    return torch.gather(ys, 0, r - 1)


@pytest.mark.parametrize('name', ['qm_detrend_equal', 'qm_detrend_longer', 'qm_detrend_f64'])
def test_qm_detrend_golden(dev, golden, name):
    """QuantileMapper(detrend=True) (quantile.py:94-98, 127-145) against the live-reference vectors."""
    g = golden(name)
    pw = pm().PointWiseDownscaler(pm().QuantileMapper(detrend=True))
    pw.fit(g['ytr'])
    got = pw.transform(g['Xp'])
    assert got.dtype == g['Xp'].dtype
    assert_close(got, g['out'].astype(g['Xp'].dtype), scale=np.std(g['ytr']))
    for c in range(g['Xp'].shape[1]):                             # per-cell API: float64 like the reference
        o = pm().QuantileMapper(detrend=True).fit(g['ytr'][:, c:c + 1]).transform(g['Xp'][:, c:c + 1])[:, 0]
        np.testing.assert_allclose(o, g['out'][:, c], rtol=1e-9, atol=1e-9)


# ------------------------------------------------------------------ QuantileMappingReressor / EquidistantCdfMatcher
@pytest.mark.parametrize('name', ['qmr_equal_len', 'qmr_pred_longer_shifted', 'qmr_f64_shorter'])
@pytest.mark.parametrize('ex', [None, 'min', 'max', 'both', '1to1'])
def test_qm_regressor_and_edcdf_golden(dev, golden, name, ex):
    """quantile.py:160-395, 556-636 for every extrapolate mode, against the live-reference vectors:
    1e-5 relative (values are float32 / float64 like the reference's) on every step the reference
    itself resolves (oracle.qmr_well_conditioned); the rest must merely be produced."""
    g = golden(name)
    ne, tag = int(g['n_endpoints']), ('none' if ex is None else ex)
    scale = np.std(g['ytr'])
    models = {'qmr': pm().QuantileMappingReressor(extrapolate=ex, n_endpoints=ne),
              'diff': pm().EquidistantCdfMatcher(kind='difference', extrapolate=ex, n_endpoints=ne),
              'ratio': pm().EquidistantCdfMatcher(kind='ratio', extrapolate=ex, n_endpoints=ne)}
    for key, model in models.items():
        pw = pm().PointWiseDownscaler(model)
        pw.fit(g['Xtr'], g['ytr'])
        got = pw.predict(g['Xp'])
        assert got.dtype == g['Xp'].dtype and got.shape == g['Xp'].shape
        for c in range(g['Xp'].shape[1]):
            st = oracle.qm_regressor_fit(g['Xtr'][:, c], g['ytr'][:, c], ex, ne)
            ok = oracle.qmr_well_conditioned(st, g['Xp'][:, c], 'regressor' if key == 'qmr' else 'edcdf')
            assert ok.sum() > 0.5 * len(ok)
            assert_close(got[ok, c], g[f'{key}_{tag}'][ok, c], scale=scale)
    # per-cell estimator API of the reference
    m = pm().QuantileMappingReressor(extrapolate=ex, n_endpoints=ne).fit(g['Xtr'][:, :1], g['ytr'][:, 0])
    o = m.predict(g['Xp'][:, :1])
    st = oracle.qm_regressor_fit(g['Xtr'][:, 0], g['ytr'][:, 0], ex, ne)
    ok = oracle.qmr_well_conditioned(st, g['Xp'][:, 0], 'regressor')
    assert o.shape == (len(g['Xp']),) and o.dtype == g['Xp'].dtype
    assert_close(o[ok], g[f'qmr_{tag}'][ok, 0], scale=scale)


@pytest.mark.parametrize('name', ['trend_aware_qmr', 'trend_aware_qmr_f64'])
def test_trend_aware_qm_regressor_golden(dev, golden, name):
    """TrendAwareQuantileMappingRegressor (quantile.py:639-716) against the live-reference vectors, batched over the
    cells and through the per-cell API."""
    g = golden(name)
    for ex, key in ((None, 'out_none'), ('1to1', 'out_1to1')):
        m = pm().TrendAwareQuantileMappingRegressor(pm().QuantileMappingReressor(extrapolate=ex, n_endpoints=6))
        m.fit_batched(eng().as_device(g['Xtr'], dev), eng().as_device(g['ytr'], dev))
        m.check_fit()
        out = m.predict_batched(eng().as_device(g['Xp'], dev)).cpu().numpy()
        assert out.dtype == np.float64
        np.testing.assert_allclose(out, g[key], rtol=1e-9, atol=1e-9)
        m1 = pm().TrendAwareQuantileMappingRegressor(pm().QuantileMappingReressor(extrapolate=ex, n_endpoints=6))
        o1 = m1.fit(g['Xtr'][:, :1], g['ytr'][:, :1]).predict(g['Xp'][:, :1])
        assert o1.shape == (g['Xp'].shape[0], 1)
        np.testing.assert_allclose(o1[:, 0], g[key][:, 0], rtol=1e-9, atol=1e-9)


def test_linear_trend_transformer_roundtrip(dev):
    """The reference's own known-answer test (test_pointwise_models.py:56-78): a pure line detrends to zero with the
    fitted slope / intercept, inverse_transform restores the data; plus the oracle's line on noisy columns."""
    n = 100
    trend, yint = 1.5, 3.0
    X = (np.arange(n) * trend + yint).reshape(-1, 1)
    lt = pm().LinearTrendTransformer().fit(X)
    np.testing.assert_allclose(lt._slope.cpu().numpy().ravel(), [trend], rtol=1e-12)
    np.testing.assert_allclose(lt._icpt.cpu().numpy().ravel(), [yint], rtol=1e-10)
    Xt = lt.transform(X)
    np.testing.assert_allclose(Xt, np.zeros((n, 1)), atol=1e-10)
    np.testing.assert_allclose(lt.inverse_transform(Xt), X, rtol=1e-12, atol=1e-10)
    rng = np.random.default_rng(4)
    A = (rng.standard_normal((257, 3)) + np.linspace(0, 3, 257)[:, None]).astype(np.float32)
    lt = pm().LinearTrendTransformer().fit(A)
    line = lt.trendline(A)
    for c in range(3):
        slope, icpt = oracle.linear_trend_fit(A[:, c])
        np.testing.assert_allclose(line[:, c], np.arange(257) * slope + icpt, rtol=1e-12, atol=1e-12)
    np.testing.assert_allclose(lt.transform(A), A.astype(np.float64) - line, rtol=0, atol=1e-12)


def test_wrapper_transformers_and_trend_aware(dev, golden):
    """PointWiseDownscaler around the remaining estimators of SURVEY 8(f)2-3: LinearTrendTransformer with
    transform / inverse_transform (core.py:340-403; output in X.dtype, NaN cells stay NaN), and
    TrendAwareQuantileMappingRegressor through fit / predict against the live-reference vectors."""
    rng = np.random.default_rng(12)
    A = (rng.standard_normal((400, 3, 5)) + np.linspace(0, 4, 400)[:, None, None]).astype(np.float32)
    A[:, 1, 2] = np.nan                                            # a masked cell
    pw = pm().PointWiseDownscaler(pm().LinearTrendTransformer())
    pw.fit(A)
    B = A[:250] + 1.0
    Bt = pw.transform(B)
    assert Bt.shape == B.shape and Bt.dtype == np.float32
    assert np.isnan(Bt[:, 1, 2]).all() and np.isfinite(np.delete(Bt.reshape(250, -1), 7, axis=1)).all()
    for (i, j) in ((0, 0), (2, 4)):
        slope, icpt = oracle.linear_trend_fit(A[:, i, j])
        want = B[:, i, j].astype(np.float64) - (np.arange(250) * slope + icpt)
        np.testing.assert_allclose(Bt[:, i, j], want.astype(np.float32), rtol=0, atol=2e-6)
    back = pw.inverse_transform(Bt)
    ok = ~np.isnan(B)
    np.testing.assert_allclose(back[ok], B[ok], rtol=0, atol=4e-6)
    with pytest.raises(AttributeError):
        pw.predict(B)
    qm = pm().PointWiseDownscaler(pm().QuantileMapper())
    qm.fit(A[:, 0])
    with pytest.raises(AttributeError):
        qm.inverse_transform(B[:, 0])

    g = golden('trend_aware_qmr')
    pw = pm().PointWiseDownscaler(pm().TrendAwareQuantileMappingRegressor(pm().QuantileMappingReressor(extrapolate='1to1', n_endpoints=6)))
    pw.fit(g['Xtr'], g['ytr'])
    got = pw.predict(g['Xp'])
    assert got.dtype == g['Xp'].dtype and got.shape == g['out_1to1'].shape
    assert_close(got, g['out_1to1'].astype(got.dtype), scale=np.std(g['ytr']))


@pytest.mark.parametrize('kind', ['difference', 'ratio'])
def test_edcdf_tied_inputs_multiset(dev, kind):
    """EquidistantCdfMatcher on inputs with exact ties (precipitation-like: rounded values, a block of equal ones).
    The reference orders tied steps by np.argsort's unstable default sort (quantile.py:607), this library by
    (value, time index): which tied step receives which plotting position is not defined by the algorithm, so the
    outputs must agree as a MULTISET inside every tie run and exactly (1e-5) on the untied steps."""
    rng = np.random.default_rng(21)
    n, C = 900, 4
    Xtr = (rng.gamma(0.8, 6.0, (n, C)) + 0.5).astype(np.float32)
    ytr = (rng.gamma(0.9, 5.0, (n, C)) + 0.5).astype(np.float32)
    Xp = np.round(rng.gamma(0.8, 6.0, (n, C)) + 0.5, 0).astype(np.float32) + 1.0      # heavy ties, strictly positive
    Xp[100:160, 1] = 3.0
    pw = pm().PointWiseDownscaler(pm().EquidistantCdfMatcher(kind=kind, extrapolate=None))
    pw.fit(Xtr, ytr)
    got = pw.predict(Xp).astype(np.float64)
    scale = np.std(ytr)
    for c in range(C):
        st = oracle.qm_regressor_fit(Xtr[:, c], ytr[:, c], None, 10)
        ref = oracle.edcdf_predict(st, Xp[:, c], kind).astype(np.float32).astype(np.float64)
        vals, inv, cnt = np.unique(Xp[:, c], return_inverse=True, return_counts=True)
        assert (cnt > 1).sum() > 5
        for v in range(len(vals)):
            sel = inv == v
            a, b = np.sort(got[sel, c]), np.sort(ref[sel])
            assert np.all(np.abs(a - b) <= 1e-5 * np.maximum(np.abs(b), scale)), (kind, c, vals[v])


def test_edcdf_known_answer(dev, golden):
    """The reference's own exact test (test_pointwise_models.py:323-344)."""
    x = golden('edcdf_known_answer')['x']
    for kind, Xt, want in (('difference', x + 2, (x + 3) + 2), ('ratio', x * 2, (x + 3) * 2)):
        m = pm().EquidistantCdfMatcher(kind=kind).fit(pd.DataFrame(x), pd.DataFrame(x + 3))
        got = m.predict(pd.DataFrame(Xt))
        assert (got.reshape(-1, 1) == want.reshape(-1, 1)).all()


def test_qm_regressor_errors(dev):
    with pytest.raises(ValueError, match='n_endpoints'):
        pm().QuantileMappingReressor(n_endpoints=1)
    with pytest.raises(NotImplementedError):
        pm().EquidistantCdfMatcher(kind='sum')
    x = np.arange(15.0).reshape(-1, 1)
    with pytest.raises(ValueError, match='minimum of 21'):
        pm().QuantileMappingReressor().fit(x, x[:, 0])                 # 2 * n_endpoints + 1 samples needed
    with pytest.raises(ValueError, match='extrapolate'):
        pm().QuantileMappingReressor(extrapolate='sideways').fit(np.arange(30.0).reshape(-1, 1), np.arange(30.0))


# ------------------------------------------------------------------ BCSD
@pytest.mark.parametrize('name,kw', [
    ('bcsd_t_month_anoms', {}),
    ('bcsd_t_month_abs', {'return_anoms': False}),
    ('bcsd_t_month_future', {}),
    ('bcsd_t_month_future_qt', {'qm_kwargs': {'qt_kwargs': dict(alpha=0.3, beta=0.5, n_endpoints=5, extrapolate='max')}}),
    ('bcsd_t_month_detrend', {'qm_kwargs': {'detrend': True}}),
    ('bcsd_t_month_detrend_future', {'qm_kwargs': {'detrend': True}, 'return_anoms': False}),
    ('bcsd_t_month_f64', {}),
    ('bcsd_t_month_30yr', {}),
    ('bcsd_t_nasanex', {'time_grouper': 'daily_nasa-nex', 'return_anoms': False}),
])
def test_bcsd_temperature_golden(dev, golden, name, kw):
    g = golden(name)
    idx_f0 = synth.daily_index(len(g['Xtr']), str(g['start_fit']))
    idx_p0 = synth.daily_index(len(g['Xp']), str(g['start_pred']))
    with force_generic('scalar'):        # the unaligned-row form of the tile kernels gives the same field
        pw0 = pm().PointWiseDownscaler(pm().BcsdTemperature(**kw))
        pw0.fit(g['Xtr'], g['ytr'], time=idx_f0)
        assert_close(pw0.predict(g['Xp'], time=idx_p0), g['out'], scale=np.nanstd(g['ytr']))
    idx_f = synth.daily_index(len(g['Xtr']), str(g['start_fit']))
    idx_p = synth.daily_index(len(g['Xp']), str(g['start_pred']))
    pw = pm().PointWiseDownscaler(pm().BcsdTemperature(**kw))
    pw.fit(g['Xtr'], g['ytr'], time=idx_f)
    got = pw.predict(g['Xp'], time=idx_p)
    assert got.dtype == g['Xp'].dtype and got.shape == g['out'].shape
    assert_close(got, g['out'], scale=np.nanstd(g['ytr']))
    # per-cell estimator API (config[0] of BASELINE.json): float64 DataFrame like the reference
    c = 0
    m = pm().BcsdTemperature(**kw)
    m.fit(pd.DataFrame({'x': g['Xtr'][:, c]}, index=idx_f), pd.DataFrame({'x': g['ytr'][:, c]}, index=idx_f))
    o = m.predict(pd.DataFrame({'x': g['Xp'][:, c]}, index=idx_p))
    assert isinstance(o, pd.DataFrame) and o.shape == (len(idx_p), 1) and o.values.dtype == np.float64
    assert o.index.equals(idx_p)
    np.testing.assert_allclose(o.values[:, 0], g['out64'][:, c], rtol=0, atol=2e-6 * np.nanstd(g['ytr']))


@pytest.mark.parametrize('name,kw', [
    ('bcsd_p_month_anoms', {}),
    ('bcsd_p_month_30yr', {}),
    ('bcsd_p_month_detrend', {'qm_kwargs': {'detrend': True}}),
    ('bcsd_p_month_abs_future', {'return_anoms': False}),
    ('bcsd_p_nasanex', {'time_grouper': 'daily_nasa-nex', 'return_anoms': False}),
])
def test_bcsd_precipitation_golden(dev, golden, name, kw):
    g = golden(name)
    idx_f = synth.daily_index(len(g['Xtr']), str(g['start_fit']))
    idx_p = synth.daily_index(len(g['Xp']), str(g['start_pred']))
    pw = pm().PointWiseDownscaler(pm().BcsdPrecipitation(**kw))
    pw.fit(g['Xtr'], g['ytr'], time=idx_f)
    got = pw.predict(g['Xp'], time=idx_p)
    assert_close(got, g['out'], scale=np.std(g['ytr']))


def test_bcsd_errors(dev):
    idx = synth.daily_index(800)
    z = np.zeros((800, 3), np.float32)
    pw = pm().PointWiseDownscaler(pm().BcsdPrecipitation())
    with pytest.raises(ValueError, match='Invalid value in target climatology'):      # bcsd.py:140-141
        pw.fit(z + 1, z, time=idx)
    Xtr, ytr, Xp = synth.temperature(800, 3, 5)
    bad = ytr.copy()
    bad[100, 1] = np.nan                                                               # NaN inside an unmasked cell
    pw = pm().PointWiseDownscaler(pm().BcsdTemperature())
    with pytest.raises(ValueError, match='NaN'):                                       # base.py:18-20
        pw.fit(Xtr, bad, time=idx)
    m = pm().BcsdTemperature(time_grouper='daily_nasa-nex')                            # return_anoms=True
    pw = pm().PointWiseDownscaler(m)
    pw.fit(Xtr, ytr, time=idx)
    with pytest.raises(ValueError, match='shape of climo'):                            # bcsd.py:267,279-280
        pw.predict(Xp, time=idx)
    with pytest.raises(ValueError, match='1 feature'):
        pm().BcsdTemperature().fit(pd.DataFrame(np.zeros((10, 2)), index=idx[:10]), pd.DataFrame(np.zeros((10, 1)), index=idx[:10]))
    with pytest.raises(AssertionError):                                                # base.py:17
        pm().BcsdTemperature().fit(pd.DataFrame(np.zeros((10, 1)), index=idx[:10]), pd.DataFrame(np.zeros((10, 1)), index=idx[5:15]))


@pytest.mark.parametrize('generic', ['tile', 'generic', 'scalar'])
@pytest.mark.parametrize('anoms', [True, False])
def test_bcsd_temperature_vs_oracle_ranks(dev, anoms, generic):
    """30-year daily series, ragged cell count, NaN cells; ranks bit-exact, values 1e-5.
    Both kernel families (float32 tile kernels / generic kernels) must give the same answer."""
    with force_generic(generic):
        _bcsd_temperature_vs_oracle_ranks(dev, anoms)


def _bcsd_temperature_vs_oracle_ranks(dev, anoms):
    T, C = 10950, 37
    idx = synth.daily_index(T)
    Xtr, ytr, Xp = synth.temperature(T, C, seed=3)
    for c in (4, 36):
        Xtr[:, c] = np.nan
    m = pm().BcsdTemperature(return_anoms=anoms)
    xtr = eng().as_device(Xtr, dev)
    m.fit_batched(xtr, eng().as_device(ytr, dev), idx, valid=eng().cell_mask(xtr[0]))
    out, rank = m.predict_batched(eng().as_device(Xp, dev), idx, want_rank=True)
    out, rank = out.cpu().numpy(), rank.cpu().numpy()
    groups = oracle.groups_from_keys(oracle.month_keys(idx))
    for c in range(0, C, 3):
        if np.isnan(Xtr[0, c]):
            assert np.isnan(out[:, c]).all()
            continue
        st = oracle.bcsd_temperature_fit(Xtr[:, c], ytr[:, c], groups)
        o, r = oracle.bcsd_temperature_predict(st, Xp[:, c], groups, groups, anoms, return_rank=True)
        assert np.array_equal(rank[:, c], r), f'rank mismatch cell {c}'
        assert_close(out[:, c], o.astype(np.float32), scale=np.std(ytr[:, c]))
    assert np.isnan(out[:, 4]).all() and np.isnan(out[:, 36]).all()
    # fitted climatologies are the reference's float32 Kahan means, bit for bit
    yc = m._state.y_climo.cpu().numpy()
    for gi, (key, rows) in enumerate(groups):
        assert yc[gi, 0] == oracle.bcsd._group_mean_like_pandas(ytr[rows, 0])


@pytest.mark.parametrize('family', ['tile', 'scalar'])
@pytest.mark.parametrize('case', ['outlier', 'clusters', 'constant', 'two_values', 'zeros_and_tiny', 'ties_and_pairs'])
@pytest.mark.parametrize('model', ['T', 'P'])
def test_tile_kernel_bucket_fixups(dev, case, model, family):
    with force_generic(family):
        _tile_kernel_bucket_fixups(dev, case, model)


def _tile_kernel_bucket_fixups(dev, case, model):
    """Inputs built to defeat the 22-bit key quantisation of the tile kernels: a huge outlier
    squeezing every other value into one bucket (→ exact 64-bit fallback sort), clusters of
    near-duplicates one float32 ulp apart (→ local exact fix-up), constant series and
    two-valued series (→ exact-tie runs).  Ranks must still be bit-exact."""
    T, C = 2922, 9
    idx = synth.daily_index(T)
    rng = np.random.default_rng(77)
    Xtr, ytr, Xp = synth.temperature(T, C, seed=21)
    if case == 'outlier':
        Xp[::365, :] = 1.0e9          # still exactly summable in float64 next to O(10) values
        Xp[100::400, 1] = -5.0e8
    elif case == 'clusters':
        base = np.float32(12.5)
        for c in range(C):
            k = rng.integers(0, 40, T)                # ~40 adjacent float32 values
            Xp[:, c] = (base.view(np.int32) + k.astype(np.int32)).view(np.float32)
        Xp[5::7, 0] += 3.0
    elif case == 'constant':
        Xp[:] = np.float32(7.25)
    elif case == 'two_values':
        Xp[:] = np.where(rng.random((T, C)) < 0.6, np.float32(0.0), np.float32(1e-7))
    elif case == 'zeros_and_tiny':
        # zero-inflated series whose smallest positive values would share the zeros' bucket under a
        # plain linear quantisation (range ~60 → bucket width 1.4e-5)
        wet = rng.random((T, C)) > 0.55
        Xp[:] = np.where(wet, rng.gamma(0.8, 6.0, (T, C)), 0.0).astype(np.float32)
        Xp[3::97, :] = np.float32(3e-6)
        Xp[5::101, 2] = np.float32(7e-6)
    elif case == 'ties_and_pairs':
        # long exact-tie runs (values on a 0.5 grid) next to isolated near-duplicates one ulp apart
        Xp[:] = (np.round(Xp * 2) / 2).astype(np.float32)
        near = (np.float32(11.25).view(np.int32) + np.arange(1, 6, dtype=np.int32)).view(np.float32)
        for k, v in enumerate(near):
            Xp[40 + 31 * k::360, 1::2] = v
    groups = oracle.groups_from_keys(oracle.month_keys(idx))
    if model == 'T':
        m = pm().BcsdTemperature(return_anoms=False)
    else:
        m = pm().BcsdPrecipitation(return_anoms=False)
    m.fit_batched(eng().as_device(Xtr, dev), eng().as_device(ytr, dev), idx)
    out, rank = m.predict_batched(eng().as_device(Xp, dev), idx, want_rank=True)
    out, rank = out.cpu().numpy(), rank.cpu().numpy()
    for c in range(C):
        if model == 'T':
            st = oracle.bcsd_temperature_fit(Xtr[:, c], ytr[:, c], groups)
            with np.errstate(all='ignore'):
                o, r = oracle.bcsd_temperature_predict(st, Xp[:, c], groups, groups, False, return_rank=True)
        else:
            st = oracle.bcsd_precipitation_fit(ytr[:, c], groups, False)
            o, r = oracle.bcsd_precipitation_predict(st, Xp[:, c], groups, False, return_rank=True)
        assert np.array_equal(rank[:, c], r), f'{case}/{model}: rank mismatch in cell {c}'
        if case != 'outlier':
            assert_close(out[:, c], o.astype(np.float32), scale=np.std(ytr[:, c]))
    # the same call without rank instrumentation takes the register paths (one member per bucket, tie
    # runs / isolated pairs patched in registers, exact-sort fallback): identical field
    out2 = m.predict_batched(eng().as_device(Xp, dev), idx).cpu().numpy()
    assert np.array_equal(out2, out, equal_nan=True), f'{case}/{model}: register path differs from the staged path'


@pytest.mark.parametrize('generic', ['tile', 'generic', 'scalar'])
def test_bcsd_precipitation_vs_oracle(dev, generic):
    with force_generic(generic):
        _bcsd_precipitation_vs_oracle(dev)


def _bcsd_precipitation_vs_oracle(dev):
    T, C = 3653, 16
    idx = synth.daily_index(T)
    Xtr, ytr, Xp = synth.precipitation(T, C, seed=9)
    m = pm().BcsdPrecipitation()
    m.fit_batched(eng().as_device(Xtr, dev), eng().as_device(ytr, dev), idx)
    m.check_fit()
    out, rank = m.predict_batched(eng().as_device(Xp, dev), idx, want_rank=True)
    out, rank = out.cpu().numpy(), rank.cpu().numpy()
    groups = oracle.groups_from_keys(oracle.month_keys(idx))
    for c in range(C):
        st = oracle.bcsd_precipitation_fit(ytr[:, c], groups)
        o, r = oracle.bcsd_precipitation_predict(st, Xp[:, c], groups, True, return_rank=True)
        assert np.array_equal(rank[:, c], r)          # all zeros of a group share the highest rank
        assert_close(out[:, c], o.astype(np.float32), scale=1.0)


def test_bcsd_custom_grouper_and_strided_views(dev):
    """callable grouper (quarters) + inputs that are column slices of a wider array (ld > C)."""
    T, C = 2000, 12
    idx = synth.daily_index(T)
    Xtr, ytr, Xp = synth.temperature(T, C + 5, seed=11)
    q = lambda ts: ts.quarter   # noqa: E731
    m = pm().BcsdTemperature(time_grouper=q, climate_trend=q)
    d = lambda a: eng().as_device(a, dev)[:, 2:2 + C]   # noqa: E731
    m.fit_batched(d(Xtr), d(ytr), idx)
    out = m.predict_batched(d(Xp), idx).cpu().numpy()
    groups = oracle.groups_from_keys(np.asarray(idx.quarter))
    for c in (0, C - 1):
        st = oracle.bcsd_temperature_fit(Xtr[:, 2 + c], ytr[:, 2 + c], groups)
        o = oracle.bcsd_temperature_predict(st, Xp[:, 2 + c], groups, groups, True)
        assert_close(out[:, c], o.astype(np.float32), scale=np.std(ytr))
    # climate_trend left at MONTH while mapping by quarter → neighbour-table kernel
    m2 = pm().BcsdTemperature(time_grouper=q)
    m2.fit_batched(d(Xtr), d(ytr), idx)
    out2 = m2.predict_batched(d(Xp), idx).cpu().numpy()
    roll = oracle.groups_from_keys(oracle.month_keys(idx))
    st = oracle.bcsd_temperature_fit(Xtr[:, 2], ytr[:, 2], groups)
    o2 = oracle.bcsd_temperature_predict(st, Xp[:, 2], roll, groups, True)
    assert_close(out2[:, 0], o2.astype(np.float32), scale=np.std(ytr))


def test_bcsd_abs_multiset_property(dev):
    """pure per-month quantile mapping (BcsdPrecipitation, return_anoms=False) of X onto y's
    distribution: every element equals the fitted order statistic of its month at its tie-max
    rank — exact equality on a larger block, expectation computed independently with torch."""
    T, C = 10950, 1024
    idx = synth.daily_index(T)
    gen = torch.Generator(device=dev).manual_seed(1)
    s = torch.sin(2 * torch.pi * torch.arange(T, device=dev) / 365.25)[:, None]
    Xtr = 15 + 10 * s + 3 * torch.randn((T, C), device=dev, generator=gen)
    ytr = 14 + 12 * s + 2 * torch.randn((T, C), device=dev, generator=gen)
    m = pm().BcsdPrecipitation(return_anoms=False)      # pure per-month QM of X onto y's distribution
    m.fit_batched(Xtr, ytr, idx)
    out = m.predict_batched(Xtr, idx)
    month = torch.as_tensor(np.asarray(idx.month), device=dev)
    for mo in (1, 2, 7, 12):
        sel = month == mo
        torch.testing.assert_close(out[sel], _expected_rank_map(Xtr[sel], ytr[sel]), rtol=0, atol=0)


def test_full_shard_headline_and_precipitation(dev):
    """BASELINE.json's full per-GPU shard (129 600 cells x 10 950 days, the 8-GPU slice of the 720x1440
    grid): (a) BcsdTemperature fit+predict — cells sampled at the ends of the shard, around tile
    boundaries and at random are compared with the oracle (1e-5), which exercises the full-size launch
    geometry and 64-bit addressing; (b) the per-month quantile map of X onto y's distribution
    (BcsdPrecipitation, return_anoms=False) — for two whole months EVERY cell must equal the fitted
    order statistic at its tie-max rank (exact, computed independently with torch), on ZERO-INFLATED fields."""
    if torch.cuda.get_device_properties(dev).total_memory < 60e9:
        pytest.skip('needs ~35 GB of device memory')
    T, C = 10950, 129600
    idx = synth.daily_index(T)
    gen = torch.Generator(device=dev).manual_seed(2024)
    s = torch.sin(2 * torch.pi * torch.arange(T, device=dev, dtype=torch.float32) / 365.25)[:, None]

    def field(mean, amp, sd):
        x = torch.randn((T, C), device=dev, dtype=torch.float32, generator=gen)
        return x.mul_(sd).add_(mean + amp * s)

    Xtr, ytr, Xp = field(15.0, 10.0, 3.0), field(14.0, 12.0, 2.0), field(16.5, 10.0, 3.0)
    m = pm().BcsdTemperature(return_anoms=True)
    m.fit_batched(Xtr, ytr, idx)
    out = m.predict_batched(Xp, idx)
    m._state.check_finite()
    rng = np.random.default_rng(5)
    cells = sorted({0, 1, 7, 8, 9, C // 2, C - 9, C - 8, C - 1, *rng.integers(0, C, 12).tolist()})
    groups = oracle.groups_from_keys(oracle.month_keys(idx))
    got = out[:, cells].cpu().numpy()
    xs, ys, ps = (a[:, cells].cpu().numpy() for a in (Xtr, ytr, Xp))
    for k in range(len(cells)):
        st = oracle.bcsd_temperature_fit(xs[:, k], ys[:, k], groups)
        o = oracle.bcsd_temperature_predict(st, ps[:, k], groups, groups, True)
        assert_close(got[:, k], o.astype(np.float32), scale=np.std(ys[:, k]))
    del out, m, Xtr, ytr, Xp
    torch.cuda.empty_cache()
    # (b) ZERO-INFLATED gamma fields (Gamma(.8, 6) where U >= p_dry, else 0): the zeros of a month
    # are one exact-tie run of hundreds that takes the run's highest rank — the register tie-run path of the tile
    # kernels at the full launch geometry
    def precip(p_dry):
        x = torch._standard_gamma(torch.full((T, C), 0.8, device=dev, dtype=torch.float32)).mul_(6.0)
        x[torch.rand((T, C), device=dev, generator=gen) < p_dry] = 0.0
        return x

    ytr, Xp = precip(0.5), precip(0.55)
    q = pm().BcsdPrecipitation(return_anoms=False)
    q.fit_batched(ytr, ytr, idx)
    out = q.predict_batched(Xp, idx)
    q._state.check_finite()
    month = torch.as_tensor(np.asarray(idx.month), device=dev)
    for mo in (2, 8):
        sel = month == mo
        torch.testing.assert_close(out[sel], _expected_rank_map(Xp[sel], ytr[sel]), rtol=0, atol=0)
    frac_zero = float((Xp == 0).float().mean())
    assert 0.5 < frac_zero < 0.6


# ------------------------------------------------------------------ GARD
@pytest.mark.parametrize('kind', ['best_analog', 'mean_analogs', 'weight_analogs', 'sample_analogs'])
@pytest.mark.parametrize('suffix,thresh', [('', None), ('_thresh', 0.0)])
def test_pure_analog_golden(dev, golden, kind, suffix, thresh):
    g = golden(f'pure_{kind}{suffix}')
    model = pm().PureAnalog(n_analogs=10, kind=kind, thresh=thresh)
    pw = pm().PointWiseDownscaler(model)
    pw.fit(g['Xtr'], g['ytr'])
    if kind == 'sample_analogs':
        np.random.seed(1234)                     # same global-RNG stream as the reference run (gard.py:315)
    got = pw.predict(g['Xq'])
    assert got.shape == g['out'].shape and got.dtype == g['Xq'].dtype
    assert_close(got, g['out'], scale=np.std(g['ytr']))


def test_pure_analog_k200_golden(dev, golden):
    g = golden('pure_mean_analogs_k200')
    pw = pm().PointWiseDownscaler(pm().PureAnalog(n_analogs=200, kind='mean_analogs'))
    pw.fit(g['Xtr'], g['ytr'])
    assert_close(pw.predict(g['Xq']), g['out'], scale=np.std(g['ytr']))


@pytest.mark.parametrize('name,k', [('analogreg_k10', 10), ('analogreg_k200', 200), ('analogreg_k10_30yr', 10)])
def test_analog_regression_golden(dev, golden, name, k):
    g = golden(name)
    pw = pm().PointWiseDownscaler(pm().AnalogRegression(n_analogs=k))
    pw.fit(g['Xtr'], g['ytr'])
    got = pw.predict(g['Xq'])
    assert_close(got, g['out'], scale=np.std(g['ytr']))
    # per-cell estimator API, float64
    m = pm().AnalogRegression(n_analogs=k).fit(pd.DataFrame(g['Xtr'][..., 0]), pd.DataFrame(g['ytr'][:, 0]))
    o = m.predict(pd.DataFrame(g['Xq'][..., 0]))
    assert list(o.columns) == ['pred', 'exceedance_prob', 'prediction_error']
    np.testing.assert_allclose(o.values, g['out64'][:, :, 0], rtol=1e-7, atol=1e-8)


@pytest.mark.parametrize('name,k', [('analogreg_thresh_k20', 20), ('analogreg_thresh_k10_C', 10),
                                    ('analogreg_thresh_k200', 200)])
def test_analog_regression_thresh_golden(dev, golden, name, k):
    """AnalogRegression(thresh=...): prediction / RMSE to 1e-7 relative; exceedance probability to 1e-6
    against the tightly converged reference and to 2e-3 against the reference's default lbfgs run
    (stopped at tol = 1e-4 by scikit-learn, so only defined to that level)."""
    g = golden(name)
    th, C_reg = float(g['thresh']), float(g['C_reg'])
    kw = {} if C_reg == 1.0 else {'logistic_kwargs': {'C': C_reg}}
    m = pm().AnalogRegression(n_analogs=k, thresh=th, **kw)
    m.fit(pd.DataFrame(g['Xtr'][..., 0]), pd.DataFrame(g['ytr'][:, 0]))
    o = m.predict(pd.DataFrame(g['Xq'][..., 0])).values
    ref, tight = g['out64'][:, :, 0], g['out64_tight'][:, :, 0]
    np.testing.assert_allclose(o[:, [0, 2]], ref[:, [0, 2]], rtol=1e-7, atol=1e-8)
    np.testing.assert_allclose(o[:, 1], tight[:, 1], rtol=0, atol=1e-6)
    np.testing.assert_allclose(o[:, 1], ref[:, 1], rtol=0, atol=2e-3)
    # batched wrapper path (float32 in / out) against the oracle
    pw = pm().PointWiseDownscaler(pm().AnalogRegression(n_analogs=k, thresh=th, **kw))
    pw.fit(g['Xtr'], g['ytr'])
    got = pw.predict(g['Xq'])
    want = oracle.analog_regression_predict(g['Xtr'][..., 0], g['ytr'][:, 0], g['Xq'][..., 0], k, thresh=th,
                                            logistic_C=C_reg)
    np.testing.assert_allclose(got[:, :, 0], want.astype(np.float32), rtol=2e-5, atol=2e-6)


def test_analog_regression_thresh_one_class_raises(dev):
    """All analogs of a query at or below thresh: the reference's LogisticRegression raises (gard.py:208-209)."""
    Xtr, ytr, Xq = synth.analog(200, 20, 1, 3, seed=3)
    m = pm().AnalogRegression(n_analogs=5, thresh=1e6).fit(pd.DataFrame(Xtr[..., 0]), pd.DataFrame(ytr[:, 0]))
    with pytest.raises(ValueError, match='at least 2 classes'):
        m.predict(pd.DataFrame(Xq[..., 0]))


@pytest.mark.parametrize('name', ['pure_regression', 'pure_regression_thresh', 'pure_regression_f64_thresh'])
def test_pure_regression_golden(dev, golden, name):
    """PureRegression (gard.py:367-504): prediction and RMSE to 1e-5 of the target spread (the reference
    fits float32 inputs in float32), 1e-9 for float64 inputs; exceedance probability 1e-6 against the
    tightly converged reference and 2e-3 against its default lbfgs run."""
    g = golden(name)
    th = float(g['thresh']) if 'thresh' in g else None
    f64 = g['Xtr'].dtype == np.float64
    pw = pm().PointWiseDownscaler(pm().PureRegression(thresh=th))
    pw.fit(g['Xtr'], g['ytr'])
    got = pw.predict(g['Xq'])
    assert got.shape == g['out'].shape and got.dtype == g['Xq'].dtype
    tol = 1e-9 if f64 else 1e-5 * np.std(g['ytr'])
    np.testing.assert_allclose(got[:, [0, 2]], g['out'][:, [0, 2]], rtol=0, atol=tol)
    if th is None:
        assert (got[:, 1] == 1.0).all()
    else:
        np.testing.assert_allclose(got[:, 1], g['prob_tight'], rtol=0, atol=1e-6)
        np.testing.assert_allclose(got[:, 1], g['out'][:, 1], rtol=0, atol=2e-3)
    # per-cell estimator API
    m = pm().PureRegression(thresh=th).fit(pd.DataFrame(g['Xtr'][..., 0]), pd.DataFrame(g['ytr'][:, 0]))
    o = m.predict(pd.DataFrame(g['Xq'][..., 0]))
    assert list(o.columns) == ['pred', 'exceedance_prob', 'prediction_error']
    np.testing.assert_allclose(o.values[:, [0, 2]], g['out'][:, [0, 2], 0], rtol=0, atol=tol)
    want = oracle.pure_regression_fit_predict(g['Xtr'][..., 0].astype(np.float64), g['ytr'][:, 0].astype(np.float64),
                                              g['Xq'][..., 0].astype(np.float64), th)
    np.testing.assert_allclose(o.values, want, rtol=1e-9, atol=1e-9)       # the float64 arithmetic itself
    assert abs(m.fit_error_ - want[0, 2]) <= 1e-9


def test_pure_regression_no_exceeding_row_raises(dev):
    Xtr, ytr, Xq = synth.analog(100, 10, 2, 3, seed=8)
    pw = pm().PointWiseDownscaler(pm().PureRegression(thresh=1e9))
    with pytest.raises(ValueError, match='0 sample'):                    # gard.py:435
        pw.fit(Xtr, ytr)


@pytest.mark.parametrize('p,k,T,Tq', [(1, 5, 700, 300), (3, 10, 2500, 600), (3, 16, 1100, 257), (3, 17, 900, 100),
                                      (4, 10, 1030, 90), (6, 8, 600, 70), (3, 200, 1500, 64)])
def test_analog_indices_bit_exact(dev, p, k, T, Tq):
    """kNN indices against the float64 brute-force oracle: bit-exact, every k / p code path."""
    C = 3
    Xtr, ytr, Xq = synth.analog(T, Tq, C, p, seed=100 + p + k)
    m = pm().AnalogRegression(n_analogs=k)
    m.fit_batched(eng().as_device(Xtr, dev), eng().as_device(ytr, dev))
    out, idx = m.predict_batched(eng().as_device(Xq, dev), out_dtype=torch.float64, want_idx=True)
    out, idx = out.cpu().numpy(), idx.cpu().numpy()
    for c in range(C):
        o, inds = oracle.analog_regression_predict(Xtr[..., c], ytr[:, c], Xq[..., c], k, return_inds=True)
        assert np.array_equal(idx[:, :, c], inds), f'kNN index mismatch cell {c}'
        np.testing.assert_allclose(out[:, :, c], o, rtol=1e-7, atol=1e-8)


def test_analog_masked_cells_and_small_train(dev):
    Xtr, ytr, Xq = synth.analog(6, 20, 4, 3, seed=5)
    Xtr[:, :, 2] = np.nan
    pw = pm().PointWiseDownscaler(pm().PureAnalog(n_analogs=10, kind='mean_analogs'))
    with pytest.warns(UserWarning, match='less than n_analogs'):      # gard.py:75-79
        pw.fit(Xtr, ytr)
    got = pw.predict(Xq)
    assert np.isnan(got[:, :, 2]).all()
    ref = oracle.pointwise_fit_predict({'name': 'PureAnalog', 'n_analogs': 10, 'kind': 'mean_analogs'}, Xtr, ytr, Xq)
    assert_close(got, ref, scale=1.0)


# ------------------------------------------------------------------ streamed host pipeline
@pytest.mark.parametrize('model_name', ['T', 'P'])
def test_streamed_host_pipeline_matches_whole_block(dev, model_name):
    """host arrays streamed through the GPU in cell chunks (copy / compute / copy-back overlapped)
    give exactly the field of the whole-block path; ragged last chunk, NaN cells, pinned `out`."""
    T, C = 1461, 77
    idx = synth.daily_index(T)
    if model_name == 'T':
        Xtr, ytr, Xp = synth.temperature(T, C, seed=31)
        mk = lambda: pm().BcsdTemperature()          # noqa: E731
    else:
        Xtr, ytr, Xp = synth.precipitation(T, C, seed=32)
        mk = lambda: pm().BcsdPrecipitation()        # noqa: E731
    for c in (3, 40, 76):
        Xtr[:, c] = np.nan
    whole = pm().PointWiseDownscaler(mk(), chunk_cells=1 << 30)
    whole.fit(Xtr, ytr, time=idx)
    ref = whole.predict(Xp, time=idx)
    streamed = pm().PointWiseDownscaler(mk(), chunk_cells=16)
    streamed.fit(Xtr, ytr, time=idx)
    got = streamed.predict(Xp, time=idx)
    assert isinstance(got, np.ndarray) and got.dtype == Xp.dtype and got.shape == ref.shape
    np.testing.assert_array_equal(got, ref)
    # torch host tensors + caller-provided pinned output buffer, 3-D cell shape
    out = torch.empty((T, C), dtype=torch.float32, pin_memory=True)
    streamed.fit(torch.from_numpy(Xtr).pin_memory(), torch.from_numpy(ytr).pin_memory(), time=idx)
    got2 = streamed.predict(torch.from_numpy(Xp).pin_memory(), time=idx, out=out)
    assert isinstance(got2, torch.Tensor) and got2.data_ptr() == out.data_ptr()
    np.testing.assert_array_equal(got2.numpy(), ref)
    # a NaN inside an unmasked cell is still an error (base.py:18-20)
    bad = ytr.copy()
    bad[7, 50] = np.nan
    with pytest.raises(ValueError, match='NaN'):
        streamed.fit(Xtr, bad, time=idx)


# ------------------------------------------------------------------ randomized sweep
@pytest.mark.parametrize('seed', range(6))
def test_random_small_cases_vs_oracle(dev, seed):
    """random shapes / lengths / ties / NaN cells / dtypes through every BCSD + QM entry point;
    ranks bit-exact, values 1e-5, all kernel families."""
    rng = np.random.default_rng(1000 + seed)
    years = int(rng.integers(1, 9))
    Tf = 365 * years + int(rng.integers(0, 40))
    Tp = int(rng.integers(20, 3 * Tf))
    C = int(rng.integers(1, 20))
    dtype = np.float32 if seed % 3 else np.float64
    idx_f = synth.daily_index(Tf, '1983-03-05')
    idx_p = synth.daily_index(Tp, '1990-11-17')
    Xtr, ytr, _ = synth.temperature(Tf, C, seed=seed, dtype=dtype)
    _, _, Xp = synth.temperature(Tp, C, seed=seed + 50, dtype=dtype)
    if seed % 2:                                   # quantised data: plenty of exact ties
        Xp = (np.round(Xp * 4) / 4).astype(dtype)
        ytr = (np.round(ytr * 2) / 2).astype(dtype)
    nan_cells = [c for c in range(C) if rng.random() < 0.15]
    for c in nan_cells:
        Xtr[:, c] = np.nan
    fit_g = oracle.groups_from_keys(oracle.month_keys(idx_f))
    pred_g = oracle.groups_from_keys(oracle.month_keys(idx_p))
    for family in ('tile', 'generic', 'scalar'):
        with force_generic(family):
            for anoms in (True, False):
                m = pm().BcsdTemperature(return_anoms=anoms)
                xtr = eng().as_device(Xtr, dev)
                m.fit_batched(xtr, eng().as_device(ytr, dev), idx_f, valid=eng().cell_mask(xtr[0]))
                out, rank = m.predict_batched(eng().as_device(Xp, dev), idx_p, want_rank=True)
                out, rank = out.cpu().numpy(), rank.cpu().numpy()
                for c in range(C):
                    if c in nan_cells:
                        assert np.isnan(out[:, c]).all()
                        continue
                    st = oracle.bcsd_temperature_fit(Xtr[:, c], ytr[:, c], fit_g)
                    o, r = oracle.bcsd_temperature_predict(st, Xp[:, c], pred_g, pred_g, anoms, return_rank=True)
                    assert np.array_equal(rank[:, c], r), (family, anoms, c)
                    assert_close(out[:, c], o.astype(dtype), scale=np.std(ytr[:, c]))
            # whole-series QuantileMapper on the same data
            q = pm().QuantileMapper()
            q.fit_batched(eng().as_device(ytr, dev))
            qo, qr = q.transform_batched(eng().as_device(Xp, dev), want_rank=True)
            qo, qr = qo.cpu().numpy(), qr.cpu().numpy()
            for c in range(0, C, 3):
                o, r = oracle.quantile_mapper_transform(Xp[:, c], oracle.quantile_mapper_fit(ytr[:, c]), return_rank=True)
                assert np.array_equal(qr[:, c], r)
                assert_close(qo[:, c], o.astype(dtype), scale=np.std(ytr[:, c]))
