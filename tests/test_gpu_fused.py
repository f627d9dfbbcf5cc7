"""GPU parity tests of the fused fit+predict entry (``sdb_bcsd_fit_predict``, csrc/qm_fused.cuh): the
counting-rank kernel must give BIT-IDENTICAL fields to the split path (``sdb_qm_fit`` + ``sdb_qm_predict``,
itself held to the oracle / the live-reference goldens in test_gpu_parity.py) for every input — including
the inputs built to defeat its bucket quantisation — and the oracle's values within 1e-5.
Nothing here reads /root/reference."""

import numpy as np
import pytest
import torch

import oracle
import synth
from test_gpu_parity import assert_close, eng, pm, _expected_rank_map

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def dev():
    assert torch.cuda.is_available(), 'GPU tests need a CUDA device'
    import skdownscale_b200  # noqa: F401
    return torch.device('cuda:0')


class debug_flags:
    def __init__(self, flags):
        self.flags = flags

    def __enter__(self):
        from skdownscale_b200 import _lib
        self.old = _lib.load().sdb_set_debug_flags(self.flags)

    def __exit__(self, *a):
        from skdownscale_b200 import _lib
        _lib.load().sdb_set_debug_flags(self.old)


def _model(name, anoms):
    if name == 'T':
        return pm().BcsdTemperature(return_anoms=anoms)
    return pm().BcsdPrecipitation(return_anoms=anoms)


def _split(name, anoms, Xtr, ytr, Xp, idx, valid=None):
    m = _model(name, anoms)
    m.fit_batched(Xtr, ytr, idx, valid=valid)
    return m.predict_batched(Xp, idx), m


def _fused(name, anoms, Xtr, ytr, Xp, idx, valid=None, keep_state=True, stats=None):
    m = _model(name, anoms)
    out = m.fit_predict_batched(Xtr, ytr, Xp, idx, valid=valid, keep_state=keep_state, stats=stats, fused=True)
    return out, m


def _same_state(st_a, st_b, what, sel=None):
    """fitted sorted values of every group (the 16-byte alignment gaps between groups are never written)."""
    a, b = st_a.sorted_state, st_b.sorted_state
    if sel is not None:
        a, b = a[sel], b[sel]
    assert np.array_equal(st_a.state_off, st_b.state_off) and st_a.state_ld == st_b.state_ld
    for off, n in zip(st_a.state_off, st_a.sort_table.len):
        _same(a[:, off:off + n], b[:, off:off + n], what)


def _same(a, b, what):
    a, b = a.cpu().numpy(), b.cpu().numpy()
    assert np.array_equal(a, b, equal_nan=True), f'{what}: {np.sum(~((a == b) | (np.isnan(a) & np.isnan(b))))} elements differ'


@pytest.mark.parametrize('flags', [0, 2, 4, 8, 12])
@pytest.mark.parametrize('anoms', [True, False])
@pytest.mark.parametrize('name', ['T', 'P'])
def test_fused_equals_split(dev, name, anoms, flags):
    """30-year daily series, ragged cell count (not a multiple of the 8-cell tile), NaN cells.  flags: 2 = 4-byte
    row accesses, 4 / 8 / 12 = training / prediction / both sides on the sorting-network path."""
    T, C = 10950, 43
    idx = synth.daily_index(T)
    Xtr, ytr, Xp = synth.temperature(T, C, seed=3) if name == 'T' else synth.precipitation(T, C, seed=5)
    for c in (4, 42):
        Xtr[:, c] = np.nan
    d = lambda a: eng().as_device(a, dev)   # noqa: E731
    xtr, yt, xp = d(Xtr), d(ytr), d(Xp)
    valid = eng().cell_mask(xtr[0])
    ref, m_ref = _split(name, anoms, xtr, yt, xp, idx, valid)
    m_ref._state.check_finite()
    stats = torch.zeros(8, dtype=torch.int64, device=dev)
    with debug_flags(flags):
        got, m = _fused(name, anoms, xtr, yt, xp, idx, valid, stats=stats)
    m._state.check_finite()
    _same(got, ref, f'{name} anoms={anoms} flags={flags}')
    # the call leaves the same fitted model behind as fit()
    sel = valid.bool()
    _same_state(m._state, m_ref._state, 'fitted state', sel)
    _same(m._state.y_climo, m_ref._state.y_climo, 'y_climo')
    if name == 'T':
        _same(m._state.x_climo, m_ref._state.x_climo, 'x_climo')
    _same(m.predict_batched(xp, idx), ref, 'predict after the fused call')
    st = stats.cpu().numpy()
    assert st[0] == 12 * (C - 2)
    if flags in (0, 2):
        # the network path is the exception: a prediction series takes it when two members of one bucket agree in
        # all 16 fraction bits of their quantised key (about one series in a hundred)
        assert st[1] <= 0.02 * st[0] and st[2] <= 0.03 * st[0], f'counting rank fell back to the network: {st}'
    if flags == 12:
        assert st[1] == st[0] and st[2] == st[0]


def test_fused_vs_oracle(dev):
    """Values against the numpy oracle directly (1e-5 of max(|ref|, sigma_y)), both BCSD models."""
    T, C = 10950, 16
    idx = synth.daily_index(T)
    groups = oracle.groups_from_keys(oracle.month_keys(idx))
    d = lambda a: eng().as_device(a, dev)   # noqa: E731
    Xtr, ytr, Xp = synth.temperature(T, C, seed=13)
    out, _ = _fused('T', True, d(Xtr), d(ytr), d(Xp), idx)
    out = out.cpu().numpy()
    for c in range(0, C, 3):
        st = oracle.bcsd_temperature_fit(Xtr[:, c], ytr[:, c], groups)
        o = oracle.bcsd_temperature_predict(st, Xp[:, c], groups, groups, True)
        assert_close(out[:, c], o.astype(np.float32), scale=np.std(ytr[:, c]))
    Xtr, ytr, Xp = synth.precipitation(T, C, seed=14)
    out, m = _fused('P', True, d(Xtr), d(ytr), d(Xp), idx)
    m.check_fit()
    out = out.cpu().numpy()
    for c in range(0, C, 3):
        st = oracle.bcsd_precipitation_fit(ytr[:, c], groups)
        o = oracle.bcsd_precipitation_predict(st, Xp[:, c], groups, True)
        assert_close(out[:, c], o.astype(np.float32), scale=1.0)


CASES = ['outlier', 'clusters', 'constant', 'two_values', 'zeros_and_tiny', 'ties_and_pairs', 'grid_0p1', 'dense_pairs']


@pytest.mark.parametrize('case', CASES)
@pytest.mark.parametrize('side', ['x', 'y', 'both'])
@pytest.mark.parametrize('name', ['T', 'P'])
def test_fused_adversarial_inputs(dev, name, side, case):
    """Inputs built to defeat the bucket quantisation of the counting rank (a huge outlier squeezing every
    other value into a few buckets, clusters one float32 ulp apart, constant and two-valued series, values
    on a coarse grid = heavy exact ties, zero-inflation with tiny positives): whatever path a series takes
    (dirty-entry queue or network fallback) the field must equal the split path's bit for bit."""
    T, C = 2922, 11
    idx = synth.daily_index(T)
    rng = np.random.default_rng(123)
    Xtr, ytr, Xp = synth.temperature(T, C, seed=21)

    def spoil(A):
        if case == 'outlier':
            A[::365, :] = 1.0e9
            A[100::400, 1] = -5.0e8
        elif case == 'clusters':
            base = np.float32(12.5)
            for c in range(C):
                k = rng.integers(0, 40, T)
                A[:, c] = (base.view(np.int32) + k.astype(np.int32)).view(np.float32)
            A[5::7, 0] += 3.0
        elif case == 'constant':
            A[:] = np.float32(7.25)
        elif case == 'two_values':
            A[:] = np.where(rng.random((T, C)) < 0.6, np.float32(0.0), np.float32(1e-7))
        elif case == 'zeros_and_tiny':
            wet = rng.random((T, C)) > 0.55
            A[:] = np.where(wet, rng.gamma(0.8, 6.0, (T, C)), 0.0).astype(np.float32)
            A[3::97, :] = np.float32(3e-6)
            A[5::101, 2] = np.float32(7e-6)
        elif case == 'ties_and_pairs':
            A[:] = (np.round(A * 2) / 2).astype(np.float32)
            near = (np.float32(11.25).view(np.int32) + np.arange(1, 6, dtype=np.int32)).view(np.float32)
            for k, v in enumerate(near):
                A[40 + 31 * k::360, 1::2] = v
        elif case == 'grid_0p1':
            A[:] = (np.round(A * 10) / 10).astype(np.float32)        # station-data resolution: ties everywhere
        elif case == 'dense_pairs':
            # every value has a partner one or two ulps away: every occupied bucket is contested
            half = A[: T // 2].copy()
            A[: 2 * (T // 2) : 2] = half
            A[1 : 2 * (T // 2) : 2] = (half.view(np.int32) + rng.integers(1, 3, half.shape).astype(np.int32)).view(np.float32)

    if side in ('x', 'both'):
        spoil(Xp)
    if side in ('y', 'both'):
        spoil(ytr)
    d = lambda a: eng().as_device(a, dev)   # noqa: E731
    xtr, yt, xp = d(Xtr), d(ytr), d(Xp)
    ref, m_ref = _split(name, False, xtr, yt, xp, idx)
    got, m = _fused(name, False, xtr, yt, xp, idx)
    _same(got, ref, f'{name}/{side}/{case}')
    _same_state(m._state, m_ref._state, f'{name}/{side}/{case}: fitted state')


@pytest.mark.parametrize('T', [31, 59, 400, 1461, 2922])
def test_fused_short_and_ragged_groups(dev, T):
    """Group lengths from 1 member up (T = 31: January only; 59: two groups), tile-boundary cell counts."""
    idx = synth.daily_index(T)
    for C in (1, 7, 8, 9, 17):
        Xtr, ytr, Xp = synth.temperature(T, C, seed=100 + C)
        d = lambda a: eng().as_device(a, dev)   # noqa: E731
        for name in ('T', 'P'):
            ref, _ = _split(name, False, d(Xtr), d(ytr), d(Xp), idx)
            got, _ = _fused(name, False, d(Xtr), d(ytr), d(Xp), idx)
            _same(got, ref, f'T={T} C={C} {name}')


def test_fused_strided_views_and_no_state(dev):
    """Inputs / output that are column slices of wider arrays (ld > C, rows not 16-byte aligned) and
    keep_state=False (nothing but the output and the climatologies is written)."""
    T, C = 2000, 21
    idx = synth.daily_index(T)
    Xtr, ytr, Xp = synth.temperature(T, C + 6, seed=11)
    d = lambda a: eng().as_device(a, dev)[:, 3:3 + C]   # noqa: E731
    xtr, yt, xp = d(Xtr), d(ytr), d(Xp)
    ref, m_ref = _split('T', True, xtr, yt, xp, idx)
    wide = torch.full((T, C + 6), -1.0, dtype=torch.float32, device=dev)
    m = pm().BcsdTemperature(return_anoms=True)
    got = m.fit_predict_batched(xtr, yt, xp, idx, out=wide[:, 3:3 + C], keep_state=False, fused=True)
    _same(got, ref, 'strided')
    assert bool((wide[:, :3] == -1).all()) and bool((wide[:, 3 + C:] == -1).all()), 'wrote outside its columns'
    assert not hasattr(m, '_state')
    _same(m._climo_state.y_climo, m_ref._state.y_climo, 'y_climo')


def test_fused_nonfinite_is_flagged(dev):
    T, C = 1461, 9
    idx = synth.daily_index(T)
    Xtr, ytr, Xp = synth.temperature(T, C, seed=1)
    Xp[700, 3] = np.nan
    d = lambda a: eng().as_device(a, dev)   # noqa: E731
    _, m = _fused('T', True, d(Xtr), d(ytr), d(Xp), idx)
    with pytest.raises(ValueError, match='NaN or infinity'):
        m._state.check_finite()
    Xtr, ytr, Xp = synth.temperature(T, C, seed=1)
    ytr[5, 0] = np.inf
    _, m = _fused('P', False, d(Xtr), d(ytr), d(Xp), idx)
    with pytest.raises(ValueError, match='NaN or infinity'):
        m._state.check_finite()


def test_fused_rank_map_property_block(dev):
    """Pure per-month quantile map of X onto y's distribution on a larger block: every element equals the
    fitted order statistic of its month at its tie-max rank — exact, expectation computed with torch."""
    T, C = 10950, 2048
    idx = synth.daily_index(T)
    gen = torch.Generator(device=dev).manual_seed(7)
    s = torch.sin(2 * torch.pi * torch.arange(T, device=dev) / 365.25)[:, None]
    X = 15 + 10 * s + 3 * torch.randn((T, C), device=dev, generator=gen)
    y = 14 + 12 * s + 2 * torch.randn((T, C), device=dev, generator=gen)
    stats = torch.zeros(8, dtype=torch.int64, device=dev)
    m = pm().BcsdPrecipitation(return_anoms=False)
    out = m.fit_predict_batched(X, y, X, idx, stats=stats, fused=True)
    month = torch.as_tensor(np.asarray(idx.month), device=dev)
    for mo in (1, 2, 7, 12):
        sel = month == mo
        torch.testing.assert_close(out[sel], _expected_rank_map(X[sel], y[sel]), rtol=0, atol=0)
    st = stats.cpu().numpy()
    assert st[1] <= 0.01 * st[0] and st[2] <= 0.01 * st[0], f'Gaussian fields must stay on the counting-rank path: {st}'
