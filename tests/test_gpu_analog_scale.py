"""GPU parity of the analog search at the BASELINE.json shapes (T_fit = 18 250 / 10 950, 3 predictors) and of its
exact pruning (csrc/analog_kernels.cu, csrc/qm_long.cu::series_argsort_kernel):

* kNN indices bit-exact against ``oracle.knn_bruteforce`` (float64 brute force, the reference's KDTree order,
  gard.py:82,194,299) at the full window lengths, on cells at both ends of a 4 096-cell block,
* the pruned search against the brute-force search of the same library: identical indices AND outputs for every
  model kind, including inputs without pruning power (constant first predictor), duplicated training rows
  (exact distance ties → lowest training index first) and a query set that lies far outside the training cloud,
* ``sdb_series_argsort`` against ``np.argsort(kind='stable')``.
Nothing here reads /root/reference."""

import numpy as np
import pytest
import torch

import oracle
import synth
from test_gpu_parity import eng, pm

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def dev():
    assert torch.cuda.is_available(), 'GPU tests need a CUDA device'
    import skdownscale_b200  # noqa: F401
    return torch.device('cuda:0')


def _lib():
    from skdownscale_b200 import _lib
    return _lib


@pytest.mark.parametrize('n', [1, 2, 31, 1000, 4097, 10950, 18250, 32768])
def test_series_argsort_matches_numpy_stable(dev, n):
    rng = np.random.default_rng(n)
    C = 6
    x = rng.standard_normal((n, 3, C)).astype(np.float32)
    if n > 10:
        x[:, 0, 1] = np.round(x[:, 0, 1] * 4) / 4          # heavy ties
        x[:, 0, 2] = 1.5                                   # constant
        x[: n // 2, 0, 3] = x[0, 0, 3]                     # half the series equal to one value
        x[::7, 0, 4] = x[:: 7, 0, 4].min()                 # many copies of the minimum (the lower-bound class)
        x[:, 0, 5] = np.sort(x[:, 0, 5])[::-1]             # descending
    xd = eng().as_device(x, dev)
    order = eng().series_argsort(xd, 3 * C, n, C).cpu().numpy()
    for c in range(C):
        assert np.array_equal(order[:, c], np.argsort(x[:, 0, c], kind='stable')), f'cell {c}'


def _both(model_factory, Xtr, ytr, Xq, dev, **kw):
    """predict with the brute-force search and with the pruned search: (out, idx) of each."""
    res = []
    for prune in (False, True):
        m = model_factory()
        m.fit_batched(eng().as_device(Xtr, dev), eng().as_device(ytr, dev))
        if not prune:
            m._order_train = None
        else:
            assert m._order_train is not None, 'the training order was not built at fit time'
        out, idx = m.predict_batched(eng().as_device(Xq, dev), out_dtype=torch.float64, want_idx=True, **kw)
        m._check_finite()
        res.append((out.cpu().numpy(), idx.cpu().numpy()))
    return res


CASES = ['normal', 'const_x0', 'duplicates', 'far_queries', 'clustered', 'p1', 'p2', 'k16']


@pytest.mark.parametrize('case', CASES)
def test_pruned_search_equals_brute_force(dev, case):
    p, k, T, Tq, C = 3, 10, 4200, 1500, 5
    if case == 'p1':
        p = 1
    if case == 'p2':
        p = 2
    if case == 'k16':
        k = 16
    Xtr, ytr, Xq = synth.analog(T, Tq, C, p, seed=7 + len(case))
    rng = np.random.default_rng(3)
    if case == 'const_x0':
        Xtr[:, 0, :] = 0.25
        Xq[:, 0, :] = 0.25
    elif case == 'duplicates':
        Xtr[1::2] = Xtr[0:-1:2]                        # every training row twice: exact distance ties everywhere
        Xq[::5] = Xtr[: len(Xq[::5])]                  # and queries that coincide with training rows
    elif case == 'far_queries':
        Xq[:, 0, :] += 25.0
        Xq[::3, 0, :] -= 60.0
    elif case == 'clustered':
        Xtr[:, 0, :] = np.round(Xtr[:, 0, :])          # first predictor on a coarse grid
        Xq[:, 0, :] = np.round(Xq[:, 0, :] * 2) / 2
    for kind in (['regression'] if case in ('k16',) else ['regression', 'mean_analogs', 'weight_analogs', 'best_analog']):
        if kind == 'regression':
            mk = lambda: pm().AnalogRegression(n_analogs=k)     # noqa: E731
        else:
            mk = lambda: pm().PureAnalog(n_analogs=k, kind=kind)   # noqa: E731
        (o_b, i_b), (o_p, i_p) = _both(mk, Xtr, ytr, Xq, dev)
        assert np.array_equal(i_b, i_p), f'{case}/{kind}: neighbour indices differ'
        assert np.array_equal(o_b, o_p, equal_nan=True), f'{case}/{kind}: outputs differ'
        if case == 'duplicates' and kind == 'regression':   # the documented tie rule: lowest training index first
            _, inds = oracle.knn_bruteforce(Xtr[..., 0], Xq[..., 0], k)
            assert np.array_equal(i_p[:, :, 0], inds)


def test_pruned_search_with_thresh_and_masked_cells(dev):
    T, Tq, C = 3000, 700, 6
    Xtr, ytr, Xq = synth.analog(T, Tq, C, 3, seed=77)
    Xtr[:, :, 4] = np.nan
    valid = torch.tensor([1, 1, 1, 1, 0, 1], dtype=torch.uint8, device=dev)
    res = []
    for prune in (False, True):
        m = pm().AnalogRegression(n_analogs=20, thresh=-0.5)
        m.fit_batched(eng().as_device(Xtr, dev), eng().as_device(ytr, dev), valid=valid)
        if not prune:
            m._order_train = None
        res.append(m.predict_batched(eng().as_device(Xq, dev), out_dtype=torch.float64).cpu().numpy())
    assert np.array_equal(res[0], res[1], equal_nan=True)
    assert np.isnan(res[1][:, :, 4]).all()


@pytest.mark.parametrize('name,T,k', [('PureAnalog', 18250, 10), ('AnalogRegression', 10950, 10), ('AnalogRegression', 10950, 200)])   # k = 200: brute force
def test_analog_indices_at_baseline_window_lengths(dev, name, T, k):
    """BASELINE configs 4 / 5: 50-year and 30-year daily windows, 3 predictors.  A 4 096-cell block on the device;
    the kNN indices of cells 0, 1, 2047 and 4095 (first / last cell of the block) are compared bit for bit with the
    float64 brute-force oracle on a sample of query steps, the outputs with the oracle's (1e-7)."""
    C, Tq = 4096, 320
    rng = np.random.default_rng(T + k)
    gen = torch.Generator(device=dev).manual_seed(T + k)
    X = torch.randn((T, 3, C), device=dev, generator=gen)
    w = torch.tensor([1.0, 0.5, -0.3], device=dev)[None, :, None]
    y = (X * w).sum(1) + 0.3 * torch.randn((T, C), device=dev, generator=gen)
    Xq = torch.randn((Tq, 3, C), device=dev, generator=gen)
    cells = [0, 1, 2047, 4095]
    if name == 'PureAnalog':
        for kind in ('mean_analogs', 'best_analog', 'weight_analogs'):
            m = pm().PureAnalog(n_analogs=k, kind=kind).fit_batched(X, y)
            assert m._order_train is not None
            out, idx = m.predict_batched(Xq, out_dtype=torch.float64, want_idx=True)
            kk = 1 if kind == 'best_analog' else k
            for c in cells:
                xa, ya, xq = X[:, :, c].cpu().numpy(), y[:, c].cpu().numpy(), Xq[:, :, c].cpu().numpy()
                _, inds = oracle.knn_bruteforce(xa, xq, kk)
                assert np.array_equal(idx[:, :, c].cpu().numpy(), inds), f'{kind}: kNN index mismatch in cell {c}'
                ref = oracle.pure_analog_predict(xa, ya, xq, k, kind)
                np.testing.assert_allclose(out[:, :, c].cpu().numpy(), ref, rtol=1e-6, atol=1e-7)
    else:
        m = pm().AnalogRegression(n_analogs=k).fit_batched(X, y)
        out, idx = m.predict_batched(Xq, out_dtype=torch.float64, want_idx=True)
        for c in cells:
            xa, ya, xq = X[:, :, c].cpu().numpy(), y[:, c].cpu().numpy(), Xq[:, :, c].cpu().numpy()
            ref, inds = oracle.analog_regression_predict(xa, ya, xq, k, return_inds=True)
            assert np.array_equal(idx[:, :, c].cpu().numpy(), inds), f'kNN index mismatch in cell {c}'
            np.testing.assert_allclose(out[:, :, c].cpu().numpy(), ref, rtol=1e-7, atol=1e-8)
    m._check_finite()
