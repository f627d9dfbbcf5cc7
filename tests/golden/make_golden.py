"""Generate the golden vectors under tests/golden/ by running the LIVE reference.

Run in the build container only (``/root/reference`` is not on the GPU box):

    python tests/golden/make_golden.py

The reference package ``__init__`` imports xarray (absent here), so the
per-cell estimator modules are imported with stub parent packages (SURVEY.md
Appendix C); the ``PointWiseDownscaler`` cell loop (core.py:69-143) is restated
in ``_cell_loop`` below (deepcopy → fit → predict → squeeze → cast to X.dtype).
Versions used are recorded in tests/golden/VERSIONS.json.
"""

from __future__ import annotations

import copy
import importlib
import json
import os
import sys
import types
import warnings

import numpy as np
import pandas as pd

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import synth  # noqa: E402

REF = '/root/reference/skdownscale'


def load_reference():
    for name, path in [('skdownscale', REF), ('skdownscale.pointwise_models', REF + '/pointwise_models')]:
        m = types.ModuleType(name)
        m.__path__ = [path]
        sys.modules[name] = m
    mods = {}
    for sub in ('bcsd', 'quantile', 'gard', 'groupers'):
        mods[sub] = importlib.import_module(f'skdownscale.pointwise_models.{sub}')
    return mods


def _df(a, index=None):
    a = np.asarray(a)
    if a.ndim == 1:
        a = a[:, None]
    cols = [f'f{i}' for i in range(a.shape[1])]
    return pd.DataFrame(a, columns=cols, index=index)


def _cell_loop(model, Xtr, ytr, Xp, index_fit=None, index_pred=None, n_outputs=1):
    """core.py:69-143 for arrays [T, C] or [T, p, C]."""
    C = Xp.shape[-1]
    Tp = Xp.shape[0]
    out = np.full((Tp, n_outputs, C) if n_outputs > 1 else (Tp, C), np.nan, dtype=Xp.dtype)
    for c in range(C):
        first = Xtr[0, 0, c] if Xtr.ndim == 3 else Xtr[0, c]
        if np.isnan(first):
            continue
        mod = copy.deepcopy(model)
        xdf = _df(Xtr[..., c], index_fit)
        ydf = _df(ytr[:, c], index_fit)
        mod.fit(xdf, ydf)
        res = mod.predict(_df(Xp[..., c], index_pred))
        res = np.asarray(res).squeeze()
        if n_outputs > 1:
            out[:, :, c] = res
        else:
            out[:, c] = res
    return out


ONLY = [a for a in sys.argv[1:]]          # optional name prefixes: regenerate just those files


def save(name, **arrays):
    if ONLY and not any(name.startswith(o) for o in ONLY):
        return
    np.savez_compressed(os.path.join(HERE, name + '.npz'), **arrays)
    print('wrote', name, {k: getattr(v, 'shape', None) for k, v in arrays.items()})


def main():
    warnings.simplefilter('ignore')
    ref = load_reference()
    QuantileMapper = ref['quantile'].QuantileMapper
    BcsdTemperature = ref['bcsd'].BcsdTemperature
    BcsdPrecipitation = ref['bcsd'].BcsdPrecipitation
    PureAnalog = ref['gard'].PureAnalog
    AnalogRegression = ref['gard'].AnalogRegression
    PaddedDOYGrouper = ref['groupers'].PaddedDOYGrouper

    # --- 1. the reference's own known-answer test (test_pointwise_models.py:81-90)
    n = 100
    expected = (np.sin(np.linspace(-10 * np.pi, 10 * np.pi, n)) * 10).reshape(-1, 1)
    with_bias = expected + 2
    actual = QuantileMapper().fit(expected).transform(with_bias)
    np.testing.assert_almost_equal(actual, expected)
    save('qm_known_answer', fit=expected, x=with_bias, out=actual)

    # --- 2. QuantileMapper on whole series, several length relations, ties, dtypes
    def qm_case(name, Tf, Tp, C, seed, dtype=np.float32, quant=None):
        _, ytr, _ = synth.temperature(Tf, C, seed, dtype)
        _, _, Xp = synth.temperature(Tp, C, seed + 100, dtype)
        if quant:
            ytr = (np.round(ytr / quant) * quant).astype(dtype)
            Xp = (np.round(Xp / quant) * quant).astype(dtype)
        out = np.empty((Tp, C), dtype=np.float64)
        for c in range(C):
            out[:, c] = QuantileMapper().fit(ytr[:, c:c + 1]).transform(Xp[:, c:c + 1])[:, 0]
        save(name, ytr=ytr, Xp=Xp, out=out)

    # non-default CunnaneTransformer settings through qt_kwargs (quantile.py:420-432)
    QT_VARIANTS = {
        'ab': dict(alpha=0.3, beta=0.5, n_endpoints=5),
        'none': dict(extrapolate=None),
        'min': dict(extrapolate='min', n_endpoints=4),
        'max': dict(extrapolate='max', alpha=0.0, beta=1.0),
        '1to1': dict(extrapolate='1to1'),
    }

    def qm_qt_case(tag, qt, Tf, Tp, C, seed):
        _, ytr, _ = synth.temperature(Tf, C, seed)
        _, _, Xp = synth.temperature(Tp, C, seed + 100)
        out = np.empty((Tp, C), dtype=np.float64)
        for c in range(C):
            out[:, c] = QuantileMapper(qt_kwargs=qt).fit(ytr[:, c:c + 1]).transform(Xp[:, c:c + 1])[:, 0]
        save(f'qm_qt_{tag}', ytr=ytr, Xp=Xp, out=out)

    for tag, qt in QT_VARIANTS.items():
        qm_qt_case(tag, qt, 365, 1000, 2, 16)           # T_pred > T_fit: both tails are reached

    # --- 2b. QuantileMappingReressor / EquidistantCdfMatcher, every extrapolate mode
    QMR = ref['quantile'].QuantileMappingReressor
    EDC = ref['quantile'].EquidistantCdfMatcher

    def qmr_case(name, Tf, Tp, C, seed, dtype=np.float32, ne=7, offset=0.0):
        Xtr, ytr, _ = synth.temperature(Tf, C, seed, dtype)
        _, _, Xp = synth.temperature(Tp, C, seed + 100, dtype)
        Xp = (Xp + dtype(offset)).astype(dtype)
        arrays = dict(Xtr=Xtr, ytr=ytr, Xp=Xp, n_endpoints=np.int64(ne))
        for ex in (None, 'min', 'max', 'both', '1to1'):
            tag = 'none' if ex is None else ex
            out = np.empty((Tp, C), dtype=dtype)
            od, orat = np.empty((Tp, C), dtype=dtype), np.empty((Tp, C), dtype=dtype)
            for c in range(C):
                out[:, c] = QMR(extrapolate=ex, n_endpoints=ne).fit(Xtr[:, c:c + 1], ytr[:, c]).predict(Xp[:, c:c + 1])
                od[:, c] = EDC(kind='difference', extrapolate=ex, n_endpoints=ne).fit(Xtr[:, c:c + 1], ytr[:, c]).predict(Xp[:, c:c + 1])
                orat[:, c] = EDC(kind='ratio', extrapolate=ex, n_endpoints=ne).fit(Xtr[:, c:c + 1], ytr[:, c]).predict(Xp[:, c:c + 1])
            arrays[f'qmr_{tag}'], arrays[f'diff_{tag}'], arrays[f'ratio_{tag}'] = out, od, orat
        save(name, **arrays)

    qmr_case('qmr_equal_len', 730, 730, 3, 17)
    qmr_case('qmr_pred_longer_shifted', 500, 1200, 2, 18, offset=1.5)      # values beyond both ends of the fitted range
    qmr_case('qmr_f64_shorter', 900, 400, 2, 19, dtype=np.float64, ne=10)

    # TrendAwareQuantileMappingRegressor over the regressor above (quantile.py:639-716): pinned for the
    # oracle only — the estimator is not in the product yet
    TAQ = ref['quantile'].TrendAwareQuantileMappingRegressor

    def taq_case(name, Tf, Tp, C, seed, dtype=np.float32):
        Xtr, ytr, _ = synth.temperature(Tf, C, seed, dtype)
        _, _, Xp = synth.temperature(Tp, C, seed + 100, dtype)
        Xp = (Xp + np.linspace(0, 4, Tp)[:, None]).astype(dtype)
        arrays = dict(Xtr=Xtr, ytr=ytr, Xp=Xp)
        for ex in (None, '1to1'):
            out = np.empty((Tp, C), dtype=np.float64)
            for c in range(C):
                m = TAQ(QMR(extrapolate=ex, n_endpoints=6)).fit(pd.DataFrame(Xtr[:, c]), pd.DataFrame(ytr[:, c]))
                out[:, c] = m.predict(pd.DataFrame(Xp[:, c]))[:, 0]
            arrays['out_none' if ex is None else 'out_1to1'] = out
        save(name, **arrays)

    taq_case('trend_aware_qmr', 800, 1100, 2, 70)
    taq_case('trend_aware_qmr_f64', 500, 400, 2, 71, dtype=np.float64)

    # the reference's own known-answer test (test_pointwise_models.py:323-344)
    xs = np.arange(1, 22)
    ka = {}
    for kind in ('difference', 'ratio'):
        Xt = xs + 2 if kind == 'difference' else xs * 2
        got = EDC(kind=kind).fit(pd.DataFrame(xs), pd.DataFrame(xs + 3)).predict(pd.DataFrame(Xt))
        assert (got.reshape(-1, 1) == ((xs + 3) + 2 if kind == 'difference' else (xs + 3) * 2).reshape(-1, 1)).all()
        ka[kind] = got
    save('edcdf_known_answer', x=xs, **ka)

    # EquidistantCdfMatcher on exactly tied inputs: np.argsort's unstable order decides which tied step takes which
    # plotting position (quantile.py:607) — pinned as a multiset per tie run (tests/test_oracle_golden.py)
    rng = np.random.default_rng(21)
    n, C = 900, 4
    Xtr_t = (rng.gamma(0.8, 6.0, (n, C)) + 0.5).astype(np.float32)
    ytr_t = (rng.gamma(0.9, 5.0, (n, C)) + 0.5).astype(np.float32)
    Xp_t = np.round(rng.gamma(0.8, 6.0, (n, C)) + 0.5, 0).astype(np.float32) + 1.0
    Xp_t[100:160, 1] = 3.0
    tied = dict(Xtr=Xtr_t, ytr=ytr_t, Xp=Xp_t)
    for kind in ('difference', 'ratio'):
        o = np.empty((n, C), dtype=np.float32)
        for c in range(C):
            o[:, c] = EDC(kind=kind, extrapolate=None).fit(Xtr_t[:, c:c + 1], ytr_t[:, c]).predict(Xp_t[:, c:c + 1])
        tied[kind] = o
    save('edcdf_tied', **tied)

    # --- 2c. detrending mappers: QuantileMapper(detrend=True), BCSD qm_kwargs detrend
    def qm_detrend_case(name, Tf, Tp, C, seed, dtype=np.float32):
        _, ytr, _ = synth.temperature(Tf, C, seed, dtype)
        _, _, Xp = synth.temperature(Tp, C, seed + 100, dtype)
        ytr = (ytr + np.linspace(0, 2, Tf)[:, None]).astype(dtype)         # give the series trends to remove
        Xp = (Xp + np.linspace(-1, 3, Tp)[:, None]).astype(dtype)
        out = np.empty((Tp, C), dtype=np.float64)
        for c in range(C):
            out[:, c] = QuantileMapper(detrend=True).fit(ytr[:, c:c + 1]).transform(Xp[:, c:c + 1])[:, 0]
        save(name, ytr=ytr, Xp=Xp, out=out)

    qm_detrend_case('qm_detrend_equal', 900, 900, 2, 26)
    qm_detrend_case('qm_detrend_longer', 400, 1100, 2, 27)
    qm_detrend_case('qm_detrend_f64', 500, 450, 2, 28, dtype=np.float64)

    qm_case('qm_equal_len', 730, 730, 3, 10)
    qm_case('qm_pred_longer', 365, 1000, 3, 11)      # exercises both OLS tails
    qm_case('qm_pred_shorter', 1000, 300, 3, 12)
    qm_case('qm_ties', 500, 700, 3, 13, quant=0.5)   # heavy ties, max-rank rule
    qm_case('qm_f64', 400, 450, 2, 14, dtype=np.float64)
    qm_case('qm_tiny', 7, 25, 2, 15)                 # fewer fit points than n_endpoints

    # --- 3. BcsdTemperature, monthly groups
    def bcsd_t_case(name, Tf, Tp, C, seed, start_fit='1981-01-01', start_pred=None, dtype=np.float32,
                    nan_cells=(), **kw):
        idx_f = synth.daily_index(Tf, start_fit)
        idx_p = synth.daily_index(Tp, start_pred or start_fit)
        Xtr, ytr, _ = synth.temperature(Tf, C, seed, dtype)
        _, _, Xp = synth.temperature(Tp, C, seed + 100, dtype)
        for c in nan_cells:
            Xtr[:, c] = np.nan
            ytr[:, c] = np.nan
            Xp[:, c] = np.nan
        out = _cell_loop(BcsdTemperature(**kw), Xtr, ytr, Xp, idx_f, idx_p)
        out64 = np.full((Tp, C), np.nan)
        for c in range(C):
            if np.isnan(Xtr[0, c]):
                continue
            m = BcsdTemperature(**kw).fit(_df(Xtr[:, c], idx_f), _df(ytr[:, c], idx_f))
            out64[:, c] = m.predict(_df(Xp[:, c], idx_p)).values[:, 0]
        save(name, Xtr=Xtr, ytr=ytr, Xp=Xp, out=out, out64=out64,
             start_fit=np.array(start_fit), start_pred=np.array(start_pred or start_fit))

    bcsd_t_case('bcsd_t_month_anoms', 1461, 1461, 4, 20, nan_cells=(2,))
    bcsd_t_case('bcsd_t_month_abs', 1461, 1461, 3, 21, return_anoms=False)
    bcsd_t_case('bcsd_t_month_future', 1461, 2192, 3, 22, start_pred='1985-01-01')   # T_pred > T_fit → tails
    bcsd_t_case('bcsd_t_month_future_qt', 1096, 1826, 2, 25, start_pred='1984-01-01',
                qm_kwargs={'qt_kwargs': dict(alpha=0.3, beta=0.5, n_endpoints=5, extrapolate='max')})
    bcsd_t_case('bcsd_t_month_detrend', 1461, 1461, 3, 29, nan_cells=(1,), qm_kwargs={'detrend': True})
    bcsd_t_case('bcsd_t_month_detrend_future', 1096, 1826, 2, 31, start_pred='1984-01-01', return_anoms=False,
                qm_kwargs={'detrend': True})
    bcsd_t_case('bcsd_t_month_f64', 1096, 1096, 2, 23, dtype=np.float64)
    bcsd_t_case('bcsd_t_month_30yr', 10950, 10950, 1, 0)                              # BASELINE config[0]
    bcsd_t_case('bcsd_t_nasanex', 1096, 1096, 2, 24, start_fit='1980-01-01',
                time_grouper='daily_nasa-nex', return_anoms=False)

    # --- 4. BcsdPrecipitation
    def bcsd_p_case(name, Tf, Tp, C, seed, start_fit='1981-01-01', start_pred=None, **kw):
        idx_f = synth.daily_index(Tf, start_fit)
        idx_p = synth.daily_index(Tp, start_pred or start_fit)
        Xtr, ytr, _ = synth.precipitation(Tf, C, seed)
        _, _, Xp = synth.precipitation(Tp, C, seed + 100)
        out = _cell_loop(BcsdPrecipitation(**kw), Xtr, ytr, Xp, idx_f, idx_p)
        out64 = np.empty((Tp, C))
        for c in range(C):
            m = BcsdPrecipitation(**kw).fit(_df(Xtr[:, c], idx_f), _df(ytr[:, c], idx_f))
            out64[:, c] = m.predict(_df(Xp[:, c], idx_p)).values[:, 0]
        save(name, Xtr=Xtr, ytr=ytr, Xp=Xp, out=out, out64=out64,
             start_fit=np.array(start_fit), start_pred=np.array(start_pred or start_fit))

    bcsd_p_case('bcsd_p_month_detrend', 1461, 1461, 2, 32, qm_kwargs={'detrend': True})
    bcsd_p_case('bcsd_p_month_anoms', 1461, 1461, 3, 30)
    bcsd_p_case('bcsd_p_month_30yr', 10950, 10950, 2, 33)                  # BASELINE config 3's series length
    bcsd_p_case('bcsd_p_month_abs_future', 1096, 1461, 3, 31, start_pred='1984-01-01', return_anoms=False)
    bcsd_p_case('bcsd_p_nasanex', 1096, 1096, 2, 32, start_fit='1980-01-01',
                time_grouper='daily_nasa-nex', return_anoms=False)

    # --- 5. PaddedDOYGrouper membership (test_pointwise_models.py:302-312) + full table
    index = pd.date_range(start='1980-01-01', end='1982-12-31')
    Xg = pd.DataFrame({'foo': np.arange(len(index), dtype=np.float64)}, index=index)
    groups = dict(list(PaddedDOYGrouper(Xg)))
    np.testing.assert_array_equal(np.unique(groups[123].index.dayofyear), np.arange(108, 139))
    lens = np.array([len(groups[d]) for d in range(1, 367)])
    rows = np.full((366, lens.max()), -1, dtype=np.int64)
    for d in range(1, 367):
        rows[d - 1, :lens[d - 1]] = groups[d]['foo'].values.astype(np.int64)
    save('padded_doy_1980_1982', rows=rows, lens=lens)

    # --- 6. GARD
    def pure_case(name, T, Tq, C, seed, n_analogs, kind, thresh=None):
        Xtr, ytr, Xq = synth.analog(T, Tq, C, 3, seed)
        rand = None
        if kind == 'sample_analogs':
            np.random.seed(1234)          # the reference draws from the GLOBAL numpy RNG (gard.py:315)
            rand = np.stack([np.random.randint(0, n_analogs, size=Tq) for _ in range(C)], axis=1)
            np.random.seed(1234)
        out = _cell_loop(PureAnalog(n_analogs=n_analogs, kind=kind, thresh=thresh), Xtr, ytr, Xq, n_outputs=3)
        kw = dict(Xtr=Xtr, ytr=ytr, Xq=Xq, out=out)
        if rand is not None:
            kw['rand'] = rand
        save(name, **kw)

    for kind in ('best_analog', 'mean_analogs', 'weight_analogs', 'sample_analogs'):
        pure_case(f'pure_{kind}', 400, 150, 2, 40, 10, kind)
        pure_case(f'pure_{kind}_thresh', 400, 150, 2, 41, 10, kind, thresh=0.0)
    pure_case('pure_mean_analogs_k200', 500, 60, 1, 42, 200, 'mean_analogs')

    def ar_case(name, T, Tq, C, seed, n_analogs):
        Xtr, ytr, Xq = synth.analog(T, Tq, C, 3, seed)
        out = _cell_loop(AnalogRegression(n_analogs=n_analogs), Xtr, ytr, Xq, n_outputs=3)
        out64 = np.empty((Tq, 3, C))
        for c in range(C):
            m = AnalogRegression(n_analogs=n_analogs).fit(_df(Xtr[..., c]), _df(ytr[:, c]))
            out64[:, :, c] = np.asarray(m.predict(_df(Xq[..., c])))
        save(name, Xtr=Xtr, ytr=ytr, Xq=Xq, out=out, out64=out64)

    ar_case('analogreg_k10', 300, 120, 2, 50, 10)
    ar_case('analogreg_k200', 400, 40, 1, 51, 200)
    ar_case('analogreg_k10_30yr', 10950, 160, 2, 52, 10)                   # BASELINE config 5's training window

    # AnalogRegression(thresh=...) (gard.py:201-215).  Query steps whose analogs are ALL at or below the
    # threshold make the reference raise, so they are dropped from the query set (the error path has
    # its own test).  `out64` = the reference with its default lbfgs tolerance (1e-4), `out64_tight` =
    # the same call with logistic_kwargs tol=1e-12 (the optimum the default run approximates).
    def ar_thresh_case(name, T, Tq, seed, n_analogs, thresh, C_reg=1.0):
        Xtr, ytr, Xq = synth.analog(T, Tq, 1, 3, seed)
        A, Q = Xtr[..., 0].astype(np.float64), Xq[..., 0].astype(np.float64)
        k = min(n_analogs, T)
        d2 = ((Q[:, None, :] - A[None, :, :]) ** 2).sum(-1)
        inds = np.argsort(d2, axis=1, kind='stable')[:, :k]
        keep = (ytr[:, 0][inds] > thresh).sum(axis=1) >= 1
        Xq = Xq[keep]
        kws = {} if C_reg == 1.0 else {'C': C_reg}
        m = AnalogRegression(n_analogs=n_analogs, thresh=thresh, logistic_kwargs=kws or None)
        out64 = np.asarray(m.fit(_df(Xtr[..., 0]), _df(ytr[:, 0])).predict(_df(Xq[..., 0])))
        m = AnalogRegression(n_analogs=n_analogs, thresh=thresh, logistic_kwargs=dict(kws, tol=1e-12, max_iter=10000))
        tight = np.asarray(m.fit(_df(Xtr[..., 0]), _df(ytr[:, 0])).predict(_df(Xq[..., 0])))
        save(name, Xtr=Xtr, ytr=ytr, Xq=Xq, out64=out64[:, :, None], out64_tight=tight[:, :, None],
             thresh=np.float64(thresh), C_reg=np.float64(C_reg))

    ar_thresh_case('analogreg_thresh_k20', 300, 200, 52, 20, -0.5)
    ar_thresh_case('analogreg_thresh_k10_C', 300, 160, 54, 10, -0.8, C_reg=0.3)
    ar_thresh_case('analogreg_thresh_k200', 400, 60, 53, 200, 0.0)

    # --- 7. PureRegression: one OLS (+ logistic exceedance model) per cell
    PureRegression = ref['gard'].PureRegression

    def pr_case(name, T, Tq, C, seed, thresh=None, dtype=np.float32, p=3):
        Xtr, ytr, Xq = synth.analog(T, Tq, C, p, seed, dtype=dtype)
        out = np.empty((Tq, 3, C), dtype=np.float64)
        tight = np.empty((Tq, C), dtype=np.float64)
        for c in range(C):
            m = PureRegression(thresh=thresh).fit(_df(Xtr[..., c]), _df(ytr[:, c]))
            out[:, :, c] = np.asarray(m.predict(_df(Xq[..., c])), dtype=np.float64)
            if thresh is not None:
                mt = PureRegression(thresh=thresh, logistic_kwargs={'tol': 1e-12, 'max_iter': 10000})
                mt.fit(_df(Xtr[..., c].astype(np.float64)), _df(ytr[:, c].astype(np.float64)))
                tight[:, c] = np.asarray(mt.predict(_df(Xq[..., c].astype(np.float64))))[:, 1]
        kw = dict(Xtr=Xtr, ytr=ytr, Xq=Xq, out=out)
        if thresh is not None:
            kw.update(prob_tight=tight, thresh=np.float64(thresh))
        save(name, **kw)

    pr_case('pure_regression', 600, 150, 3, 61)
    pr_case('pure_regression_thresh', 600, 150, 2, 62, thresh=0.0)
    pr_case('pure_regression_f64_thresh', 500, 120, 2, 63, thresh=-0.3, dtype=np.float64, p=2)

    # ---- ZScoreRegressor.predict (zscore.py:68-110) through the live reference.  zscore.py imports xarray at module
    # level (absent here) but predict never uses it: a placeholder module lets the import succeed.  fit needs xarray
    # (_calc_stats) and cannot run — shift_ / scale_ come from the oracle's restatement (oracle/zscore.py header).
    sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
    import oracle
    from oracle import zscore as ozs
    if 'xarray' not in sys.modules:
        sys.modules['xarray'] = types.ModuleType('xarray')
    ZScoreRegressor = importlib.import_module('skdownscale.pointwise_models.zscore').ZScoreRegressor

    def zs_case(name, T, Tp, C, seed, window=31, dtype=np.float32, start='1999-03-01'):
        idx = pd.date_range(start, periods=T, freq='D')
        Xtr, ytr, Xp = synth.temperature(max(T, Tp), C, seed=seed)
        Xtr, ytr, Xp = Xtr[:T].astype(dtype), ytr[:T].astype(dtype), Xp[:Tp].astype(dtype)
        idx_p = pd.date_range('2031-01-01', periods=Tp, freq='D')
        shift, scale, out = [], [], np.empty((Tp, C), dtype=np.float64)
        for c in range(C):
            st = ozs.zscore_fit(Xtr[:, c], ytr[:, c], idx, window)
            m = ZScoreRegressor(window_width=window)
            m.shift_ = pd.Series(st['shift'])
            m.scale_ = pd.Series(st['scale'])
            m.n_features_in_ = 1
            out[:, c] = np.asarray(m.predict(_df(Xp[:, c], idx_p)), dtype=np.float64)[:, 0]
            shift.append(st['shift'])
            scale.append(st['scale'])
        save(name, Xtr=Xtr, ytr=ytr, Xp=Xp, start=np.array(start), window=np.int64(window),
             shift=np.stack(shift, 1), scale=np.stack(scale, 1), out=out)

    zs_case('zscore_4yr', 1461, 1461, 5, 71)
    zs_case('zscore_pred_longer', 1200, 2000, 3, 72)
    zs_case('zscore_w30_f64', 1100, 700, 3, 73, window=30, dtype=np.float64, start='1999-06-01')
    zs_case('zscore_short_pred', 1461, 300, 2, 74, window=11)

    import sklearn
    with open(os.path.join(HERE, 'VERSIONS.json'), 'w') as f:
        json.dump({'numpy': np.__version__, 'pandas': pd.__version__, 'sklearn': sklearn.__version__,
                   'reference': 'pangeo-data/scikit-downscale @ 44d0425 (/root/reference)'}, f, indent=1)


if __name__ == '__main__':
    main()
