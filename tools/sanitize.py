"""Small shapes through every kernel family, meant to run under compute-sanitizer:

    compute-sanitizer --tool memcheck  python tools/sanitize.py
    compute-sanitizer --tool racecheck python tools/sanitize.py
"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'tests')):
    sys.path.insert(0, p)
import numpy as np
import torch
import synth
import skdownscale_b200  # noqa
from skdownscale_b200 import _lib, engine
from skdownscale_b200.pointwise_models import (AnalogRegression, BcsdPrecipitation, BcsdTemperature,
                                               EquidistantCdfMatcher, PureAnalog, PureRegression,
                                               QuantileMapper, QuantileMappingReressor)

dev = torch.device('cuda:0')
T, C = 1200, 11
idx = synth.daily_index(T)
Xtr, ytr, Xp = synth.temperature(T, C, 1)
Xp[:, 3] = np.round(Xp[:, 3])                 # ties → exact-comparison paths
Xp[::50, 5] = 1e9                              # outlier → exact 64-bit fallback
Ptr, pytr, Pp = synth.precipitation(T, C, 2)
for flags in (0, 1, 2):
    _lib.load().sdb_set_debug_flags(flags)
    for model, a, b, c in ((BcsdTemperature(), Xtr, ytr, Xp), (BcsdPrecipitation(), Ptr, pytr, Pp)):
        model.fit_batched(engine.as_device(a, dev), engine.as_device(b, dev), idx)
        model.predict_batched(engine.as_device(c, dev), idx, want_rank=True)
        model.predict_batched(engine.as_device(c[:700], dev), idx[:700])          # T_pred != T_fit
    q = QuantileMapper()
    q.fit_batched(engine.as_device(ytr[:300], dev))
    q.transform_batched(engine.as_device(Xp[:500], dev))
_lib.load().sdb_set_debug_flags(0)
nz = BcsdTemperature(time_grouper='daily_nasa-nex', return_anoms=False)
nz.fit_batched(engine.as_device(Xtr, dev), engine.as_device(ytr, dev), idx)
nz.predict_batched(engine.as_device(Xp, dev), idx)
A, ya, Aq = synth.analog(300, 90, 3, 3, 4)
for m in (PureAnalog(n_analogs=10, kind='weight_analogs'), AnalogRegression(n_analogs=10), AnalogRegression(n_analogs=40)):
    m.fit_batched(engine.as_device(A, dev), engine.as_device(ya, dev))
    m.predict_batched(engine.as_device(Aq, dev), want_idx=True)
# AnalogRegression(thresh=...): logistic + min-norm epilogue (queries whose analogs are all below the
# threshold only raise the flag)
m = AnalogRegression(n_analogs=12, thresh=-0.5)
m.fit_batched(engine.as_device(A, dev), engine.as_device(ya, dev))
m.predict_batched(engine.as_device(Aq, dev))
for m in (PureRegression(), PureRegression(thresh=0.0)):
    m.fit_batched(engine.as_device(A, dev), engine.as_device(ya, dev))
    m.predict_batched(engine.as_device(Aq, dev))
# CDF-to-CDF regressors, every tail mode; detrending mappers; non-default Cunnane tails
for ex in (None, 'min', 'max', 'both', '1to1'):
    for est in (QuantileMappingReressor(extrapolate=ex, n_endpoints=5), EquidistantCdfMatcher(kind='ratio', extrapolate=ex, n_endpoints=5)):
        est.fit_batched(engine.as_device(Xtr[:400], dev), engine.as_device(ytr[:400], dev))
        est.predict_batched(engine.as_device(Xp[:700], dev))
for model, a, b, c in ((BcsdTemperature(qm_kwargs={'detrend': True}), Xtr, ytr, Xp),
                       (BcsdPrecipitation(qm_kwargs={'detrend': True, 'qt_kwargs': {'extrapolate': 'max', 'n_endpoints': 4}}), Ptr, pytr, Pp)):
    model.fit_batched(engine.as_device(a, dev), engine.as_device(b, dev), idx)
    model.predict_batched(engine.as_device(c, dev), idx)
q = QuantileMapper(detrend=True, qt_kwargs={'extrapolate': None})
q.fit_batched(engine.as_device(ytr[:300], dev))
q.transform_batched(engine.as_device(Xp[:500], dev))
# ---- round 2 kernels
from skdownscale_b200.pointwise_models import LinearTrendTransformer, TrendAwareQuantileMappingRegressor  # noqa: E402
TL = 2600                                                    # one group longer than 1024 steps → block-wide counting rank
Ltr, Lytr, Lp = synth.temperature(TL, 5, 9)
Lp[:, 1] = np.round(Lp[:, 1])                                # ties
Lp[::9, 2] = Lp[:, 2].min()                                  # the lower-bound class
Lp[:, 3] = 2.5                                               # constant series
q = QuantileMapper()
q.fit_batched(engine.as_device(Lytr, dev))
q.transform_batched(engine.as_device(Lp, dev))
engine.series_argsort(engine.as_device(Lp, dev), 5, TL, 5)
seasonal = BcsdTemperature(time_grouper=lambda t: (t.month % 12) // 3)      # four ~650-step groups at T = 2600 ... and
seasonal.fit_batched(engine.as_device(Ltr, dev), engine.as_device(Lytr, dev), synth.daily_index(TL))
seasonal.predict_batched(engine.as_device(Lp, dev), synth.daily_index(TL))
T6 = 6000                                                    # ... 1 500-step groups: the long BCSD path
L6 = synth.temperature(T6, 3, 10)
seasonal = BcsdTemperature(time_grouper=lambda t: (t.month % 12) // 3)
seasonal.fit_batched(engine.as_device(L6[0], dev), engine.as_device(L6[1], dev), synth.daily_index(T6))
seasonal.predict_batched(engine.as_device(L6[2], dev), synth.daily_index(T6))
# fused fit+predict entry (counting rank in shared memory), with ties / outliers
fm = BcsdTemperature()
fm.fit_predict_batched(engine.as_device(Xtr, dev), engine.as_device(ytr, dev), engine.as_device(Xp, dev), idx, fused=True)
# grid-pruned analog search (per-cell kernel), every list capacity, duplicates → tie rule
A2, ya2, Aq2 = synth.analog(2300, 200, 3, 3, 6)
A2[1::2] = A2[0:-1:2]
for m in (PureAnalog(n_analogs=1), PureAnalog(n_analogs=10, kind='mean_analogs'), AnalogRegression(n_analogs=14),
          AnalogRegression(n_analogs=9, thresh=-0.5)):
    m.fit_batched(engine.as_device(A2, dev), engine.as_device(ya2, dev))
    assert m._order_train is not None
    m.predict_batched(engine.as_device(Aq2, dev), want_idx=True)
for p_ in (1, 2):
    A3, ya3, Aq3 = synth.analog(2100, 100, 2, p_, 7)
    m = PureAnalog(n_analogs=5, kind='weight_analogs')
    m.fit_batched(engine.as_device(A3, dev), engine.as_device(ya3, dev))
    m.predict_batched(engine.as_device(Aq3, dev))
# trend-aware regressor, trend transformer
ta = TrendAwareQuantileMappingRegressor(QuantileMappingReressor())
ta.fit_batched(engine.as_device(Xtr[:500], dev), engine.as_device(ytr[:500], dev))
ta.predict_batched(engine.as_device(Xp[:700], dev))
lt = LinearTrendTransformer()
lt.fit_batched(engine.as_device(Xtr, dev))
lt.inverse_transform_batched(lt.transform_batched(engine.as_device(Xp, dev)))
# ZScoreRegressor: day-column sums, sliding windows, rolling predict (odd / even window, masked cell, short predict)
from skdownscale_b200.pointwise_models import ZScoreRegressor  # noqa: E402
Z = synth.temperature(1461, 7, 12)
Z[0][:, 3] = np.nan
for w_ in (31, 30, 5):
    zs = ZScoreRegressor(window_width=w_)
    zs.fit_batched(engine.as_device(Z[0], dev), engine.as_device(Z[1], dev), synth.daily_index(1461),
                   valid=engine.cell_mask(engine.as_device(Z[0], dev)[0]), want_stats=True)
    zs.predict_batched(engine.as_device(Z[2], dev))
    zs.predict_batched(engine.as_device(Z[2][:200], dev), out_dtype=torch.float64)
# the push kernels (on one device: destination = a second local buffer)
src = torch.randn((64, 256), device=dev)
d1, d2 = torch.zeros((64, 512), device=dev), torch.zeros((64, 512), device=dev)
engine.peer_copy2d(d1[:, 128:384], src, method='kernel')
import ctypes  # noqa: E402
ptrs = (ctypes.c_void_p * 2)(d1[:, 0:256].data_ptr(), d2[:, 256:512].data_ptr())
_lib.check(_lib.load().sdb_peer_bcast2d(ptrs, 2, 512 * 4, src.data_ptr(), 256 * 4, 256 * 4, 64, 8, torch.cuda.current_stream().cuda_stream), 'bcast')
torch.cuda.synchronize()
assert torch.equal(d1[:, :256], src) and torch.equal(d2[:, 256:], src)
print('sanitize workload done')
