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
torch.cuda.synchronize()
print('sanitize workload done')
