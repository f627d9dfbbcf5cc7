"""Secondary measurements (one GPU, device-resident inputs, CUDA-event timed): throughput of the
other BASELINE.json configurations in cell-timesteps/s, one JSON line each.  Not the driver's
bench contract (that is bench.py) — numbers quoted in DESIGN.md §5."""

from __future__ import annotations

import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'tests')):
    if p not in sys.path:
        sys.path.insert(0, p)

import torch  # noqa: E402

import synth  # noqa: E402
import skdownscale_b200  # noqa: F401,E402
from skdownscale_b200.pointwise_models import (AnalogRegression, BcsdPrecipitation, BcsdTemperature,  # noqa: E402
                                               EquidistantCdfMatcher, PureAnalog, QuantileMapper,
                                               QuantileMappingReressor)

dev = torch.device('cuda:0')


def timed(fn, warmup=3, steps=8):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


def report(name, cells, steps_t, ms, alg_bytes, extra=None):
    cts = cells * steps_t / (ms * 1e-3)
    line = {'config': name, 'cells': cells, 'timesteps': steps_t, 'ms': ms, 'cell_timesteps_per_s': cts,
            'algorithmic_GBps': alg_bytes * cts / 1e9}
    if extra:
        line.update(extra)
    print(json.dumps(line), flush=True)


def main():
    only_analog = '--analog' in sys.argv
    gen = torch.Generator(device=dev).manual_seed(0)
    if not only_analog:
        qm_paths(gen)
    if '--qm' not in sys.argv:
        analog_paths(gen)


def qm_paths(gen):

    # config 2: QuantileMapper, whole series as one group (generic 16384-point kernel)
    T, C = 10950, 10000
    y = torch.randn((T, C), device=dev, generator=gen) * 2 + 14
    x = torch.randn((T, C), device=dev, generator=gen) * 3 + 15
    qm = QuantileMapper()
    report('QuantileMapper 10000 cells x 10950 (one 10950-step group per cell)', C, T,
           timed(lambda: (qm.fit_batched(y), qm.transform_batched(x))), 12)
    # the "next" estimators on the same block
    qd = QuantileMapper(detrend=True)
    report('QuantileMapper(detrend=True) 10000 cells x 10950', C, T,
           timed(lambda: (qd.fit_batched(y), qd.transform_batched(x))), 12)
    qr = QuantileMappingReressor(extrapolate='1to1')
    report('QuantileMappingReressor(1to1) fit+predict 10000 cells x 10950', C, T,
           timed(lambda: (qr.fit_batched(x, y), qr.predict_batched(x))), 16)
    qr.fit_batched(x, y)
    report('QuantileMappingReressor(1to1) predict only', C, T, timed(lambda: qr.predict_batched(x)), 8)
    ed = EquidistantCdfMatcher(kind='difference')
    report('EquidistantCdfMatcher(difference) fit+predict 10000 cells x 10950', C, T,
           timed(lambda: (ed.fit_batched(x, y), ed.predict_batched(x))), 16)
    del y, x

    # config 3 (per-GPU shard): BcsdPrecipitation, zero-inflated
    T, C = 10950, 129600
    idx = synth.daily_index(T)

    def precip(p_dry):
        wet = torch.rand((T, C), device=dev, generator=gen) >= p_dry
        g = torch.distributions.Gamma(torch.tensor(0.8, device=dev), torch.tensor(1 / 6.0, device=dev))
        v = g.sample((T, C)).float()
        return torch.where(wet, v, torch.zeros((), device=dev))
    ytr, xp = precip(0.5), precip(0.55)
    bp = BcsdPrecipitation()
    out = torch.empty((T, C), device=dev)
    report('BcsdPrecipitation 129600 cells x 10950 (zero-inflated)', C, T,
           timed(lambda: (bp.fit_batched(ytr, ytr, idx), bp.predict_batched(xp, idx, out=out))), 12)
    del ytr, xp, out
    torch.cuda.empty_cache()

    # headline shard again for reference in the same process
    season = torch.sin(2 * torch.pi * torch.arange(T, device=dev) / 365.25)[:, None]
    mk = lambda m, a, s: torch.randn((T, C), device=dev, generator=gen) * s + m + a * season   # noqa: E731
    xtr, ytr, xp = mk(15, 10, 3), mk(14, 12, 2), mk(16.5, 10, 3)
    bt = BcsdTemperature()
    out = torch.empty((T, C), device=dev)
    report('BcsdTemperature 129600 cells x 10950', C, T,
           timed(lambda: (bt.fit_batched(xtr, ytr, idx), bt.predict_batched(xp, idx, out=out))), 16)
    # future period three times as long as the fit period (T_pred != T_fit: interpolation path)
    bt2 = BcsdTemperature()
    idx_f = synth.daily_index(3650)
    bt2.fit_batched(xtr[:3650], ytr[:3650], idx_f)
    report('BcsdTemperature predict only, fit 3650 days / predict 10950 days (interpolation + tails)', C, T,
           timed(lambda: bt2.predict_batched(xp, idx, out=out)), 8)
    del xtr, ytr, xp, out
    torch.cuda.empty_cache()


def analog_paths(gen):
    # config 4/5 style: analog models, k = 10, 3 predictors
    for name, model, T, Tq, C in (('PureAnalog(k=10, mean_analogs) 1024 cells x 18250', PureAnalog(n_analogs=10, kind='mean_analogs'), 18250, 18250, 1024),
                                  ('AnalogRegression(k=10) 1024 cells x 10950', AnalogRegression(n_analogs=10), 10950, 10950, 1024),
                                  ('AnalogRegression(k=10, thresh=0) 1024 cells x 10950', AnalogRegression(n_analogs=10, thresh=0.0), 10950, 10950, 1024),
                                  ('AnalogRegression(k=200) 128 cells x 10950', AnalogRegression(n_analogs=200), 10950, 10950, 128)):
        X = torch.randn((T, 3, C), device=dev, generator=gen)
        w = torch.tensor([1.0, 0.5, -0.3], device=dev)[None, :, None]
        yv = (X * w).sum(1) + 0.3 * torch.randn((T, C), device=dev, generator=gen)
        Xq = torch.randn((Tq, 3, C), device=dev, generator=gen)
        model.fit_batched(X, yv)
        ms = timed(lambda: model.predict_batched(Xq), warmup=1, steps=2)
        report(name, C, Tq, ms, 40, {'distance_evals_per_s': C * Tq * T / (ms * 1e-3)})
        del X, yv, Xq


if __name__ == '__main__':
    main()
