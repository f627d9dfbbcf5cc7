"""Times the split path (sdb_qm_fit + sdb_qm_predict) against the fused entry (sdb_bcsd_fit_predict) on the bench
shard, CUDA events, and prints the counting-rank statistics.  python tools/time_fused.py [cells] [T|P|both]"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'tests')):
    sys.path.insert(0, p)
import numpy as np
import torch
import skdownscale_b200  # noqa
from skdownscale_b200.pointwise_models import BcsdTemperature, BcsdPrecipitation
import synth

C = int(sys.argv[1]) if len(sys.argv) > 1 else 129600
which = sys.argv[2] if len(sys.argv) > 2 else 'both'
T = 10950
dev = torch.device('cuda:0')
idx = synth.daily_index(T)
gen = torch.Generator(device=dev).manual_seed(1234)
season = torch.sin(2 * torch.pi * torch.arange(T, device=dev, dtype=torch.float32) / 365.25)[:, None]


def field(mean, amp, sd):
    x = torch.randn((T, C), device=dev, dtype=torch.float32, generator=gen)
    return x.mul_(sd).add_(mean + amp * season)


def precip(p_dry):
    u = torch.rand((T, C), device=dev, generator=gen)
    g = torch.distributions.Gamma(torch.tensor(0.8, device=dev), torch.tensor(1.0 / 6.0, device=dev))
    x = torch._standard_gamma(torch.full((T, C), 0.8, device=dev, dtype=torch.float32)).mul_(6.0)
    return torch.where(u < p_dry, torch.zeros((), device=dev), x)


def timeit(f, n=5, w=2):
    for _ in range(w):
        f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        f()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


for name in (['T', 'P'] if which == 'both' else [which]):
    if name == 'T':
        Xtr, ytr, Xp = field(15.0, 10.0, 3.0), field(14.0, 12.0, 2.0), field(16.5, 10.0, 3.0)
        model = BcsdTemperature(return_anoms=True)
    else:
        Xtr, ytr, Xp = precip(0.6), precip(0.5), precip(0.55)
        model = BcsdPrecipitation(return_anoms=True)
    out = torch.empty((T, C), device=dev, dtype=torch.float32)

    def split():
        model.fit_batched(Xtr, ytr, idx)
        model.predict_batched(Xp, idx, out=out)

    def fused(keep=False, stats=None):
        model.fit_predict_batched(Xtr, ytr, Xp, idx, out=out, keep_state=keep, stats=stats, fused=True)

    ms_split = timeit(split)
    ref = out.clone()
    ms_fused = timeit(lambda: fused(False))
    same = bool(torch.equal(out, ref))
    ms_fused_state = timeit(lambda: fused(True))
    stats = torch.zeros(8, dtype=torch.int64, device=dev)
    fused(False, stats)
    torch.cuda.synchronize()
    st = stats.cpu().numpy().tolist()
    print(json.dumps({'model': name, 'cells': C, 'days': T, 'ms_split': ms_split, 'ms_fused': ms_fused,
                      'ms_fused_keep_state': ms_fused_state, 'bit_identical': same,
                      'series': st[0], 'y_network': st[1], 'x_network': st[2],
                      'y_queued_per_series': st[3] / max(st[0], 1), 'x_queued_per_series': st[4] / max(st[0], 1)}))
    del Xtr, ytr, Xp, out, ref
    torch.cuda.empty_cache()
