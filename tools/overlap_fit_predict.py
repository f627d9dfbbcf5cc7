"""Experiment: does running the fit kernel of the next cell chunk beside the predict kernel of the current one (two
streams) shorten the BcsdTemperature step?  fit is ALU-pipe bound (VIMNMX 88 %), predict is latency bound (issue 55 %).
One JSON line."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'tests')):
    sys.path.insert(0, p)
import torch
import synth
import skdownscale_b200  # noqa
from skdownscale_b200.pointwise_models import BcsdTemperature

dev = torch.device('cuda:0')
T, C = 10950, 129600
idx = synth.daily_index(T)
g = torch.Generator(device=dev).manual_seed(0)
season = torch.sin(2 * torch.pi * torch.arange(T, device=dev, dtype=torch.float32) / 365.25)[:, None]
mk = lambda m_, a_, s_: torch.randn((T, C), device=dev, generator=g) * s_ + m_ + a_ * season   # noqa: E731
Xtr, ytr, Xp = mk(15, 10, 3), mk(14, 12, 2), mk(16.5, 10, 3)
out = torch.empty((T, C), device=dev)
out2 = torch.empty((T, C), device=dev)


def timed(fn, n=5):
    fn(); fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


m = BcsdTemperature()
def serial():
    m.fit_batched(Xtr, ytr, idx)
    m.predict_batched(Xp, idx, out=out)
res = {'serial_ms': timed(serial)}
s1, s2 = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
for chunk in (16200, 32400, 64800):
    spans = [(c0, min(C, c0 + chunk)) for c0 in range(0, C, chunk)]
    models = [BcsdTemperature() for _ in spans]

    def piped():
        cur = torch.cuda.current_stream()
        s1.wait_stream(cur); s2.wait_stream(cur)
        for mdl, (c0, c1) in zip(models, spans):
            with torch.cuda.stream(s1):
                mdl.fit_batched(Xtr[:, c0:c1], ytr[:, c0:c1], idx)
                ev = torch.cuda.Event(); ev.record(s1)
            with torch.cuda.stream(s2):
                s2.wait_event(ev)
                mdl.predict_batched(Xp[:, c0:c1], idx, out=out2[:, c0:c1])
        cur.wait_stream(s1); cur.wait_stream(s2)
    res[f'two_streams_chunk_{chunk}_ms'] = timed(piped)
    res[f'identical_{chunk}'] = bool(torch.equal(out, out2))
print(json.dumps(res))
