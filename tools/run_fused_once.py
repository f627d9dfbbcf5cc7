"""One fused fit+predict call (and optionally one split call) on synthetic temperature — the ncu target.
python tools/run_fused_once.py [cells] [T|P] [split]"""
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'tests')):
    sys.path.insert(0, p)
import torch
import skdownscale_b200  # noqa
from skdownscale_b200.pointwise_models import BcsdTemperature, BcsdPrecipitation
import synth
C = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
name = sys.argv[2] if len(sys.argv) > 2 else 'T'
T = 10950
dev = torch.device('cuda:0')
idx = synth.daily_index(T)
gen = torch.Generator(device=dev).manual_seed(1234)
season = torch.sin(2 * torch.pi * torch.arange(T, device=dev, dtype=torch.float32) / 365.25)[:, None]
def field(mean, amp, sd):
    x = torch.randn((T, C), device=dev, dtype=torch.float32, generator=gen)
    return x.mul_(sd).add_(mean + amp * season)
Xtr, ytr, Xp = field(15.0, 10.0, 3.0), field(14.0, 12.0, 2.0), field(16.5, 10.0, 3.0)
if name == 'P':
    Xtr, ytr, Xp = [torch.where(torch.rand_like(a) < p, torch.zeros((), device=dev), (a - 10).abs()) for a, p in ((Xtr, .6), (ytr, .5), (Xp, .55))]
model = (BcsdTemperature if name == 'T' else BcsdPrecipitation)(return_anoms=True)
out = torch.empty((T, C), device=dev, dtype=torch.float32)
for _ in range(2):
    model.fit_predict_batched(Xtr, ytr, Xp, idx, out=out, keep_state=False, fused=True)
    if 'split' in sys.argv:
        model.fit_batched(Xtr, ytr, idx)
        model.predict_batched(Xp, idx, out=out)
torch.cuda.synchronize()
print('done')
