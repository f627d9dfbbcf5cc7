"""One predict launch of a secondary configuration, for `ncu -k regex:qm_predict_tile -c 1`:
    python tools/profile_one.py precip|interp|temp [cells]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'tests')):
    if p not in sys.path:
        sys.path.insert(0, p)

import torch  # noqa: E402

import synth  # noqa: E402
import skdownscale_b200  # noqa: F401,E402
from skdownscale_b200.pointwise_models import BcsdPrecipitation, BcsdTemperature  # noqa: E402

case = sys.argv[1]
C = int(sys.argv[2]) if len(sys.argv) > 2 else 32768
T = 10950
dev = torch.device('cuda:0')
gen = torch.Generator(device=dev).manual_seed(0)
idx = synth.daily_index(T)
out = torch.empty((T, C), device=dev)
if case == 'precip':
    def precip(p_dry):
        wet = torch.rand((T, C), device=dev, generator=gen) >= p_dry
        g = torch.distributions.Gamma(torch.tensor(0.8, device=dev), torch.tensor(1 / 6.0, device=dev))
        return torch.where(wet, g.sample((T, C)).float(), torch.zeros((), device=dev))
    ytr, xp = precip(0.5), precip(0.55)
    m = BcsdPrecipitation().fit_batched(ytr, ytr, idx)
    m.predict_batched(xp, idx, out=out)
else:
    season = torch.sin(2 * torch.pi * torch.arange(T, device=dev) / 365.25)[:, None]
    mk = lambda mu, a, s: torch.randn((T, C), device=dev, generator=gen) * s + mu + a * season   # noqa: E731
    xtr, ytr, xp = mk(15, 10, 3), mk(14, 12, 2), mk(16.5, 10, 3)
    m = BcsdTemperature()
    if case == 'interp':
        m.fit_batched(xtr[:3650], ytr[:3650], synth.daily_index(3650))
    else:
        m.fit_batched(xtr, ytr, idx)
    m.predict_batched(xp, idx, out=out)
torch.cuda.synchronize()
print('ok', case, C)
