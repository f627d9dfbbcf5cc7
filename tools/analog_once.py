"""One small AnalogRegression(k=10) call (256 cells x 10950 steps, 3 predictors) — target for an ncu capture."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import skdownscale_b200  # noqa
from skdownscale_b200.pointwise_models import AnalogRegression
dev = torch.device('cuda:0')
g = torch.Generator(device=dev).manual_seed(0)
T, C = 10950, 256
X = torch.randn((T, 3, C), device=dev, generator=g)
y = (X * torch.tensor([1.0, .5, -.3], device=dev)[None, :, None]).sum(1) + .3 * torch.randn((T, C), device=dev, generator=g)
Xq = torch.randn((T, 3, C), device=dev, generator=g)
m = AnalogRegression(n_analogs=10)
m.fit_batched(X, y)
for _ in range(2):
    out = m.predict_batched(Xq)
torch.cuda.synchronize()
print(out.shape)
