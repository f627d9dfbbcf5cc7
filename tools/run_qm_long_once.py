"""QuantileMapper 10000 cells x 10950 (BASELINE config 2) once — ncu target."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'tests')):
    sys.path.insert(0, p)
import torch
import skdownscale_b200  # noqa
from skdownscale_b200.pointwise_models import QuantileMapper
C = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
dev = torch.device('cuda:0')
gen = torch.Generator(device=dev).manual_seed(0)
y = torch.randn((10950, C), device=dev, generator=gen) * 2 + 14
x = torch.randn((10950, C), device=dev, generator=gen) * 3 + 15
qm = QuantileMapper()
for _ in range(3):
    qm.fit_batched(y); out = qm.transform_batched(x)
torch.cuda.synchronize(); print('done')
