import sys, json; sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import torch, synth, skdownscale_b200
from skdownscale_b200.pointwise_models import ZScoreRegressor
dev = torch.device('cuda:0')
T, C = 10950, 129600
g = torch.Generator(device=dev).manual_seed(1)
X = 15 + 3 * torch.randn((T, C), device=dev, generator=g)
y = 14 + 2 * torch.randn((T, C), device=dev, generator=g)
Xp = 16 + 3 * torch.randn((T, C), device=dev, generator=g)
out = torch.empty((T, C), device=dev)
idx = synth.daily_index(T)
m = ZScoreRegressor()
def t(fn, n=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
tf = t(lambda: m.fit_batched(X, y, idx))
tp = t(lambda: m.predict_batched(Xp, out=out))
print(json.dumps({'case': 'ZScoreRegressor 129600 cells x 10950 days f32', 'fit_ms': tf, 'predict_ms': tp,
                  'fit_GBps_algorithmic': 8 * T * C / tf / 1e6, 'predict_GBps_algorithmic': 8 * T * C / tp / 1e6,
                  'step_cell_timesteps_per_s': T * C / ((tf + tp) * 1e-3)}))
