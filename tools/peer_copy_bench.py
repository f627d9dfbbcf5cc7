"""torchrun --nproc-per-node 2 tools/peer_copy_bench.py — bandwidth of pushing a [10950, 129600] float32 column block
into a peer's [10950, 259200] replica: cudaMemcpy2DAsync (copy engines), sdb_peer_copy2d (SM kernel, several CTA
counts), torch copy_ on the strided views, and a contiguous cudaMemcpyAsync for reference."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch, torch.distributed as dist
import skdownscale_b200  # noqa
from skdownscale_b200 import distributed as D, engine
rank, world = int(os.environ['RANK']), int(os.environ['WORLD_SIZE'])
dev = torch.device('cuda', int(os.environ['LOCAL_RANK'])); torch.cuda.set_device(dev)
dist.init_process_group('nccl', device_id=dev)
T, C = 10950, 129600
g = D.PeerGather(T, world * C, torch.float32, dev)
g.local.normal_()
peer = g.peers[(rank + 1) % world]
print('mapped', rank, flush=True)
src = g.full[:, g.a:g.b]; dst = peer[:, g.a:g.b]
nbytes = src.numel() * 4
def timeit(fn, n=3):
    fn(); torch.cuda.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize(); dist.barrier()
    return e0.elapsed_time(e1) / n
res = {}
res['cudaMemcpy2DAsync'] = timeit(lambda: engine.peer_copy2d(dst, src, method='ce'), 1)
for ctas in (16, 32, 64, 148, 296):
    res[f'kernel_{ctas}ctas'] = timeit(lambda: engine.peer_copy2d(dst, src, n_ctas=ctas))
res['torch_copy_strided'] = timeit(lambda: dst.copy_(src, non_blocking=True))
flat_src = torch.empty(T * C, device=dev); 
res['chunks8_kernel_32ctas'] = timeit(lambda: [engine.peer_copy2d(dst[:, k * 16200:(k + 1) * 16200], src[:, k * 16200:(k + 1) * 16200], n_ctas=32) for k in range(8)])
if rank == 0:
    print(json.dumps({k: {'ms': v, 'GB/s': nbytes / v / 1e6} for k, v in res.items()}))
dist.barrier(); dist.destroy_process_group()
