"""Multi-GPU tests (NCCL, one process per GPU; skipped on a single-GPU box): the sharded path with the gather
of the predicted field inside it (``PeerGather``: copy-engine pushes into every peer's replica while the next
cell chunk is computed) must reproduce the single-GPU field bit for bit on every rank, and so must the plain
NCCL all-gather (``gather_cells``)."""

import os
import socket

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    import sys
    here = os.path.dirname(os.path.abspath(__file__))
    for p in (os.path.dirname(here), here):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    import torch.distributed as dist
    dev = torch.device('cuda', rank)
    torch.cuda.set_device(dev)
    dist.init_process_group('nccl', rank=rank, world_size=world, device_id=dev)
    import synth
    import skdownscale_b200  # noqa: F401
    from skdownscale_b200 import distributed as D, engine
    from skdownscale_b200.pointwise_models import BcsdTemperature
    ok, why = True, ''
    try:
        T, C = 1461, 75                                   # uneven shards (38 + 37), not multiples of the 8-cell tile
        idx = synth.daily_index(T)
        Xtr, ytr, Xp = (engine.as_device(a, dev) for a in synth.temperature(T, C, seed=31))
        ref_model = BcsdTemperature().fit_batched(Xtr, ytr, idx)
        ref = ref_model.predict_batched(Xp, idx)          # the whole field on this GPU
        a, b = D.cell_range(C, world, rank)
        m = BcsdTemperature().fit_batched(Xtr[:, a:b], ytr[:, a:b], idx)
        g = D.PeerGather(T, C, torch.float32, dev)
        field = m.predict_gathered(Xp[:, a:b], idx, g, chunk_cells=16)
        if not torch.equal(field, ref):
            ok, why = False, f'PeerGather field differs on rank {rank}: {int((field != ref).sum())} elements'
        # second use of the same replica (pushes overwrite in place)
        field2 = m.predict_gathered(Xp[:, a:b], idx, g, chunk_cells=64)
        if ok and not torch.equal(field2, ref):
            ok, why = False, 'second predict_gathered differs'
        # the other push methods, on 16-byte aligned column blocks (SM kernel / one-read-n-writes kernel)
        for method in ('kernel', 'bcast'):
            C2 = 64
            g2 = D.PeerGather(T, C2, torch.float32, dev, method=method)
            a2, b2 = D.cell_range(C2, world, rank)
            m2 = BcsdTemperature().fit_batched(Xtr[:, a2:b2], ytr[:, a2:b2], idx)
            f2 = m2.predict_gathered(Xp[:, a2:b2], idx, g2, chunk_cells=16)
            if ok and not torch.equal(f2, ref[:, :C2]):
                ok, why = False, f'PeerGather(method={method}) differs'
            f2 = None
            g2.close()
        ref = ref.clone()
        field = field2 = None
        g.close()
        got = D.gather_cells(m.predict_batched(Xp[:, a:b], idx), C)
        if ok and not torch.equal(got, ref):
            ok, why = False, 'NCCL gather_cells differs'
        dist.barrier()
    except Exception as e:      # noqa: BLE001
        ok, why = False, repr(e)
    q.put((rank, ok, why))
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_predict_with_peer_gather():
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip('needs at least 2 GPUs')
    world = 2
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=300) for _ in range(world))
    for p in procs:
        p.join(timeout=120)
    assert all(ok for _, ok, _ in res), res
    assert all(p.exitcode == 0 for p in procs)
