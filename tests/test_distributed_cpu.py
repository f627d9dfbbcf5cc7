"""world_size-2 gloo test (CPU) of the multi-GPU host logic: cell sharding + the final gather."""

import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_cells, q):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    import skdownscale_b200  # noqa: F401
    from skdownscale_b200 import distributed as D
    T = 7
    full = torch.arange(T * n_cells, dtype=torch.float32).reshape(T, n_cells)
    local = D.shard_cells(full).clone() + 0.0
    a, b = D.cell_range(n_cells, world, rank)
    assert local.shape == (T, b - a)
    got = D.gather_cells(local, n_cells)
    ok = bool(torch.equal(got, full))
    full3 = torch.arange(T * 3 * n_cells, dtype=torch.float64).reshape(T, 3, n_cells)     # GARD-style 3 outputs
    got3 = D.gather_cells(D.shard_cells(full3).contiguous(), n_cells)
    ok = ok and bool(torch.equal(got3, full3))
    q.put((rank, ok))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize('n_cells', [10, 11])
def test_shard_and_gather_gloo(n_cells):
    world = 2
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_cells, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res == [(0, True), (1, True)]


def test_cell_range_partition():
    import skdownscale_b200  # noqa: F401
    from skdownscale_b200.distributed import cell_range
    for n in (0, 1, 7, 8, 1036800):
        for w in (1, 2, 3, 8):
            spans = [cell_range(n, w, r) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))
            assert max(b - a for a, b in spans) - min(b - a for a, b in spans) <= 1
    assert cell_range(1036800, 8, 3) == (388800, 518400)      # the bench shard: 129 600 cells per GPU
