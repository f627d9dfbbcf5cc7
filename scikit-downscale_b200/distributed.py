"""Multi-GPU plumbing: cells shard embarrassingly (every cell is an independent series in the
reference — skdownscale/pointwise_models/core.py:86-96, 137-141), so each rank fits and predicts
its own contiguous range of the flattened cell axis with NO data-path collective.  The only
exchange is the optional gather of the predicted field at the end (NCCL over NVLink on GPUs;
gloo works for CPU tensors, which is how the host logic is tested).

torch.distributed is plumbing here: one process per GPU, launched with torchrun.
"""

from __future__ import annotations

import torch
import torch.distributed as dist


def cell_range(n_cells: int, world_size: int, rank: int) -> tuple[int, int]:
    """Contiguous [start, stop) of the flattened cell axis owned by ``rank``; the first
    ``n_cells % world_size`` ranks take one extra cell."""
    if not 0 <= rank < world_size:
        raise ValueError(f'rank {rank} outside world of {world_size}')
    base, extra = divmod(n_cells, world_size)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def shard_cells(field: torch.Tensor, world_size: int | None = None, rank: int | None = None) -> torch.Tensor:
    """This rank's column block of a ``[..., cells]`` array (a view, no copy)."""
    world_size = dist.get_world_size() if world_size is None else world_size
    rank = dist.get_rank() if rank is None else rank
    a, b = cell_range(field.shape[-1], world_size, rank)
    return field[..., a:b]


def gather_cells(local: torch.Tensor, n_cells: int, group=None, stacked: bool = False) -> torch.Tensor:
    """All-gather the per-rank column blocks ``[T, (k,) C_rank]`` into the full ``[T, (k,) n_cells]``
    field on every rank (one NCCL all-gather over NVLink + one local re-interleave).

    ``stacked=True`` skips the re-interleave and returns the collective's own layout
    ``[world, T, (k,) C_max]`` (rank-major; uneven shards zero-padded to the widest)."""
    world = dist.get_world_size(group)
    spans = [cell_range(n_cells, world, r) for r in range(world)]
    wmax = max(b - a for a, b in spans)
    lead = tuple(local.shape[:-1])
    mine = local
    if local.shape[-1] != wmax or not local.is_contiguous():
        mine = torch.zeros(lead + (wmax,), dtype=local.dtype, device=local.device)
        mine[..., : local.shape[-1]] = local
    # concatenation form along dim 0 (accepted by both NCCL and gloo), viewed as [world, ...] afterwards
    flat = torch.empty((world * mine.shape[0],) + tuple(mine.shape[1:]), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(flat, mine, group=group)
    full = flat.view((world,) + tuple(mine.shape))
    if stacked:
        return full
    if all(b - a == wmax for a, b in spans):
        return full.movedim(0, -2).reshape(lead + (world * wmax,))     # [.., world, C_rank] → cells contiguous
    return torch.cat([full[r][..., : (b - a)] for r, (a, b) in enumerate(spans)], dim=-1)


class PeerGather:
    """The gather of the predicted field INSIDE the path: every rank owns a full ``[T, (k,) n_cells]`` replica,
    predicts straight into its own column block of it (``ld_out = n_cells``, no staging buffer) and pushes
    finished cell chunks into the same columns of every peer's replica (peer memory mapped through CUDA IPC, one
    stream per peer) while the next chunk is still being computed.  ``method='ce'``: pitched ``cudaMemcpy2DAsync``
    on the copy engines, no SM time — 557 GB/s per direction measured between two B200s for the 129 600-cell block
    (profiles/r02_peer_copy_bench.json); ``method='kernel'``: ``sdb_peer_copy2d``, 16-byte loads / stores by
    ``n_ctas`` CTAs (637 GB/s with 148 CTAs, 686 GB/s with 296); ``method='bcast'``: ``sdb_peer_bcast2d``, ONE kernel
    per chunk that reads the chunk once and stores it into every peer's replica (for many peers: the seven
    concurrent copy-engine transfers of an 8-GPU box only reach 353 GB/s received per GPU).  No re-interleave, no
    second full-size buffer.
    The replica is allocated by the library (``sdb_peer_alloc``: an IPC handle names a whole cudaMalloc
    allocation) and wrapped as a torch tensor without a copy.

    Usage (one process per GPU, default process group initialised)::

        g = PeerGather(T, n_cells, torch.float32, device)
        for c0, c1 in g.chunks(16200):                 # local cell numbers of this rank's shard
            model.predict_cells(c0, c1, out=g.local[:, c0:c1])
            g.push(c0, c1)
        field = g.finish()                               # [T, n_cells] complete on every rank
    """

    def __init__(self, n_steps: int, n_cells: int, dtype, device, n_outputs: int | None = None, group=None,
                 method: str = 'ce', n_ctas: int = 296):
        import ctypes
        from . import _lib
        lib = _lib.load()
        self._lib, self._libmod = lib, _lib
        self.group = group
        self.method, self.n_ctas = method, n_ctas
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self.device = torch.device(device)
        self.shape = (n_steps, n_cells) if n_outputs is None else (n_steps, n_outputs, n_cells)
        self.dtype = dtype
        nbytes = int(torch.tensor(self.shape).prod().item()) * torch.empty((), dtype=dtype).element_size()
        self._own = ctypes.c_void_p()
        handle = ctypes.create_string_buffer(64)
        with torch.cuda.device(self.device):
            _lib.check(lib.sdb_peer_alloc(nbytes, ctypes.byref(self._own), handle), 'sdb_peer_alloc')
            self.full = self._wrap(self._own.value)
            self.a, self.b = cell_range(n_cells, self.world, self.rank)
            self.local = self.full[..., self.a:self.b]
            self.peers = [None] * self.world
            self._opened = []
            if self.world > 1:
                handles = [None] * self.world
                dist.all_gather_object(handles, (self.device.index, handle.raw), group=group)
                for r, (peer_dev, h) in enumerate(handles):
                    if r == self.rank:
                        continue
                    p = ctypes.c_void_p()
                    _lib.check(lib.sdb_peer_open(ctypes.create_string_buffer(h, 64), ctypes.byref(p)), 'sdb_peer_open')
                    self._opened.append(p)
                    self.peers[r] = self._wrap(p.value)      # the peer's replica, mapped for access from THIS device
                self.streams = [torch.cuda.Stream(self.device, priority=-1) if r != self.rank else None for r in range(self.world)]
        self.bytes_pushed = 0

    def _wrap(self, ptr: int) -> torch.Tensor:
        """A torch view of raw device memory (zero copy, through ``__cuda_array_interface__``)."""
        typestr = {torch.float32: '<f4', torch.float64: '<f8'}[self.dtype]

        class _Raw:
            pass
        raw = _Raw()
        raw.__cuda_array_interface__ = {'shape': tuple(self.shape), 'typestr': typestr, 'data': (ptr, False), 'version': 2}
        return torch.as_tensor(raw, device=self.device)

    def close(self) -> None:
        """Unmap the peers' replicas and free the own one (after a barrier: nobody may still be pushing into it)."""
        if getattr(self, '_own', None) is None:
            return
        torch.cuda.synchronize(self.device)
        if self.world > 1:
            dist.barrier(group=self.group)
        self.full = self.local = None
        self.peers = []
        with torch.cuda.device(self.device):
            for p in self._opened:
                self._libmod.check(self._lib.sdb_peer_close(p), 'sdb_peer_close')
            self._opened = []
            if self.world > 1:
                dist.barrier(group=self.group)
            self._libmod.check(self._lib.sdb_peer_free(self._own), 'sdb_peer_free')
        self._own = None

    def chunks(self, chunk_cells: int):
        w = self.b - self.a
        return [(c0, min(c0 + chunk_cells, w)) for c0 in range(0, w, chunk_cells)]

    def push(self, c0: int, c1: int) -> None:
        """Send local cells [c0, c1) (already written into ``self.local`` on the current stream) to every peer."""
        if self.world == 1:
            return
        from . import engine
        ready = torch.cuda.Event()
        ready.record(torch.cuda.current_stream(self.device))
        lead = self.full.shape[:-1]
        rows = int(torch.tensor(lead).prod().item()) if len(lead) > 1 else lead[0]
        src = self.full.view(rows, self.full.shape[-1])[:, self.a + c0:self.a + c1]
        es = src.element_size()
        aligned = not ((src.shape[1] * es | src.stride(0) * es | src.data_ptr()) & 15)
        if self.method == 'bcast' and aligned and self.world <= 9:
            # one kernel reads the chunk once and stores it into every peer's replica
            import ctypes
            with torch.cuda.device(self.device):
                st = self.streams[(self.rank + 1) % self.world]
                st.wait_event(ready)
                order = [(self.rank + k) % self.world for k in range(1, self.world)]
                ptrs = (ctypes.c_void_p * len(order))(*[self.peers[r].view(rows, self.full.shape[-1])[:, self.a + c0:self.a + c1].data_ptr()
                                                       for r in order])
                self._libmod.check(self._lib.sdb_peer_bcast2d(ptrs, len(order), self.full.shape[-1] * es, src.data_ptr(), src.stride(0) * es,
                                                              src.shape[1] * es, rows, self.n_ctas, st.cuda_stream), 'sdb_peer_bcast2d')
                self.bytes_pushed += len(order) * src.numel() * es
            return
        with torch.cuda.device(self.device):
            # start with the next rank so that at any moment the eight senders aim at eight different receivers
            for k in range(1, self.world):
                r = (self.rank + k) % self.world
                st = self.streams[r]
                st.wait_event(ready)
                dst = self.peers[r].view(rows, self.full.shape[-1])[:, self.a + c0:self.a + c1]
                engine.peer_copy2d(dst, src, stream=st.cuda_stream, n_ctas=self.n_ctas, method=self.method)
                self.bytes_pushed += src.numel() * src.element_size()

    def finish(self) -> torch.Tensor:
        """Wait for this rank's pushes, then for everybody's: the replica is complete."""
        if self.world > 1:
            for st in self.streams:
                if st is not None:
                    st.synchronize()
            torch.cuda.current_stream(self.device).synchronize()
            dist.barrier(group=self.group)
        return self.full
