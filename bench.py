"""bench.py — headline benchmark: BcsdTemperature fit+predict throughput in cell-timesteps/s.

    python bench.py --gpus N --steps K --warmup W            (N > 1: launched under torchrun)
    python bench.py --impl reference ...                       CPU arm (oracle port on host cores)

Workload (BASELINE.json north_star / SURVEY.md §8(d)): the 720x1440 global 0.25-degree grid x
10 950 daily steps, BcsdTemperature (monthly groups, return_anoms=True), float32 in / out.
The full grid (4 arrays x 45.4 GB) does not fit one 180 GB GPU, so the per-GPU workload is
the 8-GPU shard of that grid — 129 600 cells x 10 950 days — held FIXED as N grows (weak
scaling; at N = 8 the job is exactly the headline configuration).  A step = one fit
(two climatology kernels + the per-group sort) plus one predict over the shard, inputs
resident in HBM.  Inputs (17 GB per GPU) are far larger than the 126 MB L2, so no flush is
needed between steps.  Prints ONE JSON line (rank 0).
"""

from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, 'tests')):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402

METRIC = 'cell_timesteps_per_sec_fit_predict_bcsd_temperature'
UNIT = 'cell-timesteps/s'
DAYS = 10950
CELLS_PER_GPU = 129600           # 1/8 of 720 x 1440
ALG_BYTES_STEP = 16              # read X_train, y_train, X_pred, write out (f32)
ALG_BYTES_PREDICT = 8            # dominant kernel (qm_predict): read X_pred, write out


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=5)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--cells-per-gpu', type=int, default=CELLS_PER_GPU)
    ap.add_argument('--days', type=int, default=DAYS)
    ap.add_argument('--e2e-steps', type=int, default=2)
    ap.add_argument('--cpu-seconds', type=float, default=10.0, help='target CPU time of the cpu_baseline sample')
    ap.add_argument('--no-e2e', action='store_true')
    ap.add_argument('--no-cpu', action='store_true')
    ap.add_argument('--debug-flags', type=int, default=None, help='sdb_set_debug_flags value (kernel variant experiments)')
    ap.add_argument('--no-secondary', action='store_true', help='skip the other BASELINE.json configurations')
    ap.add_argument('--secondary-seconds', type=float, default=60.0, help='time budget of the secondary list')
    ap.add_argument('--gather-chunk', type=int, default=16200, help='cells per pushed chunk of the in-path gather')
    ap.add_argument('--gather-method', default='auto',
                    help='how finished chunks reach the peers: ce | kernel | bcast, optionally method:chunk_cells, or a comma list '
                         '(every entry is timed, the fastest is reported; auto: copy engines for 2 GPUs, one broadcast kernel beyond)')
    ap.add_argument('--fused', action='store_true', help='also time the fused counting-rank entry (sdb_bcsd_fit_predict)')
    return ap.parse_args()


# ---------------------------------------------------------------------------------- CPU arm
def _oracle_cells(args):
    """worker: oracle BcsdTemperature fit+predict on a block of cells (numpy port of the reference)."""
    import oracle
    import synth
    T, c0, n, seed = args
    idx = synth.daily_index(T)
    Xtr, ytr, Xp = synth.temperature(T, n, seed=seed + c0)
    t0 = time.perf_counter()
    oracle.pointwise_fit_predict({'name': 'BcsdTemperature', 'return_anoms': True}, Xtr, ytr, Xp, idx, idx)
    return time.perf_counter() - t0


_POOL = {}


def _pool(cores: int):
    """One persistent joblib/loky pool per run, warmed up (worker start-up and imports are not part
    of the measured sample)."""
    from joblib import Parallel, delayed
    if cores not in _POOL:
        par = Parallel(n_jobs=cores)
        par(delayed(_oracle_cells)((365, i, 1, 7)) for i in range(cores))
        _POOL[cores] = par
    return _POOL[cores]


def cpu_sample(T: int, n_cells: int, cores: int):
    """Time the oracle port on ``n_cells`` cells spread over ``cores`` worker processes."""
    from joblib import delayed
    par = _pool(cores)
    per = max(1, n_cells // cores)
    jobs = [(T, i * per, per, 1000) for i in range(cores)]
    t0 = time.perf_counter()
    par(delayed(_oracle_cells)(j) for j in jobs)
    wall = time.perf_counter() - t0
    done = per * cores
    return done * T / wall, done, wall


def cpu_cells(T: int, cores: int, target_seconds: float) -> int:
    """Sample size for about ``target_seconds`` of wall time, calibrated THROUGH the pool (the host
    cores of a shared box deliver far less in parallel than one process alone suggests)."""
    _, done, wall = cpu_sample(T, 2 * cores, cores)           # 2 cells per worker
    per_worker = int(max(2, min(400, 2 * target_seconds / max(wall, 1e-3))))
    return per_worker * cores


def cpu_baseline(T: int, target_seconds: float):
    cores = os.cpu_count() or 1
    n_cells = cpu_cells(T, cores, target_seconds)
    value, done, wall = cpu_sample(T, n_cells, cores)
    return {'value': value, 'unit': UNIT, 'cores': cores, 'kind': 'port',
            'sample': f'oracle (numpy port of the reference algorithm) BcsdTemperature fit+predict on {done} cells x {T} days, '
                      f'{cores} joblib worker processes (pool warmed up), {wall:.1f} s wall'}


def run_reference(a):
    """--impl reference: the reference's CPU implementation of the path on the host cores.  The
    Python reference cannot travel to the GPU box, so this is the oracle port (kind 'port')."""
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    budget = 120.0 / max(1, a.steps + a.warmup)        # keep the whole run within a few minutes
    n_cells = cpu_cells(a.days, cores, min(budget, 15.0))
    for _ in range(a.warmup):
        cpu_sample(a.days, n_cells, cores)
    vals, walls, done = [], [], 0
    for _ in range(a.steps):
        v, done, w = cpu_sample(a.days, n_cells, cores)
        vals.append(v)
        walls.append(w)
    value = done * a.days * len(walls) / sum(walls)
    sample = (f'oracle port, BcsdTemperature fit+predict, {done} cells x {a.days} days per step, '
              f'{cores} joblib processes')
    print(json.dumps({
        'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': a.gpus,
        'steps': a.steps, 'warmup': a.warmup, 'ms_per_step': 1e3 * sum(walls) / len(walls),
        'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32 data, f64 keys/interp',
        'data': 'synthetic',
        'config': {'workload': f'BcsdTemperature fit+predict, monthly groups, {a.days} daily steps; bounded CPU sample of the '
                               f'{a.cells_per_gpu}-cells-per-GPU shard of the 720x1440 grid', 'cells_per_step': done,
                   'days': a.days},
        'cpu_baseline': {'value': value, 'unit': UNIT, 'cores': cores, 'kind': 'port', 'sample': sample},
        'e2e': {'value': value, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }))


# ---------------------------------------------------------------------------------- GPU arm
class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,'
         'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
         'clocks_event_reasons.sw_power_cap')

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.lines = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', f'--query-gpu={self.Q}', '--format=csv,noheader,nounits',
                                          '-lms', '100', '-i', str(self.gpu)], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=self._pump, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append((time.perf_counter(), line.strip()))

    def stop(self, t0, t1):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        for ts, line in self.lines:
            f = [x.strip() for x in line.split(',')]
            if len(f) < 8 or not (t0 <= ts <= t1 + 0.2):
                continue
            try:
                sm.append(float(f[1]))
                mx = float(f[2])
            except ValueError:
                continue
            for name, v in zip(('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap'), f[4:8]):
                if v.lower().startswith('active'):
                    reasons.add(name)
        return {'sm_mhz': float(np.median(sm)) if sm else None, 'sm_max_mhz': mx, 'reasons': sorted(reasons),
                'samples': len(sm)}


def ncu_traffic(kernel: str, units: float):
    """DRAM bytes per launch of ``kernel`` from the committed ncu capture (profiles/r01_traffic.json),
    scaled to this run's number of cell-timesteps per launch (traffic is linear in it)."""
    try:
        with open(os.path.join(ROOT, 'profiles', 'r01_traffic.json')) as f:
            t = json.load(f)
        k = t['kernels'][kernel]
        per_unit = (k['dram_bytes_read'] + k['dram_bytes_write']) / (t['cells'] * t['days'])
        return per_unit * units
    except Exception:
        return None


def ncu_limiter(kernel: str):
    """What actually bounds ``kernel`` according to the committed ncu capture: pipe / issue utilisation
    (the path is a sorting network — the ALU pipe that executes VIMNMX and dependency latency, not HBM)."""
    try:
        with open(os.path.join(ROOT, 'profiles', 'r01_traffic.json')) as f:
            k = json.load(f)['kernels'][kernel]
        return {m: k[m] for m in ('alu_pipe_pct', 'fma_pipe_pct', 'issue_active_pct', 'dram_throughput_pct', 'registers')
                if m in k}
    except Exception:
        return None


def gpu_numa_affinity(dev_index: int):
    """Bind this process (and so its pinned allocations, first touched here) to the CPUs of the NUMA node the GPU
    hangs off: with eight ranks staging 23 GB per step each through one host, crossing the socket interconnect is
    what round 1 measured as 0.30 end-to-end scaling efficiency.  Returns (description, previous affinity)."""
    try:
        import torch
        bus = torch.cuda.get_device_properties(dev_index)
        pci = f'{bus.pci_domain_id:04x}:{bus.pci_bus_id:02x}:{bus.pci_device_id:02x}.0'
        base = f'/sys/bus/pci/devices/{pci}'
        node = int(open(f'{base}/numa_node').read())
        cpus = open(f'{base}/local_cpulist').read().strip()
        ids = set()
        for part in cpus.split(','):
            lo, _, hi = part.partition('-')
            ids.update(range(int(lo), int(hi or lo) + 1))
        old = os.sched_getaffinity(0)
        ids &= old
        if ids:
            os.sched_setaffinity(0, ids)
        return {'pci': pci, 'numa_node': node, 'cpus': cpus, 'bound': bool(ids)}, old
    except Exception as ex:      # noqa: BLE001
        return {'error': repr(ex)[:200]}, None


def measured_peak():
    try:
        with open(os.path.join(ROOT, 'MEASURED_PEAKS.json')) as f:
            return float(json.load(f)['hbm_gbs']), 'MEASURED_PEAKS.json hbm_gbs (measured copy bandwidth)'
    except Exception:
        return 6650.0, 'fallback 6.65 TB/s (B200_PROFILING.md)'


# ---------------------------------------------------------------------------------- secondary: BASELINE configs 2-5
def run_secondary(a, dev, world, rank, barrier, reduce_max, peak):
    """The other BASELINE.json configurations, measured in the same run after the headline (device-resident
    inputs, CUDA events, max over ranks; per-GPU cell counts as BASELINE states them, or the largest that fits
    the time budget — stated per entry).  Under --gpus N every rank runs its own shard (weak scaling)."""
    import torch
    import synth
    from skdownscale_b200.pointwise_models import AnalogRegression, BcsdPrecipitation, PureAnalog, QuantileMapper
    t_begin = time.perf_counter()
    out_list = []
    gen = torch.Generator(device=dev).manual_seed(99 + rank)

    def left():
        return a.secondary_seconds - (time.perf_counter() - t_begin)

    def timed(fn, warmup, steps):
        for _ in range(warmup):
            fn()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        return reduce_max(e0.elapsed_time(e1) / steps)

    def entry(name, cells, T, ms, alg_bytes, **extra):
        cts = world * cells * T / (ms * 1e-3)
        gbs = alg_bytes * cells * T / (ms * 1e-3) / 1e9
        e = {'config': name, 'cells_per_gpu': cells, 'timesteps': T, 'ms': ms, 'value': cts, 'unit': UNIT,
             'algorithmic_bytes_per_cell_timestep': alg_bytes, 'algorithmic_GBps_per_gpu': gbs, 'frac_of_hbm_peak': gbs / peak}
        e.update(extra)
        out_list.append(e)

    def guarded(name, fn):
        try:
            fn()
        except Exception as ex:      # noqa: BLE001 - a failed secondary entry must not lose the headline line
            out_list.append({'config': name, 'error': repr(ex)[:300]})
        torch.cuda.empty_cache()

    T = 10950
    idx = synth.daily_index(T)

    def precip(shape, p_dry):
        g = torch._standard_gamma(torch.full(shape, 0.8, device=dev, dtype=torch.float32)).mul_(6.0)
        g[torch.rand(shape, device=dev, generator=gen) < p_dry] = 0.0
        return g

    # -- config 2: QuantileMapper, 10 000 cells x 10 950 days, one 10 950-step group per cell
    def c2():
        C = 10000
        y = torch.randn((T, C), device=dev, generator=gen) * 2 + 14
        x = torch.randn((T, C), device=dev, generator=gen) * 3 + 15
        qm = QuantileMapper()
        ms = timed(lambda: (qm.fit_batched(y), qm.transform_batched(x)), 2, 5)
        entry('QuantileMapper fit+transform, 10000 cells x 10950 days, one 10950-step group per cell (BASELINE config 2)',
              C, T, ms, 12, kernels='qm_fit_long_kernel + qm_predict_long_kernel (counting rank)')
    guarded('QuantileMapper 10000 x 10950', c2)

    # -- config 3 per GPU: BcsdPrecipitation, zero-inflated gamma, the 129 600-cell shard
    def c3():
        C = a.cells_per_gpu
        ytr, xp = precip((T, C), 0.5), precip((T, C), 0.55)
        bp = BcsdPrecipitation()
        out = torch.empty((T, C), device=dev)
        ms = timed(lambda: (bp.fit_batched(ytr, ytr, idx), bp.predict_batched(xp, idx, out=out)), 2, 4)
        bp.check_fit()
        entry(f'BcsdPrecipitation fit+predict, zero-inflated gamma, {C} cells x 10950 days (BASELINE config 3, per-GPU shard)', C, T, ms, 12)
    guarded('BcsdPrecipitation shard', c3)

    # -- config 3 as stated: the whole 720 x 1440 grid on ONE GPU, resident, processed in 8 cell tiles
    def c3_full():
        tiles, Ct = 8, 129600
        C = tiles * Ct
        free, _ = torch.cuda.mem_get_info(dev)
        need = 3 * T * C * 4 + 4 * T * Ct * 4
        if free < need + (4 << 30):
            raise RuntimeError(f'needs {need / 1e9:.0f} GB of device memory, {free / 1e9:.0f} GB free')
        ytr = torch.empty((T, C), device=dev)
        xp = torch.empty((T, C), device=dev)
        out = torch.empty((T, C), device=dev)
        for k in range(tiles):
            ytr[:, k * Ct:(k + 1) * Ct] = precip((T, Ct), 0.5)
            xp[:, k * Ct:(k + 1) * Ct] = precip((T, Ct), 0.55)
        bp = BcsdPrecipitation()

        def step():
            for k in range(tiles):
                s = slice(k * Ct, (k + 1) * Ct)
                bp.fit_batched(ytr[:, s], ytr[:, s], idx)
                bp.predict_batched(xp[:, s], idx, out=out[:, s])
        ms = timed(step, 1, 2)
        entry('BcsdPrecipitation fit+predict, zero-inflated gamma, 720x1440 = 1036800 cells x 10950 days on ONE GPU '
              '(BASELINE config 3: y_train / X_pred / out resident = 136 GB, 8 tiles of 129600 cells, row stride 1036800)', C, T, ms, 12)
    if world == 1 and left() > 25:
        guarded('BcsdPrecipitation full grid on one GPU', c3_full)

    # -- 'daily_nasa-nex' (PaddedDOYGrouper, groupers.py:19-89): 366 overlapping +-15-day pools at fit, day-of-month keyed
    #    predict with the month-grouped 9-sample window (neighbour-table kernel) — still on the generic kernels
    def nasanex():
        from skdownscale_b200.pointwise_models import BcsdTemperature as BT
        C = 4096
        season = torch.sin(2 * torch.pi * torch.arange(T, device=dev, dtype=torch.float32) / 365.25)[:, None]
        mk = lambda m_, a_, s_: torch.randn((T, C), device=dev, generator=gen) * s_ + m_ + a_ * season   # noqa: E731
        xtr, ytr, xp = mk(15, 10, 3), mk(14, 12, 2), mk(16.5, 10, 3)
        m = BT(time_grouper='daily_nasa-nex', return_anoms=False)
        out = torch.empty((T, C), device=dev)
        ms_fit = timed(lambda: m.fit_batched(xtr, ytr, idx), 1, 2)
        ms_pred = timed(lambda: m.predict_batched(xp, idx, out=out), 1, 2)
        entry(f"BcsdTemperature('daily_nasa-nex', return_anoms=False) fit+predict, {C} cells x 10950 days (generic kernels)", C, T,
              ms_fit + ms_pred, 16, ms_fit=ms_fit, ms_predict=ms_pred)
    if left() > 30:
        guarded('daily_nasa-nex', nasanex)

    # -- ZScoreRegressor (zscore.py): the HBM-streaming estimator of the family, per-GPU shard
    def zscore():
        from skdownscale_b200.pointwise_models import ZScoreRegressor
        C = a.cells_per_gpu
        mk = lambda m_, s_: torch.randn((T, C), device=dev, generator=gen) * s_ + m_   # noqa: E731
        xtr, ytr, xp = mk(15, 3), mk(14, 2), mk(16.5, 3)
        m = ZScoreRegressor()
        out = torch.empty((T, C), device=dev)
        ms_fit = timed(lambda: m.fit_batched(xtr, ytr, idx), 1, 3)
        ms_pred = timed(lambda: m.predict_batched(xp, out=out), 1, 3)
        m.check_fit()
        entry(f'ZScoreRegressor(window_width=31) fit+predict, {C} cells x 10950 days', C, T, ms_fit + ms_pred, 16,
              ms_fit=ms_fit, ms_predict=ms_pred, kernels='zscore_daysum_kernel + zscore_window_kernel + zscore_predict_kernel')
    if left() > 25:
        guarded('ZScoreRegressor', zscore)

    # -- configs 4 / 5: analog models, k = 10, 3 predictors; cell count per GPU = BASELINE's, or what the budget allows
    def analog(name, make, Tn, want_cells, probe_cells=2048):
        def run(Cn, steps):
            X = torch.randn((Tn, 3, Cn), device=dev, generator=gen)
            w = torch.tensor([1.0, 0.5, -0.3], device=dev)[None, :, None]
            yv = (X * w).sum(1) + 0.3 * torch.randn((Tn, Cn), device=dev, generator=gen)
            Xq = torch.randn((Tn, 3, Cn), device=dev, generator=gen)
            m = make()
            m.fit_batched(X, yv)
            return timed(lambda: m.predict_batched(Xq), 1, steps)
        ms_probe = run(probe_cells, 1)
        budget_ms = max(2.0, min(left() - 8.0, 20.0)) * 1e3 / 2.2          # warm-up + 1 timed step + generation
        fit_cells = budget_ms / ms_probe * probe_cells
        cells = want_cells if fit_cells >= want_cells else max(probe_cells, int(fit_cells) // 1024 * 1024)
        ms = run(cells, 1) if cells != probe_cells else ms_probe
        entry(f'{name}, {cells} cells/GPU x {Tn} days' + ('' if cells == want_cells else f' (BASELINE: {want_cells} cells/GPU; bounded by the bench time budget)'),
              cells, Tn, ms, 40, distance_evals_per_s=world * cells * float(Tn) * Tn / (ms * 1e-3))
    if left() > 12:
        guarded('PureAnalog', lambda: analog('PureAnalog(n_analogs=10, kind=mean_analogs), 3 predictors (BASELINE config 4)',
                                           lambda: PureAnalog(n_analogs=10, kind='mean_analogs'), 18250, 25000))
    if left() > 12:
        guarded('AnalogRegression', lambda: analog('AnalogRegression(n_analogs=10), 3 predictors (BASELINE config 5)',
                                                 lambda: AnalogRegression(n_analogs=10), 10950, 129600))
    return out_list


def run_b200(a):
    import torch
    import torch.distributed as dist
    import skdownscale_b200  # noqa: F401
    from skdownscale_b200 import _lib, engine
    from skdownscale_b200.pointwise_models import BcsdTemperature, PointWiseDownscaler
    import synth

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if world != a.gpus:
        raise SystemExit(f'--gpus {a.gpus} but WORLD_SIZE={world}: launch with torchrun --nproc-per-node {a.gpus}')
    dev = torch.device('cuda', local)
    torch.cuda.set_device(dev)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    _lib.load()
    if a.debug_flags is not None:
        _lib.load().sdb_set_debug_flags(a.debug_flags)

    T, C = a.days, a.cells_per_gpu
    idx = synth.daily_index(T)
    # synthetic temperature shard, generated on the device (SURVEY.md §8(d) formulas), seed + rank
    gen = torch.Generator(device=dev).manual_seed(1234 + rank)
    season = torch.sin(2 * torch.pi * torch.arange(T, device=dev, dtype=torch.float32) / 365.25)[:, None]

    def field(mean, amp, sd):
        x = torch.randn((T, C), device=dev, dtype=torch.float32, generator=gen)
        x.mul_(sd).add_(mean + amp * season)
        return x

    Xtr, ytr, Xp = field(15.0, 10.0, 3.0), field(14.0, 12.0, 2.0), field(16.5, 10.0, 3.0)
    out = torch.empty((T, C), device=dev, dtype=torch.float32)
    model = BcsdTemperature(return_anoms=True)

    ev_pred = []

    def step(time_predict=False):
        model.fit_batched(Xtr, ytr, idx)
        if time_predict:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
        model.predict_batched(Xp, idx, out=out)
        if time_predict:
            e1.record()
            ev_pred.append((e0, e1))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(a.warmup):
        step()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)
    barrier()
    t_wall0 = time.perf_counter()
    e_start, e_stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e_start.record()
    for _ in range(a.steps):
        step(time_predict=True)
    e_stop.record()
    barrier()
    t_wall1 = time.perf_counter()
    ms = torch.tensor([e_start.elapsed_time(e_stop)], device=dev, dtype=torch.float64)
    pred_ms = torch.tensor([sum(x.elapsed_time(y) for x, y in ev_pred) / len(ev_pred)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(pred_ms, op=dist.ReduceOp.MAX)
    clocks = sampler.stop(t_wall0, t_wall1) if rank == 0 else None
    ms_total = float(ms.item())
    ms_step = ms_total / a.steps
    value = world * C * T / (ms_step * 1e-3)
    model._state.check_finite()

    # ---- parity of the field that was just timed: sampled cells against the oracle (numpy port of the reference)
    parity = None
    if rank == 0:
        import oracle
        cells = [0, C // 2 + 3, C - 1]
        xs, ys, ps = (t[:, cells].cpu().numpy() for t in (Xtr, ytr, Xp))
        got = out[:, cells].cpu().numpy().astype(np.float64)
        groups = oracle.groups_from_keys(oracle.month_keys(idx))
        worst = 0.0
        for k in range(len(cells)):
            st = oracle.bcsd_temperature_fit(xs[:, k], ys[:, k], groups)
            ref = oracle.bcsd_temperature_predict(st, ps[:, k], groups, groups, True).astype(np.float32).astype(np.float64)
            worst = max(worst, float(np.max(np.abs(got[:, k] - ref) / np.maximum(np.abs(ref), np.std(ys[:, k])))))
        parity = {'cells': cells, 'max_err_rel_to_max(|ref|,sigma_y)': worst, 'tolerance': 1e-5, 'ok': bool(worst <= 1e-5)}
        if not parity['ok']:
            raise SystemExit(f'bench: the timed field is outside the parity tolerance: {parity}')

    # ---- the fused counting-rank entry on the same inputs (opt-in: measured slower on B200)
    fused = None
    if a.fused:
        out_f = torch.empty_like(out)
        fm = BcsdTemperature(return_anoms=True)
        for _ in range(2):
            fm.fit_predict_batched(Xtr, ytr, Xp, idx, out=out_f, keep_state=False, fused=True)
        barrier()
        f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        f0.record()
        for _ in range(a.steps):
            fm.fit_predict_batched(Xtr, ytr, Xp, idx, out=out_f, keep_state=False, fused=True)
        f1.record()
        barrier()
        fused = {'ms_per_step': f0.elapsed_time(f1) / a.steps, 'bit_identical_to_split': bool(torch.equal(out_f, out)),
                 'what': 'sdb_bcsd_fit_predict: climatologies + ONE kernel (counting rank in shared memory, no fitted state in HBM)'}
        del out_f

    # ---- the gather of the predicted field INSIDE the path: every finished cell chunk is pushed by the copy
    # engines into every peer's replica of the full field while the next chunk is computed (PeerGather)
    gather = None
    if world > 1:
        from skdownscale_b200.distributed import PeerGather
        specs = a.gather_method if a.gather_method != 'auto' else ('ce' if world <= 2 else 'bcast:32400')
        tried = []
        for spec in specs.split(','):
            gmethod, _, gchunk = spec.partition(':')
            gchunk = int(gchunk) if gchunk else a.gather_chunk
            pg = PeerGather(T, world * C, torch.float32, dev, method=gmethod)

            def gstep():
                model.fit_batched(Xtr, ytr, idx)
                return model.predict_gathered(Xp, idx, pg, chunk_cells=gchunk)

            gstep()                                          # warm-up (peer mappings, streams)
            barrier()
            g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            tg0 = time.perf_counter()
            g0.record()
            for _ in range(a.steps):
                fullf = gstep()
            g1.record()
            barrier()
            wall_g = (time.perf_counter() - tg0) * 1e3
            gms = torch.tensor([max(g0.elapsed_time(g1), wall_g) / a.steps], device=dev, dtype=torch.float64)
            dist.all_reduce(gms, op=dist.ReduceOp.MAX)
            # every rank's replica must hold what each peer computed: checksum of every rank's own block ...
            rows = torch.arange(0, T, 97, device=dev)
            mine = pg.local.index_select(0, rows).double().sum().reshape(1)
            sums = [torch.zeros(1, device=dev, dtype=torch.float64) for _ in range(world)]
            dist.all_gather(sums, mine)
            ok_g = True
            for r in range(world):                           # ... against the same block of MY replica
                blk = fullf[:, r * C:(r + 1) * C].index_select(0, rows).double().sum()
                ok_g = ok_g and bool(blk == sums[r][0])
            ok_t = torch.tensor([1 if ok_g and torch.equal(pg.local, out) else 0], device=dev)
            dist.all_reduce(ok_t, op=dist.ReduceOp.MIN)
            tried.append({'method': gmethod, 'chunk_cells': gchunk, 'ms_per_step_with_gather': float(gms.item()),
                          'field_verified_on_every_rank': bool(ok_t.item())})
            del fullf
            pg.close()
            torch.cuda.empty_cache()
        recv = (world - 1) * T * C * 4
        best = min((t for t in tried if t['field_verified_on_every_rank']), key=lambda t: t['ms_per_step_with_gather'], default=None)
        if best is None:
            raise SystemExit(f'bench: the gathered field is wrong on some rank: {tried}')
        step_g = best['ms_per_step_with_gather']
        gather = {'ms_per_step_with_gather': step_g, 'value_with_gather': world * C * T / (step_g * 1e-3),
                  'bytes_received_per_gpu': recv, 'GB/s_per_gpu_over_the_step': recv / (step_g * 1e-3) / 1e9,
                  'field_verified_on_every_rank': True, 'method': best['method'], 'chunk_cells': best['chunk_cells'],
                  'timed': tried, 'floor_ms': recv / 770e9 * 1e3,
                  'what': 'fit + predict in cell chunks written straight into the rank\'s columns of a full [T, n_cells] replica '
                          '(IPC-mapped on every peer); finished chunks are pushed to all peers while the next chunk computes — '
                          'ce: cudaMemcpy2DAsync per peer on its own stream, kernel: one SM copy kernel per peer, bcast: ONE kernel '
                          'that reads the chunk once and stores it into all peers; floor_ms = bytes received / 770 GB/s measured '
                          'peer-copy bandwidth'}

    # ---- end to end through the public API: pinned host inputs, H2D + fit + predict + D2H per step
    e2e = None
    if not a.no_e2e:
        numa, old_affinity = gpu_numa_affinity(local)
        h = [torch.empty((T, C), dtype=torch.float32, pin_memory=True) for _ in range(4)]
        for dst, src in zip(h[:3], (Xtr, ytr, Xp)):
            dst.copy_(src)
        torch.cuda.synchronize()
        Xtr = ytr = Xp = out = None
        model._state = None
        torch.cuda.empty_cache()
        pw = PointWiseDownscaler(BcsdTemperature(return_anoms=True), device=dev)

        def e2e_step():
            pw.fit(h[0], h[1], time=idx)                 # chunked H2D of X_train, y_train overlapped with the fit kernels
            pw.predict(h[2], time=idx, out=h[3])         # chunked H2D / predict / D2H on three streams, result in h[3]
            torch.cuda.synchronize()

        e2e_step()                                       # warm-up
        barrier()
        t0 = time.perf_counter()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(a.e2e_steps):
            e2e_step()
        e1.record()
        barrier()
        wall = time.perf_counter() - t0
        tms = torch.tensor([max(e0.elapsed_time(e1), wall * 1e3)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(tms, op=dist.ReduceOp.MAX)
        step_s = float(tms.item()) * 1e-3 / a.e2e_steps
        e2e = {'value': world * C * T / step_s, 'unit': UNIT,
               'h2d_bytes_per_step': 3 * T * C * 4, 'd2h_bytes_per_step': T * C * 4, 'steps': a.e2e_steps,
               'h2d_GBps_per_rank': 3 * T * C * 4 / step_s / 1e9, 'd2h_GBps_per_rank': T * C * 4 / step_s / 1e9,
               'host_binding_rank0': numa,
               'path': 'PointWiseDownscaler.fit(X, y) + .predict(X, out=pinned) on pinned host tensors: 16384-cell chunks, H2D / kernels / D2H overlapped on three streams'}

        # the host-link ceiling of the same step: the same bytes over the same pinned buffers in the same 16384-cell
        # chunks (3 arrays H2D on one stream, 1 array D2H on another), no kernels at all — every rank at once
        dbuf = [torch.empty((T, 16384), dtype=torch.float32, device=dev) for _ in range(2)]
        s_in, s_out = torch.cuda.Stream(dev), torch.cuda.Stream(dev)

        def bare_step():
            for c0 in range(0, C, 16384):
                c1 = min(C, c0 + 16384)
                for k in range(3):
                    engine.copy2d(dbuf[0][:, :c1 - c0], h[k][:, c0:c1], True, stream=s_in.cuda_stream)
                engine.copy2d(h[3][:, c0:c1], dbuf[1][:, :c1 - c0], False, stream=s_out.cuda_stream)
            torch.cuda.synchronize()

        bare_step()
        barrier()
        tb0 = time.perf_counter()
        for _ in range(a.e2e_steps):
            bare_step()
        barrier()
        bt = torch.tensor([(time.perf_counter() - tb0) / a.e2e_steps], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(bt, op=dist.ReduceOp.MAX)
        bare_s = float(bt.item())
        e2e['host_link_ceiling'] = {'ms_per_step': bare_s * 1e3, 'value': world * C * T / bare_s,
                                    'h2d_GBps_per_rank': 3 * T * C * 4 / bare_s / 1e9, 'd2h_GBps_per_rank': T * C * 4 / bare_s / 1e9,
                                    'e2e_over_ceiling': bare_s / step_s,
                                    'what': 'bare cudaMemcpy2DAsync loop (sdb_memcpy2d_async): the step\'s 17 GB H2D + 5.7 GB D2H per rank over the same pinned '
                                            'buffers and chunks, concurrently on every rank, no kernels'}
        dbuf = None
        h = pw = None
        if old_affinity:
            os.sched_setaffinity(0, old_affinity)        # the CPU baseline below uses every core again

    secondary = None
    if not a.no_secondary:
        Xtr = ytr = Xp = out = None                      # (already released when the e2e section ran)
        model._state = None
        torch.cuda.empty_cache()

        def reduce_max(v):
            t = torch.tensor([v], device=dev, dtype=torch.float64)
            if world > 1:
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t.item())
        secondary = run_secondary(a, dev, world, rank, barrier, reduce_max, measured_peak()[0])

    if rank == 0:
        peak, peak_src = measured_peak()
        pms = float(pred_ms.item())
        ach = ALG_BYTES_PREDICT * C * T / (pms * 1e-3) / 1e9
        ach_step = ALG_BYTES_STEP * C * T / (ms_step * 1e-3) / 1e9
        line = {
            'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': a.steps, 'warmup': a.warmup,
            'ms_per_step': ms_step, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
            'dtype': 'f32 data, f64 keys/interp', 'data': 'synthetic',
            'config': {'workload': f'BcsdTemperature fit+predict, monthly groups, return_anoms=True, {C} cells/GPU '
                                   f'(the 8-GPU shard of the 720x1440 grid) x {T} daily steps, f32',
                       'cells_per_gpu': C, 'days': T, 'parallelism': f'cells sharded over {world} GPU(s), no data-path collective',
                       'l2': 'inputs 17 GB/GPU >> 126 MB L2 (no flush needed)'},
            'clocks': clocks,
            'e2e': e2e,
            'gather': gather,
            'value_with_gather': gather['value_with_gather'] if gather else None,
            'parity_check': parity,
            'fused_counting_rank': fused,
            'gpu_launches': 4 * a.steps,
            'roofline': {'bound': 'hbm', 'kernel': 'qm_predict_tile_kernel<32,true> (dominant)', 'achieved': ach, 'peak': peak, 'unit': 'GB/s',
                         'frac': ach / peak, 'traffic': ncu_traffic('qm_predict_tile_kernel<32,true>', float(C) * T),
                         'traffic_source': 'profiles/r01_traffic.json (ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum per launch)',
                         'peak_source': peak_src,
                         'limiter_from_ncu': {'qm_predict_tile_kernel<32,true>': ncu_limiter('qm_predict_tile_kernel<32,true>'),
                                              'qm_fit_tile_kernel<32>': ncu_limiter('qm_fit_tile_kernel<32>')},
                         'algorithmic_bytes_per_cell_timestep': ALG_BYTES_PREDICT, 'kernel_ms': pms,
                         'whole_step': {'achieved': ach_step, 'frac': ach_step / peak,
                                        'algorithmic_bytes_per_cell_timestep': ALG_BYTES_STEP}},
        }
        if not a.no_cpu:
            line['cpu_baseline'] = cpu_baseline(T, a.cpu_seconds)
        line['secondary'] = secondary
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == '__main__':
    args = parse()
    if args.impl == 'reference':
        run_reference(args)
    else:
        run_b200(args)
