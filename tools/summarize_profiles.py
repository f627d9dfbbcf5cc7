"""Turn the evidence run of tools/gpu_round_e.sh (gpurun_out/e_*) into the tracked files under profiles/:
launch list, raw metrics of the full-shard capture, sanitizer logs and the traffic / limiter json that
bench.py reads.   python tools/summarize_profiles.py [tag]      (tag default r01)"""
import csv
import json
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else 'r01'
src = os.path.join(ROOT, 'gpurun_out')
dst = os.path.join(ROOT, 'profiles')

shutil.copy(os.path.join(src, 'e_launches.csv'), os.path.join(dst, f'{tag}_launches_tile_final.csv'))
shutil.copy(os.path.join(src, 'e_full_shard_raw.csv'), os.path.join(dst, f'{tag}_full_shard_ncu_raw.csv'))
for tool in ('memcheck', 'racecheck'):
    with open(os.path.join(src, f'e_{tool}.log')) as f:
        lines = [l for l in f if l.startswith('=========') or 'sanitize workload' in l]
    with open(os.path.join(dst, f'{tag}_compute_sanitizer_{tool}.log'), 'w') as f:
        f.writelines(lines)

rows = list(csv.reader(open(os.path.join(src, 'e_full_shard_raw.csv'))))
hdr = rows[0]
col = {h: i for i, h in enumerate(hdr)}
bench = json.loads(open(os.path.join(src, 'e_bench.json')).read().strip().splitlines()[-1])
kern = {}
for r in rows[2:]:
    name = r[col['Kernel Name']]
    short = ('qm_predict_tile_kernel<32,true>' if 'qm_predict_tile' in name else
             'qm_fit_tile_kernel<32>' if 'qm_fit_tile' in name else 'group_mean_kernel<float>')
    g = lambda m: float(r[col[m]])      # noqa: E731
    kern[short] = {
        'dram_bytes_read': g('dram__bytes_read.sum') * 1e9, 'dram_bytes_write': g('dram__bytes_write.sum') * 1e9,
        'duration_ms_under_ncu': g('gpu__time_duration.sum'),
        'alu_pipe_pct': g('sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active'),
        'fma_pipe_pct': g('sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active'),
        'issue_active_pct': g('smsp__issue_active.avg.pct_of_peak_sustained_active'),
        'dram_throughput_pct': g('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed'),
        'registers': int(float(r[col['launch__registers_per_thread']])),
        'warp_instructions': g('smsp__inst_executed.sum'),
    }
out = {'source': f'ncu --set full --clock-control none, bench.py --steps 1 --warmup 0 --no-e2e --no-cpu '
                 f'({bench["config"]["cells_per_gpu"]} cells x {bench["config"]["days"]} days), profiles/{tag}_full_shard_ncu_raw.csv',
       'cells': bench['config']['cells_per_gpu'], 'days': bench['config']['days'], 'kernels': kern}
json.dump(out, open(os.path.join(dst, f'{tag}_traffic.json'), 'w'), indent=1)
print(json.dumps(out, indent=1))
