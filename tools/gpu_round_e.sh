#!/bin/bash
# final evidence run: full default bench, launch list, full-shard ncu capture of the step's kernels, sanitizers
mkdir -p gpurun_out
timeout 900 python bench.py > gpurun_out/e_bench.json 2> gpurun_out/e_bench.err; tail -c 600 gpurun_out/e_bench.json
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/e_bench_ref.json 2>> gpurun_out/e_bench.err; tail -c 300 gpurun_out/e_bench_ref.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/e_launches.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu > gpurun_out/e_launches.log 2>&1
timeout 900 ncu --set full --clock-control none -k regex:'qm_predict_tile|qm_fit_tile|group_mean' -c 4 -o gpurun_out/e_full_shard -f python bench.py --steps 1 --warmup 0 --no-e2e --no-cpu > gpurun_out/e_ncu_full.log 2>&1
ncu -i gpurun_out/e_full_shard.ncu-rep --page raw --csv > gpurun_out/e_full_shard_raw.csv 2>/dev/null
timeout 900 compute-sanitizer --tool memcheck python tools/sanitize.py > gpurun_out/e_memcheck.log 2>&1; tail -3 gpurun_out/e_memcheck.log
timeout 900 compute-sanitizer --tool racecheck python tools/sanitize.py > gpurun_out/e_racecheck.log 2>&1; tail -3 gpurun_out/e_racecheck.log
