#!/bin/bash
# dram traffic + time of the tile kernels for library variants: tools/gpu_round_f.sh v1 v2 ...
mkdir -p gpurun_out; rm -f gpurun_out/f_*
run() {  # label
  timeout 300 python bench.py --no-cpu --no-e2e --steps 10 --warmup 3 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$1 bench ms', d['ms_per_step'], d['roofline']['kernel_ms'])" | tee -a gpurun_out/f_bench.log
  timeout 300 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:'qm_predict_tile|qm_fit_tile' -c 2 --csv --log-file gpurun_out/f_ncu_$1.csv python bench.py --steps 1 --warmup 0 --no-e2e --no-cpu > /dev/null 2>&1
}
unset SDB_LIBRARY; run main
for v in "$@"; do export SDB_LIBRARY=$PWD/scikit-downscale_b200/csrc/variants/libsdb_$v.so; run $v; done
unset SDB_LIBRARY
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -1
