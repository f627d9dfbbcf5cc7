#!/bin/bash
# gpu tests + bench of the main build, then source-level ncu captures of the two tile kernels
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/c_pytest.log 2>&1
tail -3 gpurun_out/c_pytest.log
timeout 300 python bench.py --no-cpu --no-e2e --steps 10 --warmup 3 2>&1 | tail -1 | tee gpurun_out/c_bench.json | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['roofline']['kernel_ms'])"
for v in "$@"; do
  export SDB_LIBRARY=$PWD/scikit-downscale_b200/csrc/variants/libsdb_$v.so
  echo -n "variant=$v "
  timeout 300 python bench.py --no-cpu --no-e2e --steps 10 --warmup 3 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['roofline']['kernel_ms'])"
  timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -1
done
unset SDB_LIBRARY
timeout 600 ncu --set full --import-source on --clock-control none -k regex:qm_predict_tile -c 1 -o gpurun_out/c_predict_src -f python tools/profile_one.py temp 16384 > gpurun_out/c_ncu.log 2>&1
timeout 300 ncu --set full --import-source on --clock-control none -k regex:qm_fit_tile -c 1 -o gpurun_out/c_fit_src -f python tools/profile_one.py temp 16384 > gpurun_out/c_ncu_fit.log 2>&1
tail -1 gpurun_out/c_ncu_fit.log
