#!/bin/bash
# gpu tests on the main build, then bench the main build and the variants given as arguments
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/b_pytest.log 2>&1
tail -3 gpurun_out/b_pytest.log
rm -f gpurun_out/b_variants.log
for v in "" "$@"; do
  if [ -z "$v" ]; then unset SDB_LIBRARY; else export SDB_LIBRARY=$PWD/scikit-downscale_b200/csrc/variants/libsdb_$v.so; fi
  echo -n "variant=$v " | tee -a gpurun_out/b_variants.log
  timeout 300 python bench.py --no-cpu --no-e2e --steps 10 --warmup 3 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['roofline']['kernel_ms'])" | tee -a gpurun_out/b_variants.log
  if [ -n "$v" ]; then SDB_LIBRARY=$SDB_LIBRARY timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -1 | tee -a gpurun_out/b_variants.log; fi
done
