#!/bin/bash
# one gpurun call: gpu tests, baseline + variant benches, source-level ncu capture of the predict kernel
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/a_smi.txt 2>&1
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/a_pytest.log 2>&1
tail -3 gpurun_out/a_pytest.log
for v in "" h1 h2 h3 h7; do
  if [ -z "$v" ]; then unset SDB_LIBRARY; else export SDB_LIBRARY=$PWD/scikit-downscale_b200/csrc/variants/libsdb_$v.so; fi
  echo -n "variant=$v " | tee -a gpurun_out/a_variants.log
  timeout 300 python bench.py --no-cpu --no-e2e --steps 10 --warmup 3 2>&1 | tail -1 | tee -a gpurun_out/a_variants.log
done
unset SDB_LIBRARY
timeout 600 ncu --set full --import-source on --clock-control none -k regex:qm_predict_tile -c 1 -o gpurun_out/a_predict_src -f python tools/profile_one.py temp 16384 > gpurun_out/a_ncu.log 2>&1
tail -2 gpurun_out/a_ncu.log
timeout 300 ncu --set full --import-source on --clock-control none -k regex:qm_fit_tile -c 1 -o gpurun_out/a_fit_src -f python tools/profile_one.py temp 16384 > gpurun_out/a_ncu_fit.log 2>&1
tail -2 gpurun_out/a_ncu_fit.log
