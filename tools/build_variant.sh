#!/bin/bash
# Build an experimental libsdb variant: tools/build_variant.sh NAME -DSDB_FOO=1 ...  → csrc/variants/libsdb_NAME.so
# (only the 1024-member tile TU is recompiled; run with SDB_LIBRARY=<that file>).
set -e
cd "$(dirname "$0")/../scikit-downscale_b200/csrc"
name=$1; shift
mkdir -p variants
nvcc -O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC -fmad=false \
     --expt-relaxed-constexpr -Xptxas -v "$@" -c qm_np1024.cu -o variants/qm_np1024_$name.o 2> variants/$name.ptxas.log
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o variants/libsdb_$name.so \
     sdb_api.o qm_api.o qm_np256.o variants/qm_np1024_$name.o qm_np4096.o qm_np16384.o analog_kernels.o qmr_kernels.o trend_kernels.o
echo variants/libsdb_$name.so
