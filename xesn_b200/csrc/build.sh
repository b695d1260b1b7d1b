#!/bin/bash
# Build libxesn_b200.so (sm_100a only) in-tree.  Usage: build.sh [extra nvcc flags]
# XESN_OUT / XESN_BUILD_DIR: other output library / object directory (debug builds: XESN_OUT=../libxesn_b200_trace.so
# XESN_BUILD_DIR=build_trace build.sh -DXESN_RECUR_TRACE; run with XESN_B200_LIB=<that file>)
set -euo pipefail
cd "$(dirname "$0")"
OUT=${XESN_OUT:-../libxesn_b200.so}
BUILD=${XESN_BUILD_DIR:-build}
NVCC=${NVCC:-nvcc}
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -Xcompiler -O3"
mkdir -p $BUILD
pids=()
for f in gemm_f64 recurrence cholesky lu_solve fused_train gram_i8 predict_loop predict_pull predict_flow cost bank_layout api; do
  $NVCC $FLAGS "$@" -c $f.cu -o $BUILD/$f.o &
  pids+=($!)
done
for p in "${pids[@]}"; do wait $p; done
$NVCC -shared -o $OUT $BUILD/gemm_f64.o $BUILD/recurrence.o $BUILD/cholesky.o $BUILD/lu_solve.o $BUILD/fused_train.o $BUILD/gram_i8.o $BUILD/predict_loop.o $BUILD/predict_pull.o $BUILD/predict_flow.o $BUILD/cost.o $BUILD/bank_layout.o $BUILD/api.o -lcudart
echo "built $(realpath $OUT)"
