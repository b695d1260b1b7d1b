#!/bin/bash
# Build libxesn_b200.so (sm_100a only) in-tree.  Usage: build.sh [extra nvcc flags]
set -euo pipefail
cd "$(dirname "$0")"
OUT=../libxesn_b200.so
NVCC=${NVCC:-nvcc}
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -Xcompiler -O3"
mkdir -p build
pids=()
for f in gemm_f64 recurrence cholesky fused_train gram_i8 predict_loop predict_pull cost api; do
  $NVCC $FLAGS "$@" -c $f.cu -o build/$f.o &
  pids+=($!)
done
for p in "${pids[@]}"; do wait $p; done
$NVCC -shared -o $OUT build/gemm_f64.o build/recurrence.o build/cholesky.o build/fused_train.o build/gram_i8.o build/predict_loop.o build/predict_pull.o build/cost.o build/api.o -lcudart
echo "built $(realpath $OUT)"
