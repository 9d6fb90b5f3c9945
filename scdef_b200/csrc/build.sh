#!/usr/bin/env bash
# Builds libscdef_b200.so for sm_100a (in-tree; the .so travels to the GPU box with gpurun).
set -euo pipefail
cd "$(dirname "$0")"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
FLAGS="${SCDEF_EXTRA_FLAGS:-} -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -Xptxas -v"
mkdir -p build
for f in api lik_tc; do
  $NVCC $FLAGS -c $f.cu -o build/$f.o 2> build/$f.ptxas.log || { cat build/$f.ptxas.log; exit 1; }
done
$NVCC -shared -gencode arch=compute_100a,code=sm_100a -o ../libscdef_b200.so build/api.o build/lik_tc.o
echo "built $(cd .. && pwd)/libscdef_b200.so"
