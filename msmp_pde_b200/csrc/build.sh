#!/bin/bash
# Builds libmsmp_b200.so (hand-written sm_100a kernels + C ABI) in-tree.
set -e
cd "$(dirname "$0")"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -I../../include -I. ${MSMP_NVCC_EXTRA}"
mkdir -p build
pids=()
for f in api pack optim prep g2 linear linear_tc linear_tma linear_ts wgrad_tc wgrad_ws edge edge_tc edge_ws norm lem lem_tc decoder decoder_rt; do
  [ -f $f.cu ] || continue
  if [ ! -f build/$f.o ] || [ $f.cu -nt build/$f.o ] || [ common.cuh -nt build/$f.o ] || [ umma.cuh -nt build/$f.o ] || [ linear_common.cuh -nt build/$f.o ] || [ ../../include/msmp_b200.h -nt build/$f.o ]; then
    $NVCC $FLAGS -Xptxas -v -c $f.cu -o build/$f.o > build/$f.ptxas.log 2>&1 &
    pids+=($!)
  fi
done
for p in "${pids[@]}"; do wait $p || { cat build/*.ptxas.log | grep -v "^ptxas info" | head -50; exit 1; }; done
$NVCC -gencode arch=compute_100a,code=sm_100a -shared -o libmsmp_b200.so build/*.o -lcudart
echo "built $(pwd)/libmsmp_b200.so"
