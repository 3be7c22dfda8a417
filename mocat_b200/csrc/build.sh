#!/bin/bash
# Build libmocat_b200.so for sm_100a in-tree (the .so travels to the GPU box; it is git-ignored).
set -e
cd "$(dirname "$0")"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
FLAGS="$MB_EXTRA_FLAGS -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -Xptxas -v --expt-relaxed-constexpr"
mkdir -p build
pids=()
for f in abi comm reduce resample resample_fused propagate pf_l96 enkf backward abc svgd svgd_tc ksd teki; do
  [ -f $f.cu ] || continue
  if [ ! -f build/$f.o ] || [ $f.cu -nt build/$f.o ] || [ common.cuh -nt build/$f.o ] || [ select.cuh -nt build/$f.o ] || [ comm.cuh -nt build/$f.o ] || [ rng.cuh -nt build/$f.o ] || [ pf_common.cuh -nt build/$f.o ] || [ gk.cuh -nt build/$f.o ] || [ ../../include/mocat_b200.h -nt build/$f.o ]; then
    ( $NVCC $FLAGS -c $f.cu -o build/$f.o > build/$f.log 2>&1 || { cat build/$f.log; exit 1; } ) &
    pids+=($!)
  fi
done
for p in "${pids[@]}"; do wait $p; done
$NVCC -gencode arch=compute_100a,code=sm_100a -shared -o ../libmocat_b200.so build/*.o -lcudart
echo "built $(cd ..; pwd)/libmocat_b200.so"
