#!/bin/bash
# Build A/B variants of the Newton kernel (ROOT_LOOKAHEAD x ROOT_GJ, see newton_kernel.cuh) into tools/bin/ for the GPU box:
#   B200OPT_LIB=tools/bin/libb200opt_N11.so python bench.py --workload newton --no-e2e --no-cpu
# usage: tools/newton_variants.sh 00 10 01 11   (first digit: look-ahead, second: Gauss-Jordan)
set -e
cd "$(dirname "$0")/.."
mkdir -p tools/bin
python -m optimistix_b200.build
OBJS=$(ls optimistix_b200/csrc/build/*.o | grep -v "/newton.o")
for v in "$@"; do
  nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC --expt-relaxed-constexpr \
    -DROOT_LOOKAHEAD=${v:0:1} -DROOT_GJ=${v:1:1} -c optimistix_b200/csrc/newton.cu -o tools/bin/newton_N$v.o
  nvcc -shared -o tools/bin/libb200opt_N$v.so $OBJS tools/bin/newton_N$v.o -gencode arch=compute_100a,code=sm_100a -ldl
  echo built tools/bin/libb200opt_N$v.so
done
