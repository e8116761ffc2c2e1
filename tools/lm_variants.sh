#!/bin/bash
# Build A/B variants of the packed LM kernel (LSQ_P2_FAST bit mask, see lsq_kernel_p2.cuh) into tools/bin/ for the GPU box:
#   B200OPT_LIB=tools/bin/libb200opt_F7.so python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu
# usage: tools/lm_variants.sh 0 1 3 7 15  ["-DLSQ_P2_MINBLOCKS=5" as last argument starting with -D is passed through]
set -e
cd "$(dirname "$0")/.."
mkdir -p tools/bin
python -m optimistix_b200.build
OBJS=$(ls optimistix_b200/csrc/build/*.o | grep -v lsq_gauss2)
for v in "$@"; do
  nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC --expt-relaxed-constexpr \
    -DLSQ_P2_FAST=$v -c optimistix_b200/csrc/lsq_gauss2.cu -o tools/bin/lsq_gauss2_F$v.o
  nvcc -shared -o tools/bin/libb200opt_F$v.so $OBJS tools/bin/lsq_gauss2_F$v.o -gencode arch=compute_100a,code=sm_100a -ldl
  echo built tools/bin/libb200opt_F$v.so
done
