#!/bin/bash
# Build A/B variants of the BFGS kernel (bfgs_kernel.cuh switches) into tools/bin/ for the GPU box:
#   B200OPT_LIB=tools/bin/libb200opt_B_<tag>.so python bench.py --workload bfgs --no-e2e --no-cpu
# usage: tools/bfgs_variants.sh tag1:"-DBFGS_EARLY_B=0" tag2:"-DBFGS_NACC=2 -DBFGS_EARLY_B=1" ...
set -e
cd "$(dirname "$0")/.."
mkdir -p tools/bin
python -m optimistix_b200.build
OBJS=$(ls optimistix_b200/csrc/build/*.o | grep -v "/bfgs.o")
for spec in "$@"; do
  tag=${spec%%:*}; defs=${spec#*:}
  nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC --expt-relaxed-constexpr -Xptxas -v \
    $defs -c optimistix_b200/csrc/bfgs.cu -o tools/bin/bfgs_$tag.o 2>&1 | grep -A1 "RosenbrockIdEEd" | grep -E "registers|spill" || true
  nvcc -shared -o tools/bin/libb200opt_B_$tag.so $OBJS tools/bin/bfgs_$tag.o -gencode arch=compute_100a,code=sm_100a -ldl
  echo built tools/bin/libb200opt_B_$tag.so
done
