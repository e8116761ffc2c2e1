#!/bin/bash
# Build A/B variants of ONE translation unit of the library into tools/bin/ for the GPU box (the other objects are the default build):
#   tools/kernel_variants.sh lsq_gauss2 o0:"-DLSQ_P2_OPT=0" o7:"-DLSQ_P2_OPT=7"      -> tools/bin/libb200opt_lsq_gauss2_o0.so ...
#   B200OPT_LIB=tools/bin/libb200opt_lsq_gauss2_o0.so python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu --no-secondary
set -e
cd "$(dirname "$0")/.."
src=$1; shift
mkdir -p tools/bin
python -m optimistix_b200.build
OBJS=$(ls optimistix_b200/csrc/build/*.o | grep -v "/$src.o")
for spec in "$@"; do
  tag=${spec%%:*}; defs=${spec#*:}
  nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC --expt-relaxed-constexpr -Xptxas -v \
    $defs -c optimistix_b200/csrc/$src.cu -o tools/bin/${src}_$tag.o > tools/bin/${src}_$tag.log 2>&1
  nvcc -shared -o tools/bin/libb200opt_${src}_$tag.so $OBJS tools/bin/${src}_$tag.o -gencode arch=compute_100a,code=sm_100a -ldl
  echo built tools/bin/libb200opt_${src}_$tag.so
done
