#!/bin/bash
# Build tuning variants of the Gram kernel (tile rows x pipeline stages) into tools/bin/ for A/B runs on the GPU box:
#   B200OPT_LIB=tools/bin/libb200opt_K64_S3.so python bench.py --workload gram ...
set -e
cd "$(dirname "$0")/.."
mkdir -p tools/bin
OBJS=$(ls optimistix_b200/csrc/build/*.o | grep -v gram_tcgen05)
for v in "$@"; do
  K=${v%x*}; S=${v#*x}
  nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC --expt-relaxed-constexpr \
    -DGRAM_KROWS=$K -DGRAM_STAGES=$S -c optimistix_b200/csrc/gram_tcgen05.cu -o tools/bin/gram_K${K}_S${S}.o
  nvcc -shared -o tools/bin/libb200opt_K${K}_S${S}.so $OBJS tools/bin/gram_K${K}_S${S}.o -gencode arch=compute_100a,code=sm_100a -lcuda
  echo built tools/bin/libb200opt_K${K}_S${S}.so
done
