#!/bin/bash
# BFGS kernel A/B on one B200: CUDA-event timings of the given libraries (tags under tools/bin/, `default` = the in-tree build).
R=${1:-r2v}; shift
mkdir -p gpurun_out
for rep in 1 2; do
for v in "$@" default; do
  lib=tools/bin/libb200opt_$v.so; [ "$v" = default ] && lib=optimistix_b200/libb200opt.so
  B200OPT_LIB=$lib timeout 300 python bench.py --workload bfgs --steps 5 --warmup 3 --no-e2e --no-cpu 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$v'.ljust(10), 'ms', round(d['ms_per_step'],3), d.get('checks'))" | tee -a gpurun_out/bfgs_ab_$R.log
done; done
