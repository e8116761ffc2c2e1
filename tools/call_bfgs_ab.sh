#!/bin/bash
# BFGS kernel A/B on one B200: tests with the default build, then CUDA-event timings of each variant library.
R=${1:-r2o}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_bfgs_gpu.py -q -m gpu -x -s > gpurun_out/pytest_bfgs_$R.log 2>&1; echo "pytest exit $?" | tee -a gpurun_out/pytest_bfgs_$R.log
for v in base "" B_e0 B_n2 B_n1 base ""; do
  lib=optimistix_b200/libb200opt.so; [ -n "$v" ] && lib=tools/bin/libb200opt_$v.so; [ "$v" = base ] && lib=tools/bin/libb200opt_base.so
  B200OPT_LIB=$lib timeout 300 python bench.py --workload bfgs --steps 5 --warmup 3 --no-e2e --no-cpu 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$v'.ljust(6), 'ms', round(d['ms_per_step'],3), d.get('checks'))" | tee -a gpurun_out/bfgs_ab_$R.log
done
