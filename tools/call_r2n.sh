#!/bin/bash
# Re-entry check on one B200: GPU test suite, smoke, and --set full captures (with source) of the three batched kernels.
R=${1:-r2n}
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -m gpu -x > gpurun_out/pytest_$R.log 2>&1; echo "pytest exit $?" | tee -a gpurun_out/pytest_$R.log
timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke_$R.log 2>&1; echo "smoke exit $?" | tee -a gpurun_out/smoke_$R.log
timeout 200 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu --no-secondary > gpurun_out/plain_lm_$R.log 2>&1; echo "plain lm $?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:lsq_solve_kernel -s 3 -c 1 -f -o gpurun_out/lsq_$R python bench.py --batch 131072 --steps 2 --warmup 3 --no-e2e --no-cpu --no-secondary > gpurun_out/ncu_lsq_$R.log 2>&1; echo "lsq $?"
timeout 200 python bench.py --workload bfgs --batch 4096 --steps 2 --warmup 3 --no-e2e --no-cpu > gpurun_out/plain_bfgs_$R.log 2>&1; echo "plain bfgs $?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:bfgs_solve_kernel -s 3 -c 1 -f -o gpurun_out/bfgs_$R python bench.py --workload bfgs --batch 4096 --steps 2 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_bfgs_$R.log 2>&1; echo "bfgs $?"
timeout 200 python bench.py --workload newton --batch 32768 --steps 2 --warmup 3 --no-e2e --no-cpu > gpurun_out/plain_newton_$R.log 2>&1; echo "plain newton $?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:root_solve_kernel -s 3 -c 1 -f -o gpurun_out/newton_$R python bench.py --workload newton --batch 32768 --steps 2 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_newton_$R.log 2>&1; echo "newton $?"
tail -3 gpurun_out/pytest_$R.log; tail -2 gpurun_out/smoke_$R.log; tail -1 gpurun_out/plain_lm_$R.log | cut -c1-400
