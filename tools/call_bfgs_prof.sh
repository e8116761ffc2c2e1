#!/bin/bash
R=${1:-r2p}
mkdir -p gpurun_out
timeout 200 python bench.py --workload bfgs --steps 5 --warmup 3 --no-e2e --no-cpu > gpurun_out/plain_bfgs_$R.log 2>&1; echo "plain bfgs $?"; tail -1 gpurun_out/plain_bfgs_$R.log | cut -c1-200
timeout 300 ncu --set full --clock-control none --import-source on -k regex:bfgs_solve_kernel -s 3 -c 1 -f -o gpurun_out/bfgs_$R python bench.py --workload bfgs --batch 4096 --steps 2 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_bfgs_$R.log 2>&1; echo "bfgs $?"
