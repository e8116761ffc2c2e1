#!/bin/bash
# One --set full capture (with source) of the headline kernel at 131,072 fits, after a plain run of the same command.
R=${1:-r2y}
mkdir -p gpurun_out
python bench.py --batch 131072 --steps 2 --warmup 3 --no-e2e --no-cpu --no-secondary > gpurun_out/plain_lsq_$R.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:lsq_solve_kernel -s 3 -c 1 -f -o gpurun_out/lsq_$R python bench.py --batch 131072 --steps 2 --warmup 3 --no-e2e --no-cpu --no-secondary > gpurun_out/ncu_lsq_$R.log 2>&1; echo "lsq $?"
