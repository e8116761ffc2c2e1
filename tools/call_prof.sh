# ncu captures of the three persistent solve kernels (each only after the same command exited 0 without ncu)
set -x
R=${1:-r09}
python bench.py --batch 131072 --steps 2 --warmup 3 --no-e2e --no-cpu > gpurun_out/plain_lsq.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:lsq_solve_kernel -s 3 -c 1 -f -o gpurun_out/lsq_$R python bench.py --batch 131072 --steps 2 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_lsq.log 2>&1; echo "lsq $?"
python bench.py --workload newton --batch 32768 --steps 2 --warmup 3 --no-e2e --no-cpu > gpurun_out/plain_newton.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:root_solve_kernel -s 3 -c 1 -f -o gpurun_out/newton_$R python bench.py --workload newton --batch 32768 --steps 2 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_newton.log 2>&1; echo "newton $?"
python bench.py --workload bfgs --batch 4096 --steps 2 --warmup 3 --no-e2e --no-cpu > gpurun_out/plain_bfgs.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:bfgs_solve_kernel -s 3 -c 1 -f -o gpurun_out/bfgs_$R python bench.py --workload bfgs --batch 4096 --steps 2 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_bfgs.log 2>&1; echo "bfgs $?"
ls -la gpurun_out/*_$R.ncu-rep
