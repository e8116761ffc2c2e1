#!/bin/bash
# Round-2 verification on one B200: full GPU test suite, smoke, the bench line, the reference arm, launch lists and
# `ncu --set full` captures of the kernels that changed this round (each only after the same command exited 0 without ncu).
R=${1:-r2z}
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -m gpu > gpurun_out/pytest_$R.log 2>&1; echo "pytest exit $?" | tee -a gpurun_out/pytest_$R.log
timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke_$R.log 2>&1; echo "smoke exit $?" | tee -a gpurun_out/smoke_$R.log
timeout 900 python bench.py > gpurun_out/bench_$R.json 2> gpurun_out/bench_$R.err; echo "bench exit $?" | tee -a gpurun_out/bench_$R.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_$R.json 2> gpurun_out/bench_ref_$R.err; echo "ref exit $?"
# launch list of the headline command (share of the step per kernel)
python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-secondary > gpurun_out/plain_lm_$R.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"lsq_solve|root_solve|bfgs_solve|gram|lsq_step|gn_" -c 40 --csv --log-file gpurun_out/launches_lm_$R.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-secondary > gpurun_out/ncu_l1_$R.log 2>&1; echo "launch list lm $?"
# launch list of the large problem through the public API (device loop) and the kernel-only loops
python tools/gram_api_probe.py --rows 8388608 --iters 4 --reps 1 > gpurun_out/plain_gram_$R.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_gram_$R.csv python tools/gram_api_probe.py --rows 8388608 --iters 4 --reps 1 > gpurun_out/ncu_l2_$R.log 2>&1; echo "launch list gram $?"
timeout 300 python tools/gram_api_probe.py --out gpurun_out/gram_probe_$R.json > gpurun_out/gram_probe_$R.log 2>&1
# --set full captures
ncu --set full --clock-control none --import-source on -k regex:lsq_solve_kernel -s 3 -c 1 -f -o gpurun_out/lsq_$R python bench.py --batch 131072 --steps 2 --warmup 3 --no-e2e --no-cpu --no-secondary > gpurun_out/ncu_lsq_$R.log 2>&1; echo "lsq $?"
ncu --set full --clock-control none --import-source on -k regex:gn_step_kernel -s 2 -c 1 -f -o gpurun_out/gnstep_$R python tools/gram_api_probe.py --rows 4194304 --iters 4 --reps 1 > gpurun_out/ncu_step_$R.log 2>&1; echo "step $?"
ncu --set full --clock-control none --import-source on -k regex:gram_kernel -s 14 -c 1 -f -o gpurun_out/gramstate_$R python tools/gram_api_probe.py --rows 8388608 --iters 4 --reps 1 > gpurun_out/ncu_gram_$R.log 2>&1; echo "gram $?"
python bench.py --workload bfgs --batch 4096 --steps 2 --warmup 3 --no-e2e --no-cpu > gpurun_out/plain_bfgs_$R.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:bfgs_solve_kernel -s 3 -c 1 -f -o gpurun_out/bfgs_$R python bench.py --workload bfgs --batch 4096 --steps 2 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_bfgs_$R.log 2>&1; echo "bfgs $?"
python bench.py --workload newton --batch 32768 --steps 2 --warmup 3 --no-e2e --no-cpu > gpurun_out/plain_newton_$R.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:root_solve_kernel -s 3 -c 1 -f -o gpurun_out/newton_$R python bench.py --workload newton --batch 32768 --steps 2 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_newton_$R.log 2>&1; echo "newton $?"
python bench.py --workload exp_decay --steps 2 --warmup 3 --no-e2e --no-cpu > gpurun_out/plain_exp_$R.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:lsq_solve_kernel -s 3 -c 1 -f -o gpurun_out/lsqexp_$R python bench.py --workload exp_decay --steps 2 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_exp_$R.log 2>&1; echo "exp $?"
tail -3 gpurun_out/pytest_$R.log; tail -2 gpurun_out/smoke_$R.log
