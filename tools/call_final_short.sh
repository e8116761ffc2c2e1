#!/bin/bash
# Final check of the tree on one B200: GPU test suite, smoke, bench line, reference arm, launch list and one capture of the headline kernel.
R=${1:-r2zz}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/pytest_$R.log 2>&1; echo "pytest exit $?" | tee -a gpurun_out/pytest_$R.log
timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke_$R.log 2>&1; echo "smoke exit $?" | tee -a gpurun_out/smoke_$R.log
timeout 900 python bench.py > gpurun_out/bench_$R.json 2> gpurun_out/bench_$R.err; echo "bench exit $?" | tee -a gpurun_out/bench_$R.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_$R.json 2> gpurun_out/bench_ref_$R.err; echo "ref exit $?"
python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-secondary > gpurun_out/plain_lm_$R.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"lsq_solve|root_solve|bfgs_solve|gram|lsq_step|gn_" -c 40 --csv --log-file gpurun_out/launches_lm_$R.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-secondary > gpurun_out/ncu_l1_$R.log 2>&1; echo "launch list lm $?"
python bench.py --batch 131072 --steps 2 --warmup 3 --no-e2e --no-cpu --no-secondary > gpurun_out/plain_lsq_$R.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:lsq_solve_kernel -s 3 -c 1 -f -o gpurun_out/lsq_$R python bench.py --batch 131072 --steps 2 --warmup 3 --no-e2e --no-cpu --no-secondary > gpurun_out/ncu_lsq_$R.log 2>&1; echo "lsq $?"
tail -3 gpurun_out/pytest_$R.log; tail -2 gpurun_out/smoke_$R.log
