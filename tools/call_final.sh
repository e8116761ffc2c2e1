# Round-end verification: full GPU test suite, the bench lines quoted in DESIGN.md, ncu captures of the solve kernels.
set -x
R=${1:-r10}
timeout 900 python -m pytest tests -q -m gpu > gpurun_out/pytest_$R.log 2>&1; echo "pytest exit $?" | tee -a gpurun_out/pytest_$R.log
timeout 300 python bench.py > gpurun_out/bench_default.log 2>&1; echo "bench exit $?"
for W in newton bfgs exp_decay; do timeout 300 python bench.py --workload $W --steps 5 --warmup 3 > gpurun_out/bench_$W.log 2>&1; echo "$W exit $?"; done
timeout 200 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference.log 2>&1; echo "ref exit $?"
python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu > gpurun_out/plain_lm.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"lsq_solve|root_solve|bfgs_solve|gram|lsq_step|linear_residual|gn_solve" -c 40 --csv --log-file gpurun_out/launches_lm_$R.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_l1.log 2>&1; echo "launch list $?"
bash tools/call_prof.sh $R
tail -3 gpurun_out/pytest_$R.log
