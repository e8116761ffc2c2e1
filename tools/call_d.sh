set -x
timeout 200 python bench.py --workload bfgs --steps 5 --warmup 3 --no-e2e --no-cpu > gpurun_out/bv_spec.log 2>&1; echo "spec exit $?"
B200OPT_LIB=tools/bin/libb200opt_B1.so timeout 200 python bench.py --workload bfgs --steps 5 --warmup 3 --no-e2e --no-cpu > gpurun_out/bv_B1b.log 2>&1; echo "B1 exit $?"
timeout 600 python -m pytest tests/test_bfgs_gpu.py -q -m gpu > gpurun_out/pytest_r2d.log 2>&1; echo "pytest exit $?"
tail -5 gpurun_out/pytest_r2d.log
