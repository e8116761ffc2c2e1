set -x
timeout 200 python bench.py --workload bfgs --steps 5 --warmup 3 --no-e2e --no-cpu > gpurun_out/bv_ilp.log 2>&1; echo "ilp exit $?"
B200OPT_LIB=tools/bin/libb200opt_spec.so timeout 200 python bench.py --workload bfgs --steps 5 --warmup 3 --no-e2e --no-cpu > gpurun_out/bv_spec2.log 2>&1; echo "spec exit $?"
B200OPT_LIB=tools/bin/libb200opt_M3.so timeout 100 python bench.py --workload newton --steps 5 --warmup 3 --no-e2e --no-cpu > gpurun_out/nv_M3.log 2>&1; echo "M3 exit $?"
timeout 100 python bench.py --workload newton --steps 5 --warmup 3 --no-e2e --no-cpu > gpurun_out/nv_def.log 2>&1; echo "newton default exit $?"
timeout 600 python -m pytest tests/test_bfgs_gpu.py -q -m gpu > gpurun_out/pytest_r2e.log 2>&1; echo "pytest exit $?"
tail -5 gpurun_out/pytest_r2e.log
