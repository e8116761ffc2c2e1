set -x
timeout 100 python bench.py --workload newton --steps 5 --warmup 3 --no-e2e --no-cpu > gpurun_out/nv_mask.log 2>&1; echo "mask exit $?"
B200OPT_LIB=tools/bin/libb200opt_N11.so timeout 100 python bench.py --workload newton --steps 5 --warmup 3 --no-e2e --no-cpu > gpurun_out/nv_N11b.log 2>&1; echo "N11 exit $?"
timeout 200 python bench.py --workload bfgs --steps 5 --warmup 3 --no-e2e --no-cpu > gpurun_out/bv_redux.log 2>&1; echo "bfgs exit $?"
timeout 600 python -m pytest tests/test_newton_gpu.py tests/test_bfgs_gpu.py tests/test_lm_gpu.py tests/test_jit_models_gpu.py -q -m gpu > gpurun_out/pytest_r2f.log 2>&1; echo "pytest exit $?"
tail -5 gpurun_out/pytest_r2f.log
