set -x
timeout 100 python bench.py --workload newton --steps 5 --warmup 3 --no-e2e --no-cpu > gpurun_out/nv_rcp.log 2>&1; echo "rcp exit $?"
B200OPT_LIB=tools/bin/libb200opt_mask.so timeout 100 python bench.py --workload newton --steps 5 --warmup 3 --no-e2e --no-cpu > gpurun_out/nv_mask2.log 2>&1; echo "mask exit $?"
timeout 600 python -m pytest tests/test_newton_gpu.py tests/test_jit_models_gpu.py -q -m gpu > gpurun_out/pytest_r2g.log 2>&1; echo "pytest exit $?"
tail -3 gpurun_out/pytest_r2g.log
