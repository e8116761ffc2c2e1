set -x
for V in 00 11 10 01; do
  B200OPT_LIB=tools/bin/libb200opt_N$V.so timeout 120 python bench.py --workload newton --steps 5 --warmup 3 --no-e2e --no-cpu > gpurun_out/nv_N$V.log 2>&1; echo "N$V exit $?"
done
timeout 600 python -m pytest tests/test_newton_gpu.py tests/test_pytree.py tests/test_jit_models_gpu.py tests/test_generic_gpu.py tests/test_adjoint_gpu.py tests/test_lm_gpu.py -q -m gpu > gpurun_out/pytest_r2b.log 2>&1; echo "pytest exit $?"
tail -5 gpurun_out/pytest_r2b.log
