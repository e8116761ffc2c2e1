set -x
for V in 0 1; do
  B200OPT_LIB=tools/bin/libb200opt_B$V.so timeout 200 python bench.py --workload bfgs --steps 5 --warmup 3 --no-e2e --no-cpu > gpurun_out/bv_B$V.log 2>&1; echo "B$V exit $?"
done
timeout 600 python -m pytest tests/test_bfgs_gpu.py -q -m gpu > gpurun_out/pytest_r2c.log 2>&1; echo "pytest exit $?"
tail -5 gpurun_out/pytest_r2c.log
