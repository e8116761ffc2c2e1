set -x
for F in 0 7 1 3 15; do
  B200OPT_LIB=tools/bin/libb200opt_F$F.so timeout 120 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu > gpurun_out/lmv_F$F.log 2>&1; echo "F$F exit $?"
done
for F in 7 15; do
  B200OPT_LIB=tools/bin/libb200opt_F$F.so timeout 200 python bench.py --steps 3 --warmup 3 --no-e2e > gpurun_out/lmv_cpu_F$F.log 2>&1; echo "F$F cpu exit $?"
done
timeout 300 python -m pytest tests/test_lm_gpu.py -x -q -m gpu > gpurun_out/pytest_lm_r2a.log 2>&1; echo "pytest exit $?"
tail -3 gpurun_out/pytest_lm_r2a.log
