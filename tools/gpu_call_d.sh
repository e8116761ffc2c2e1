#!/bin/bash
# round-2 GPU call D (1 GPU): re-run of the tests that failed, gram API probe + ncu launch list of it
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_protocol_gpu.py tests/test_bfgs_gpu.py tests/test_lm_gpu.py tests/test_generic_gpu.py -m gpu -q -s > gpurun_out/r2d_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2d_pytest.log
timeout 300 python tools/gram_api_probe.py --out gpurun_out/r2d_gram_probe.json > gpurun_out/r2d_gram_probe.log 2>&1; echo "probe exit $?" >> gpurun_out/r2d_gram_probe.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2d_launches_gram_api.csv python tools/gram_api_probe.py --rows 8388608 --iters 4 --reps 1 > gpurun_out/r2d_ncu_probe.log 2>&1; echo "ncu exit $?" >> gpurun_out/r2d_ncu_probe.log
grep -E "passed|failed|exit" gpurun_out/r2d_pytest.log | tail -5
cat gpurun_out/r2d_gram_probe.log | tail -40
