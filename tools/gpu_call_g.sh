#!/bin/bash
# round-2 GPU call G (1 GPU): 16-bit data tests, gram tests (folded vs separate reduction), probe, bench line
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_lm_gpu.py tests/test_gram_gpu.py tests/test_jit_models_gpu.py -m gpu -q > gpurun_out/r2g_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2g_pytest.log
timeout 300 python tools/gram_api_probe.py --out gpurun_out/r2g_gram_probe.json > gpurun_out/r2g_gram_probe.log 2>&1; echo "probe exit $?" >> gpurun_out/r2g_gram_probe.log
timeout 900 python bench.py > gpurun_out/r2g_bench.json 2> gpurun_out/r2g_bench.err; echo "bench exit $?" >> gpurun_out/r2g_bench.err
grep -E "passed|failed|exit" gpurun_out/r2g_pytest.log | tail -3
tail -2 gpurun_out/r2g_bench.err
