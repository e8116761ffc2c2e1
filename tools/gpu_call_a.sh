#!/bin/bash
# round-2 GPU call A (1 GPU): tests, LM parity report, full bench line
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm --format=csv > gpurun_out/r2a_smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -q -x --deselect tests/test_gram_gpu.py > gpurun_out/r2a_pytest_rest.log 2>&1; echo "pytest rest exit $?" >> gpurun_out/r2a_pytest_rest.log
timeout 900 python -m pytest tests/test_gram_gpu.py -m gpu -q -s > gpurun_out/r2a_pytest_gram.log 2>&1; echo "pytest gram exit $?" >> gpurun_out/r2a_pytest_gram.log
timeout 600 python tools/lm_parity_report.py --out gpurun_out/r2a_lm_parity.json > gpurun_out/r2a_lm_parity.log 2>&1; echo "parity exit $?" >> gpurun_out/r2a_lm_parity.log
timeout 900 python bench.py > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err; echo "bench exit $?" >> gpurun_out/r2a_bench.err
tail -3 gpurun_out/r2a_pytest_rest.log gpurun_out/r2a_pytest_gram.log gpurun_out/r2a_lm_parity.log gpurun_out/r2a_bench.err
