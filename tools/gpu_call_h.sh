#!/bin/bash
# round-2 GPU call H (1 GPU): 16-bit tests again, bench line, reference arm
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_lm_gpu.py -m gpu -q > gpurun_out/r2h_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2h_pytest.log
timeout 900 python bench.py > gpurun_out/r2h_bench.json 2> gpurun_out/r2h_bench.err; echo "bench exit $?" >> gpurun_out/r2h_bench.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2h_bench_ref.json 2> gpurun_out/r2h_bench_ref.err
grep -E "passed|failed|exit" gpurun_out/r2h_pytest.log | tail -3
tail -2 gpurun_out/r2h_bench.err
