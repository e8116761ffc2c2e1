#!/bin/bash
# round-2 GPU call C (1 GPU): protocol / QR-mode / refactored generic path tests, parity report with the QR mode
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_protocol_gpu.py tests/test_lm_gpu.py tests/test_generic_gpu.py tests/test_adjoint_gpu.py tests/test_pytree.py tests/test_jit_models_gpu.py -m gpu -q -s > gpurun_out/r2c_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2c_pytest.log
timeout 600 python tools/lm_parity_report.py --out gpurun_out/r2c_lm_parity.json > gpurun_out/r2c_lm_parity.log 2>&1; echo "parity exit $?" >> gpurun_out/r2c_lm_parity.log
grep -E "passed|failed|error|exit" gpurun_out/r2c_pytest.log | tail -15
tail -2 gpurun_out/r2c_lm_parity.log
