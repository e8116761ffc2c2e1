#!/bin/bash
# round-2 GPU call B (2 GPUs): one thread driving two devices, the row-sharded device loop over the peer exchange, the 2-GPU bench line
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_multi_device_gpu.py tests/test_protocol_gpu.py tests/test_bfgs_gpu.py "tests/test_lm_gpu.py::test_qr_mode_solves_what_the_normal_equations_cannot" tests/test_lm_gpu.py::test_lm_exp_decay_fp64_vs_oracle tests/test_lm_gpu.py::test_lm_gauss2_fp32_vs_oracle tests/test_generic_gpu.py::test_generic_limits_are_reported -m gpu -q -s > gpurun_out/r2b_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2b_pytest.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/check_gn_large_multi.py > gpurun_out/r2b_multi.log 2>&1; echo "multi exit $?" >> gpurun_out/r2b_multi.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r2b_bench_n2.json 2> gpurun_out/r2b_bench_n2.err; echo "bench exit $?" >> gpurun_out/r2b_bench_n2.err
grep -E "passed|failed|exit" gpurun_out/r2b_pytest.log | tail -5
tail -4 gpurun_out/r2b_multi.log
tail -3 gpurun_out/r2b_bench_n2.err
