#!/bin/bash
# Multi-GPU verification on one box (run with `gpurun --gpus N -- bash tools/scale_check.sh N`): every exchange path of the
# row-sharded problem gives the same bits on every rank, one thread drives two devices, and the full bench line at N GPUs.
N=${1:-2}
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 tools/check_gn_large_multi.py > gpurun_out/multi_n$N.log 2>&1; echo "multi exit $?" >> gpurun_out/multi_n$N.log
timeout 300 python -m pytest tests/test_multi_device_gpu.py -m gpu -q > gpurun_out/multidev_n$N.log 2>&1; echo "pytest exit $?" >> gpurun_out/multidev_n$N.log
timeout 800 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29532 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "bench exit $?" >> gpurun_out/bench_n$N.err
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --impl reference --gpus $N --steps 2 --warmup 1 > gpurun_out/bench_ref_n$N.json 2> gpurun_out/bench_ref_n$N.err; echo "ref exit $?" >> gpurun_out/bench_ref_n$N.err
tail -2 gpurun_out/multi_n$N.log; tail -2 gpurun_out/multidev_n$N.log; tail -1 gpurun_out/bench_n$N.err; tail -1 gpurun_out/bench_ref_n$N.err
