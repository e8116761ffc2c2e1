#!/bin/bash
# round-2 GPU call F (8 GPUs): row-sharded device loop over 8 peers, the full 8-GPU bench line (weak headline + strong secondary + gram)
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r2f_topo.txt 2>&1
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 tools/check_gn_large_multi.py > gpurun_out/r2f_multi.log 2>&1; echo "multi exit $?" >> gpurun_out/r2f_multi.log
timeout 800 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/r2f_bench_n8.json 2> gpurun_out/r2f_bench_n8.err; echo "bench exit $?" >> gpurun_out/r2f_bench_n8.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29523 bench.py --gpus 8 --steps 5 --warmup 3 --no-secondary --no-numa > gpurun_out/r2f_bench_n8_nonuma.json 2> gpurun_out/r2f_bench_n8_nonuma.err; echo "bench exit $?" >> gpurun_out/r2f_bench_n8_nonuma.err
tail -3 gpurun_out/r2f_multi.log; tail -2 gpurun_out/r2f_bench_n8.err
