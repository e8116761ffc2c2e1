#!/bin/bash
# A/B of the packed LM kernels on the GPU box: two problems per warp (default build) against one problem per warp
# (tools/bin/libb200opt_p2.so, built with -DLSQ_P2X=0): bitwise comparison of the results, then the headline bench twice each.
mkdir -p gpurun_out
timeout 200 python tools/lm_ab_check.py save gpurun_out/ab_p2x.npz > gpurun_out/lm_ab.log 2>&1
B200OPT_LIB=tools/bin/libb200opt_p2.so timeout 200 python tools/lm_ab_check.py save gpurun_out/ab_p2.npz >> gpurun_out/lm_ab.log 2>&1
python tools/lm_ab_check.py cmp gpurun_out/ab_p2.npz gpurun_out/ab_p2x.npz >> gpurun_out/lm_ab.log 2>&1
tail -12 gpurun_out/lm_ab.log
show='import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(sys.argv[1], round(d["value"]/1e6,2), "M/s", round(d["ms_per_step"],3), "ms", d["checks"]["mean_num_steps"], d["checks"]["successful_frac"])'
for i in 1 2; do
  timeout 200 python bench.py --no-secondary --no-e2e --no-cpu 2>/dev/null | python -c "$show" p2x | tee -a gpurun_out/lm_ab.log
  B200OPT_LIB=tools/bin/libb200opt_p2.so timeout 200 python bench.py --no-secondary --no-e2e --no-cpu 2>/dev/null | python -c "$show" p2 | tee -a gpurun_out/lm_ab.log
done
for b in 131072; do
  timeout 200 python bench.py --no-secondary --no-e2e --no-cpu --batch $b 2>/dev/null | python -c "$show" p2x_131k | tee -a gpurun_out/lm_ab.log
  B200OPT_LIB=tools/bin/libb200opt_p2.so timeout 200 python bench.py --no-secondary --no-e2e --no-cpu --batch $b 2>/dev/null | python -c "$show" p2_131k | tee -a gpurun_out/lm_ab.log
done
rm -f gpurun_out/ab_p2.npz gpurun_out/ab_p2x.npz
