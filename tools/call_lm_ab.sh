#!/bin/bash
# LM kernel A/B on one B200: bitwise check of the default build against the round's baseline, then CUDA-event timings per variant.
R=${1:-r2q}; shift
mkdir -p gpurun_out
B200OPT_LIB=tools/bin/libb200opt_base.so python tools/lm_ab_check.py save /tmp/ab_base.npz > gpurun_out/lm_ab_$R.log 2>&1
python tools/lm_ab_check.py save /tmp/ab_new.npz >> gpurun_out/lm_ab_$R.log 2>&1
python tools/lm_ab_check.py cmp /tmp/ab_base.npz /tmp/ab_new.npz >> gpurun_out/lm_ab_$R.log 2>&1; echo "bitwise cmp exit $?" | tee -a gpurun_out/lm_ab_$R.log
for rep in 1 2; do
for v in base "$@" default; do
  lib=tools/bin/libb200opt_lsq_gauss2_$v.so; [ "$v" = base ] && lib=tools/bin/libb200opt_base.so; [ "$v" = default ] && lib=optimistix_b200/libb200opt.so
  B200OPT_LIB=$lib timeout 300 python bench.py --steps 8 --warmup 3 --no-e2e --no-cpu --no-secondary 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$v'.ljust(8), 'ms', round(d['ms_per_step'],3), d.get('checks',{}).get('mean_num_steps'), d['clocks']['sm_mhz'])" | tee -a gpurun_out/lm_ab_$R.log
done; done
