set -x
timeout 120 python tools/coop_check.py save /tmp/ref.npz 2>&1 | tail -2
B200OPT_LSQ_COOP=8 timeout 120 python tools/coop_check.py save /tmp/c8.npz 2>&1 | tail -2; echo "c8 exit $?"
B200OPT_LSQ_COOP=4 timeout 120 python tools/coop_check.py save /tmp/c4.npz 2>&1 | tail -2; echo "c4 exit $?"
python tools/coop_check.py cmp /tmp/ref.npz /tmp/c8.npz > gpurun_out/coop_cmp8.log 2>&1; tail -25 gpurun_out/coop_cmp8.log
python tools/coop_check.py cmp /tmp/ref.npz /tmp/c4.npz > gpurun_out/coop_cmp4.log 2>&1; tail -3 gpurun_out/coop_cmp4.log
B200OPT_LSQ_COOP=8 timeout 90 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu > gpurun_out/lmv_coop8.log 2>&1; echo "bench8 exit $?"
B200OPT_LSQ_COOP=4 timeout 90 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu > gpurun_out/lmv_coop4.log 2>&1; echo "bench4 exit $?"
grep -o '"ms_per_step": [0-9.]*' gpurun_out/lmv_coop8.log gpurun_out/lmv_coop4.log
