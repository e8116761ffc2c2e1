set -x
timeout 100 python tools/lm_ab_check.py save /tmp/ref.npz 2>&1 | tail -1
B200OPT_LSQ_LEAN=1 B200OPT_LSQ_VERBOSE=1 timeout 100 python tools/lm_ab_check.py save /tmp/lean.npz 2>&1 | tail -2; echo "lean exit $?"
python tools/lm_ab_check.py cmp /tmp/ref.npz /tmp/lean.npz > gpurun_out/lean_cmp.log 2>&1; grep -c same gpurun_out/lean_cmp.log; grep -v "^same" gpurun_out/lean_cmp.log | tail -12
B200OPT_LSQ_LEAN=1 timeout 60 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu > gpurun_out/lmv_lean.log 2>&1; echo "bench lean exit $?"
grep -o '"ms_per_step": [0-9.]*' gpurun_out/lmv_lean.log; grep -o '"mean_num_steps": [0-9.]*' gpurun_out/lmv_lean.log
