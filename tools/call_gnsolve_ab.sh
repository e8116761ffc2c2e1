for v in tools/bin/libb200opt_base.so optimistix_b200/libb200opt.so; do B200OPT_LIB=$v python bench.py --workload gram --rows 4194304 --steps 3 --warmup 2 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$v', d['solve_kernel_ms'], round(d['ms_per_step'],3))"; done
timeout 600 python -m pytest tests/test_gram_gpu.py -q -m gpu -x 2>&1 | tail -2
