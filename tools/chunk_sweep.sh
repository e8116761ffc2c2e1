for c in 16384 32768 65536 131072 262144; do python bench.py --workload newton --no-cpu --chunk $c --steps 5 --warmup 3 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print($c, 'device', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['ms_per_step'],3), d['e2e'].get('pageable_inputs',{}).get('ms_per_step'))"; done
for c in 16384 32768 65536; do python bench.py --workload bfgs --no-cpu --chunk $c --steps 3 --warmup 2 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('bfgs', $c, 'device', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['ms_per_step'],3))"; done
