#!/bin/bash
# Repeat the headline measurement (device-resident leg only) and print per-step device and host times.
for i in $(seq 1 ${1:-12}); do python bench.py --no-e2e --no-cpu --no-secondary 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(d['ms_per_step'],3), round(d['wall_ms_timed_region'],1), d['ms_each_step'], d['host_us_each_step'], d['clocks'].get('samples'))"; done
