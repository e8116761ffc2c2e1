#!/bin/bash
# A/B of the 256 x 256 normal-equation solvers (bench.py's `solve_kernel_ms`) for the libraries given as tags under tools/bin/
# (`default` = the in-tree build), then the Gram test file.
for rep in 1 2; do for v in "$@" default; do lib=tools/bin/libb200opt_$v.so; [ "$v" = default ] && lib=optimistix_b200/libb200opt.so
B200OPT_LIB=$lib python bench.py --workload gram --rows 4194304 --steps 3 --warmup 2 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$v', d['solve_kernel_ms'])"; done; done
timeout 600 python -m pytest tests/test_gram_gpu.py -q -m gpu -x 2>&1 | tail -2
