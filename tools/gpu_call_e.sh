#!/bin/bash
# round-2 GPU call E (1 GPU): full GPU suite, bench line, gram probe, ncu captures of the changed kernels
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r2e_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2e_pytest.log
timeout 900 python bench.py > gpurun_out/r2e_bench.json 2> gpurun_out/r2e_bench.err; echo "bench exit $?" >> gpurun_out/r2e_bench.err
timeout 300 python tools/gram_api_probe.py --out gpurun_out/r2e_gram_probe.json > gpurun_out/r2e_gram_probe.log 2>&1; echo "probe exit $?" >> gpurun_out/r2e_gram_probe.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2e_launches_gram_api.csv python tools/gram_api_probe.py --rows 8388608 --iters 4 --reps 1 > gpurun_out/r2e_ncu_probe.log 2>&1
# --set full captures (each command has already exited 0 without ncu above)
timeout 300 ncu --set full --clock-control none --import-source on -k regex:lsq_solve_kernel -s 3 -c 1 -f -o gpurun_out/lsq_r2e python bench.py --batch 131072 --steps 2 --warmup 3 --no-e2e --no-cpu --no-secondary > gpurun_out/r2e_ncu_lsq.log 2>&1; echo "ncu lsq $?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gn_step_kernel -s 2 -c 1 -f -o gpurun_out/gnstep_r2e python tools/gram_api_probe.py --rows 4194304 --iters 4 --reps 1 > gpurun_out/r2e_ncu_step.log 2>&1; echo "ncu step $?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gram_kernel -s 14 -c 1 -f -o gpurun_out/gramfold_r2e python tools/gram_api_probe.py --rows 8388608 --iters 4 --reps 1 > gpurun_out/r2e_ncu_gram.log 2>&1; echo "ncu gram $?"
grep -E "passed|failed|exit" gpurun_out/r2e_pytest.log | tail -3
tail -2 gpurun_out/r2e_bench.err
