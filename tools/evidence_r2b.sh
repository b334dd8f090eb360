#!/bin/bash
# Round-2 (second session) end-of-round evidence, one GPU:  gpurun --timeout 1200 -- 'bash tools/evidence_r2b.sh 2>&1 | tail -20'
set -x
timeout 200 python -m pytest tests/test_optim.py -q -m gpu > gpurun_out/r2b_test_optim.log 2>&1; tail -2 gpurun_out/r2b_test_optim.log
python bench.py --steps 10 --warmup 3 > gpurun_out/r2b_bench_n1.json 2> gpurun_out/r2b_bench_n1.err
timeout 300 python bench.py --mode train --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r2b_plain.log 2>&1 && timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none -c 8000 --csv --log-file gpurun_out/r2b_launches_bench_steps2.csv python bench.py --mode train --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r2b_ncu_launch.log 2>&1
timeout 200 python tools/one_step.py 1 > gpurun_out/r2b_plain2.log 2>&1 && timeout 500 ncu --set full --clock-control none --import-source on -k regex:"norm_act_bwd_u2|plane_reduce_u2|norm_act_fwd_un|adamw_clip|grad_sumsq" -s 2 -c 14 -o gpurun_out/r2b_prof_hbm python tools/one_step.py 1 > gpurun_out/r2b_ncu_full.log 2>&1
ls -la gpurun_out | tail -6
