#!/bin/bash
# Round-2 end-of-round evidence (one GPU):  gpurun --timeout 1700 -- 'bash tools/evidence_r2.sh 2>&1 | tail -20'
set -x
python bench.py --steps 10 --warmup 3 > gpurun_out/r2_bench_final.json 2> gpurun_out/r2_bench_final.err
python tools/conv_bench.py --iters 5 --out gpurun_out/r2_convbench.json > gpurun_out/r2_convbench.txt 2>&1
python bench.py --patch 192 --batch 1 --mode train --no-cpu-baseline --steps 5 --warmup 3 > gpurun_out/r2_bench_cfg5_192.json 2> gpurun_out/r2_bench_cfg5.err
python bench.py --patch 64 --batch 1 --mode train --no-cpu-baseline --steps 5 --warmup 3 > gpurun_out/r2_bench_cfg1_64.json 2> gpurun_out/r2_bench_cfg1.err
timeout 300 python bench.py --mode train --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/plain_r2.log 2>&1 && timeout 700 ncu --metrics gpu__time_duration.sum --clock-control none -c 8000 --csv --log-file gpurun_out/r2_launches_bench_steps2.csv python bench.py --mode train --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_launch_r2.log 2>&1
timeout 200 python tools/one_step.py 1 > gpurun_out/plain2_r2.log 2>&1 && timeout 700 ncu --set full --clock-control none --import-source on -k regex:"tc5t_gather_conv|slab_conv|norm_act_fwd|norm_act_bwd|tc5_wgrad2" -s 2 -c 16 -o gpurun_out/r2_prof_conv2 python tools/one_step.py 1 > gpurun_out/ncu_full_r2.log 2>&1
ls -la gpurun_out | tail -8
