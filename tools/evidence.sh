#!/bin/bash
# End-of-round evidence capture (round 1: 640 s of box time, of which ~8 min are the two ncu passes - ncu first re-runs
# the command once without profiling, then replays every kernel; budget for it).
#   gpurun --timeout 1500 -- 'bash tools/evidence.sh 2>&1 | tail -15'
set -x
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r1x.json 2> gpurun_out/bench_r1x.err
python tools/conv_bench.py --iters 5 --out gpurun_out/convbench_r1x.json > gpurun_out/convbench_r1x.txt 2>&1
STEP_PROFILE_LAUNCHES=1 timeout 300 python tools/step_profile.py > gpurun_out/step_profile_r1x.txt 2>&1
timeout 300 python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/plain_r1x.log 2>&1 && timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/launches_r1x.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_launch_r1x.log 2>&1
timeout 200 python tools/one_step.py 1 > gpurun_out/plain2_r1x.log 2>&1 && timeout 600 ncu --set full --clock-control none --import-source on -k regex:tc5t_gather_conv -s 6 -c 2 -o gpurun_out/prof_tc5t_r1x python tools/one_step.py 1 > gpurun_out/ncu_full_r1x.log 2>&1
ls -la gpurun_out | tail -8
