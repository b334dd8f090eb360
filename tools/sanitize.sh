#!/bin/bash
# compute-sanitizer passes over the kernel parity tests (race detection / sanitizers).  Run on a GPU box:
#   gpurun --timeout 1500 -- 'bash tools/sanitize.sh > gpurun_out/sanitize.log 2>&1'
# memcheck on the small-shape kernel tests, racecheck + synccheck on the shared-memory / mbarrier heavy ones.  The
# tests are the tiny-shape subsets: every kernel family is launched at least once, a sanitised launch is ~50x slower.
set -x
SEL='test_tc5_conv_fwd_bwd and (n2_32to32_16x16x16 or n1_64to64_16x16x16 or n2_256to256_4x4x4) or test_tc5_fused_statistics or test_instance_norm_act or test_split_precision_layout_kernels or test_head or test_avg_pool or test_weight_pack_unpack_kernels'
for tool in memcheck racecheck synccheck; do
  timeout 1200 compute-sanitizer --tool $tool --error-exitcode 3 --print-limit 20 \
      python -m pytest tests/test_gpu_tc5.py tests/test_gpu_ops.py tests/test_gpu_blend.py -x -q -m gpu -k "$SEL" \
      > gpurun_out/sanitize_$tool.log 2>&1
  echo "$tool exit code $?"
  tail -5 gpurun_out/sanitize_$tool.log
done
