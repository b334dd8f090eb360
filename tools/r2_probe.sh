#!/bin/bash
# Round-2 measurement batch (run on the GPU box): component timing of the deep-layer tap-split kernel and the 32-channel
# full-resolution layers at widths the slab kernel does not take.  Output: gpurun_out/r2_probe_*.txt
out=gpurun_out/r2_probe_conv.txt
: > $out
run() { echo "### $*" >> $out; "$@" >> $out 2>&1; }
D="--iters 5 --skip-bwd"
run python tools/conv_bench.py $D --layer "s4 512@8,conv3,512,512,8,1" --layer "s5 512@4,conv3,512,512,4,1" --layer "s3 256@16,conv3,256,256,16,1"
for dbg in 1 2 4 5 6 7; do
  run env RESENC_TC5T_DEBUG=$dbg python tools/conv_bench.py $D --layer "s4 512@8 dbg$dbg,conv3,512,512,8,1" --layer "s5 512@4 dbg$dbg,conv3,512,512,4,1" --layer "s3 256@16 dbg$dbg,conv3,256,256,16,1"
done
run env RESENC_NO_TC5T=1 python tools/conv_bench.py $D --layer "s4 512@8 noT,conv3,512,512,8,1" --layer "s3 256@16 noT,conv3,256,256,16,1"
# 32-channel layers at W = 96 / 192: slab (96), h-major gather, per-tap gather
B1="--iters 3 --batch 1"
run python tools/conv_bench.py $B1 --layer "32@96,conv3,32,32,96,1" --layer "32@192,conv3,32,32,192,1" --layer "cat32@96,cat3,32,32,96,1" --layer "cat32@192,cat3,32,32,192,1"
run env RESENC_NO_SLAB=1 python tools/conv_bench.py $B1 --layer "32@96 noslab,conv3,32,32,96,1" --layer "cat32@96 noslab,cat3,32,32,96,1"
run env RESENC_NO_SLAB=1 RESENC_NO_TC5T_HM32=1 python tools/conv_bench.py $B1 --layer "32@96 pertap,conv3,32,32,96,1" --layer "32@192 pertap,conv3,32,32,192,1" --layer "cat32@192 pertap,cat3,32,32,192,1"
cat $out | grep -v "^per-step\|^layer " | tail -60
