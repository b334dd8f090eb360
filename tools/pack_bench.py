#!/usr/bin/env python
"""Weight (un)packing micro-benchmark: `rb_pack_conv_weights` (fp32 OIDHW -> bf16 [tap][co][ci] + flipped
[tap][ci][co]) and `rb_unpack_wgrad` on every conv weight shape of the 128^3 network, CUDA events, L2 flushed between
iterations.  Prints per-shape time and achieved GB/s against the algorithmic bytes (4 B read + 2 x 2 B written per
weight for the pack, 4 + 4 B for the unpack), and the count-weighted per-step total (round-1 step profile: 1.68 ms of
packing, 0.61 ms of unpacking per step).

    python tools/pack_bench.py [--iters 20]
"""
import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import resenc_b200 as rb   # noqa: E402

# (Cout, Cin, taps, count per step) of the 6-stage two-task network
SHAPES = [(32, 32, 27, 2), (64, 32, 27, 1), (64, 64, 27, 5), (128, 64, 27, 1), (128, 128, 27, 7), (256, 128, 27, 1),
          (256, 256, 27, 11), (512, 256, 27, 1), (512, 512, 27, 23), (512, 1024, 27, 2), (256, 512, 27, 2),
          (128, 256, 27, 2), (64, 128, 27, 2), (32, 64, 27, 2), (64, 32, 1, 1), (128, 64, 1, 1), (256, 128, 1, 1),
          (512, 256, 1, 1)]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=20)
    args = ap.parse_args()
    dev = torch.device("cuda")
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    tot_p = tot_u = 0.0
    print(f"{'shape':>22} {'cnt':>3} | {'pack us':>8} {'GB/s':>7} | {'unpack us':>9} {'GB/s':>7}")
    for co, ci, t, cnt in SHAPES:
        k = {27: (3, 3, 3), 1: (1, 1, 1)}[t]
        w = torch.randn(co, ci, *k, device=dev)
        dw = torch.randn(t, co, ci, device=dev)
        res = []
        for fn in (lambda: rb.ops._pack_kernel(w, True, True), lambda: rb.ops.unpack_wgrad(dw, co, ci, k)):
            for _ in range(3):
                fn()
            ms = 0.0
            for _ in range(args.iters):
                flush.zero_()
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                fn()
                b.record()
                torch.cuda.synchronize()
                ms += a.elapsed_time(b)
            res.append(ms / args.iters)
        n = co * ci * t
        print(f"{co:>5}x{ci:<5}x{t:<3}{'':>6} {cnt:>3} | {res[0] * 1e3:8.1f} {n * 8 / res[0] / 1e6:7.0f} | "
              f"{res[1] * 1e3:9.1f} {n * 8 / res[1] / 1e6:7.0f}")
        tot_p += cnt * res[0]
        tot_u += cnt * res[1]
    print(f"count-weighted per step: pack {tot_p:.3f} ms, unpack {tot_u:.3f} ms")
    rb._lib.device_error_check()


if __name__ == "__main__":
    main()
