#!/usr/bin/env python
"""In-graph timing of ONE conv call (memset + tap-split tcgen05 conv + finish) for the deep layers: the call is captured
20 times back to back into a CUDA graph, so host launch overhead is out of the picture (conv_bench.py times eager calls).
RESENC_TC5T_DEBUG=1|2|4 (skip MMAs / TMA loads / the epilogue body) and =8 (per-role cycle counters of CTA 0) apply."""
import ctypes, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import resenc_b200 as rb
ops = rb.ops
shapes = [(512, 512, 4), (512, 512, 8), (256, 256, 16)] if len(sys.argv) < 4 else [tuple(int(v) for v in sys.argv[1:4])]
for cin, cout, dim in shapes:
    x = ops.as_cl(torch.randn(2, cin, dim, dim, dim, device="cuda"))
    w = torch.randn(cout, cin, 3, 3, 3, device="cuda") * 0.05
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    f = lambda: ops._conv_forward(w, (1, 1, 1), None, x, None, out_f32=True, want_stats=True)
    for _ in range(3):
        f()
    torch.cuda.synchronize()
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        f()
    torch.cuda.current_stream().wait_stream(s)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(20):
            y, st = f()
    for _ in range(2):
        g.replay()
    torch.cuda.synchronize()
    tot = 0.0
    for _ in range(5):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); g.replay(); b.record(); torch.cuda.synchronize()
        tot += a.elapsed_time(b)
    us = tot / 5 / 20 * 1e3
    gf = 2.0 * 2 * dim ** 3 * cin * cout * 27 / 1e9
    line = f"{cin}->{cout} @{dim}^3 x2: {us:7.1f} us per call in-graph ({gf / us * 1e-3:6.0f} TFLOP/s) dbg={os.environ.get('RESENC_TC5T_DEBUG', '0')}"
    if int(os.environ.get("RESENC_TC5T_DEBUG", "0")) & 8:
        buf = (ctypes.c_ulonglong * 16)()
        rb._lib.check(rb._lib.load().rb_debug_counters(buf), "dbg")
        v = [int(c) for c in buf]
        t = max(1, v[11])
        line += (f" | CTA0: tiles {t}; per tile cycles: producer {v[1]/t:.0f} (wait empty {v[0]/t:.0f}) | mma {v[4]/t:.0f} (wait full {v[2]/t:.0f}, "
                 f"wait tmem-empty {v[3]/t:.0f}) | epilogue {v[10]/t:.0f} (wait tmem-full {v[8]/t:.0f})")
    print(line)
