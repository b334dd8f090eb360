#!/usr/bin/env python
"""Per-role cycle counters of CTA 0 for one weights-on-M conv launch (RESENC_TC5T_DEBUG=8 [+1,2,4])."""
import ctypes, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import resenc_b200 as rb
ops = rb.ops
cin, cout, dim = (int(v) for v in (sys.argv[1:4] if len(sys.argv) > 3 else (64, 64, 64)))
x = ops.as_cl(torch.randn(2, cin, dim, dim, dim, device="cuda"))
w = torch.randn(cout, cin, 3, 3, 3, device="cuda") * 0.05
for _ in range(2):
    ops._conv_forward(w, (1, 1, 1), None, x, None, out_f32=True, want_stats=True)
torch.cuda.synchronize()
buf = (ctypes.c_ulonglong * 16)()
rb._lib.check(rb._lib.load().rb_debug_counters(buf), "dbg")
v = [int(c) for c in buf]
t = max(1, v[11])
print(f"tiles {t}; per tile: producer total {v[1]/t:.0f} (wait empty {v[0]/t:.0f}) | mma total {v[4]/t:.0f} (wait full {v[2]/t:.0f}, "
      f"wait tmem-empty {v[3]/t:.0f}) | epilogue total {v[10]/t:.0f} (wait tmem-full {v[8]/t:.0f})")
