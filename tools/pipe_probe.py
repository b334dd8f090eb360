#!/usr/bin/env python
"""Pipeline cycle counters of CTA 0 for one tcgen05 conv launch (RESENC_TC5_DEBUG=8 [+1..4])."""
import ctypes, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import resenc_b200 as rb
ops = rb.ops
cin, cout, dim = (int(v) for v in (sys.argv[1:4] if len(sys.argv) > 3 else (32, 32, 128)))
x = ops.as_cl(torch.randn(2, cin, dim, dim, dim, device="cuda"))
w = torch.randn(cout, cin, 3, 3, 3, device="cuda") * 0.05
for _ in range(2):
    ops._conv_forward(w, (1, 1, 1), "tc5", x, None, out_f32=True, want_stats=True)
buf = (ctypes.c_ulonglong * 16)()
rb._lib.check(rb._lib.load().rb_debug_counters(buf), "dbg")
names = ["prod_wait_empty", "prod_total", "mma_wait_full", "mma_wait_tmem_empty", "mma_total", "epi_wait_full", "epi_total", "tiles"]
v = list(buf)[:8]
print({n: int(c) for n, c in zip(names, v)})
t = max(1, v[7])
print(f"per tile: prod total {v[1]/t:.0f} (wait {v[0]/t:.0f}), mma total {v[4]/t:.0f} (wait full {v[2]/t:.0f}, wait tmem {v[3]/t:.0f}), "
      f"epi total {v[6]/t:.0f} (wait {v[5]/t:.0f})")
