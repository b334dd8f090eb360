#!/usr/bin/env python
"""Per-role cycle counters of CTA 0 for one slab conv launch (RESENC_SLAB_DEBUG=32 [+1,2,4,8,16])."""
import ctypes, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import resenc_b200 as rb
ops = rb.ops
dim = int(sys.argv[1]) if len(sys.argv) > 1 else 128
f32 = len(sys.argv) > 2 and sys.argv[2] == "f32"
x = ops.as_cl(torch.randn(2, 32, dim, dim, dim, device="cuda"))
w = torch.randn(32, 32, 3, 3, 3, device="cuda") * 0.05
for _ in range(2):
    ops._conv_forward(w, (1, 1, 1), None, x, None, out_f32=f32, want_stats=f32)
torch.cuda.synchronize()
buf = (ctypes.c_ulonglong * 16)()
rb._lib.check(rb._lib.load().rb_debug_counters(buf), "dbg")
v = [int(c) for c in buf]
t = max(1, v[11])
print(f"tiles {t}; per tile: producer total {v[1]/t:.0f} (wait empty {v[0]/t:.0f}) | mma total {v[4]/t:.0f} (wait full {v[2]/t:.0f}, "
      f"wait tmem-empty {v[3]/t:.0f}) | side total {v[7]/t:.0f} (wait tmem-full {v[5]/t:.0f}, wait staging-free {v[6]/t:.0f}) | "
      f"centre total {v[10]/t:.0f} (wait tmem-full {v[8]/t:.0f}, wait staging-full {v[9]/t:.0f}, tmem loads {v[12]/t:.0f}, "
      f"add+store {v[13]/t:.0f})")
