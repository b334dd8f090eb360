#!/usr/bin/env python
"""Device time and effective bandwidth (bytes this design moves / time) of the InstanceNorm apply / reduce passes at the
two largest geometries of config 2.  RESENC_NORM_VARIANT=0|1 selects the kernel generation, RESENC_PRENORM=f32|f16 the
pre-norm element type."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import resenc_b200 as rb
ops = rb.ops
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")


def timeit(fn, iters=5):
    fn(); torch.cuda.synchronize()
    tot = 0.0
    for _ in range(iters):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        tot += a.elapsed_time(b)
    return tot / iters


print(f"variant {os.environ.get('RESENC_NORM_VARIANT', '1')}, pre-norm {ops.PRENORM_DTYPE}")
for n, c, dim in ((2, 32, 128), (2, 64, 64), (2, 128, 32)):
    E = n * c * dim ** 3
    yb = 4 if ops.PRENORM_DTYPE == torch.float32 else 2
    y = torch.randn((n, dim, dim, dim, c), device="cuda").to(ops.PRENORM_DTYPE).permute(0, 4, 1, 2, 3)
    res = ops.as_cl(torch.randn(n, c, dim, dim, dim, device="cuda"))
    dz = ops.as_cl(torch.randn(n, c, dim, dim, dim, device="cuda"))
    A = torch.rand(n, c, device="cuda") + 0.5
    Bv = torch.randn(n, c, device="cuda")
    k = [torch.randn(n, c, device="cuda") for _ in range(3)]
    rows = []
    t = timeit(lambda: ops._apply_fwd(y, None, A, Bv, False, True, 0.01)); rows.append(("apply fwd", t, E * (yb + 2)))
    t = timeit(lambda: ops._apply_fwd(y, res, A, Bv, False, True, 0.01)); rows.append(("apply fwd + res", t, E * (yb + 4)))
    z = ops._apply_fwd(y, res, A, Bv, False, True, 0.01)
    t = timeit(lambda: ops._apply_bwd(dz, None, y, k[0], k[1], k[2], False, True, 0.01, False, (A, Bv))); rows.append(("apply bwd (sign from y)", t, E * (2 + yb + 2)))
    t = timeit(lambda: ops._apply_bwd(dz, z, y, k[0], k[1], k[2], False, True, 0.01, True)); rows.append(("apply bwd + z + dres", t, E * (2 + yb + 2 + 2 + 2)))
    t = timeit(lambda: ops._plane_reduce(1, y, dz, None, False, 0.01, (A, Bv))); rows.append(("reduce (sign from y)", t, E * (2 + yb)))
    t = timeit(lambda: ops._plane_reduce(1, y, dz, z, False, 0.01)); rows.append(("reduce + z", t, E * (4 + yb)))
    for name, ms, nbytes in rows:
        print(f"  {c:4d}ch @{dim}^3 x{n}  {name:26s} {ms * 1e3:8.1f} us  {nbytes / ms / 1e6:7.0f} GB/s")
