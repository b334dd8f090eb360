#!/usr/bin/env python
"""Device time of the inference tail of a decoder at config 3's geometry (4 patches of 128^3, 32 channels): the two-pass
path (rb_norm_act_fwd + rb_head_fwd) against the fused pass (rb_norm_act_head_fwd), K = 1 (sheet) and 3 (normals).
L2 flushed between iterations; bytes = what each path moves."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import resenc_b200 as rb
ops, L = rb.ops, rb._lib
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


n, c, dim = int(os.environ.get("NB", 4)), 32, 128
S = dim ** 3
E = n * c * S
y = torch.randn((n, dim, dim, dim, c), device="cuda").to(ops.PRENORM_DTYPE).permute(0, 4, 1, 2, 3)
A = torch.rand(n, c, device="cuda") + 0.5
Bv = torch.randn(n, c, device="cuda")
yb = y.element_size()
for k in (1, 3):
    hw = (torch.randn(k, c, device="cuda") * 0.2).contiguous()
    hb = torch.randn(k, device="cuda") * 0.1
    out = torch.empty((n, k, dim, dim, dim), device="cuda")
    z = ops._apply_fwd(y, None, A, Bv, False, True, 0.01)

    def two_pass():
        zz = ops._apply_fwd(y, None, A, Bv, False, True, 0.01)
        L.check(L.load().rb_head_fwd(zz.data_ptr(), hw.data_ptr(), hb.data_ptr(), out.data_ptr(), n, S, c, k, 0, L.stream_ptr()), "head")

    def head_only():
        L.check(L.load().rb_head_fwd(z.data_ptr(), hw.data_ptr(), hb.data_ptr(), out.data_ptr(), n, S, c, k, 0, L.stream_ptr()), "head")

    def fused():
        L.check(L.load().rb_norm_act_head_fwd(y.data_ptr(), ops._YMODE[y.dtype], None, A.data_ptr(), Bv.data_ptr(), hw.data_ptr(),
                                             hb.data_ptr(), out.data_ptr(), n, S, c, k, 1, 0.01, 0, L.stream_ptr()), "fused")

    ref = out.clone(); two_pass(); ref.copy_(out); fused()
    err = float((out - ref).norm() / ref.norm())
    for name, fn, nbytes in (("norm_act_fwd + head_fwd", two_pass, E * (yb + 2 + 2) + 4 * n * k * S),
                             ("head_fwd alone", head_only, E * 2 + 4 * n * k * S),
                             ("norm_act_head_fwd (fused)", fused, E * yb + 4 * n * k * S)):
        ms = timeit(fn)
        print(f"K={k} {c}ch @{dim}^3 x{n}  {name:28s} {ms * 1e3:8.1f} us  {nbytes / ms / 1e6:7.0f} GB/s")
    print(f"K={k} fused vs two-pass rel-L2 {err:.2e}")
