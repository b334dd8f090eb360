#!/usr/bin/env python
"""Per-layer device timing of every implicit-GEMM shape of the ResEncM 128^3 batch-2 network:
fprop, data gradient and weight gradient in isolation, CUDA events on the launching stream, L2 flushed
between iterations.  Prints a table and writes JSON (profiles/).  Usage: python tools/conv_bench.py [--out f.json]"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import resenc_b200 as rb   # noqa: E402

ops = rb.ops

# (name, kind, cin, cout, in_dim, stride, count per forward [2 tasks])
LAYERS = [
    ("enc_s0 32->32 @128", "conv3", 32, 32, 128, 1, 2),
    ("enc_s1 32->64 s2", "conv3", 32, 64, 128, 2, 1),
    ("enc_s1 64->64 @64", "conv3", 64, 64, 64, 1, 5),
    ("enc_s2 64->128 s2", "conv3", 64, 128, 64, 2, 1),
    ("enc_s2 128->128 @32", "conv3", 128, 128, 32, 1, 7),
    ("enc_s3 128->256 s2", "conv3", 128, 256, 32, 2, 1),
    ("enc_s3 256->256 @16", "conv3", 256, 256, 16, 1, 11),
    ("enc_s4 256->512 s2", "conv3", 256, 512, 16, 2, 1),
    ("enc_s4 512->512 @8", "conv3", 512, 512, 8, 1, 11),
    ("enc_s5 512->512 s2", "conv3", 512, 512, 8, 2, 1),
    ("enc_s5 512->512 @4", "conv3", 512, 512, 4, 1, 11),
    ("skip 32->64 @64 k1", "conv1", 32, 64, 64, 1, 1),
    ("skip 256->512 @8 k1", "conv1", 256, 512, 8, 1, 1),
    ("dec 1024->512 @8", "cat3", 512, 512, 8, 1, 2),
    ("dec 512->256 @16", "cat3", 256, 256, 16, 1, 2),
    ("dec 256->128 @32", "cat3", 128, 128, 32, 1, 2),
    ("dec 128->64 @64", "cat3", 64, 64, 64, 1, 2),
    ("dec 64->32 @128", "cat3", 32, 32, 128, 1, 2),
    ("up 512->512 4->8", "convT", 512, 512, 4, 2, 2),
    ("up 512->256 8->16", "convT", 512, 256, 8, 2, 2),
    ("up 256->128 16->32", "convT", 256, 128, 16, 2, 2),
    ("up 128->64 32->64", "convT", 128, 64, 32, 2, 2),
    ("up 64->32 64->128", "convT", 64, 32, 64, 2, 2),
]


def timeit(fn, iters, flush):
    fn()
    torch.cuda.synchronize()
    tot = 0.0
    for _ in range(iters):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        tot += a.elapsed_time(b)
    return tot / iters


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=None)
    ap.add_argument("--iters", type=int, default=3)
    ap.add_argument("--batch", type=int, default=2)
    ap.add_argument("--only", default=None)
    ap.add_argument("--prenorm", default="f32", choices=["f32", "bf16"])
    ap.add_argument("--stats", type=int, default=1)
    ap.add_argument("--skip-bwd", action="store_true")
    ap.add_argument("--layer", action="append", default=[],
                    help="extra layer 'name,kind,cin,cout,dim,stride' (kind conv3|conv1|cat3|convT); replaces the built-in table")
    args = ap.parse_args()
    global LAYERS
    if args.layer:
        LAYERS = []
        for spec in args.layer:
            name, kind, cin, cout, dim, s_ = spec.split(",")
            LAYERS.append((name, kind, int(cin), int(cout), int(dim), int(s_), 1))
    torch.manual_seed(0)
    dev = "cuda"
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)   # > 126 MB L2
    rows = []
    N = args.batch
    for name, kind, cin, cout, dim, s, cnt in LAYERS:
        if args.only and args.only not in name:
            continue
        dims = (dim,) * 3
        if kind == "convT":
            x = ops.as_cl(torch.randn(N, cin, *dims, device=dev))
            w = torch.randn(cin, cout, 2, 2, 2, device=dev) * 0.05
            od = tuple(d * 2 for d in dims)
            flops = 2.0 * N * dim ** 3 * cin * cout * 8
            dy = ops.as_cl(torch.randn(N, cout, *od, device=dev))
            ctx = type("C", (), {})()
            f_fwd = lambda: ops._ConvT3dFn.apply(w, (2, 2, 2), None, x)
            xg = x.clone().requires_grad_(True)
            wg = w.clone().requires_grad_(True)
            y = ops._ConvT3dFn.apply(wg, (2, 2, 2), None, xg)

            def f_bwd_all():
                torch.autograd.grad(y, [xg, wg], dy, retain_graph=True)
            t_f = timeit(f_fwd, args.iters, flush)
            t_b = timeit(f_bwd_all, args.iters, flush)
            rows.append(dict(layer=name, count=cnt, gflop=flops / 1e9, fprop_ms=t_f, dgrad_wgrad_ms=t_b))
            continue
        k = 3 if kind != "conv1" else 1
        x0 = ops.as_cl(torch.randn(N, cin, *dims, device=dev))
        x1 = ops.as_cl(torch.randn(N, cin, *dims, device=dev)) if kind == "cat3" else None
        ctot = cin * (2 if kind == "cat3" else 1)
        w = torch.randn(cout, ctot, k, k, k, device=dev) * 0.05
        st = (s,) * 3
        od = ops._conv_out_dims(dims, (k,) * 3, st)
        flops = 2.0 * N * od[0] * od[1] * od[2] * ctot * cout * k ** 3
        dy = ops.as_cl(torch.randn(N, cout, *od, device=dev))
        f_f = lambda: ops._conv_forward(w, st, None, x0, x1, out_f32=args.prenorm == "f32", want_stats=bool(args.stats))
        f_d = lambda: ops._conv_backward(w, st, None, x0, x1, dy, False, True, x1 is not None)
        f_w = lambda: ops._conv_backward(w, st, None, x0, x1, dy, True, False, False)
        if args.skip_bwd:
            t_f = timeit(f_f, args.iters, flush)
            t_d = t_w = float('nan')
        else:
            t_f, t_d, t_w = (timeit(f, args.iters, flush) for f in (f_f, f_d, f_w))
        rows.append(dict(layer=name, count=cnt, gflop=flops / 1e9, fprop_ms=t_f, dgrad_ms=t_d, wgrad_ms=t_w))
    tot = {"fprop": 0.0, "dgrad": 0.0, "wgrad": 0.0}
    print(f"{'layer':26s} {'cnt':>3s} {'GFLOP':>8s} | {'fprop ms':>9s} {'TF/s':>6s} | {'dgrad ms':>9s} {'TF/s':>6s} | {'wgrad ms':>9s} {'TF/s':>6s}")
    for r in rows:
        g = r["gflop"]
        if "dgrad_wgrad_ms" in r:
            print(f"{r['layer']:26s} {r['count']:3d} {g:8.1f} | {r['fprop_ms']:9.3f} {g / r['fprop_ms']:6.0f} | "
                  f"{r['dgrad_wgrad_ms']:9.3f} {2 * g / r['dgrad_wgrad_ms']:6.0f} (dgrad+wgrad)")
            tot["fprop"] += r["fprop_ms"] * r["count"]
            tot["dgrad"] += r["dgrad_wgrad_ms"] * r["count"]
        else:
            print(f"{r['layer']:26s} {r['count']:3d} {g:8.1f} | {r['fprop_ms']:9.3f} {g / r['fprop_ms']:6.0f} | "
                  f"{r['dgrad_ms']:9.3f} {g / r['dgrad_ms']:6.0f} | {r['wgrad_ms']:9.3f} {g / r['wgrad_ms']:6.0f}")
            for kk in tot:
                tot[kk] += r[kk + "_ms"] * r["count"]
    print("per-step totals (ms, count-weighted):", {k: round(v, 2) for k, v in tot.items()})
    if args.out:
        with open(args.out, "w") as f:
            json.dump({"rows": rows, "totals_ms": tot, "batch": N}, f, indent=1)


if __name__ == "__main__":
    main()
