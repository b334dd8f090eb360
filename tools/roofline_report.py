#!/usr/bin/env python
"""Per-layer roofline report from a tools/conv_bench.py JSON: achieved TFLOP/s, fraction of the measured sustained
bf16 peak (MEASURED_PEAKS.json), and the milliseconds per training step each row would give back at a target
fraction - the ranked to-do list for the conv kernels.

    python tools/roofline_report.py profiles/r1_convbench.json [--target 0.7]
"""
import argparse
import json
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("json")
    ap.add_argument("--target", type=float, default=0.7, help="fraction of the sustained peak used as the attainable bar")
    args = ap.parse_args()
    try:
        peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["bf16_tflops_sustained"]
    except Exception:
        peak = 1351.6
    d = json.load(open(args.json))
    rows = []
    for r in d["rows"]:
        g = r["gflop"]
        for kind in ("fprop", "dgrad", "wgrad"):
            key = kind + "_ms"
            if key not in r:
                continue
            ms = r[key]
            tf = g / ms
            ideal = g / (peak * args.target)
            rows.append((r["count"] * max(0.0, ms - ideal), r["layer"], kind, r["count"], ms, tf, tf / peak, r["count"] * ms))
    rows.sort(reverse=True)
    tot = sum(x[-1] for x in rows)
    print(f"peak (sustained bf16) {peak:.1f} TFLOP/s, target {args.target:.0%}; listed launches sum to {tot:.2f} ms per step")
    print(f"{'layer':<24}{'pass':<7}{'cnt':>4}{'ms':>8}{'TF/s':>8}{'% peak':>8}{'ms/step':>9}{'recoverable':>13}")
    acc = 0.0
    for rec, layer, kind, cnt, ms, tf, frac, per_step in rows:
        acc += rec
        print(f"{layer:<24}{kind:<7}{cnt:>4}{ms:>8.3f}{tf:>8.0f}{100 * frac:>7.0f}%{per_step:>9.2f}{rec:>13.2f}")
    print(f"total recoverable at {args.target:.0%} of peak: {acc:.2f} ms per step")


if __name__ == "__main__":
    main()
