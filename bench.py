#!/usr/bin/env python
"""Headline benchmark (BASELINE.json): ResEncM-autoconfig 128^3 multi-task training step,
batch 2 per GPU, bf16 kernels, synthetic data, 1..8 B200 data-parallel.

    python bench.py --gpus 1 --steps K --warmup W                 # this repo's CUDA path
    torchrun ... bench.py --gpus N --steps K --warmup W            # one rank per GPU (NCCL)
    python bench.py --impl reference --steps K --warmup W          # CPU baseline arm (oracle port)

A "step" = forward + multi-task loss + backward + grad-clip + AdamW update on one batch that is
already resident in HBM (`value`), and the same step with the batch copied from pinned host memory
and the loss read back every step (`e2e`).  Prints ONE JSON line on rank 0.
"""
import argparse
import contextlib
import io
import json
import os
import subprocess
import sys
import threading
import time
from types import SimpleNamespace

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

TASKS = {"sheet": {"channels": 1, "activation": "sigmoid"}, "normals": {"channels": 3, "activation": "none"}}
FWD_FLOP_PER_VOXEL = {6: 937.9e3, 5: 918.6e3}       # conv + transposed-conv MACs x 2, per patch voxel


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--patch", type=int, default=128)
    ap.add_argument("--batch", type=int, default=2, help="per-GPU batch")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-sample-patch", type=int, default=64)
    ap.add_argument("--profile-kernels", action="store_true", default=True)
    ap.add_argument("--no-graph", action="store_true", help="time eager launches instead of a captured CUDA graph")
    return ap.parse_args()


def make_mgr(patch, batch):
    return SimpleNamespace(tasks=TASKS, train_patch_size=[patch] * 3, train_batch_size=batch, in_channels=1,
                           vram_max=16.0, autoconfigure=True, model_config={})


def synthetic_batch(batch, patch, device, seed):
    g = torch.Generator(device="cpu").manual_seed(seed)
    x = torch.rand(batch, 1, patch, patch, patch, generator=g)
    sheet = (torch.rand(batch, 1, patch, patch, patch, generator=g) > 0.8).float()
    normals = torch.nn.functional.normalize(torch.randn(batch, 3, patch, patch, patch, generator=g), dim=1)
    return x, {"sheet": sheet, "normals": normals}


def losses(out, tgt, O):
    """CPU arm: the oracle's restatement of the reference losses."""
    return O.bce_dice_loss(out["sheet"], tgt["sheet"]) + O.masked_cosine_loss(out["normals"], tgt["normals"])


def gpu_losses(out, tgt, crit):
    """GPU arm: the package's own loss modules (training/losses/losses.py mirrors); nothing under oracle/ runs here."""
    return sum(crit[t](out[t], tgt[t]) for t in out)


# ------------------------------------------------------------------------------------------
# CPU arms (the oracle port of the reference's PyTorch path, timed on the host cores)
# ------------------------------------------------------------------------------------------
def cpu_step_rate(patch, batch, steps, warmup, topology_patch=None):
    """voxels/s of fwd + loss + bwd + AdamW with the oracle's functional network on the CPU.  `topology_patch`
    selects the network (the 128^3 autoconfiguration: 6 stages, 235.5 M parameters) independently of the size of
    the sample that is pushed through it."""
    topology_patch = topology_patch or patch
    from oracle import resenc_oracle as O
    import resenc_b200 as rb
    torch.set_num_threads(os.cpu_count())
    torch.manual_seed(0)
    with contextlib.redirect_stdout(io.StringIO()):
        shell = rb.NetworkFromConfig(make_mgr(topology_patch, batch))        # parameter container only (CPU)
    params = {n: p.detach().clone().requires_grad_(True) for n, p in shell.named_parameters()}
    topo = O.autoconfig([topology_patch] * 3)
    opt = torch.optim.AdamW(list(params.values()), lr=1e-3, weight_decay=1e-4)
    x, tgt = synthetic_batch(batch, patch, "cpu", 0)
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        out = O.net_forward(params, topo, x, TASKS, training=True)
        loss = losses(out, tgt, O)
        opt.zero_grad(set_to_none=True)
        loss.backward()
        torch.nn.utils.clip_grad_norm_([p for p in params.values() if p.grad is not None], 3.0)
        opt.step()
        if i >= warmup:
            times.append(time.perf_counter() - t0)
    sec = sum(times) / len(times)
    return batch * patch ** 3 / sec, sec


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    p = args.cpu_sample_patch
    rate, sec = cpu_step_rate(p, 1, args.steps, args.warmup, topology_patch=args.patch)
    cores = os.cpu_count()
    sample = (f"{p}^3 x1 crop per timed step through the same {args.patch}^3-autoconfigured network "
              f"(6 stages, 235.5 M parameters), fwd+loss+bwd+clip+AdamW, fp32 eager, all host threads")
    line = {
        "impl": "reference", "metric": "train voxels/s", "value": rate, "unit": "voxels/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"ResEncM-autoconfig {args.patch}^3 batch {args.batch}/GPU multi-task train step "
                               f"(sheet 1ch BCEDice + normals 3ch MaskedCosine, AdamW)"},
        "cpu_baseline": {"value": rate, "unit": "voxels/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": rate, "unit": "voxels/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------
# clock sampling during the timed region
# ------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [v.strip() for v in ln.split(",")]
            if len(f) < 6:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for nm, v in zip(names, f[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------
# main arm
# ------------------------------------------------------------------------------------------
def main():
    args = parse()
    if args.impl == "reference":
        run_reference_arm(args)
        return

    import torch.distributed as dist
    import resenc_b200 as rb
    import importlib
    par = importlib.import_module(rb._pkg.__name__ + ".parallel")
    loss_mod = importlib.import_module(rb._pkg.__name__ + ".losses")

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the B200 path has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # NCCL prints its version banner on stdout when NCCL_DEBUG is set in the environment: route fd 1 to stderr
        # while the communicator comes up so that stdout carries the one JSON line only
        sys.stdout.flush()
        saved_fd = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=dev)
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved_fd, 1)
            os.close(saved_fd)
    rb._lib.load()

    P, B = args.patch, args.batch
    torch.manual_seed(0)
    with contextlib.redirect_stdout(io.StringIO()):
        model = rb.NetworkFromConfig(make_mgr(P, B)).to(dev)
    n_stages = model.num_stages
    crit = loss_mod.task_losses(make_mgr(P, B).tasks)
    model.train()
    use_graph = not args.no_graph   # N > 1: the bucketed NCCL all-reduces on the side stream are captured with the step
    opt = torch.optim.AdamW(model.parameters(), lr=1e-3, weight_decay=1e-4, capturable=use_graph, fused=True)
    buckets = par.GradientBuckets(model) if world > 1 else None
    params = [p for p in model.parameters()]

    x_h, tgt_h = synthetic_batch(B, P, "cpu", 100 + rank)
    x_h = x_h.pin_memory()
    tgt_h = {k: v.pin_memory() for k, v in tgt_h.items()}
    x_d = x_h.to(dev)
    tgt_d = {k: v.to(dev) for k, v in tgt_h.items()}
    h2d = x_h.numel() * 4 + sum(v.numel() * 4 for v in tgt_h.values())

    def step(x, tgt):
        out = model(x)
        loss = gpu_losses(out, tgt, crit)
        if buckets is not None:
            buckets.zero_grad()
        else:
            opt.zero_grad(set_to_none=True)
        loss.backward()
        if buckets is not None:
            buckets.finish()
        torch.nn.utils.clip_grad_norm_([p for p in params if p.grad is not None], 3.0)
        opt.step()
        return loss

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step(x_d, tgt_d)
    barrier()

    # ---- whole-step CUDA graph: the step is ~700 of our launches plus torch glue; replaying it as one graph removes
    # the host launch latency that otherwise dominates the deep 4^3 / 8^3 layers.  Under torchrun the gradient
    # all-reduces (NCCL, side stream, overlapped with backward) are part of the captured graph ----
    graph, g_loss, graph_note = None, None, "eager"
    if use_graph:
        try:
            rb.ops.PACK_CACHE = False
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                step(x_d, tgt_d)
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                g_loss = step(x_d, tgt_d)
            torch.cuda.synchronize()
            graph.replay()
            torch.cuda.synchronize()
            rb._lib.device_error_check()
            graph_note = ("cuda-graph replay of the whole step (fwd + loss + bwd + clip + AdamW)" if world == 1 else
                          "cuda-graph replay of the whole step (fwd + loss + bwd + bucketed NCCL all-reduce + clip + AdamW)")
        except Exception as e:   # pragma: no cover - capture is an optimisation, eager is the contract
            print(f"[bench] CUDA graph capture failed, timing eager launches: {e!r}", file=sys.stderr)
            graph, g_loss = None, None
            torch.cuda.synchronize()
        finally:
            rb.ops.PACK_CACHE = True

    def timed_step(x, tgt):
        if graph is None:
            return step(x, tgt)
        if x is not x_d:
            x_d.copy_(x, non_blocking=True)
            for k in tgt_d:
                tgt_d[k].copy_(tgt[k], non_blocking=True)
        graph.replay()
        return g_loss

    # ---- timed region 1: inputs resident in HBM -------------------------------------------------
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        loss = timed_step(x_d, tgt_d)
    e1.record()
    barrier()
    ms_dev = e0.elapsed_time(e1)
    clocks = sampler.stop() if rank == 0 else None

    # ---- per-kernel CUDA-event spans: the same K steps launched eagerly (events cannot be recorded inside a
    # graph replay); also counts our launches per step ----
    rb.ops.KERNEL_TIMER.enable(args.profile_kernels)
    l0 = rb._lib.launch_count()
    p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    p0.record()
    for _ in range(args.steps):
        step(x_d, tgt_d)
    p1.record()
    barrier()
    launches = rb._lib.launch_count() - l0
    ms_eager = p0.elapsed_time(p1)
    kstat = rb.ops.KERNEL_TIMER.summary()
    rb.ops.KERNEL_TIMER.enable(False)

    # ---- timed region 2: end to end (pinned host batch in, loss out, every step) -----------------
    # (a prefetching variant - next batch copied on a side stream during the step, loss read one step late - measured
    # slower here, 38.9 vs 33.5 ms per step, so the plain in-order loop stays)
    barrier()
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    f0.record()
    last = 0.0
    for _ in range(args.steps):
        if graph is not None:
            last = float(timed_step(x_h, tgt_h).item())      # pinned host -> the graph's static input buffers
        else:
            xb = x_h.to(dev, non_blocking=True)
            tb = {k: v.to(dev, non_blocking=True) for k, v in tgt_h.items()}
            last = float(step(xb, tb).item())
    f1.record()
    barrier()
    ms_e2e = f0.elapsed_time(f1)

    t = torch.tensor([ms_dev, ms_e2e], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_dev, ms_e2e = float(t[0]), float(t[1])
    rb._lib.device_error_check()

    if rank == 0:
        vox = B * P ** 3 * args.steps * world
        value = vox / (ms_dev * 1e-3)
        e2e = vox / (ms_e2e * 1e-3)
        peaks = {}
        try:
            with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
                peaks = json.load(f)
        except Exception:
            pass
        peak_tf = peaks.get("bf16_tflops_sustained", 1400.0)
        peak_src = "MEASURED_PEAKS.json bf16_tflops_sustained" if peaks else "fallback 1.4 PFLOP/s sustained (B200_PROFILING.md)"
        conv = kstat.get("conv", {"flops": 0.0, "ms": 0.0, "launches": 0})
        ach = conv["flops"] / (conv["ms"] * 1e-3) / 1e12 if conv["ms"] > 0 else None
        step_flops = 3 * FWD_FLOP_PER_VOXEL.get(n_stages, 937.9e3) * B * P ** 3
        line = {
            "metric": "train voxels/s", "value": value, "unit": "voxels/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_dev / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": f"ResEncM-autoconfig {P}^3 batch {B}/GPU multi-task train step (sheet 1ch BCEDice + "
                                   f"normals 3ch MaskedCosine, grad-clip 3, AdamW); {n_stages} stages",
                       "parallelism": f"dp{world}", "global_batch": B * world, "launch": graph_note,
                       "eager_ms_per_step": ms_eager / args.steps,
                       "l2": "per-step working set (activations + weights, several GB) far exceeds the 126 MB L2"},
            "e2e": {"value": e2e, "unit": "voxels/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4,
                    "ms_per_step": ms_e2e / args.steps, "last_loss": last},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": {"bound": "tensor", "kernel": "conv gather implicit GEMM (fprop + dgrad; tcgen05 and mma.sync launches)",
                         "achieved": ach, "peak": peak_tf, "unit": "TFLOP/s",
                         "frac": (ach / peak_tf) if ach else None, "traffic": None, "peak_source": peak_src,
                         "launches_per_step": conv["launches"] / max(args.steps, 1),
                         "kernel_ms_per_step": conv["ms"] / max(args.steps, 1),
                         "share_of_step": conv["ms"] / ms_dev if ms_dev else None,
                         "measured_in": "eager pass of the same K steps, CUDA-event span around every launch, the wgrad side-stream "
                                        "fork switched off so that each span times one kernel running alone",
                         "whole_step_tflops": step_flops * args.steps / (ms_dev * 1e-3) / 1e12},
            "kernel_ms_per_step": {k: v["ms"] / max(args.steps, 1) for k, v in kstat.items()},
        }
        try:    # counters of the round's ncu --set full capture (never measured in this run: a pointer to the evidence)
            with open(os.path.join(ROOT, "profiles", "ncu_summary.json")) as f:
                line["roofline"]["ncu"] = json.load(f)
        except Exception:
            pass
        if world == 1 and not args.no_cpu_baseline:
            p = args.cpu_sample_patch
            rate, sec = cpu_step_rate(p, 1, 3, 1, topology_patch=P)
            line["cpu_baseline"] = {"value": rate, "unit": "voxels/s", "cores": os.cpu_count(), "kind": "port",
                                    "sample": f"{p}^3 x1 crop through the same {P}^3-autoconfigured network (fwd+loss+bwd+"
                                              f"clip+AdamW), oracle port of the reference, fp32 eager, 1 warm-up + 3 "
                                              f"timed, {sec:.2f} s/step"}
        print(json.dumps(line))
    if world > 1:
        # a communicator whose collectives were captured into a CUDA graph can block in destroy_process_group():
        # drop the graph first, and never let tear-down outlive the measurement (the line above is already printed)
        sys.stdout.flush()
        if graph is not None:
            graph.reset()
        torch.cuda.synchronize()
        dist.barrier()
        killer = threading.Timer(20.0, lambda: os._exit(0))
        killer.daemon = True
        killer.start()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
