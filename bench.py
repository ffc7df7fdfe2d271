#!/usr/bin/env python
"""bench.py -- restored images/s of the UniRestore hot path (DiffUIE.forward) on B200.

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...   # the reference's algorithm on the host CPU (oracle port)

A "step" is one pass of the hot path over one batch: B=8 synthetic 512x512 images per GPU through
VAE-encode(+CFRM) -> 20 x {Controller, ControlledUNet(+SC-Tuner), DDIM} -> VAE-decode(+TFA), bf16 activations,
random-init weights of the sd-turbo / UniRestore architecture (BASELINE.json configs[1]).  Batches shard over
images (weak scaling, 8 images per GPU); the only collective is the NCCL all-gather of the decoded images.

One JSON line is printed by rank 0 (contract: see the task statement / DESIGN.md "Measurement").
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "restored images/sec @512x512, 20 DDIM steps"
UNIT = "images/s"
CFG = (dict(type="CFRM"), dict(type="scedit", num_inference_steps=20),
       dict(type="TFA", prompt_len=1, task=["ir", "cls", "seg"]))


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=8, help="images per GPU")
    ap.add_argument("--size", type=int, default=512)
    ap.add_argument("--ddim-steps", type=int, default=20)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


def cheap_init_(model, seed=0):
    """Random-init weights without the reference's zero-initialised sub-graphs (SURVEY.md 8c hazard 2)."""
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for n, p in model.named_parameters():
            if p.dim() <= 1 or "beta" in n or "gamma" in n or "task_prompts" in n:
                mean = 1.0 if (p.dim() == 1 and n.endswith("weight")) else 0.0
                p.copy_(torch.randn(p.shape, generator=g) * 0.05 + mean)
            elif float(p.abs().max()) == 0.0:
                p.copy_(torch.randn(p.shape, generator=g) * 0.02)
    return model


# ------------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons of one GPU every 200 ms while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self._stop = index, [], threading.Event()
        self.t = threading.Thread(target=self._run, daemon=True)

    def _run(self):
        while not self._stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                self.rows.append([c.strip() for c in out.strip().split(",")])
            except Exception:
                pass
            self._stop.wait(0.2)

    def __enter__(self):
        self.t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        self.t.join(timeout=6)

    def summary(self):
        sm = [float(r[0]) for r in self.rows if len(r) >= 6 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 6 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 6 for i in range(4) if r[2 + i].lower() == "active"})
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


# ------------------------------------------------------------------------------------------------ CPU baseline
def cpu_reference_times(threads=None, size=512):
    """Times the oracle port of the reference's PyTorch path on the host cores for ONE image:
    (seconds for VAE-encode+CFRM and VAE-decode+TFA, seconds for one Controller+UNet(+SC-Tuner)+DDIM step)."""
    from oracle import unirestore as O
    torch.set_num_threads(threads or os.cpu_count())
    m = cheap_init_(O.DiffUIE(*CFG)).eval().requires_grad_(False)
    img = torch.rand(1, 3, size, size, generator=torch.Generator().manual_seed(42))
    with torch.no_grad():
        t0 = time.perf_counter()
        z0, mids = m.ae.encode(img, enable_fr=True)
        t_enc = time.perf_counter() - t0
        zt, _, _ = m.diffuse(z0, torch.full((1,), 999, dtype=torch.long))
        ts = m.scheduler.timesteps[:1]
        t0 = time.perf_counter()
        eps = m.base_model(zt, m.controller(z0, ts), ts)
        zt = m.scheduler.step(eps, ts[0], zt).prev_sample
        t_step = time.perf_counter() - t0
        t0 = time.perf_counter()
        m.ae.decode(zt, mids, "ir")
        t_dec = time.perf_counter() - t0
    return t_enc + t_dec, t_step, m


def run_reference(a):
    """--impl reference: the reference's own algorithm (oracle port, fp32 PyTorch) on the host CPU.
    Each step is a bounded sample of the workload: one Controller+UNet+DDIM step for one 512x512 image; the
    once-per-image part (VAE/CFRM/TFA) is timed once during warm-up.  images/s = 1 / (t_once + 20 * t_step)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count()
    t_once, t_step, m = cpu_reference_times(cores, a.size)
    from oracle import unirestore as O  # noqa: F401
    img = torch.rand(1, 3, a.size, a.size, generator=torch.Generator().manual_seed(42))
    with torch.no_grad():
        z0, _ = m.ae.encode(img, enable_fr=True) if a.warmup > 1 else (torch.randn(1, 4, a.size // 8, a.size // 8), None)
        zt = torch.randn_like(z0)
        ts = m.scheduler.timesteps[:1]
        times = []
        for i in range(max(0, a.warmup - 1) + a.steps):
            t0 = time.perf_counter()
            eps = m.base_model(zt, m.controller(z0, ts), ts)
            zt2 = m.scheduler.step(eps, ts[0], zt).prev_sample
            dt = time.perf_counter() - t0
            if i >= max(0, a.warmup - 1):
                times.append(dt)
            del zt2
    t_step = sum(times) / len(times)
    per_img = t_once + a.ddim_steps * t_step
    value = 1.0 / per_img
    sample = ("1 image %dx%d, fp32 CPU oracle port of the reference path: VAE-encode+CFRM and VAE-decode+TFA timed once "
              "(%.2f s), one Controller+UNet+SC-Tuner+DDIM step per bench step (%.2f s); images/s = 1/(once + %d*step)"
              % (a.size, a.size, t_once, t_step, a.ddim_steps))
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps,
        "warmup": a.warmup, "ms_per_step": 1e3 * per_img, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "batch=%d/GPU %dx%d, %d DDIM steps, bf16, PIR prompt ('ir'), random-init weights"
                               % (a.batch, a.size, a.size, a.ddim_steps),
                   "reference_arm": "host CPU, fp32, bounded sample of 1 image (see cpu_baseline.sample)"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def claim_stdout():
    """stdout must carry exactly ONE JSON line, but libraries write to file descriptor 1 behind Python's back (NCCL
    prints "NCCL version ..." there when the first communicator is created).  Point fd 1 at stderr for the rest of the
    process and return a writer for the real stdout."""
    sys.stdout.flush()
    real = os.dup(1)
    os.dup2(2, 1)

    def emit(text):
        os.write(real, (text.rstrip("\n") + "\n").encode())
    return emit


# ------------------------------------------------------------------------------------------------ ours
def kernel_roofline(dev):
    """The dominant kernel (tcgen05 implicit-GEMM conv, UNet 320->320 3x3 @64x64, B=8: 60.4 GFLOP per launch)
    timed alone with CUDA events on the launching stream; operands rotate over 8 buffer sets (> L2).
    ``traffic`` = dram__bytes_read.sum + dram__bytes_write.sum of that launch from the committed ncu --set full
    capture (profiles/ncu_r1f_gemm_unet320.txt: 22.89 MB + 0.14 MB; the output stays in L2 within the capture)."""
    from unirestore_b200 import ops
    B, H, W, C = 8, 64, 64, 320
    sets = []
    for i in range(8):
        x = torch.randn(B, H, W, C, device=dev).to(torch.bfloat16)
        sets.append((x, torch.empty_like(x)))
    w = (torch.randn(C, 9 * C, device=dev) * (9 * C) ** -0.5).to(torch.bfloat16)
    bias = torch.randn(C, device=dev)
    for x, o in sets[:3]:
        ops.conv_gemm(x, w, C, taps=ops.TAPS_3x3, bias=bias, out=o)
    torch.cuda.synchronize()
    reps = 40
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for i in range(reps):
        x, o = sets[i % 8]
        ops.conv_gemm(x, w, C, taps=ops.TAPS_3x3, bias=bias, out=o)
    e.record()
    torch.cuda.synchronize()
    t = s.elapsed_time(e) * 1e-3 / reps
    flops = 2.0 * B * H * W * C * C * 9
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("bf16_tflops", 1590.0))
    ach = flops / t / 1e12
    return {"bound": "tensor", "achieved": ach, "peak": peak, "unit": "TFLOP/s", "frac": ach / peak,
            "traffic": 23.03e6, "kernel": "ur::conv_gemm_persistent_kernel<160, PAIR> conv3x3 320->320 @64x64 B=8",
            "launch_us": t * 1e6, "peak_source": "MEASURED_PEAKS.json bf16_tflops (burst)" if peaks else "fallback"}


def run_ours(a):
    emit = claim_stdout()
    import torch.distributed as dist
    from unirestore_b200 import _cabi
    from unirestore_b200.diffuie import DiffUIE
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the product path has no CPU fallback)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    cfg = (CFG[0], dict(CFG[1], num_inference_steps=a.ddim_steps), CFG[2])
    model = cheap_init_(DiffUIE(*cfg)).eval().requires_grad_(False).to(dev)
    model.use_cuda_graph = True
    B = a.batch
    g = torch.Generator().manual_seed(42 + rank)
    host_img = torch.rand(B, 3, a.size, a.size, generator=g).pin_memory()
    host_out = torch.empty(B, 3, a.size, a.size).pin_memory()
    img = host_img.to(dev)
    gathered = [torch.empty(B, 3, a.size, a.size, device=dev) for _ in range(world)] if world > 1 else None
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def step(x):
        y = model(x, "ir")
        if world > 1:
            dist.all_gather(gathered, y)            # the only collective on the path
        return y

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(a.warmup, 3)):
        step(img)
    barrier()
    # ---------------- device-resident throughput: K steps, CUDA events per step, L2 flushed in between
    launches0 = _cabi.launch_count
    evs = []
    with ClockSampler(local) as clk:
        barrier()
        for _ in range(a.steps):
            flush.zero_()
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            step(img)
            e.record()
            evs.append((s, e))
        barrier()
    total_ms = sum(s.elapsed_time(e) for s, e in evs)
    t = torch.tensor([total_ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms = float(t.item())
    value = world * B * a.steps / (total_ms * 1e-3)
    # kernels launched inside the timed region: the graph replays exactly the launches recorded at capture
    graph_launches = getattr(next(iter(model._graphs.values())), "n_launches", None)
    gpu_launches = (graph_launches or 0) * a.steps + (_cabi.launch_count - launches0)
    # ---------------- end to end: pinned host images in, restored images back to pinned host memory
    barrier()
    t0 = time.perf_counter()
    for _ in range(a.steps):
        x = host_img.to(dev, non_blocking=True)
        y = step(x)
        host_out.copy_(y, non_blocking=True)
        torch.cuda.synchronize()
    barrier()
    e2e_s = time.perf_counter() - t0
    te = torch.tensor([e2e_s], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = world * B * a.steps / float(te.item())
    nbytes = host_img.numel() * 4
    if rank == 0:
        roof = kernel_roofline(dev)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": max(a.warmup, 3),
            "ms_per_step": total_ms / a.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16", "data": "synthetic",
            "config": {"workload": "batch=%d/GPU %dx%d, %d DDIM steps, bf16, PIR prompt ('ir'), random-init weights"
                                   % (B, a.size, a.size, a.ddim_steps),
                       "l2": "256 MiB flush between timed steps; per-step working set >> 126 MB L2",
                       "parallelism": "dp%d (batch-sharded, final NCCL all-gather)" % world},
            "clocks": clk.summary(),
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": nbytes, "d2h_bytes_per_step": nbytes},
            "gpu_launches": gpu_launches,
            "roofline": roof,
        }
        if world == 1 and not a.no_cpu_baseline:
            try:
                cores = os.cpu_count()
                t_once, t_step, _ = cpu_reference_times(cores, a.size)
                v = 1.0 / (t_once + a.ddim_steps * t_step)
                line["cpu_baseline"] = {
                    "value": v, "unit": UNIT, "cores": cores, "kind": "port",
                    "sample": "1 image %dx%d fp32 on the host CPU (oracle port): VAE/CFRM/TFA once %.2f s + one "
                              "Controller+UNet step %.2f s, extrapolated to %d steps" % (a.size, a.size, t_once, t_step,
                                                                                         a.ddim_steps)}
            except Exception as ex:   # the oracle is test infrastructure; never let it break the GPU numbers
                line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": os.cpu_count(), "kind": "port",
                                        "sample": "failed: %r" % (ex,)}
        emit(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)
