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

METRIC = "restored images/sec @512x512, 20 DDIM steps"        # BASELINE.json metric (configs[1]); other configs: metric_name()
UNIT = "images/s"
CFG = (dict(type="CFRM"), dict(type="scedit", num_inference_steps=20),
       dict(type="TFA", prompt_len=1, task=["ir", "cls", "seg"]))


# BASELINE.json configs[0..4] (c2 is the headline the metric is quoted on and the default).  ``global_batch`` /
# ``gpus`` describe the configuration as BASELINE states it; a run with fewer GPUs than that measures the PER-GPU
# SLICE of it (global_batch / gpus images on every GPU present -- the path shards over images with no data-path
# collective, so each GPU's work is exactly that slice) and says so in config.workload.
CONFIGS = {
    "c1": dict(global_batch=1, gpus=1, size=256, ddim=1, task="ir", name="configs[0]: single 256x256 LQ image, 1 DDIM step"),
    "c2": dict(global_batch=8, gpus=1, size=512, ddim=20, task="ir", name="configs[1]: batch=8 512x512, 20 DDIM steps, PIR prompt"),
    "c3": dict(global_batch=32, gpus=8, size=512, ddim=20, task="seg", name="configs[2]: batch=32 512x512, 20 steps, 8 GPUs, Segmentation TFA head"),
    "c4": dict(global_batch=16, gpus=4, size=1024, ddim=50, task="ir", name="configs[3]: batch=16 1024x1024, 50 DDIM steps, 4 GPUs"),
    "c5": dict(global_batch=64, gpus=8, size=512, ddim=20, task="ir", name="configs[4]: batch=64 512x512 on 8 GPUs, DDIM-step sweep point"),
}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="c2", choices=sorted(CONFIGS), help="BASELINE.json configuration (default: the headline c2)")
    ap.add_argument("--batch", type=int, default=None, help="images per GPU (default: the configuration's per-GPU slice)")
    ap.add_argument("--size", type=int, default=None)
    ap.add_argument("--ddim-steps", type=int, default=None)
    ap.add_argument("--task", default=None)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--torch-eager", action="store_true",
                    help="also time stock PyTorch eager bf16 (the oracle under autocast) on this GPU: informational key")
    a = ap.parse_args()
    c = CONFIGS[a.config]
    a.batch = a.batch if a.batch is not None else max(1, c["global_batch"] // c["gpus"])
    a.size = a.size if a.size is not None else c["size"]
    a.ddim_steps = a.ddim_steps if a.ddim_steps is not None else c["ddim"]
    a.task = a.task or c["task"]
    return a


def metric_name(a):
    return "restored images/sec @%dx%d, %d DDIM steps" % (a.size, a.size, a.ddim_steps)


def workload_name(a, world):
    c = CONFIGS[a.config]
    w = "%s -- batch=%d/GPU x %d GPU(s) %dx%d, %d DDIM steps, bf16, task '%s', random-init weights" % (
        c["name"], a.batch, world, a.size, a.size, a.ddim_steps, a.task)
    if world < c["gpus"]:
        w += " [per-GPU slice of the %d-GPU configuration: %d of %d images]" % (c["gpus"], a.batch * world, c["global_batch"])
    return w


def env_switches():
    """Development switches of the library.  UR_* variables change kernel selection (never results, but the bench line
    must describe the default build): bench.py refuses to run with any of them set.  UNIRESTORE_* variables select
    documented host options (side streams, graph) and are echoed in ``config``."""
    bad = sorted(k for k in os.environ if k.startswith("UR_"))
    if bad:
        raise SystemExit("bench.py: refusing to run with development switches set: %s" % ", ".join(bad))
    return {k: os.environ[k] for k in sorted(os.environ) if k.startswith("UNIRESTORE_") and k != "UNIRESTORE_B200_LIB"}


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
def cpu_reference_forward(a, timed, warm, threads=None):
    """The oracle port of the reference's PyTorch path on the host cores: FULL forwards (VAE-encode+CFRM, all DDIM
    steps of Controller+UNet+SC-Tuner, VAE-decode+TFA) of ONE image of the workload's size -- a bounded sample of the
    batch, nothing extrapolated.  ``warm`` untimed forwards first (oneDNN primitive creation, allocator warm-up).
    Returns the list of timed seconds per forward."""
    from oracle import unirestore as O
    torch.set_num_threads(threads or os.cpu_count())
    cfg = (CFG[0], dict(CFG[1], num_inference_steps=a.ddim_steps), CFG[2])
    m = cheap_init_(O.DiffUIE(*cfg)).eval().requires_grad_(False)
    img = torch.rand(1, 3, a.size, a.size, generator=torch.Generator().manual_seed(42))
    times = []
    with torch.no_grad():
        for i in range(warm + timed):
            t0 = time.perf_counter()
            m(img, a.task)
            if i >= warm:
                times.append(time.perf_counter() - t0)
    return times


def run_reference(a):
    """--impl reference: the reference's own algorithm (oracle port, fp32 PyTorch) on the host CPU, all host threads.
    One bench step = ONE full forward of ONE image of the configuration (a bounded sample of the batch: the CPU path
    is batch-linear); ``ms_per_step`` is the measured time of that forward, value = 1 / that."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count()
    times = cpu_reference_forward(a, a.steps, a.warmup, cores)
    t_fwd = sum(times) / len(times)
    value = 1.0 / t_fwd
    sample = ("%d timed + %d warm-up full forwards of 1 image %dx%d, %d DDIM steps, task '%s', fp32 CPU oracle port of the "
              "reference path on %d threads (bounded sample: 1 image instead of the batch; nothing extrapolated)"
              % (a.steps, a.warmup, a.size, a.size, a.ddim_steps, a.task, cores))
    print(json.dumps({
        "impl": "reference", "metric": metric_name(a), "value": value, "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps,
        "warmup": a.warmup, "ms_per_step": 1e3 * t_fwd, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(a, max(1, a.gpus)),
                   "reference_arm": "host CPU, fp32; one step = one full forward of 1 image (see cpu_baseline.sample)"},
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
def traffic_from_profile():
    """dram__bytes_read.sum + dram__bytes_write.sum of the dominant kernel from the newest committed ncu --set full
    summary of that kernel (profiles/ncu_r*_gemm_unet320.txt), parsed -- not a constant.  (file, bytes) or (None, None)."""
    import glob
    import re
    files = sorted(glob.glob(os.path.join(ROOT, "profiles", "ncu_r*_gemm_unet320.txt")), key=os.path.getmtime)
    unit = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    for f in reversed(files):
        tot, seen = 0.0, 0
        for line in open(f):
            m = re.match(r"\s*dram__bytes_(read|write)\.sum\s+([0-9.]+)\s+(\w+)\s*$", line)
            if m and m.group(3) in unit:
                tot += float(m.group(2)) * unit[m.group(3)]
                seen += 1
            if seen == 2:
                return os.path.relpath(f, ROOT), tot
    return None, None


def kernel_roofline(dev):
    """The dominant kernel (tcgen05 implicit-GEMM conv, UNet 320->320 3x3 @64x64, B=8: 60.4 GFLOP per launch = SURVEY 8d
    counting rule 2*B*Ho*Wo*Cout*taps*Cin) timed alone with CUDA events on the launching stream; operands rotate over 8
    buffer sets (> L2); median of 5 trials of 40 launches."""
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
    reps, trials = 40, []
    for _ in range(5):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for i in range(reps):
            x, o = sets[i % 8]
            ops.conv_gemm(x, w, C, taps=ops.TAPS_3x3, bias=bias, out=o)
        e.record()
        torch.cuda.synchronize()
        trials.append(s.elapsed_time(e) * 1e-3 / reps)
    t = statistics.median(trials)
    flops = 2.0 * B * H * W * C * C * 9
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("bf16_tflops", 1590.0))
    ach = flops / t / 1e12
    tfile, traffic = traffic_from_profile()
    return {"bound": "tensor", "achieved": ach, "peak": peak, "unit": "TFLOP/s", "frac": ach / peak,
            "traffic": traffic, "traffic_source": tfile,
            "kernel": "ur::conv_gemm_persistent_kernel<160, PAIR> conv3x3 320->320 @64x64 B=8",
            "launch_us": t * 1e6, "algorithmic_flop_per_launch": flops,
            "peak_source": "MEASURED_PEAKS.json bf16_tflops (burst: kernel timed alone)" if peaks else "fallback 1590 (burst)"}


def torch_eager_same_gpu(a, dev, img):
    """Informational (VERDICT r1 item 2 / SURVEY 8d "kernel to beat"): stock PyTorch eager on the SAME GPU -- the oracle
    restatement under torch.autocast(bf16), cuDNN / cuBLASLt / SDPA kernels -- on the same workload.  Not a parity
    oracle here and not part of value / e2e."""
    from oracle import unirestore as O
    cfg = (CFG[0], dict(CFG[1], num_inference_steps=a.ddim_steps), CFG[2])
    om = cheap_init_(O.DiffUIE(*cfg)).eval().requires_grad_(False).to(dev)
    om.scheduler.set_timesteps(a.ddim_steps, device=dev)            # as Lightning would (unifie.py:73-75)
    with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
        for _ in range(2):
            om(img, a.task)
        torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(3):
            om(img, a.task)
        e.record()
        torch.cuda.synchronize()
    ms = s.elapsed_time(e) / 3
    del om
    torch.cuda.empty_cache()
    return {"value": img.shape[0] / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms,
            "what": "stock PyTorch %s eager, autocast bf16 (oracle restatement on cuda), same batch, 2 warm-up + 3 timed forwards"
                    % torch.__version__}


def run_ours(a):
    switches = env_switches()
    emit = claim_stdout()
    import torch.distributed as dist
    from unirestore_b200 import _cabi
    from unirestore_b200 import dist as urdist
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
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    out_box = [None]

    def step(x):
        y = model(x, a.task)
        if world > 1:                                # the only collective on the path: the tested helper
            out_box[0] = urdist.gather_restored(y, world * B)
        return y

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(a.warmup, 3)):
        y = step(img)
    barrier()
    gather_check = None
    if world > 1:                                    # self-check of the plumbing that was just timed-in
        lo, hi = urdist.shard_bounds(world * B, world, rank)
        ok = bool(torch.equal(out_box[0][lo:hi], y)) and bool(torch.isfinite(out_box[0]).all())
        t_ok = torch.tensor([1 if ok else 0], device=dev)
        dist.all_reduce(t_ok, op=dist.ReduceOp.MIN)
        gather_check = "ok" if int(t_ok.item()) == 1 else "MISMATCH"
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
    gf = next(iter(model._graphs.values()))
    graph_launches = getattr(gf, "n_launches", None)
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
            "metric": metric_name(a), "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": max(a.warmup, 3),
            "ms_per_step": total_ms / a.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16", "data": "synthetic",
            "config": {"workload": workload_name(a, world), "baseline_config": a.config,
                       "l2": "256 MiB flush between timed steps; per-step working set >> 126 MB L2",
                       "parallelism": "dp%d (batch-sharded, final NCCL all-gather)" % world,
                       "path": "CUDA-graph replay, Controller and SC-Tuner side streams (the path tests/test_path_gpu.py checks)",
                       "env": switches},
            "clocks": clk.summary(),
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": nbytes, "d2h_bytes_per_step": nbytes},
            "gpu_launches": gpu_launches, "launches_per_forward": graph_launches,
            "launches_by_entry_point": getattr(gf, "launches_by_entry", None),
            "roofline": roof,
        }
        if gather_check is not None:
            line["gather_check"] = gather_check
        if world == 1 and not a.no_cpu_baseline:
            try:
                cores = os.cpu_count()
                times = cpu_reference_forward(a, 1, 1, cores)
                line["cpu_baseline"] = {
                    "value": 1.0 / times[0], "unit": UNIT, "cores": cores, "kind": "port",
                    "sample": "1 warm-up + 1 timed full forward of 1 image %dx%d, %d DDIM steps, fp32 oracle port of the "
                              "reference path on the host CPU (%.1f s; nothing extrapolated)"
                              % (a.size, a.size, a.ddim_steps, times[0])}
            except Exception as ex:   # the oracle is test infrastructure; never let it break the GPU numbers
                line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": os.cpu_count(), "kind": "port",
                                        "sample": "failed: %r" % (ex,)}
        if world == 1 and a.torch_eager:
            try:
                del model
                torch.cuda.empty_cache()
                line["torch_eager_bf16_same_gpu"] = torch_eager_same_gpu(a, dev, img)
            except Exception as ex:
                line["torch_eager_bf16_same_gpu"] = {"value": None, "what": "failed: %r" % (ex,)}
        emit(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)
