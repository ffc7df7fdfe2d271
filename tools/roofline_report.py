"""Per-kernel-family roofline report of one DiffUIE.forward (BASELINE configs[4] "roofline report per kernel").

The forward is captured as ONE CUDA graph on one stream with PDL off (kernels strictly back to back, warm L2 exactly as
in the benchmarked graph) and replayed under torch.profiler (CUPTI kernel durations).  Every ops.* call made during the
capture is logged with its ALGORITHMIC work, computed from the call's shapes with the counting rules of SURVEY.md 8d /
Appendix C, and matched to its kernel (the k-th conv_gemm call is the k-th conv_gemm kernel of the replay, ...):
    conv / linear        FLOPs = 2 * B*Ho*Wo * N * taps * K_per_tap     bytes = 2*(input) + out + 2*weights (+ residual)
    attention            FLOPs = 4 * B * Tq * Tk * C                     bytes = 2*(q + k + v + out)
    norm / elementwise   bytes = read + written tensors
Fractions: tensor = TFLOP/s / bf16_tflops_sustained, hbm = GB/s / hbm_gbs (MEASURED_PEAKS.json); the binding roof of a
family is the larger of the two.  Markdown on stdout:
    python tools/roofline_report.py [--batch 8 --size 512 --steps 20] > profiles/roofline_r2.md
"""
import argparse
import collections
import json
import os
import sys

os.environ["UR_PDL"] = "0"          # development switch of the library: no programmatic dependent launch (no kernel overlap)
import torch  # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import CFG, cheap_init_  # noqa: E402
from unirestore_b200 import ops  # noqa: E402
from unirestore_b200.diffuie import DiffUIE  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=8)
ap.add_argument("--size", type=int, default=512)
ap.add_argument("--steps", type=int, default=20)
ap.add_argument("--task", default="ir")
a = ap.parse_args()
dev = "cuda:0"
peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
PT, PH = float(peaks["bf16_tflops_sustained"]), float(peaks["hbm_gbs"])

cfg = (CFG[0], dict(CFG[1], num_inference_steps=a.steps), CFG[2])
m = cheap_init_(DiffUIE(*cfg)).eval().requires_grad_(False).to(dev)
m.overlap_controller = False
m.base_model.overlap_sc_tuner = False
img = torch.rand(a.batch, 3, a.size, a.size, device=dev)

calls = collections.defaultdict(list)        # kernel-name key -> [(family, flops, bytes), ...] in call order
LOG = [False]


def nbytes(t):
    return 0 if t is None else t.numel() * t.element_size()


def conv_family(x, w, n, kw):
    taps = len(kw.get("taps", ops.TAPS_1x1))
    x4 = ops._as4(x)
    B, H, W, C1 = x4.shape
    c2 = kw["x2"].shape[-1] if kw.get("x2") is not None else 0
    stride = kw.get("stride", 1)
    ho, wo = kw.get("hout") or H // stride, kw.get("wout") or W // stride
    gk = kw.get("group_kc", 0)
    kper = gk if gk else C1 + c2
    flops = 2.0 * B * ho * wo * n * taps * kper
    act = kw.get("act", 0)
    n_out = n // 2 if act in (ops.UR_ACT_GEGLU, ops.UR_ACT_GATE) else n
    osz = 4 if kw.get("out_dtype", torch.bfloat16) == torch.float32 else 2
    byts = 2.0 * B * H * W * (C1 + c2) + osz * B * ho * wo * n_out + nbytes(w) + (2.0 * B * ho * wo * n_out if kw.get("residual") is not None else 0)
    if act == ops.UR_ACT_GEGLU:
        fam = "linear + GEGLU epilogue (K=%d)" % kper
    elif act == ops.UR_ACT_GATE:
        fam = "1x1 conv + SimpleGate epilogue"
    elif gk:
        fam = "grouped conv3x3 (+GELU)"
    elif taps >= 9 and stride == 2:
        fam = "conv3x3 stride 2"
    elif taps == 4:
        fam = "upsample conv (sub-pixel 2x2 phases)"
    elif taps >= 9:
        fam = ("conv3x3 C>=128, %dx%d" % (ho, wo)) if C1 + c2 >= 128 else "conv3x3 stem/head (C<128)"
    elif kper <= 640:
        fam = "1x1 conv / linear, K<=640"
    else:
        fam = "1x1 conv / linear, K>640"
    detail = "M=%d K=%d N=%d taps=%d act=%d%s%s%s" % (B * ho * wo, kper, n, taps, act, " +res" if kw.get("residual") is not None else "",
                                                   " +stats" if kw.get("want_stats") or kw.get("stats") is not None else "",
                                                   " x2" if c2 else "")
    return fam, flops, byts, detail


def wrap(name, key, cost):
    fn = getattr(ops, name)

    def w(*args, **kw):
        r = fn(*args, **kw)
        if LOG[0]:
            calls[key].append(cost(r, *args, **kw))
        return r
    setattr(ops, name, w)


wrap("conv_gemm", "conv_gemm", lambda r, x, w, n, **kw: conv_family(x, w, n, kw))
wrap("attention", "attention", lambda r, q, k, v, heads, **kw: (
    "attention head_dim %d, %d x %d tokens" % (q.shape[-1] // heads, q.shape[1], k.shape[1]),
    4.0 * q.shape[0] * q.shape[1] * k.shape[1] * q.shape[2],
    2.0 * (q.shape[0] * q.shape[1] * q.shape[2] * 2 + 2 * k.shape[0] * k.shape[1] * q.shape[2])))
wrap("norm_apply", "norm_apply", lambda r, x, st, *a_, **kw: ("GroupNorm / InstanceNorm apply (+SiLU)", 0.0, 2.0 * nbytes(r)))
wrap("layernorm", "layernorm", lambda r, x, *a_, **kw: ("LayerNorm / LayerNorm2d", 0.0, 2.0 * nbytes(x)))
wrap("dwconv3x3_gate", "dwconv3x3_gate", lambda r, x, *a_, **kw: ("depthwise 3x3 + SimpleGate + GAP", 0.0, 1.5 * nbytes(x)))
wrap("scale_channels_", "scale_channels", lambda r, x, *a_, **kw: ("channel scale (in place)", 0.0, 2.0 * nbytes(x)))

m.use_cuda_graph = True
LOG[0] = False
for _ in range(1):
    pass
# the first graphed call warms up twice (unlogged) and captures once: log only the capture pass
from unirestore_b200.diffuie import unifie as _u  # noqa: E402
_orig_graph = torch.cuda.graph


class _LoggedGraph(_orig_graph):
    def __enter__(self):
        LOG[0] = True
        return super().__enter__()

    def __exit__(self, *e):
        LOG[0] = False
        return super().__exit__(*e)


torch.cuda.graph = _LoggedGraph
m(img, a.task)
torch.cuda.graph = _orig_graph
for _ in range(2):
    m(img, a.task)
torch.cuda.synchronize()
from torch.profiler import ProfilerActivity, profile  # noqa: E402
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    m(img, a.task)
    torch.cuda.synchronize()
evs = [ev for ev in prof.events() if ev.device_type == torch.autograd.DeviceType.CUDA]
evs.sort(key=lambda ev: ev.time_range.start)
KEYS = [("conv_gemm", "conv_gemm"), ("attention", "attention"), ("norm_apply", "norm_apply"), ("layernorm", "layernorm"),
        ("dwconv3x3_gate", "dwconv3x3_gate"), ("scale_channels", "scale_channels")]
idx = collections.Counter()
rec = collections.defaultdict(lambda: [0, 0.0, 0.0, 0.0])      # family -> [launches, flops, bytes, seconds]
shapes = collections.defaultdict(lambda: [0, 0.0, 0.0, 0.0])   # conv_gemm shape signature -> the same
unmatched = 0
for ev in evs:
    name = ev.name
    dur = (ev.device_time if hasattr(ev, "device_time") else ev.cuda_time) * 1e-6
    fam, fl, by = None, 0.0, 0.0
    for sub, key in KEYS:
        if sub in name and "splitk" not in name:
            k = idx[key]
            idx[key] += 1
            if k < len(calls[key]):
                fam, fl, by = calls[key][k][:3]
                if len(calls[key][k]) > 3:
                    S = shapes[calls[key][k][3]]
                    S[0] += 1
                    S[1] += fl
                    S[2] += by
                    S[3] += dur
            else:
                unmatched += 1
            break
    if fam is None:
        if "splitk_finish" in name:
            fam = "split-K finish (slab sum + epilogue + statistics)"
        elif "chan_stats" in name:
            fam = "channel statistics pass"
        elif name.startswith("void ur::") or "ur::" in name:
            fam = "small / latent / image kernels"
        else:
            fam = "PyTorch fill / copy nodes"
    R = rec[fam]
    R[0] += 1
    R[1] += fl
    R[2] += by
    R[3] += dur
for key in calls:
    assert idx[key] == len(calls[key]), "kernel / call count mismatch for %s: %d kernels, %d calls" % (key, idx[key], len(calls[key]))
rows = sorted(((v[3], k, v[0], v[1], v[2]) for k, v in rec.items()), reverse=True)
tot_t, tot_f = sum(r[0] for r in rows), sum(r[3] for r in rows)
print("# Per-kernel-family roofline, DiffUIE.forward B=%d %dx%d %d DDIM steps task '%s' (round 2)\n" % (a.batch, a.size, a.size, a.steps, a.task))
print("One captured CUDA graph, one stream, PDL off (`tools/roofline_report.py`): kernels run back to back with the L2 state of "
      "the benchmarked graph; durations from CUPTI (torch.profiler); algorithmic FLOPs / bytes per call by the SURVEY.md 8d "
      "counting rules.  Peaks (MEASURED_PEAKS.json): %.0f TFLOP/s sustained bf16, %.0f GB/s HBM.  `frac` = the larger of the "
      "two fractions = the family's binding roof (BASELINE.md section 2).\n" % (PT, PH))
print("Summed kernel time %.1f ms per forward; %.1f TFLOP algorithmic -> %.0f TFLOP/s overall = %.1f %% of the sustained "
      "tensor roof.\n" % (tot_t * 1e3, tot_f / 1e12, tot_f / tot_t / 1e12, 100 * tot_f / tot_t / 1e12 / PT))
print("| family | launches | time ms | share | avg us | GFLOP | MB | TFLOP/s | GB/s | tensor frac | HBM frac | bound | frac |")
print("|---|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|---|---:|")
for t, fam, n, fl, by in rows:
    tf, gb = fl / t / 1e12, by / t / 1e9
    ft, fh = tf / PT, gb / PH
    print("| %s | %d | %.2f | %.1f %% | %.1f | %.0f | %.0f | %.0f | %.0f | %.2f | %.2f | %s | %.2f |"
          % (fam, n, t * 1e3, 100 * t / tot_t, t / n * 1e6, fl / 1e9, by / 1e6, tf, gb, ft, fh,
             "tensor" if ft >= fh else "hbm", max(ft, fh)))

print("\n## GEMM / conv launches by shape (top 45 by time)\n")
print("| shape | launches | time ms | avg us | TFLOP/s | GB/s | tensor frac | HBM frac |")
print("|---|---:|---:|---:|---:|---:|---:|---:|")
for sig, (n, fl, by, t) in sorted(shapes.items(), key=lambda kv: -kv[1][3])[:45]:
    print("| %s | %d | %.2f | %.1f | %.0f | %.0f | %.2f | %.2f |" % (sig, n, t * 1e3, t / n * 1e6, fl / t / 1e12, by / t / 1e9,
                                                                fl / t / 1e12 / PT, by / t / 1e9 / PH))
