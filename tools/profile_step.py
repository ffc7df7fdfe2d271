"""Per-op CUDA-event breakdown of one DiffUIE.forward (development tool, not the bench)."""
import argparse
import collections
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from unirestore_b200 import ops  # noqa: E402
from unirestore_b200.diffuie import DiffUIE  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=8)
ap.add_argument("--size", type=int, default=512)
ap.add_argument("--steps", type=int, default=20)
ap.add_argument("--detail", action="store_true")
a = ap.parse_args()

dev = "cuda:0"
torch.manual_seed(0)
m = DiffUIE(dict(type="CFRM"), dict(type="scedit", num_inference_steps=a.steps),
            dict(type="TFA", prompt_len=1, task=["ir", "cls", "seg"])).eval().requires_grad_(False)
from unirestore_b200.init_utils import deterministic_init_  # noqa: E402
t0 = time.time()
for n, p in m.named_parameters():           # cheap non-degenerate init
    if p.dim() <= 1 or "beta" in n or "gamma" in n or "task_prompts" in n:
        torch.nn.init.normal_(p, 0.0 if p.dim() > 1 or "bias" in n else 1.0, 0.05)
m = m.to(dev)
for mod in m.modules():                      # re-randomise the zero-initialised Controller tensors
    for n, p in mod.named_parameters(recurse=False):
        if float(p.abs().max()) == 0.0:
            torch.nn.init.normal_(p, 0.0, 0.02)
print("model ready in %.1fs" % (time.time() - t0), flush=True)

records = collections.defaultdict(list)


def wrap(name):
    fn = getattr(ops, name)

    def w(*args, **kw):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        r = fn(*args, **kw)
        e.record()
        key = name
        if a.detail and name == "conv_gemm":
            x = args[0]
            key = "conv_gemm %s->%d taps%d" % (tuple(x.shape), args[2], len(kw.get("taps", ((0, 0),))))
        records[key].append((s, e))
        return r
    setattr(ops, name, w)


img = torch.rand(a.batch, 3, a.size, a.size, device=dev)
for _ in range(2):
    y = m(img, "ir")
torch.cuda.synchronize()
t0 = time.time()
y = m(img, "ir")
torch.cuda.synchronize()
wall = time.time() - t0
print("forward wall %.1f ms  -> %.2f img/s (B=%d, %d steps)" % (wall * 1e3, a.batch / wall, a.batch, a.steps))

for n in ["conv_gemm", "chan_stats", "norm_apply", "layernorm", "scale_channels_", "softmax_rows", "transpose_tokens",
          "dwconv3x3_gate", "small_linear", "adanaf_scales", "tfa_gates", "posterior_sample", "latent_axpby",
          "ddim_step_", "image_to_nhwc8", "nhwc_to_image", "timestep_embedding"]:
    wrap(n)
# group_norm / attention call through module globals: rebind them to the wrapped leaf ops
y = m(img, "ir")
torch.cuda.synchronize()
tot = 0.0
rows = []
for k, v in records.items():
    t = sum(s.elapsed_time(e) for s, e in v)
    rows.append((t, k, len(v)))
    tot += t
rows.sort(reverse=True)
print("sum of per-op device time: %.1f ms over %d launches" % (tot, sum(r[2] for r in rows)))
for t, k, n in rows[:40]:
    print("%9.2f ms %6.1f%% %6d  %s" % (t, 100 * t / tot, n, k))
