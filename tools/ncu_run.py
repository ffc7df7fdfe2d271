"""Tiny driver for ncu captures: builds the model with cheap random weights and runs forward once or twice."""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from unirestore_b200.diffuie import DiffUIE  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=8)
ap.add_argument("--size", type=int, default=512)
ap.add_argument("--steps", type=int, default=1)
ap.add_argument("--reps", type=int, default=1)
a = ap.parse_args()
dev = "cuda:0"
torch.manual_seed(0)
m = DiffUIE(dict(type="CFRM"), dict(type="scedit", num_inference_steps=a.steps),
            dict(type="TFA", prompt_len=1, task=["ir", "cls", "seg"])).eval().requires_grad_(False)
for n, p in m.named_parameters():
    if p.dim() <= 1 or "beta" in n or "gamma" in n or "task_prompts" in n:
        torch.nn.init.normal_(p, 0.0 if p.dim() > 1 or "bias" in n else 1.0, 0.05)
m = m.to(dev)
for mod in m.modules():
    for n, p in mod.named_parameters(recurse=False):
        if float(p.abs().max()) == 0.0:
            torch.nn.init.normal_(p, 0.0, 0.02)
img = torch.rand(a.batch, 3, a.size, a.size, device=dev)
torch.cuda.synchronize()
for _ in range(a.reps):
    y = m(img, "ir")
torch.cuda.synchronize()
print("done", tuple(y.shape), float(y.mean()))
