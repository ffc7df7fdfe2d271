"""Micro-benchmark of ur_attention (CUDA events)."""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from unirestore_b200 import ops  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--iters", type=int, default=10)
ap.add_argument("--cases", default="self64,self32,cross64,ctrl64")
ap.add_argument("--poly", type=int, default=3, help="exp2 pairs of every 8 on the FMA pipe (attention2)")
ap.add_argument("--impl", type=int, default=2, help="attention kernel generation (1 or 2)")
ap.add_argument("--kv1", type=int, default=1, help="query-tile-loop kernel for single-KV-tile launches")
ap.add_argument("--graph", action="store_true", help="time 32 launches inside one CUDA graph (no host gaps)")
a = ap.parse_args()
from unirestore_b200 import _cabi  # noqa: E402
_cabi.lib().ur_debug_set_attention_impl(a.impl)
_cabi.lib().ur_debug_set_attention_poly(a.poly)
_cabi.lib().ur_debug_set_attention_kv1(a.kv1)
CASES = {"self64": (8, 5, 64, 4096, 4096), "self32": (8, 10, 64, 1024, 1024), "cross64": (8, 5, 64, 4096, 77),
         "ctrl64": (8, 4, 64, 4096, 4096), "self16": (8, 20, 64, 256, 256), "self128": (4, 5, 64, 16384, 16384),
         "ctrl128": (8, 4, 128, 256, 256), "cross32": (8, 10, 64, 1024, 77), "cross16": (8, 20, 64, 256, 77)}
dev = "cuda:0"
for name in a.cases.split(","):
    B, h, d, Tq, Tk = CASES[name]
    C = h * d
    q = torch.randn(B, Tq, C, device=dev).to(torch.bfloat16)
    kb = 1 if Tk == 77 else B
    k = torch.randn(kb, Tk, C, device=dev).to(torch.bfloat16)
    v = torch.randn(kb, Tk, C, device=dev).to(torch.bfloat16)
    out = torch.empty_like(q)
    for _ in range(3):
        ops.attention(q, k, v, h, out=out)
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if a.graph:
        st = torch.cuda.Stream()
        with torch.cuda.stream(st):
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=st):
                for _ in range(32):
                    ops.attention(q, k, v, h, out=out)
        g.replay()
        torch.cuda.synchronize()
        s.record()
        g.replay()
        e.record()
        torch.cuda.synchronize()
        t = s.elapsed_time(e) * 1e-3 / 32
    else:
        s.record()
        for _ in range(a.iters):
            ops.attention(q, k, v, h, out=out)
        e.record()
        torch.cuda.synchronize()
        t = s.elapsed_time(e) * 1e-3 / a.iters
    fl = 4.0 * B * Tq * Tk * C
    print("impl%d poly%d %-8s B=%d h=%d d=%d Tq=%d Tk=%d  %8.1f us  %7.1f TF/s" % (a.impl, a.poly, name, B, h, d, Tq, Tk, t * 1e6, fl / t / 1e12), flush=True)
