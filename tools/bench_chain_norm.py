"""Back-to-back timing of ur_norm_apply / ur_layernorm / ur_chan_stats inside one CUDA graph (see tools/bench_chain.py for
why: no host gaps, no flush artefacts).  `ring` buffer sets are cycled (2 = the forward's usual state).

usage: python tools/bench_chain_norm.py [--ring 2] [--reps 48]"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from unirestore_b200 import ops  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--ring", type=int, default=2)
ap.add_argument("--reps", type=int, default=48)
a = ap.parse_args()
dev = "cuda:0"
# (kind, B, pixels, C, silu)
CASES = [("gn", 8, 4096, 320, True), ("gn", 8, 4096, 320, False), ("gn", 8, 4096, 640, True), ("gn", 8, 4096, 960, True),
         ("gn", 8, 1024, 640, True), ("gn", 8, 1024, 1280, True), ("gn", 8, 1024, 1920, True), ("gn", 8, 256, 1280, True),
         ("gn", 8, 256, 2560, True), ("gn", 8, 64, 1280, True), ("gn", 8, 65536, 128, True), ("gn", 8, 262144, 128, True),
         ("ln", 8, 4096, 320, False), ("ln", 8, 1024, 640, False), ("ln", 8, 256, 1280, False),
         ("stats", 8, 4096, 320, False), ("stats", 8, 262144, 128, False)]
for kind, B, P, C, silu in CASES:
    xs = [torch.randn(B, P, C, device=dev).to(torch.bfloat16) for _ in range(a.ring)]
    outs = [torch.empty_like(xs[0]) for _ in range(a.ring)]
    gamma, beta = torch.rand(C, device=dev) + 0.5, torch.randn(C, device=dev)
    st = ops.chan_stats(xs[0])

    def chain():
        for i in range(a.reps):
            k = i % a.ring
            if kind == "gn":
                ops.norm_apply(xs[k], st, 32, gamma, beta, 1e-5, silu=silu, out=outs[k])
            elif kind == "ln":
                ops.layernorm(xs[k], gamma, beta, 1e-5)
            else:
                ops.chan_stats(xs[k], stats=st)
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        chain()
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=s):
            chain()
    ts = []
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        g.replay()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3 / a.reps)
    ts.sort()
    t = ts[len(ts) // 2]
    by = B * P * C * 2 * (1 if kind == "stats" else 2)
    print("%-5s B=%d P=%6d C=%4d silu=%d  %7.1f us  %5.0f GB/s  (%5.1f MB)" % (kind, B, P, C, silu, t, by / t * 1e-3, by / 1e6), flush=True)
    del g
