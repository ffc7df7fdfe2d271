"""Micro-benchmark of ur_conv_gemm on the dominant shapes (CUDA events, L2 flushed between launches)."""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from unirestore_b200 import ops  # noqa: E402

from tools.bench_gemm_shapes import SHAPES  # noqa: E402
ap = argparse.ArgumentParser()
ap.add_argument("--shapes", default=",".join(SHAPES))
ap.add_argument("--iters", type=int, default=10)
ap.add_argument("--noflush", action="store_true")
a = ap.parse_args()
dev = "cuda:0"
import subprocess
from unirestore_b200 import _cabi
print(subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,clocks.max.sm,power.draw,power.limit", "--format=csv,noheader"],
                     capture_output=True, text=True).stdout.strip(), flush=True)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for name in a.shapes.split(","):
    B, H, W, Ci, Co, nt = SHAPES[name]
    x = torch.randn(B, H, W, Ci, device=dev).to(torch.bfloat16)
    w = (torch.randn(Co, nt * Ci, device=dev) * (nt * Ci) ** -0.5).to(torch.bfloat16)
    bias = torch.randn(Co, device=dev)
    taps = ops.TAPS_3x3 if nt == 9 else ops.TAPS_1x1
    out = torch.empty(B, H, W, Co, device=dev, dtype=torch.bfloat16)
    res = {}
    for v1 in (0, 1, 2, 3):   # 0: auto (CTA pairs when the cost model says so), 1: v1 kernel, 2: persistent single-CTA, 3: forced pairs
        _cabi.lib().ur_debug_force_gemm_v1(1 if v1 == 1 else 0)
        _cabi.lib().ur_debug_set_gemm_pair_mode(0 if v1 == 2 else (1 if v1 == 3 else -1))
        for _ in range(3):
            ops.conv_gemm(x, w, Co, taps=taps, bias=bias, out=out)
        ts = []
        for _ in range(a.iters):
            if not a.noflush:
                flush.zero_()
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            ops.conv_gemm(x, w, Co, taps=taps, bias=bias, out=out)
            e.record()
            torch.cuda.synchronize()
            ts.append(s.elapsed_time(e))
        ts.sort()
        res[v1] = ts[len(ts) // 2] * 1e-3
    _cabi.lib().ur_debug_force_gemm_v1(0)
    _cabi.lib().ur_debug_set_gemm_pair_mode(-1)
    fl = 2.0 * B * H * W * Ci * Co * nt
    by = 2.0 * (B * H * W * (Ci + Co) + Co * nt * Ci)
    t = res[0]
    print("%-20s auto %8.1f us %7.1f TF/s %6.0f GB/s | 1-CTA persistent %8.1f us %7.1f TF/s | forced pairs %8.1f us %7.1f TF/s | v1 %8.1f us %7.1f TF/s" % (
        name, t * 1e6, fl / t / 1e12, by / t / 1e9, res[2] * 1e6, fl / res[2] / 1e12, res[3] * 1e6, fl / res[3] / 1e12,
        res[1] * 1e6, fl / res[1] / 1e12), flush=True)
