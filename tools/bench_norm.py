"""Micro-benchmark of the HBM-bound normalisation kernels (CUDA events; cold = 256 MiB L2 flush before each launch)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from unirestore_b200 import _cabi, ops  # noqa: E402

dev = "cuda:0"
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timeit(fn, cold, iters=10):
    for _ in range(3):
        fn()
    ts = []
    for _ in range(iters):
        if cold:
            flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        fn()
        e.record()
        torch.cuda.synchronize()
        ts.append(s.elapsed_time(e) * 1e3)
    ts.sort()
    return ts[len(ts) // 2]


for (B, H, W, C, G) in [(8, 64, 64, 320, 32), (8, 32, 32, 640, 32), (8, 16, 16, 1280, 32), (8, 64, 64, 960, 32),
                        (8, 256, 256, 256, 32), (8, 512, 512, 128, 32)]:
    x = torch.randn(B, H, W, C, device=dev).to(torch.bfloat16)
    g = torch.ones(C, device=dev)
    b = torch.zeros(C, device=dev)
    st = ops.chan_stats(x)
    out = torch.empty_like(x)
    mb = x.numel() * 2 / 1e6
    for cold in (False, True):
        t_s = timeit(lambda: ops.chan_stats(x, stats=st), cold)
        t_a = timeit(lambda: ops.norm_apply(x, st, G, g, b, 1e-5, True, None, out), cold)
        ops.FUSED_GN_MAX_BYTES = 1 << 40
        tf = {}
        for cl in (4, 8, 16):
            _cabi.lib().ur_debug_set_group_norm_cluster(cl)
            tf[cl] = timeit(lambda: ops.group_norm(x, G, g, b, 1e-5, True), cold)
        _cabi.lib().ur_debug_set_group_norm_cluster(0)
        t_f = tf[16]
        line = "%-22s %s  stats %7.1f us %6.0f GB/s | apply+silu %7.1f us %6.0f GB/s | fused cluster GN %7.1f us %6.0f GB/s" % (
            (B, H, W, C), "cold" if cold else "warm", t_s, mb / t_s * 1e3, t_a, 2 * mb / t_a * 1e3, t_f, 2 * mb / t_f * 1e3)
        line += " [CL4 %.1f CL8 %.1f]" % (tf[4], tf[8])
        if C <= 1280 and H * W <= 4096:
            t_l = timeit(lambda: ops.layernorm(x.view(B, H * W, C), g, b, 1e-5), cold)
            line += " | layernorm %7.1f us %6.0f GB/s" % (t_l, 2 * mb / t_l * 1e3)
        print(line, flush=True)
