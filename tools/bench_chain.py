"""Back-to-back timing of ur_conv_gemm shapes inside one CUDA graph (no host launch gaps, no L2 flush artefacts: a
flush leaves 126 MB of dirty lines whose write-back lands inside the next short kernel).  `ring` buffer sets are cycled:
ring 1 = everything L2-resident, ring 6 = inputs / outputs larger than L2 (HBM-cold), 2 = the forward's usual case (the
input was just written by the previous kernel).

usage: python tools/bench_chain.py [--ring 1,2,6] [--reps 48]"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from unirestore_b200 import ops  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--ring", default="1,2,6")
ap.add_argument("--reps", type=int, default=48)
ap.add_argument("--cases", default="")
ap.add_argument("--pair", type=int, default=-1, help="CTA-pair mode: -1 library rule, 0 never, 1 whenever legal")
ap.add_argument("--gbn", type=int, default=0, help="force the N tile of the gated cases (0 = library choice)")
ap.add_argument("--bn", type=int, default=0, help="force the N tile of the non-gated cases (0 = library choice)")
a = ap.parse_args()
dev = "cuda:0"
from unirestore_b200 import _cabi  # noqa: E402
_cabi.lib().ur_debug_set_gemm_pair_mode(a.pair)
CASES = [(32768, 320, 320, ops.UR_ACT_NONE, False, False, "lin64"), (32768, 320, 320, ops.UR_ACT_NONE, True, False, "lin64+res"),
         (32768, 320, 320, ops.UR_ACT_NONE, True, True, "lin64+res+stats"), (32768, 320, 960, ops.UR_ACT_NONE, False, False, "qkv64"),
         (32768, 320, 320, ops.UR_ACT_GELU, False, False, "gelu64"), (32768, 320, 2560, ops.UR_ACT_GEGLU, False, False, "geglu64"),
         (8192, 640, 5120, ops.UR_ACT_GEGLU, False, False, "geglu32"), (32768, 1280, 320, ops.UR_ACT_NONE, True, False, "ffout64+res"),
         (8192, 640, 640, ops.UR_ACT_NONE, True, False, "lin32+res"), (8192, 640, 1920, ops.UR_ACT_NONE, False, False, "qkv32"),
         (2048, 1280, 1280, ops.UR_ACT_NONE, True, False, "lin16+res"), (2048, 1280, 10240, ops.UR_ACT_GEGLU, False, False, "geglu16"),
         (32768, 320, 2560, ops.UR_ACT_NONE, False, False, "wide64"), (2048, 1280, 3840, ops.UR_ACT_NONE, False, False, "qkv16"),
         (32768, 256, 768, ops.UR_ACT_NONE, False, False, "naf768")]
if a.cases:
    CASES = [c for c in CASES if c[-1] in a.cases.split(",")]
for (M, K, N, act, res, stats, name) in CASES:
    line = "%-16s M=%5d K=%4d N=%5d " % (name, M, K, N)
    for ring in [int(v) for v in a.ring.split(",")]:
        xs = [torch.randn(8, M // 8, K, device=dev).to(torch.bfloat16) for _ in range(ring)]
        w = (torch.randn(N, K, device=dev) * K ** -0.5).to(torch.bfloat16)
        b = torch.randn(N, device=dev)
        bn = (a.gbn or ops.pick_bn(N, True)) if act == ops.UR_ACT_GEGLU else a.bn
        if act == ops.UR_ACT_GEGLU:
            w, b = ops.pack_gated_weight(w, b, bn)
        n_out = N // 2 if act == ops.UR_ACT_GEGLU else N
        outs = [torch.empty(8, M // 8, n_out, device=dev, dtype=torch.bfloat16) for _ in range(ring)]
        rs = [torch.randn(8, M // 8, n_out, device=dev).to(torch.bfloat16) for _ in range(ring)] if res else None
        st = torch.zeros(8, n_out, 2, device=dev, dtype=torch.float64) if stats else None

        def chain():
            for i in range(a.reps):
                k = i % ring
                ops.conv_gemm(xs[k], w, N, bias=b, act=act, bn=bn, out=outs[k], residual=rs[k] if res else None, stats=st)
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
        by = 2.0 * (M * K + N * K + M * n_out * (2 if res else 1))
        line += "| ring %d: %6.1f us %5.0f GB/s %5.0f TF/s " % (ring, t, by / t * 1e-3, 2.0 * M * K * N / t * 1e-6)
        del g
    print(line, flush=True)
