"""Per-role timeline (clock64) of CTA 0 of the persistent GEMM kernel for one shape.

usage: python tools/trace_gemm.py shape[:bn[:pair]] ...   (shapes from tools/bench_gemm_shapes.py)"""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from unirestore_b200 import _cabi, ops  # noqa: E402
from tools.bench_gemm_shapes import SHAPES  # noqa: E402

dev = "cuda:0"
for spec in sys.argv[1:]:
    parts = spec.split(":")
    name, bn = parts[0], int(parts[1]) if len(parts) > 1 else 0
    pair = int(parts[2]) if len(parts) > 2 else -1
    B, H, W, Ci, Co, nt = SHAPES[name]
    x = torch.randn(B, H, W, Ci, device=dev).to(torch.bfloat16)
    w = (torch.randn(Co, nt * Ci, device=dev) * (nt * Ci) ** -0.5).to(torch.bfloat16)
    bias = torch.randn(Co, device=dev)
    taps = ops.TAPS_3x3 if nt == 9 else ops.TAPS_1x1
    out = torch.empty(B, H, W, Co, device=dev, dtype=torch.bfloat16)
    _cabi.lib().ur_debug_set_gemm_pair_mode(pair)
    for _ in range(3):
        ops.conv_gemm(x, w, Co, taps=taps, bias=bias, out=out, bn=bn)
    tr = torch.zeros(512, dtype=torch.int64, device=dev)
    _cabi.lib().ur_debug_set_gemm_trace(C.c_void_p(tr.data_ptr()))
    ops.conv_gemm(x, w, Co, taps=taps, bias=bias, out=out, bn=bn)
    torch.cuda.synchronize()
    _cabi.lib().ur_debug_set_gemm_trace(C.c_void_p(0))
    _cabi.lib().ur_debug_set_gemm_pair_mode(-1)
    full = tr.cpu()
    t = full[:128].view(8, 16)
    t0 = int(t[7, 0])
    rel = lambda v: ("%7d" % (int(v) - t0)) if int(v) else "      -"
    print("== %s bn=%d pair=%d: prologue %d cycles" % (name, bn, pair, int(t[7, 1]) - t0))
    labels = ["producer tile start", "mma: acc stage free", "mma: first operands landed", "mma: last MMA issued",
              "epi: bias staged", "epi: accumulator ready", "epi: tile done"]
    for i, lab in enumerate(labels):
        print("  %-28s %s" % (lab, " ".join(rel(v) for v in t[i, :6])))
    print("  mma thread, tile 1, kb 8..23 [before wait(full), after wait, after MMA issue, after commit]  (d = loop period)")
    mm = full[128:192].view(16, 4)
    prev = None
    for r in mm:
        d = (int(r[0]) - prev) if prev else 0
        prev = int(r[0])
        print("     ", " ".join(rel(v) for v in r), "  d=%d wait=%d issue=%d commit=%d" % (
            d, int(r[1]) - int(r[0]), int(r[2]) - int(r[1]), int(r[3]) - int(r[2])))
    for base, lab, step in ((192, "A producer warp 0 (every 3rd kb)", 3), (256, "W producer warp 3 (every 2nd kb)", 2)):
        print("  %s, tile 1 [before wait(empty), after wait, after TMA issue]" % lab)
        pr = full[base:base + 64].view(16, 4)
        prev = None
        for r in pr:
            if not int(r[0]):
                continue
            d = (int(r[0]) - prev) if prev else 0
            prev = int(r[0])
            print("     ", " ".join(rel(v) for v in r[:3]), "  d=%d wait=%d issue=%d" % (
                d, int(r[1]) - int(r[0]), int(r[2]) - int(r[1])))
    print("  epilogue group 0 thread 0, tile 1, sub-blocks [loop top, after barrier A, after tmem ld wait, after staging write, after barrier C, after global stores]")
    ep = full[320:344].view(3, 8)
    for r in ep:
        if int(r[0]):
            print("     ", " ".join(rel(v) for v in r[:6]))
