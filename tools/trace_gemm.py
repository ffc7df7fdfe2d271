"""Per-role timeline (clock64) of CTA 0 of the persistent GEMM kernel for one shape."""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from unirestore_b200 import _cabi, ops  # noqa: E402
from tools.bench_gemm_shapes import SHAPES  # noqa: E402

dev = "cuda:0"
for name in sys.argv[1:]:
    B, H, W, Ci, Co, nt = SHAPES[name]
    x = torch.randn(B, H, W, Ci, device=dev).to(torch.bfloat16)
    w = (torch.randn(Co, nt * Ci, device=dev) * (nt * Ci) ** -0.5).to(torch.bfloat16)
    bias = torch.randn(Co, device=dev)
    taps = ops.TAPS_3x3 if nt == 9 else ops.TAPS_1x1
    out = torch.empty(B, H, W, Co, device=dev, dtype=torch.bfloat16)
    for _ in range(3):
        ops.conv_gemm(x, w, Co, taps=taps, bias=bias, out=out)
    tr = torch.zeros(256, dtype=torch.int64, device=dev)
    _cabi.lib().ur_debug_set_gemm_trace(C.c_void_p(tr.data_ptr()))
    ops.conv_gemm(x, w, Co, taps=taps, bias=bias, out=out)
    torch.cuda.synchronize()
    _cabi.lib().ur_debug_set_gemm_trace(C.c_void_p(0))
    full = tr.cpu()
    t = full[:128].view(8, 16)
    t0 = int(t[7, 0])
    print("== %s: prologue %d cycles" % (name, int(t[7, 1]) - t0))
    labels = ["producer tile start", "mma: acc stage free", "mma: first operands landed", "mma: last MMA issued",
              "epi: bias staged", "epi: accumulator ready", "epi: tile done"]
    for i, lab in enumerate(labels):
        print("  %-28s %s" % (lab, " ".join("%7d" % (int(v) - t0) if int(v) else "      -" for v in t[i, :6])))
    iss, land = full[128:144], full[144:160]
    print("  tile 1 k-blocks: TMA issue   ", " ".join("%6d" % (int(v) - t0) for v in iss))
    print("  tile 1 k-blocks: landed(seen)", " ".join("%6d" % (int(v) - t0) for v in land))
    pr, mm = full[192:224].view(8, 4), full[224:256].view(8, 4)
    print("  producer kb16..23 [before wait(empty), after wait, after A issue, after W issue]:")
    for r in pr:
        print("     ", " ".join("%7d" % (int(v) - t0) for v in r))
    print("  mma kb16..23 [before wait(full), after wait, after 4 MMA issue, after commit]:")
    for r in mm:
        print("     ", " ".join("%7d" % (int(v) - t0) for v in r))
    ep = full[160:176].view(4, 4)
    print("  epilogue tile 1, warp 2, sub-blocks [before ld wait, after ld wait, after math, after stores]:")
    for r in ep:
        print("     ", " ".join("%7d" % (int(v) - t0) if int(v) else "      -" for v in r))
