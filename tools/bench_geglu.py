"""Micro-benchmark of the gated (GEGLU) and GELU epilogues of ur_conv_gemm on the transformer feed-forward shapes
(CUDA events, queue kept full by a leading L2 flush)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from unirestore_b200 import ops  # noqa: E402

NOFLUSH = "--noflush" in sys.argv
dev = "cuda:0"
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for (M, K, N, act, name) in [(32768, 320, 2560, ops.UR_ACT_GEGLU, "geglu64"), (8192, 640, 5120, ops.UR_ACT_GEGLU, "geglu32"),
                             (2048, 1280, 10240, ops.UR_ACT_GEGLU, "geglu16"), (32768, 320, 320, ops.UR_ACT_GELU, "gelu64"),
                             (32768, 320, 320, ops.UR_ACT_NONE, "lin64"), (32768, 320, 960, ops.UR_ACT_NONE, "qkv64")]:
    x = torch.randn(8, M // 8, K, device=dev).to(torch.bfloat16)
    w = (torch.randn(N, K, device=dev) * K ** -0.5).to(torch.bfloat16)
    b = torch.randn(N, device=dev)
    bn = ops.pick_bn(N, True) if act == ops.UR_ACT_GEGLU else 0
    if act == ops.UR_ACT_GEGLU:
        w, b = ops.pack_gated_weight(w, b, bn)
    for _ in range(3):
        ops.conv_gemm(x, w, N, bias=b, act=act, bn=bn)
    ts = []
    for _ in range(10):
        if not NOFLUSH:
            flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        ops.conv_gemm(x, w, N, bias=b, act=act, bn=bn)
        e.record()
        torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
    ts.sort()
    t = ts[len(ts) // 2] * 1e-3
    fl = 2.0 * M * K * N
    print("%-8s M=%d K=%d N=%d  %8.1f us  %7.1f TF/s" % (name, M, K, N, t * 1e6, fl / t / 1e12), flush=True)
