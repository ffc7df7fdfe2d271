"""One linear / conv shape, a handful of launches (for ncu captures of the persistent GEMM kernel).

usage: python tools/ncu_gemm_one.py M K N [act]      act: none | gelu | geglu"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from unirestore_b200 import ops  # noqa: E402

M, K, N = (int(v) for v in sys.argv[1:4])
act = {"none": ops.UR_ACT_NONE, "gelu": ops.UR_ACT_GELU, "geglu": ops.UR_ACT_GEGLU}[sys.argv[4] if len(sys.argv) > 4 else "none"]
dev = "cuda:0"
x = torch.randn(8, M // 8, K, device=dev).to(torch.bfloat16)
w = (torch.randn(N, K, device=dev) * K ** -0.5).to(torch.bfloat16)
b = torch.randn(N, device=dev)
bn = ops.pick_bn(N, True) if act == ops.UR_ACT_GEGLU else 0
if act == ops.UR_ACT_GEGLU:
    w, b = ops.pack_gated_weight(w, b, bn)
for _ in range(5):
    ops.conv_gemm(x, w, N, bias=b, act=act, bn=bn)
torch.cuda.synchronize()
