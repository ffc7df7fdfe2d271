"""clock64 timeline of attention3_kernel (CTA 0): softmax thread (row 0, half 0) of both query tiles and the MMA warp,
KV tiles 4..7.  python tools/trace_attn3.py [--poly 3]"""
import argparse
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from unirestore_b200 import _cabi, ops  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--poly", type=int, default=3)
a = ap.parse_args()
dev = "cuda:0"
B, h, d, T = 8, 5, 64, 4096
q = torch.randn(B, T, h * d, device=dev).to(torch.bfloat16)
k = torch.randn(B, T, h * d, device=dev).to(torch.bfloat16)
v = torch.randn(B, T, h * d, device=dev).to(torch.bfloat16)
_cabi.lib().ur_debug_set_attention_impl(3)
_cabi.lib().ur_debug_set_attention_poly(a.poly)
for _ in range(2):
    ops.attention(q, k, v, h)
tr = torch.zeros(128, dtype=torch.int64, device=dev)
_cabi.lib().ur_debug_set_attention_trace(C.c_void_p(tr.data_ptr()))
ops.attention(q, k, v, h)
torch.cuda.synchronize()
_cabi.lib().ur_debug_set_attention_trace(C.c_void_p(0))
t = tr.cpu()
t0 = int(t[0])
for x in (0, 1):
    print("softmax thread tile %d, per KV tile: [wait S, S ready, S in regs, max exchanged, O/P free, exps done, arrived]" % x)
    for j in range(4):
        print("   ", " ".join("%7d" % (int(z) - t0) for z in t[(x * 4 + j) * 8:(x * 4 + j) * 8 + 7]))
for x in (0, 1):
    print("MMA warp tile %d, per KV tile: [wait s_empty, got, S(j+1) issued, P ready, PV issued]" % x)
    for j in range(4):
        print("   ", " ".join("%7d" % (int(z) - t0) for z in t[64 + (x * 4 + j) * 8:64 + (x * 4 + j) * 8 + 5]))
