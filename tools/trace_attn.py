"""Timeline (clock64) of one softmax thread and the MMA thread of CTA (0,0,0) of ur_attention, KV tiles 4..7."""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from unirestore_b200 import _cabi, ops  # noqa: E402

dev = "cuda:0"
B, h, d, T = 8, 5, 64, 4096
qkv = torch.randn(B, T, 3 * h * d, device=dev).to(torch.bfloat16)
C_ = h * d
q, k, v = qkv[..., :C_], qkv[..., C_:2 * C_], qkv[..., 2 * C_:]
for _ in range(2):
    ops.attention(q, k, v, h)
tr = torch.zeros(64, dtype=torch.int64, device=dev)
_cabi.lib().ur_debug_set_attention_trace(C.c_void_p(tr.data_ptr()))
ops.attention(q, k, v, h)
torch.cuda.synchronize()
_cabi.lib().ur_debug_set_attention_trace(C.c_void_p(0))
t = tr.cpu()
t0 = int(t[0])
print("softmax thread, per KV tile: [wait S start, S ready, pass1 done, T ready, O updated, pass2 done, arrived]")
for j in range(4):
    print("   ", " ".join("%7d" % (int(x) - t0) for x in t[j * 8:j * 8 + 7]))
print("MMA thread, per KV tile: [start, K+S-buffer ready, S issued, P ready, PV issued]")
for j in range(4):
    print("   ", " ".join("%7d" % (int(x) - t0) for x in t[32 + j * 8:32 + j * 8 + 5]))
