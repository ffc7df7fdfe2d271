"""SC-Tuner (``CSCEAdapter``) -- reference scedit.py:24-38, applied to the 12 UNet skips at base_model.py:233-238.

    p = proj(condition);  out = tuner(x + p) + p + x,   tuner = 1x1 -> GELU -> 1x1

Three tcgen05 GEMM launches, no other passes over the activation:
    s   = proj(condition) + x                (bias + residual epilogue)
    h   = gelu(tuner.0(s))                   (bias + exact-erf GELU epilogue)
    out = tuner.2(h) + s                     (bias + residual epilogue)
"""
import torch.nn as nn

from .. import ops
from .layout import to_nchw, to_nhwc
from .sd_blocks import UrModule, pack_conv


class CSCEAdapter(UrModule):
    def __init__(self, c_in, c_emb, c_cond):
        super().__init__()
        self.c_in, self.c_emb = c_in, c_emb
        self.proj = nn.Conv2d(c_cond, c_in, 1)
        self.tuner = nn.Sequential(nn.Conv2d(c_in, c_emb, 1), nn.GELU(), nn.Conv2d(c_emb, c_in, 1))

    def _pack(self):
        p = {}
        p["wp"], p["bp"] = pack_conv(self.proj)
        p["w0"], p["b0"] = pack_conv(self.tuner[0])
        p["w2"], p["b2"] = pack_conv(self.tuner[2])
        return p

    def run(self, x, condition):
        p = self.pk
        s = ops.conv_gemm(condition, p["wp"], self.c_in, bias=p["bp"], residual=x)
        h = ops.conv_gemm(s, p["w0"], self.c_emb, bias=p["b0"], act=ops.UR_ACT_GELU)
        return ops.conv_gemm(h, p["w2"], self.c_in, bias=p["b2"], residual=s, want_stats=True)   # skip -> up-block GroupNorm

    def forward(self, x, condition):
        return to_nchw(self.run(to_nhwc(x), to_nhwc(condition)), x.dtype)
