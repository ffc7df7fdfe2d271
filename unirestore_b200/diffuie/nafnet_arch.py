"""``NAFBlock`` of the CFRM latent-feature restorer -- reference nafnet_arch.py:28-131 (SimpleGate :22-25).

Per block, on bf16 channels-last ``x [B,H,W,c]``:
    n1 = LayerNorm2d(x)                         ur_layernorm (per-pixel LN over c == timm LayerNorm2d)
    t  = conv1(n1)               c -> 2c        tcgen05 GEMM
    g  = SimpleGate(dwconv3x3(t)) ; pool(g)     ur_dwconv3x3_gate (depthwise + gate + GAP sums in one pass)
    s  = sca(pool)               [B,c]          ur_small_linear
    y  = x + conv3(g * s) * beta                ur_scale_channels + GEMM (chscale = beta, residual = x)
    n2 = LayerNorm2d(y)
    u  = SimpleGate(conv4(n2))   c -> 2c -> c   GEMM with the gate fused in the epilogue (paired-column packing)
    out = y + conv5(u) * gamma                  GEMM (chscale = gamma, residual = y)
"""
import torch
import torch.nn as nn

from .. import ops
from .layout import to_nchw, to_nhwc
from .sd_blocks import UrModule, _f32, pack_conv


class LayerNorm2d(nn.LayerNorm):
    """Parameter container for ``timm.layers.LayerNorm2d`` (eps 1e-6); computed by ur_layernorm."""

    def __init__(self, num_channels, eps=1e-6):
        super().__init__(num_channels, eps=eps)


class SimpleGate(nn.Module):
    pass


class NAFBlock(UrModule):
    def __init__(self, c, DW_Expand=2, FFN_Expand=2, drop_out_rate=0.0):
        super().__init__()
        if DW_Expand != 2 or FFN_Expand != 2:
            raise ValueError("only the reference configuration (DW_Expand = FFN_Expand = 2) is implemented")
        self.c = c
        self.conv1 = nn.Conv2d(c, 2 * c, 1)
        self.conv2 = nn.Conv2d(2 * c, 2 * c, 3, padding=1, groups=2 * c)
        self.conv3 = nn.Conv2d(c, c, 1)
        self.sca = nn.Sequential(nn.AdaptiveAvgPool2d(1), nn.Conv2d(c, c, 1))
        self.sg = SimpleGate()
        self.conv4 = nn.Conv2d(c, 2 * c, 1)
        self.conv5 = nn.Conv2d(c, c, 1)
        self.norm1 = LayerNorm2d(c)
        self.norm2 = LayerNorm2d(c)
        self.beta = nn.Parameter(torch.zeros((1, c, 1, 1)), requires_grad=True)
        self.gamma = nn.Parameter(torch.zeros((1, c, 1, 1)), requires_grad=True)

    def _pack(self):
        c = self.c
        p = dict(n1=(_f32(self.norm1.weight), _f32(self.norm1.bias), self.norm1.eps),
                 n2=(_f32(self.norm2.weight), _f32(self.norm2.bias), self.norm2.eps),
                 beta=_f32(self.beta).reshape(c), gamma=_f32(self.gamma).reshape(c))
        p["w1"], p["b1"] = pack_conv(self.conv1)
        p["w2"], p["b2"] = _f32(self.conv2.weight).reshape(2 * c, 9), _f32(self.conv2.bias)
        p["w3"], p["b3"] = pack_conv(self.conv3)
        p["wsca"], p["bsca"] = _f32(self.sca[1].weight).reshape(c, c), _f32(self.sca[1].bias)
        p["bn4"] = ops.pick_bn(2 * c, True)
        w4, b4 = pack_conv(self.conv4)
        p["w4"], p["b4"] = ops.pack_gated_weight(w4, b4, p["bn4"])
        p["w5"], p["b5"] = pack_conv(self.conv5)
        return p

    def run(self, x):
        p, c = self.pk, self.c
        B, H, W, _ = x.shape
        t = ops.conv_gemm(ops.layernorm(x, *p["n1"]), p["w1"], 2 * c, bias=p["b1"])
        g, pooled = ops.dwconv3x3_gate(t, p["w2"], p["b2"])
        s = ops.small_linear(pooled, p["wsca"], p["bsca"], stats_scale=1.0 / (H * W))
        ops.scale_channels_(g, s)
        y = ops.conv_gemm(g, p["w3"], c, bias=p["b3"], chscale=p["beta"], residual=x)
        u = ops.conv_gemm(ops.layernorm(y, *p["n2"]), p["w4"], 2 * c, bias=p["b4"], act=ops.UR_ACT_GATE, bn=p["bn4"])
        return ops.conv_gemm(u, p["w5"], c, bias=p["b5"], chscale=p["gamma"], residual=y)

    def forward(self, inp):
        return to_nchw(self.run(to_nhwc(inp)), inp.dtype)
