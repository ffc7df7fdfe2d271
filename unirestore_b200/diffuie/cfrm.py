"""CFRM building block ``AdaNAFV2`` -- reference cfrm.py:12-54.

    t = conv_in(x)                   c -> 4c                 tcgen05 GEMM
    t = GroupNorm(16)(t)                                      epilogue statistics + ur_norm_apply
    u = gelu(group_conv3x3(t))       16 groups                grouped implicit GEMM (+GELU epilogue)
    u = u * intra(GAP(u)) ; u = u * inter(GAP(u)) per group   epilogue sums + ur_adanaf_scales + ur_scale_channels
    y = x + pwconv(u)                4c -> c                  GEMM (+residual)
    out = NAFBlock(y)
Groups narrower than the 64-channel K block of the GEMM kernel (c = 128 -> 32 channels per group) are packed
pairwise into block-diagonal 64-channel groups at weight-pack time.
"""
import torch
import torch.nn as nn

from .. import ops
from .layout import to_nchw, to_nhwc
from .nafnet_arch import NAFBlock
from .sd_blocks import UrModule, _f32, pack_conv


class AdaNAFV2(UrModule):
    def __init__(self, c):
        super().__init__()
        g, wide = 16, 4 * c
        self.c, self.groups, self.wide = c, g, wide
        self.conv_in = nn.Conv2d(c, wide, 1)
        self.group_norm = nn.GroupNorm(g, wide)
        self.group_conv = nn.Conv2d(wide, wide, 3, padding=1, groups=g)
        self.act = nn.GELU()
        self.intra_group_attn = nn.Sequential(nn.AdaptiveAvgPool2d(1), nn.Conv2d(wide, wide, 1, groups=g))
        # index 2 of the reference Sequential is a parameter-free einops Rearrange (cfrm.py:29-33)
        self.inter_group_attn = nn.Sequential(nn.AdaptiveAvgPool2d(1), nn.Conv2d(wide, g, 1))
        self.pwconv = nn.Conv2d(wide, c, 1)
        self.nafblock = NAFBlock(c)

    def _pack(self):
        g, wide = self.groups, self.wide
        cg = wide // g
        p = dict(gn_g=_f32(self.group_norm.weight), gn_b=_f32(self.group_norm.bias), gn_eps=self.group_norm.eps)
        p["wi"], p["bi"] = pack_conv(self.conv_in)
        w = self.group_conv.weight.detach().float()                       # [wide, cg, 3, 3]
        kg = cg
        while kg % 64:                                                    # merge groups into block-diagonal K blocks
            kg *= 2
        m = kg // cg
        if m > 1:
            wd = w.new_zeros(wide, kg, 3, 3)
            for o in range(wide):
                sub = (o // cg) % m
                wd[o, sub * cg:(sub + 1) * cg] = w[o]
            w = wd
        p["wg"], p["kg"] = ops.pack_conv_weight(w), kg
        p["bg"] = _f32(self.group_conv.bias)
        p["bn_g"] = 128 if kg % 128 == 0 else 64
        p["w_intra"] = _f32(self.intra_group_attn[1].weight).reshape(wide, cg)
        p["b_intra"] = _f32(self.intra_group_attn[1].bias)
        p["w_inter"] = _f32(self.inter_group_attn[1].weight).reshape(g, wide)
        p["b_inter"] = _f32(self.inter_group_attn[1].bias)
        p["wp"], p["bp"] = pack_conv(self.pwconv)
        return p

    def run(self, x):
        p = self.pk
        B, H, W, _ = x.shape
        t = ops.conv_gemm(x, p["wi"], self.wide, bias=p["bi"], want_stats=True)     # GroupNorm statistics from the epilogue
        t = ops.group_norm(t, self.groups, p["gn_g"], p["gn_b"], p["gn_eps"])
        u = ops.conv_gemm(t, p["wg"], self.wide, taps=ops.TAPS_3x3, bias=p["bg"], act=ops.UR_ACT_GELU,
                          group_kc=p["kg"], group_nc=p["kg"], bn=p["bn_g"], want_stats=True)
        # the global average pool reads the channel sums the GEMM epilogue accumulated (no pass over u)
        scale = ops.adanaf_scales(u._ur_stats, H * W, self.groups, p["w_intra"], p["b_intra"], p["w_inter"],
                                  p["b_inter"])
        ops.scale_channels_(u, scale)
        y = ops.conv_gemm(u, p["wp"], self.c, bias=p["bp"], residual=x)
        return self.nafblock.run(y)

    def forward(self, inp):
        return to_nchw(self.run(to_nhwc(inp)), inp.dtype)
