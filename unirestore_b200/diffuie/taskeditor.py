"""TFA ``TaskFeatureAdapter`` -- reference taskeditor.py:10-108 (wired at autoencoder.py:117-126).

On bf16 channels-last tensors (``x [B,s,s,c_out]``, ``skip [B,s,s,c]``) and an fp32 prompt ``cond [B,T,D=c]``:
    sn   = InstanceNorm(skip)                      shared by the three gate branches (affine=False): stats + apply
    h    = gelu(conv3x3_{f,i,c}(sn))   c -> 3c     ONE implicit GEMM, the three first convs stacked along N
    pool = GAP(conv3x3_{f,i,c}(h))     3 groups    ONE grouped implicit GEMM (channel sums from its epilogue)
    f,i  = softmax(pool_f), softmax(pool_i); cval = tanh(pool_c); cond' = f*cond + i*cval
    o    = tanh(out_gate(cond')); cond_next = gelu(prompt_trans(cond'))          ur_tfa_gates (one tiny kernel)
    skip = skip + t_gate2(o * t_gate1(skip))       two GEMMs (per-image channel scale / residual epilogues)
    x    = x + conv_out(cat[x, skip])              two-source GEMM (+residual); the concat is never materialised
"""
import torch
import torch.nn as nn

from .. import ops
from .layout import to_nchw, to_nhwc
from .sd_blocks import UrModule, _f32, pack_conv


class TaskFeatureAdapter(UrModule):
    def __init__(self, c_out=512, c_skip=256, prompt_len=1, last_layer=False):
        super().__init__()
        d, hid = c_skip, c_skip * prompt_len
        self.c_out, self.c_skip = c_out, c_skip
        self.prompt_len, self.prompt_dim, self.hidden, self.last_layer = prompt_len, d, hid, last_layer
        self.t_gate1 = nn.Conv2d(c_skip, d, 1)
        self.t_gate2 = nn.Conv2d(d, c_skip, 1)
        self.conv_out = nn.Conv2d(c_skip + c_out, c_out, 1)

        def branch(tanh):
            mods = [nn.InstanceNorm2d(c_skip), nn.Conv2d(c_skip, c_skip, 3, padding=1), nn.GELU(),
                    nn.Conv2d(c_skip, hid, 3, padding=1), nn.AdaptiveAvgPool2d(1)]
            return nn.Sequential(*(mods + ([nn.Tanh()] if tanh else [])))

        self.filter_gate, self.info_gate, self.content_trans = branch(False), branch(False), branch(True)
        self.out_gate = nn.Sequential(nn.Linear(hid, d), nn.Tanh())
        if not last_layer:
            self.prompt_trans = nn.Sequential(nn.Linear(d, d // 2), nn.GELU())

    def _pack(self):
        br = (self.filter_gate, self.info_gate, self.content_trans)
        p = dict(in_eps=self.filter_gate[0].eps)
        p["w_a"] = torch.cat([ops.pack_conv_weight(b[1].weight.detach()) for b in br], 0).contiguous()
        p["b_a"] = torch.cat([_f32(b[1].bias) for b in br])
        p["w_b"] = torch.cat([ops.pack_conv_weight(b[3].weight.detach()) for b in br], 0).contiguous()
        p["b_b"] = torch.cat([_f32(b[3].bias) for b in br])
        p["w_og"], p["b_og"] = _f32(self.out_gate[0].weight), _f32(self.out_gate[0].bias)
        if not self.last_layer:
            p["w_pt"], p["b_pt"] = _f32(self.prompt_trans[0].weight), _f32(self.prompt_trans[0].bias)
        p["w_t1"], p["b_t1"] = pack_conv(self.t_gate1)
        p["w_t2"], p["b_t2"] = pack_conv(self.t_gate2)
        p["w_co"], p["b_co"] = pack_conv(self.conv_out)
        hid = self.hidden
        p["bn_b"] = next((bn for bn in (160, 128, 64) if hid % bn == 0), 0) if self.c_skip % 64 == 0 else 0
        return p

    def run(self, x, skip, cond):
        """-> (x', cond_next or None); cond fp32 [B, T, D]."""
        p, c, hid = self.pk, self.c_skip, self.hidden
        B, H, W, _ = skip.shape
        sn = ops.norm_apply(skip, ops.chan_stats(skip), c, None, None, p["in_eps"])
        h = ops.conv_gemm(sn, p["w_a"], 3 * c, taps=ops.TAPS_3x3, bias=p["b_a"], act=ops.UR_ACT_GELU)
        if p["bn_b"]:
            g = ops.conv_gemm(h, p["w_b"], 3 * hid, taps=ops.TAPS_3x3, bias=p["b_b"], group_kc=c, group_nc=hid,
                              bn=p["bn_b"], want_stats=True)      # channel sums for the pooled gates from the epilogue
        else:       # narrow (test) configurations: one launch per branch on channel-slice views
            g = torch.empty((B, H, W, 3 * hid), device=x.device, dtype=torch.bfloat16)
            for i in range(3):
                ops.conv_gemm(h[..., i * c:(i + 1) * c], p["w_b"][i * hid:(i + 1) * hid], hid, taps=ops.TAPS_3x3,
                              bias=p["b_b"][i * hid:(i + 1) * hid].contiguous(), out=g[..., i * hid:(i + 1) * hid])
        gst = getattr(g, "_ur_stats", None)
        o, cond_next = ops.tfa_gates(gst if gst is not None else ops.chan_stats(g), H * W, cond.float().contiguous(), p["w_og"], p["b_og"],
                                     p.get("w_pt"), p.get("b_pt"))
        tg = ops.conv_gemm(skip, p["w_t1"], self.prompt_dim, bias=p["b_t1"], chscale=o)
        skip2 = ops.conv_gemm(tg, p["w_t2"], c, bias=p["b_t2"], residual=skip)
        x = ops.conv_gemm(x, p["w_co"], self.c_out, x2=skip2, bias=p["b_co"], residual=x)
        return x, cond_next

    def forward(self, x, skip, condition):
        y, cond = self.run(to_nhwc(x), to_nhwc(skip), condition)
        return to_nchw(y, x.dtype), cond
