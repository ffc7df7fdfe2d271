"""``Controller`` -- the ControlNet / StableSR-style conditioning encoder, reference controller.py:65-220.

Emits one 256-channel control tensor per latent resolution, keyed by WIDTH (controller.py:216):
``{w: [B,256,w,w]}`` for w in {h, h/2, h/4, h/8}.  ``run`` is the bf16 channels-last fast path used by
``DiffUIE``; ``forward`` keeps the reference signature (fp32 NCHW in, dict of NCHW out).
"""
import torch
import torch.nn as nn

from .. import ops
from .layout import to_nchw
from .sd_blocks import (Attention, ResnetBlock2D, TimestepEmbedding, Timesteps, UNetMidBlock2D, UrModule,
                        get_down_block, pack_conv)

stablesr_config = dict(          # controller.py:29-45
    in_channels=4, model_channels=256, out_channels=256, num_res_blocks=2, attention_resolutions=[4, 2, 1],
    dropout=0.0, channel_mult=[1, 1, 2, 2], conv_resample=True, dims=2, use_fp16=False, num_heads=4,
    down_block_types=("AttnDownBlock2D", "AttnDownBlock2D", "AttnDownBlock2D", "DownBlock2D"),
)


class Controller(UrModule):
    def __init__(self, in_channels=4, model_channels=256, out_channels=256, num_res_blocks=2, channel_mult=(1, 1, 2, 2),
                 num_heads=4, down_block_types=("AttnDownBlock2D",) * 3 + ("DownBlock2D",), **_):
        super().__init__()
        ted = model_channels * 4
        self.model_channels = model_channels
        self.time_proj = Timesteps(model_channels, True, 0)
        self.time_embedding = TimestepEmbedding(model_channels, ted)
        self.conv_in = nn.Conv2d(in_channels, model_channels, 3, padding=1)
        self.down_blocks = nn.ModuleList()
        widths, ch = [], model_channels
        for i, t in enumerate(down_block_types):
            cin, ch = ch, model_channels * channel_mult[i]
            self.down_blocks.append(get_down_block(
                t, num_layers=num_res_blocks, in_channels=cin, out_channels=ch, temb_channels=ted,
                add_downsample=i != len(channel_mult) - 1, resnet_eps=1e-5, resnet_groups=32, downsample_padding=1,
                attention_head_dim=ch // num_heads))
            widths.append(ch)
        self.middle_block = UNetMidBlock2D(in_channels=ch, temb_channels=ted, resnet_eps=1e-5, resnet_groups=32,
                                           attention_head_dim=ch // num_heads)
        self.fea_tran = nn.ModuleList([
            ResnetBlock2D(in_channels=w, out_channels=out_channels, temb_channels=ted, groups=32, eps=1e-5)
            for w in widths])
        # zero initialisation of the reference (controller.py:172-185)
        for m in self.modules():
            if isinstance(m, ResnetBlock2D):
                nn.init.zeros_(m.conv2.weight), nn.init.zeros_(m.conv2.bias)
            elif isinstance(m, Attention):
                nn.init.zeros_(m.to_out[0].weight), nn.init.zeros_(m.to_out[0].bias)

    def _pack(self):
        w, b = pack_conv(self.conv_in, pad_cin=8)
        return dict(w_in=w, b_in=b)

    def time_embed(self, timesteps):
        return self.time_embedding(self.time_proj(timesteps))

    def run(self, z8, emb):
        """z8: bf16 [B,h,w,8] (latent padded to 8 channels); emb fp32 [1 or B, 4*model_channels]."""
        h = ops.conv_gemm(z8, self.pk["w_in"], self.model_channels, taps=ops.TAPS_3x3, bias=self.pk["b_in"],
                          want_stats=True)
        taps = []
        for blk in self.down_blocks:
            h, outs = blk.run(h, emb)
            taps.append(outs[-2])                                  # controller.py:205
        taps[-1] = self.middle_block.run(h, emb)                   # controller.py:211
        return {t.shape[2]: self.fea_tran[i].run(t, emb) for i, t in enumerate(taps)}

    def forward(self, x, timesteps):
        timesteps = torch.as_tensor(timesteps, device=x.device).reshape(-1)
        out = self.run(ops.image_to_nhwc8(x.float()), self.time_embed(timesteps))
        return {k: to_nchw(v, x.dtype) for k, v in out.items()}
