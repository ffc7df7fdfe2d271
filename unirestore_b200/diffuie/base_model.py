"""``ControlledUNet`` -- the SD-2.1 UNet walked by hand with SC-Tuner feature injection, reference
base_model.py:14-245 (``control_type == "scedit"``; the SPADE variant :56-92 is out of scope).

    eps = UNet(zt, null prompt, t) with every skip tensor s_i replaced by CSCEAdapter_i(s_i, control[w(s_i)])

The skip concatenation (torch.cat at base_model.py:189,197) is never materialised: GroupNorm statistics are
taken per source, the normalised concat is written once by ur_norm_apply and the 1x1 shortcut reads both
sources through two TMA tensor maps.  Cross-attention K/V of the constant null prompt are computed once.
"""
import os

import torch
import torch.nn as nn

from .. import ops
from .layout import to_nchw, to_nhwc
from .scedit import CSCEAdapter
from .sd_blocks import UNet2DConditionModel, UrModule

_ASSET = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "assets", "sd_null_emb.pt")


class ControlledUNet(UrModule):
    def __init__(self, unet: UNet2DConditionModel, control_type: str, null_embeds=None):
        super().__init__()
        self.unet = unet
        if null_embeds is None:
            null_embeds = torch.load(_ASSET, map_location="cpu")          # base_model.py:24-27
        self.register_buffer("null_embeds", null_embeds)
        self.control_type = control_type
        if control_type == "scedit":
            chans = [320] * 4 + [640] * 3 + [1280] * 5                    # base_model.py:39
            boc = unet.config["block_out_channels"]
            if tuple(boc) != (320, 640, 1280, 1280):                      # reduced test topologies
                chans = [boc[0]] + [c for i, c in enumerate(boc) for _ in range(3 if i < len(boc) - 1 else 2)]
            self.csc_editors = nn.ModuleList([CSCEAdapter(c, c, 256) for c in chans])
        else:
            raise ValueError(f"control_type '{control_type}' not supported")
        self._ctx = None
        self._cross = None
        self._sc_streams = {}
        self.overlap_sc_tuner = os.environ.get("UNIRESTORE_OVERLAP_SCTUNER", "1") == "1"

    def _reset_cache(self):
        self._pk = None
        self._ctx = None

    def context(self):
        """bf16 [1,77,1024] null-prompt embedding (one tensor object, so the K/V caches stay valid)."""
        if self._ctx is None or self._ctx.device != self.null_embeds.device:
            self._ctx = self.null_embeds.detach().to(torch.bfloat16).contiguous()
        return self._ctx

    def begin_forward(self):
        """Drop the cross-attention K/V of the null prompt: they are recomputed (32 small GEMMs) at first use in every
        forward, so a timed forward carries no precomputed activations."""
        if self._cross is None:
            from .sd_blocks import Attention
            self._cross = [m for m in self.unet.modules() if isinstance(m, Attention) and m.is_cross]
        for m in self._cross:
            m._ctx_ref, m._ctx_kv = None, None

    def time_embed(self, timesteps):
        u = self.unet
        return u.time_embedding(u.time_proj(timesteps))                    # base_model.py:104-106

    def run(self, zt8, control, emb):
        """zt8 bf16 [B,h,w,8]; control {width: bf16 [B,w,w,256]}; emb fp32 [1|B,1280] -> eps fp32 [B,h,w,8]."""
        u, ctx = self.unet, self.context()
        x = u.run_conv_in(zt8)
        skips = [x]
        for blk in u.down_blocks:                                          # base_model.py:126-150
            attns = blk.attentions if blk.has_cross_attention else [None] * len(blk.resnets)
            for r, a in zip(blk.resnets, attns):
                x = r.run(x, emb)
                if a is not None:
                    x = a.run(x, ctx)
                skips.append(x)
            if blk.downsamplers is not None:
                x = blk.downsamplers[0].run(x)
                skips.append(x)
        # SC-Tuner (base_model.py:233-238) only touches the skips: the 12 adapters (36 small GEMMs) run on a side stream,
        # deepest skip first (the order the decoder consumes them), under the mid block's 8x8 kernels that leave most
        # SMs idle.  The decoder waits on one event per skip.
        events, old_skips = None, None
        if self.overlap_sc_tuner and x.is_cuda:
            main = torch.cuda.current_stream(x.device)
            side = self._sc_streams.get(x.device)
            if side is None:
                side = self._sc_streams[x.device] = torch.cuda.Stream(device=x.device)
            side.wait_stream(main)
            events = [None] * len(skips)
            # The un-adapted skips (and ``x``, the deepest one) were allocated on the MAIN stream and are read by the
            # adapters on the SIDE stream (GEMM operand and residual).  They must not return to the main-stream pool
            # while the side stream may still read them: keep them referenced until the up blocks have waited on
            # every event (function exit), when the main stream is ordered after all side-stream reads.
            old_skips = list(skips)
            with torch.cuda.stream(side), ops.workspace_role("sct"):      # own split-K workspace on this stream
                for i in reversed(range(len(self.csc_editors))):
                    skips[i] = self.csc_editors[i].run(skips[i], control[skips[i].shape[2]])
                    events[i] = torch.cuda.Event()
                    events[i].record(side)
        x = u.mid_block.resnets[0].run(x, emb)                             # base_model.py:153-160
        x = u.mid_block.resnets[1].run(u.mid_block.attentions[0].run(x, ctx), emb)
        if events is None:
            for i, ed in enumerate(self.csc_editors):                      # base_model.py:233-238
                skips[i] = ed.run(skips[i], control[skips[i].shape[2]])
        for blk in u.up_blocks:                                            # base_model.py:173-203
            attns = blk.attentions if blk.has_cross_attention else [None] * len(blk.resnets)
            for r, a in zip(blk.resnets, attns):
                if events is not None and events[len(skips) - 1] is not None:
                    torch.cuda.current_stream(x.device).wait_event(events[len(skips) - 1])
                x = r.run(x, emb, x2=skips.pop())
                if a is not None:
                    x = a.run(x, ctx)
            if blk.upsamplers is not None:
                x = blk.upsamplers[0].run(x)
        eps = u.run_head(x)                                                # base_model.py:206-208
        del old_skips               # (main has waited on every adapter event above: safe to recycle from here on)
        return eps

    def forward(self, sample, control, timesteps):
        timesteps = torch.as_tensor(timesteps, device=sample.device).reshape(-1)
        ctl = {k: to_nhwc(v) for k, v in control.items()}
        eps = self.run(ops.image_to_nhwc8(sample.float()), ctl, self.time_embed(timesteps))
        return ops.nhwc_to_image(eps, self.unet.config["out_channels"]).to(sample.dtype)
