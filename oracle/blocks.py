"""Restatement of the ``diffusers`` leaf blocks the UniRestore hot path calls.

TEST INFRASTRUCTURE (oracle) -- never imported by the product path.

The arithmetic of the hot path lives mostly in the un-vendored third-party package
``diffusers`` (reference requirements.txt:14-15, comment-pinned ``==0.29.0``), which is
absent from /root/reference and from this image.  This file restates the published
algorithm of every diffusers symbol the reference touches (SURVEY.md section 2.2 and
Appendix A), with the diffusers state_dict key names, in plain fp32 PyTorch.
PARITY UNPINNED for these leaves (no diffusers golden exists offline); the wiring
around them is pinned by executing the reference's own files (oracle/make_golden.py).

Reference call sites anchored on:
  ResnetBlock2D        base_model.py:54, controller.py:161-170
  Transformer2DModel   base_model.py:138,159,191
  Attention            controller.py:183-185 (zero-init), VAE mid block
  get_down_block       controller.py:101-125
  UNetMidBlock2D       controller.py:133-141
  Timesteps/TimestepEmbedding  controller.py:86-89, base_model.py:104-106
  UNet2DConditionModel internals  base_model.py:94-209
  AutoencoderKL internals         autoencoder.py:11-72,132-176
"""
from __future__ import annotations

import math
from types import SimpleNamespace

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import sd_turbo_config as CFG


# --------------------------------------------------------------------------- embeddings
class Timesteps(nn.Module):
    """Sinusoidal embedding; ``flip_sin_to_cos=True`` gives [cos, sin] (Appendix A.5)."""

    def __init__(self, num_channels: int, flip_sin_to_cos: bool, downscale_freq_shift: float):
        super().__init__()
        self.num_channels = num_channels
        self.flip_sin_to_cos = flip_sin_to_cos
        self.downscale_freq_shift = downscale_freq_shift

    def forward(self, timesteps):
        half = self.num_channels // 2
        exponent = -math.log(10000) * torch.arange(half, dtype=torch.float32, device=timesteps.device)
        exponent = exponent / (half - self.downscale_freq_shift)
        emb = timesteps[:, None].float() * torch.exp(exponent)[None, :]
        emb = torch.cat([torch.sin(emb), torch.cos(emb)], dim=-1)
        if self.flip_sin_to_cos:
            emb = torch.cat([emb[:, half:], emb[:, :half]], dim=-1)
        return emb


class TimestepEmbedding(nn.Module):
    def __init__(self, in_channels: int, time_embed_dim: int, act_fn: str = "silu"):
        super().__init__()
        self.linear_1 = nn.Linear(in_channels, time_embed_dim)
        self.act = nn.SiLU()
        self.linear_2 = nn.Linear(time_embed_dim, time_embed_dim)

    def forward(self, sample):
        return self.linear_2(self.act(self.linear_1(sample)))


# --------------------------------------------------------------------------- resnet / resample
class ResnetBlock2D(nn.Module):
    """GN -> SiLU -> conv3x3 -> (+Linear(SiLU(temb))) -> GN -> SiLU -> conv3x3 -> (+shortcut).

    Additive temb only (``time_embedding_norm == "default"``; see reference comment
    base_model.py:64-67)."""

    def __init__(self, *, in_channels, out_channels=None, dropout=0.0, temb_channels=512,
                 groups=32, eps=1e-6, non_linearity="silu", output_scale_factor=1.0, **_):
        super().__init__()
        out_channels = in_channels if out_channels is None else out_channels
        self.in_channels, self.out_channels = in_channels, out_channels
        self.output_scale_factor = output_scale_factor
        self.norm1 = nn.GroupNorm(groups, in_channels, eps=eps, affine=True)
        self.conv1 = nn.Conv2d(in_channels, out_channels, 3, stride=1, padding=1)
        self.time_emb_proj = nn.Linear(temb_channels, out_channels) if temb_channels is not None else None
        self.norm2 = nn.GroupNorm(groups, out_channels, eps=eps, affine=True)
        self.dropout = nn.Dropout(dropout)
        self.conv2 = nn.Conv2d(out_channels, out_channels, 3, stride=1, padding=1)
        self.nonlinearity = nn.SiLU()
        self.conv_shortcut = nn.Conv2d(in_channels, out_channels, 1) if in_channels != out_channels else None

    def forward(self, input_tensor, temb=None, *args, **kwargs):
        h = self.conv1(self.nonlinearity(self.norm1(input_tensor)))
        if self.time_emb_proj is not None and temb is not None:
            h = h + self.time_emb_proj(self.nonlinearity(temb))[:, :, None, None]
        h = self.conv2(self.dropout(self.nonlinearity(self.norm2(h))))
        if self.conv_shortcut is not None:
            input_tensor = self.conv_shortcut(input_tensor)
        return (input_tensor + h) / self.output_scale_factor


class Downsample2D(nn.Module):
    """conv3x3 stride 2; ``padding=0`` (VAE) pads right/bottom by one zero first (A.2)."""

    def __init__(self, channels, use_conv=True, out_channels=None, padding=1, name="conv"):
        super().__init__()
        self.padding = padding
        self.conv = nn.Conv2d(channels, out_channels or channels, 3, stride=2, padding=padding)

    def forward(self, x, *args, **kwargs):
        if self.padding == 0:
            x = F.pad(x, (0, 1, 0, 1), mode="constant", value=0)
        return self.conv(x)


class Upsample2D(nn.Module):
    """nearest x2 then conv3x3 p1 (A.2)."""

    def __init__(self, channels, use_conv=True, out_channels=None):
        super().__init__()
        self.conv = nn.Conv2d(channels, out_channels or channels, 3, padding=1)

    def forward(self, x, *args, **kwargs):
        return self.conv(F.interpolate(x, scale_factor=2.0, mode="nearest"))


# --------------------------------------------------------------------------- attention
class Attention(nn.Module):
    """diffusers ``Attention``: spatial block (with GroupNorm + residual) or token block."""

    def __init__(self, query_dim, cross_attention_dim=None, heads=8, dim_head=64, bias=False,
                 out_bias=True, norm_num_groups=None, eps=1e-5, residual_connection=False,
                 rescale_output_factor=1.0, upcast_attention=False, **_):
        super().__init__()
        inner = heads * dim_head
        self.heads, self.dim_head = heads, dim_head
        self.residual_connection = residual_connection
        self.rescale_output_factor = rescale_output_factor
        self.upcast_attention = upcast_attention
        self.group_norm = (nn.GroupNorm(norm_num_groups, query_dim, eps=eps, affine=True)
                           if norm_num_groups is not None else None)
        kv_dim = query_dim if cross_attention_dim is None else cross_attention_dim
        self.to_q = nn.Linear(query_dim, inner, bias=bias)
        self.to_k = nn.Linear(kv_dim, inner, bias=bias)
        self.to_v = nn.Linear(kv_dim, inner, bias=bias)
        self.to_out = nn.ModuleList([nn.Linear(inner, query_dim, bias=out_bias), nn.Dropout(0.0)])

    def forward(self, hidden_states, encoder_hidden_states=None, **_):
        residual = hidden_states
        spatial = hidden_states.ndim == 4
        if spatial:
            b, c, h, w = hidden_states.shape
            hidden_states = hidden_states.view(b, c, h * w).transpose(1, 2)
        if self.group_norm is not None:
            hidden_states = self.group_norm(hidden_states.transpose(1, 2)).transpose(1, 2)
        ctx = hidden_states if encoder_hidden_states is None else encoder_hidden_states
        q, k, v = self.to_q(hidden_states), self.to_k(ctx), self.to_v(ctx)
        bsz = q.shape[0]

        def split(t):
            return t.view(bsz, -1, self.heads, self.dim_head).transpose(1, 2)

        q, k, v = split(q), split(k), split(v)
        if self.upcast_attention:
            q, k = q.float(), k.float()
        o = F.scaled_dot_product_attention(q, k, v, dropout_p=0.0, is_causal=False).to(v.dtype)
        o = o.transpose(1, 2).reshape(bsz, -1, self.heads * self.dim_head)
        o = self.to_out[1](self.to_out[0](o))
        if spatial:
            o = o.transpose(-1, -2).reshape(b, c, h, w)
        if self.residual_connection:
            o = o + residual
        return o / self.rescale_output_factor


class GEGLU(nn.Module):
    def __init__(self, dim_in, dim_out):
        super().__init__()
        self.proj = nn.Linear(dim_in, dim_out * 2)

    def forward(self, x):
        a, gate = self.proj(x).chunk(2, dim=-1)
        return a * F.gelu(gate)          # exact-erf GELU


class FeedForward(nn.Module):
    def __init__(self, dim, mult=4):
        super().__init__()
        self.net = nn.ModuleList([GEGLU(dim, dim * mult), nn.Dropout(0.0), nn.Linear(dim * mult, dim)])

    def forward(self, x):
        for m in self.net:
            x = m(x)
        return x


class BasicTransformerBlock(nn.Module):
    def __init__(self, dim, heads, dim_head, cross_attention_dim, upcast_attention=False):
        super().__init__()
        self.norm1 = nn.LayerNorm(dim, eps=1e-5)
        self.attn1 = Attention(dim, None, heads, dim_head, bias=False, upcast_attention=upcast_attention)
        self.norm2 = nn.LayerNorm(dim, eps=1e-5)
        self.attn2 = Attention(dim, cross_attention_dim, heads, dim_head, bias=False,
                               upcast_attention=upcast_attention)
        self.norm3 = nn.LayerNorm(dim, eps=1e-5)
        self.ff = FeedForward(dim)

    def forward(self, x, encoder_hidden_states=None):
        x = x + self.attn1(self.norm1(x))
        x = x + self.attn2(self.norm2(x), encoder_hidden_states)
        x = x + self.ff(self.norm3(x))
        return x


class Transformer2DModel(nn.Module):
    """Continuous-input transformer with ``use_linear_projection=True`` (Appendix A.4)."""

    def __init__(self, num_attention_heads, attention_head_dim, in_channels, cross_attention_dim,
                 norm_num_groups=32, upcast_attention=False, num_layers=1, **_):
        super().__init__()
        inner = num_attention_heads * attention_head_dim
        self.norm = nn.GroupNorm(norm_num_groups, in_channels, eps=1e-6, affine=True)
        self.proj_in = nn.Linear(in_channels, inner)
        self.transformer_blocks = nn.ModuleList([
            BasicTransformerBlock(inner, num_attention_heads, attention_head_dim, cross_attention_dim,
                                  upcast_attention) for _ in range(num_layers)])
        self.proj_out = nn.Linear(inner, in_channels)

    def forward(self, hidden_states, encoder_hidden_states=None, return_dict=True, **_):
        b, c, h, w = hidden_states.shape
        residual = hidden_states
        x = self.norm(hidden_states)
        x = x.permute(0, 2, 3, 1).reshape(b, h * w, c)
        x = self.proj_in(x)
        for blk in self.transformer_blocks:
            x = blk(x, encoder_hidden_states)
        x = self.proj_out(x)
        x = x.reshape(b, h, w, c).permute(0, 3, 1, 2).contiguous()
        out = x + residual
        return (out,) if not return_dict else SimpleNamespace(sample=out)


# --------------------------------------------------------------------------- UNet blocks
class DownBlock2D(nn.Module):
    has_cross_attention = False

    def __init__(self, *, num_layers, in_channels, out_channels, temb_channels, add_downsample,
                 resnet_eps, resnet_groups=32, downsample_padding=1, **_):
        super().__init__()
        self.resnets = nn.ModuleList([
            ResnetBlock2D(in_channels=in_channels if i == 0 else out_channels, out_channels=out_channels,
                          temb_channels=temb_channels, eps=resnet_eps, groups=resnet_groups)
            for i in range(num_layers)])
        self.downsamplers = (nn.ModuleList([Downsample2D(out_channels, True, out_channels, downsample_padding, "op")])
                             if add_downsample else None)

    def forward(self, hidden_states, temb=None, *args, **kwargs):
        outs = ()
        for r in self.resnets:
            hidden_states = r(hidden_states, temb)
            outs += (hidden_states,)
        if self.downsamplers is not None:
            for d in self.downsamplers:
                hidden_states = d(hidden_states)
            outs += (hidden_states,)
        return hidden_states, outs


class AttnDownBlock2D(nn.Module):
    """resnet + spatial self-attention per layer; returns (h, (attn0, attn1[, down])) (A.7)."""

    def __init__(self, *, num_layers, in_channels, out_channels, temb_channels, add_downsample,
                 resnet_eps, resnet_groups=32, attention_head_dim=1, downsample_padding=1, **_):
        super().__init__()
        self.resnets = nn.ModuleList([
            ResnetBlock2D(in_channels=in_channels if i == 0 else out_channels, out_channels=out_channels,
                          temb_channels=temb_channels, eps=resnet_eps, groups=resnet_groups)
            for i in range(num_layers)])
        self.attentions = nn.ModuleList([
            Attention(out_channels, heads=out_channels // attention_head_dim, dim_head=attention_head_dim,
                      rescale_output_factor=1.0, eps=resnet_eps, norm_num_groups=resnet_groups,
                      residual_connection=True, bias=True)
            for _ in range(num_layers)])
        self.downsamplers = (nn.ModuleList([Downsample2D(out_channels, True, out_channels, downsample_padding, "op")])
                             if add_downsample else None)

    def forward(self, hidden_states, temb=None, *args, **kwargs):
        outs = ()
        for r, a in zip(self.resnets, self.attentions):
            hidden_states = a(r(hidden_states, temb))
            outs += (hidden_states,)
        if self.downsamplers is not None:
            for d in self.downsamplers:
                hidden_states = d(hidden_states)
            outs += (hidden_states,)
        return hidden_states, outs


class CrossAttnDownBlock2D(nn.Module):
    has_cross_attention = True

    def __init__(self, *, num_layers, in_channels, out_channels, temb_channels, add_downsample,
                 resnet_eps, resnet_groups=32, num_attention_heads=1, cross_attention_dim=1024,
                 downsample_padding=1, upcast_attention=False, **_):
        super().__init__()
        self.resnets = nn.ModuleList([
            ResnetBlock2D(in_channels=in_channels if i == 0 else out_channels, out_channels=out_channels,
                          temb_channels=temb_channels, eps=resnet_eps, groups=resnet_groups)
            for i in range(num_layers)])
        self.attentions = nn.ModuleList([
            Transformer2DModel(num_attention_heads, out_channels // num_attention_heads, out_channels,
                               cross_attention_dim, resnet_groups, upcast_attention)
            for _ in range(num_layers)])
        self.downsamplers = (nn.ModuleList([Downsample2D(out_channels, True, out_channels, downsample_padding, "op")])
                             if add_downsample else None)

    def forward(self, hidden_states, temb=None, encoder_hidden_states=None, **_):
        outs = ()
        for r, a in zip(self.resnets, self.attentions):
            hidden_states = a(r(hidden_states, temb), encoder_hidden_states, return_dict=False)[0]
            outs += (hidden_states,)
        if self.downsamplers is not None:
            for d in self.downsamplers:
                hidden_states = d(hidden_states)
            outs += (hidden_states,)
        return hidden_states, outs


def get_down_block(down_block_type, *, num_layers, in_channels, out_channels, temb_channels, add_downsample,
                   resnet_eps, resnet_groups=32, downsample_padding=1, attention_head_dim=None,
                   cross_attention_dim=None, num_attention_heads=None, upcast_attention=False, **_):
    """diffusers ``get_down_block`` for the three block types the reference instantiates
    (controller.py:101-125)."""
    common = dict(num_layers=num_layers, in_channels=in_channels, out_channels=out_channels,
                  temb_channels=temb_channels, add_downsample=add_downsample, resnet_eps=resnet_eps,
                  resnet_groups=resnet_groups, downsample_padding=downsample_padding)
    if down_block_type == "DownBlock2D":
        return DownBlock2D(**common)
    if down_block_type == "AttnDownBlock2D":
        return AttnDownBlock2D(attention_head_dim=attention_head_dim, **common)
    if down_block_type == "CrossAttnDownBlock2D":
        return CrossAttnDownBlock2D(num_attention_heads=num_attention_heads,
                                    cross_attention_dim=cross_attention_dim,
                                    upcast_attention=upcast_attention, **common)
    raise ValueError(f"{down_block_type} does not exist.")


class UNetMidBlock2D(nn.Module):
    """res, spatial-attn, res (Controller mid: controller.py:133-141; VAE mid: A.6)."""

    def __init__(self, *, in_channels, temb_channels, resnet_eps=1e-6, resnet_groups=32,
                 attention_head_dim=1, add_attention=True, **_):
        super().__init__()
        mk = lambda: ResnetBlock2D(in_channels=in_channels, out_channels=in_channels,
                                   temb_channels=temb_channels, eps=resnet_eps, groups=resnet_groups)
        self.resnets = nn.ModuleList([mk(), mk()])
        self.attentions = nn.ModuleList([
            Attention(in_channels, heads=in_channels // attention_head_dim, dim_head=attention_head_dim,
                      rescale_output_factor=1.0, eps=resnet_eps, norm_num_groups=resnet_groups,
                      residual_connection=True, bias=True)])

    def forward(self, hidden_states, temb=None):
        hidden_states = self.resnets[0](hidden_states, temb)
        for a, r in zip(self.attentions, self.resnets[1:]):
            hidden_states = r(a(hidden_states), temb)
        return hidden_states


class UNetMidBlock2DCrossAttn(nn.Module):
    has_cross_attention = True

    def __init__(self, *, in_channels, temb_channels, resnet_eps=1e-5, resnet_groups=32,
                 num_attention_heads=1, cross_attention_dim=1024, upcast_attention=False, **_):
        super().__init__()
        mk = lambda: ResnetBlock2D(in_channels=in_channels, out_channels=in_channels,
                                   temb_channels=temb_channels, eps=resnet_eps, groups=resnet_groups)
        self.resnets = nn.ModuleList([mk(), mk()])
        self.attentions = nn.ModuleList([
            Transformer2DModel(num_attention_heads, in_channels // num_attention_heads, in_channels,
                               cross_attention_dim, resnet_groups, upcast_attention)])

    def forward(self, hidden_states, temb=None, encoder_hidden_states=None, **_):
        hidden_states = self.resnets[0](hidden_states, temb)
        for a, r in zip(self.attentions, self.resnets[1:]):
            hidden_states = a(hidden_states, encoder_hidden_states, return_dict=False)[0]
            hidden_states = r(hidden_states, temb)
        return hidden_states


class _UpBlockBase(nn.Module):
    def _make_resnets(self, num_layers, in_channels, out_channels, prev_output_channel, temb_channels,
                      resnet_eps, resnet_groups):
        rs = []
        for i in range(num_layers):
            skip = in_channels if i == num_layers - 1 else out_channels
            rin = prev_output_channel if i == 0 else out_channels
            rs.append(ResnetBlock2D(in_channels=rin + skip, out_channels=out_channels,
                                    temb_channels=temb_channels, eps=resnet_eps, groups=resnet_groups))
        return nn.ModuleList(rs)


class UpBlock2D(_UpBlockBase):
    has_cross_attention = False

    def __init__(self, *, num_layers, in_channels, out_channels, prev_output_channel, temb_channels,
                 add_upsample, resnet_eps, resnet_groups=32, **_):
        super().__init__()
        self.resnets = self._make_resnets(num_layers, in_channels, out_channels, prev_output_channel,
                                          temb_channels, resnet_eps, resnet_groups)
        self.upsamplers = nn.ModuleList([Upsample2D(out_channels, True, out_channels)]) if add_upsample else None


class CrossAttnUpBlock2D(_UpBlockBase):
    has_cross_attention = True

    def __init__(self, *, num_layers, in_channels, out_channels, prev_output_channel, temb_channels,
                 add_upsample, resnet_eps, resnet_groups=32, num_attention_heads=1,
                 cross_attention_dim=1024, upcast_attention=False, **_):
        super().__init__()
        self.resnets = self._make_resnets(num_layers, in_channels, out_channels, prev_output_channel,
                                          temb_channels, resnet_eps, resnet_groups)
        self.attentions = nn.ModuleList([
            Transformer2DModel(num_attention_heads, out_channels // num_attention_heads, out_channels,
                               cross_attention_dim, resnet_groups, upcast_attention)
            for _ in range(num_layers)])
        self.upsamplers = nn.ModuleList([Upsample2D(out_channels, True, out_channels)]) if add_upsample else None


class UNet2DConditionModel(nn.Module):
    """Container with the sd-turbo UNet topology and diffusers key names (Appendix A.5).

    The reference never calls ``unet.forward``: ``ControlledUNet`` walks the children by
    hand (base_model.py:94-209), so only construction + children are restated."""

    def __init__(self, **overrides):
        super().__init__()
        c = dict(CFG.UNET)
        c.update(overrides)
        self.config = SimpleNamespace(**c)
        boc, heads = c["block_out_channels"], c["num_attention_heads"]
        ted = c["time_embed_dim"]
        self.conv_in = nn.Conv2d(c["in_channels"], boc[0], 3, padding=1)
        self.time_proj = Timesteps(c["time_embed_in"], c["flip_sin_to_cos"], c["freq_shift"])
        self.time_embedding = TimestepEmbedding(c["time_embed_in"], ted)
        self.down_blocks = nn.ModuleList()
        out_ch = boc[0]
        for i, t in enumerate(c["down_block_types"]):
            in_ch, out_ch = out_ch, boc[i]
            self.down_blocks.append(get_down_block(
                t, num_layers=c["layers_per_block"], in_channels=in_ch, out_channels=out_ch,
                temb_channels=ted, add_downsample=i != len(boc) - 1, resnet_eps=c["norm_eps"],
                resnet_groups=c["norm_num_groups"], downsample_padding=1,
                cross_attention_dim=c["cross_attention_dim"], num_attention_heads=heads[i],
                upcast_attention=c["upcast_attention"]))
        self.mid_block = UNetMidBlock2DCrossAttn(
            in_channels=boc[-1], temb_channels=ted, resnet_eps=c["norm_eps"],
            resnet_groups=c["norm_num_groups"], num_attention_heads=heads[-1],
            cross_attention_dim=c["cross_attention_dim"], upcast_attention=c["upcast_attention"])
        self.up_blocks = nn.ModuleList()
        rev, rheads = list(reversed(boc)), list(reversed(heads))
        out_ch = rev[0]
        for i, t in enumerate(c["up_block_types"]):
            prev, out_ch = out_ch, rev[i]
            in_ch = rev[min(i + 1, len(boc) - 1)]
            kw = dict(num_layers=c["layers_per_block"] + 1, in_channels=in_ch, out_channels=out_ch,
                      prev_output_channel=prev, temb_channels=ted, add_upsample=i != len(boc) - 1,
                      resnet_eps=c["norm_eps"], resnet_groups=c["norm_num_groups"])
            if t == "CrossAttnUpBlock2D":
                self.up_blocks.append(CrossAttnUpBlock2D(
                    num_attention_heads=rheads[i], cross_attention_dim=c["cross_attention_dim"],
                    upcast_attention=c["upcast_attention"], **kw))
            else:
                self.up_blocks.append(UpBlock2D(**kw))
        self.conv_norm_out = nn.GroupNorm(c["norm_num_groups"], boc[0], eps=c["norm_eps"])
        self.conv_act = nn.SiLU()
        self.conv_out = nn.Conv2d(boc[0], c["out_channels"], 3, padding=1)

    @classmethod
    def from_pretrained(cls, model_id=None, subfolder=None, **kw):
        """No weights exist offline: returns the sd-turbo topology with default init."""
        return cls()


# --------------------------------------------------------------------------- VAE
class DownEncoderBlock2D(nn.Module):
    def __init__(self, in_channels, out_channels, num_layers, add_downsample, eps, groups):
        super().__init__()
        self.resnets = nn.ModuleList([
            ResnetBlock2D(in_channels=in_channels if i == 0 else out_channels, out_channels=out_channels,
                          temb_channels=None, eps=eps, groups=groups) for i in range(num_layers)])
        self.downsamplers = (nn.ModuleList([Downsample2D(out_channels, True, out_channels, 0, "op")])
                             if add_downsample else None)

    def forward(self, x, *args, **kwargs):
        for r in self.resnets:
            x = r(x, None)
        if self.downsamplers is not None:
            for d in self.downsamplers:
                x = d(x)
        return x


class UpDecoderBlock2D(nn.Module):
    def __init__(self, in_channels, out_channels, num_layers, add_upsample, eps, groups):
        super().__init__()
        self.resnets = nn.ModuleList([
            ResnetBlock2D(in_channels=in_channels if i == 0 else out_channels, out_channels=out_channels,
                          temb_channels=None, eps=eps, groups=groups) for i in range(num_layers)])
        self.upsamplers = nn.ModuleList([Upsample2D(out_channels, True, out_channels)]) if add_upsample else None

    def forward(self, x, temb=None):
        for r in self.resnets:
            x = r(x, temb)
        if self.upsamplers is not None:
            for u in self.upsamplers:
                x = u(x)
        return x


class Encoder(nn.Module):
    def __init__(self, c):
        super().__init__()
        boc, g, eps = c["block_out_channels"], c["norm_num_groups"], c["resnet_eps"]
        self.conv_in = nn.Conv2d(c["in_channels"], boc[0], 3, padding=1)
        self.down_blocks = nn.ModuleList()
        out_ch = boc[0]
        for i in range(len(boc)):
            in_ch, out_ch = out_ch, boc[i]
            self.down_blocks.append(DownEncoderBlock2D(in_ch, out_ch, c["layers_per_block"],
                                                       i != len(boc) - 1, eps, g))
        self.mid_block = UNetMidBlock2D(in_channels=boc[-1], temb_channels=None, resnet_eps=eps,
                                        resnet_groups=g, attention_head_dim=boc[-1])
        self.conv_norm_out = nn.GroupNorm(g, boc[-1], eps=1e-6)
        self.conv_act = nn.SiLU()
        self.conv_out = nn.Conv2d(boc[-1], 2 * c["latent_channels"], 3, padding=1)

    def forward(self, sample):
        sample = self.conv_in(sample)
        for d in self.down_blocks:
            sample = d(sample)
        sample = self.mid_block(sample)
        return self.conv_out(self.conv_act(self.conv_norm_out(sample)))


class Decoder(nn.Module):
    def __init__(self, c):
        super().__init__()
        boc, g, eps = c["block_out_channels"], c["norm_num_groups"], c["resnet_eps"]
        self.conv_in = nn.Conv2d(c["latent_channels"], boc[-1], 3, padding=1)
        self.mid_block = UNetMidBlock2D(in_channels=boc[-1], temb_channels=None, resnet_eps=eps,
                                        resnet_groups=g, attention_head_dim=boc[-1])
        self.up_blocks = nn.ModuleList()
        rev = list(reversed(boc))
        out_ch = rev[0]
        for i in range(len(boc)):
            prev, out_ch = out_ch, rev[i]
            self.up_blocks.append(UpDecoderBlock2D(prev, out_ch, c["layers_per_block"] + 1,
                                                   i != len(boc) - 1, eps, g))
        self.conv_norm_out = nn.GroupNorm(g, boc[0], eps=1e-6)
        self.conv_act = nn.SiLU()
        self.conv_out = nn.Conv2d(boc[0], c["out_channels"], 3, padding=1)

    def forward(self, sample, latent_embeds=None):
        sample = self.conv_in(sample)
        sample = self.mid_block(sample, latent_embeds)
        for u in self.up_blocks:
            sample = u(sample, latent_embeds)
        return self.conv_out(self.conv_act(self.conv_norm_out(sample)))


class DiagonalGaussianDistribution:
    """``sample = mean + exp(0.5*clamp(logvar,-30,20)) * randn(mean.shape, dtype=moments.dtype)`` (A.6).

    ``noise`` may be injected for parity (RNG hazard (1) of SURVEY.md section 8c)."""

    def __init__(self, parameters, noise=None):
        self.parameters = parameters
        self.mean, logvar = torch.chunk(parameters, 2, dim=1)
        self.logvar = torch.clamp(logvar, -30.0, 20.0)
        self.std = torch.exp(0.5 * self.logvar)
        self._noise = noise

    def sample(self, generator=None):
        n = self._noise
        if n is None:
            n = torch.randn(self.mean.shape, generator=generator, device=self.parameters.device,
                            dtype=self.parameters.dtype)
        return self.mean + self.std * n.to(self.mean.dtype)

    def mode(self):
        return self.mean


class AutoencoderKL(nn.Module):
    def __init__(self, **overrides):
        super().__init__()
        c = dict(CFG.VAE)
        c.update(overrides)
        self.config = SimpleNamespace(**c)
        self.encoder = Encoder(c)
        self.decoder = Decoder(c)
        self.quant_conv = nn.Conv2d(2 * c["latent_channels"], 2 * c["latent_channels"], 1)
        self.post_quant_conv = nn.Conv2d(c["latent_channels"], c["latent_channels"], 1)
        self.posterior_noise = None      # parity hook: injected posterior noise

    @classmethod
    def from_pretrained(cls, model_id=None, subfolder=None, **kw):
        return cls()

    def encode(self, x, return_dict=True):
        moments = self.quant_conv(self.encoder(x))
        post = DiagonalGaussianDistribution(moments, self.posterior_noise)
        return (post,) if not return_dict else SimpleNamespace(latent_dist=post)

    def decode(self, z, return_dict=True, generator=None):
        dec = self.decoder(self.post_quant_conv(z))
        return (dec,) if not return_dict else SimpleNamespace(sample=dec)


# --------------------------------------------------------------------------- timm
class LayerNorm2d(nn.LayerNorm):
    """``timm.layers.LayerNorm2d``: LayerNorm over C of an NCHW tensor, eps 1e-6 (A.9)."""

    def __init__(self, num_channels, eps=1e-6, affine=True):
        super().__init__(num_channels, eps=eps, elementwise_affine=affine)

    def forward(self, x):
        x = x.permute(0, 2, 3, 1)
        x = F.layer_norm(x, self.normalized_shape, self.weight, self.bias, self.eps)
        return x.permute(0, 3, 1, 2)
