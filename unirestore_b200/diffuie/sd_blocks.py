"""B200-native counterparts of the ``diffusers`` building blocks the UniRestore hot path touches.

The reference imports these from the un-vendored ``diffusers`` package (unifie.py:6-12, base_model.py:5-6,
controller.py:3-11, autoencoder.py:5); parameter names / shapes follow the diffusers state_dict keys
 so SD-Turbo safetensors and UniRestore checkpoints load unchanged.

Every module keeps fp32 ``nn.Parameter``s (inside stock ``nn.Conv2d`` / ``nn.Linear`` / ``nn.GroupNorm``
containers that are NEVER called) and a lazily built cache of packed bf16 K-major weights.  ``run(...)``
is the fast path on bf16 channels-last activations ``[B, H, W, C]``: it only launches the hand-written
kernels behind the C-ABI (``unirestore_b200.ops``); there is no PyTorch-arithmetic fallback.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from .. import ops
from ..ops import TAPS_3x3, TAPS_3x3_NOPAD, UR_ACT_GEGLU


# ------------------------------------------------------------------------------------------------- base
_WEIGHT_GEN = [0]


def weight_generation() -> int:
    """Bumped whenever any module drops its packed weights; CUDA graphs captured under an older generation replay
    kernels that point at freed / stale weight buffers and must be re-captured (``DiffUIE._forward_graphed``)."""
    return _WEIGHT_GEN[0]


class UrModule(nn.Module):
    """nn.Module with a per-module cache of packed device weights (dropped on .to() / load_state_dict)."""

    def __init__(self):
        super().__init__()
        self._pk = None
        self.register_load_state_dict_post_hook(lambda m, _: m.invalidate())

    def invalidate(self):
        _WEIGHT_GEN[0] += 1
        for m in self.modules():
            if isinstance(m, UrModule):
                m._reset_cache()

    def _reset_cache(self):
        self._pk = None

    def _apply(self, fn, *a, **k):
        self.invalidate()
        return super()._apply(fn, *a, **k)

    @property
    def pk(self):
        if self._pk is None:
            with torch.no_grad():
                self._pk = self._pack()
        return self._pk

    def _pack(self):
        return {}


def _f32(t):
    return t.detach().float().contiguous()


def _view_keep_stats(t, shape):
    """``t.view(shape)`` that keeps the epilogue-accumulated GroupNorm statistics attached (views drop attributes)."""
    v = t.view(shape)
    st = getattr(t, "_ur_stats", None)
    if st is not None:
        v._ur_stats = st
    return v


def pack_conv(conv: nn.Conv2d | nn.Linear, pad_cin: int | None = None, pad_cout: int | None = None):
    """-> (bf16 [Cout', taps*Cin'], fp32 bias [Cout']) with optional zero padding of Cin / Cout."""
    w = conv.weight.detach().float()
    b = conv.bias.detach().float() if conv.bias is not None else None
    if w.dim() == 4 and pad_cin and w.shape[1] < pad_cin:
        w = torch.cat([w, w.new_zeros(w.shape[0], pad_cin - w.shape[1], *w.shape[2:])], 1)
    if pad_cout and w.shape[0] < pad_cout:
        w = torch.cat([w, w.new_zeros(pad_cout - w.shape[0], *w.shape[1:])], 0)
        if b is not None:
            b = torch.cat([b, b.new_zeros(pad_cout - b.shape[0])])
    return ops.pack_conv_weight(w), (b.contiguous() if b is not None else None)


# ------------------------------------------------------------------------------------------------- embeddings
class Timesteps(nn.Module):
    """diffusers ``Timesteps(dim, flip_sin_to_cos=True, downscale_freq_shift=0)`` (controller.py:86)."""

    def __init__(self, num_channels: int, flip_sin_to_cos: bool = True, downscale_freq_shift: float = 0):
        super().__init__()
        if not flip_sin_to_cos or downscale_freq_shift != 0:
            raise ValueError("only the sd-turbo configuration (cos-first, shift 0) is implemented")
        self.num_channels = num_channels

    def forward(self, timesteps):
        return ops.timestep_embedding(timesteps.to(torch.int64), self.num_channels)


class TimestepEmbedding(UrModule):
    """diffusers ``TimestepEmbedding(in, dim, "silu")``: linear_2(silu(linear_1(x)))."""

    def __init__(self, in_channels: int, time_embed_dim: int, act_fn: str = "silu"):
        super().__init__()
        self.linear_1 = nn.Linear(in_channels, time_embed_dim)
        self.linear_2 = nn.Linear(time_embed_dim, time_embed_dim)

    def _pack(self):
        return dict(w1=_f32(self.linear_1.weight), b1=_f32(self.linear_1.bias), w2=_f32(self.linear_2.weight),
                    b2=_f32(self.linear_2.bias))

    def forward(self, sample):
        p = self.pk
        h = ops.small_linear(sample.float().contiguous(), p["w1"], p["b1"], act_out="silu")
        return ops.small_linear(h, p["w2"], p["b2"])


class TimeCtx:
    """Time embeddings of ONE forward: ``emb`` fp32 [T, D] for the T scheduler timesteps (computed at the start of every
    forward, inside the timed region -- nothing input-independent is carried over between forwards except packed
    weights), ``index`` = current DDIM step.  ``proj`` holds each ResnetBlock's ``time_emb_proj(silu(emb))`` [T, C],
    filled by one launch per block at first use; ``at(i)`` shares both with a different step index."""

    def __init__(self, emb, index=0, proj=None):
        self.emb, self.index, self.proj = emb, index, ({} if proj is None else proj)

    def at(self, index):
        return TimeCtx(self.emb, index, self.proj)


# ------------------------------------------------------------------------------------------------- resnet
class ResnetBlock2D(UrModule):
    """GN -> SiLU -> conv3x3 (+time_emb_proj(silu(temb))) -> GN -> SiLU -> conv3x3 (+1x1 shortcut) + input.

    diffusers ResnetBlock2D as called at base_model.py:54 / controller.py:161-170 (additive temb only)."""

    def __init__(self, *, in_channels, out_channels=None, temb_channels=512, groups=32, eps=1e-6, **_):
        super().__init__()
        out_channels = in_channels if out_channels is None else out_channels
        self.in_channels, self.out_channels, self.groups, self.eps = in_channels, out_channels, groups, eps
        self.norm1 = nn.GroupNorm(groups, in_channels, eps=eps, affine=True)
        self.conv1 = nn.Conv2d(in_channels, out_channels, 3, padding=1)
        self.time_emb_proj = nn.Linear(temb_channels, out_channels) if temb_channels is not None else None
        self.norm2 = nn.GroupNorm(groups, out_channels, eps=eps, affine=True)
        self.conv2 = nn.Conv2d(out_channels, out_channels, 3, padding=1)
        self.conv_shortcut = nn.Conv2d(in_channels, out_channels, 1) if in_channels != out_channels else None

    def _pack(self):
        p = dict(g1=_f32(self.norm1.weight), b1=_f32(self.norm1.bias), g2=_f32(self.norm2.weight),
                 b2=_f32(self.norm2.bias))
        p["w1"], p["c1b"] = pack_conv(self.conv1)
        p["w2"], p["c2b"] = pack_conv(self.conv2)
        if self.conv_shortcut is not None:
            p["ws"], p["sb"] = pack_conv(self.conv_shortcut)
        if self.time_emb_proj is not None:
            p["wt"], p["tb"] = _f32(self.time_emb_proj.weight), _f32(self.time_emb_proj.bias)
        return p

    def run(self, x, temb=None, x2=None):
        """x (and optional channel-concatenated x2) bf16 NHWC; temb fp32 [1 or B, temb_channels] or a ``TimeCtx``."""
        p = self.pk
        co = self.out_channels
        h = ops.group_norm(x, self.groups, p["g1"], p["b1"], self.eps, silu=True, x2=x2)
        tvec = None
        if temb is not None and self.time_emb_proj is not None:
            if isinstance(temb, TimeCtx):
                # all T scheduler timesteps are projected in ONE launch at first use in this forward, then sliced
                pr = temb.proj.get(id(self))
                if pr is None:
                    pr = temb.proj[id(self)] = ops.small_linear(temb.emb, p["wt"], p["tb"], act_in="silu")
                tvec = pr[temb.index:temb.index + 1]
            else:
                tvec = ops.small_linear(temb, p["wt"], p["tb"], act_in="silu")
        # want_stats: the GEMM epilogue accumulates the statistics of the GroupNorm that consumes its output
        h = ops.conv_gemm(h, p["w1"], co, taps=TAPS_3x3, bias=p["c1b"], rowvec=tvec, want_stats=True)
        h = ops.group_norm(h, self.groups, p["g2"], p["b2"], self.eps, silu=True)
        if self.conv_shortcut is not None:
            res = ops.conv_gemm(x, p["ws"], co, x2=x2, bias=p["sb"])
        else:
            res = x
        return ops.conv_gemm(h, p["w2"], co, taps=TAPS_3x3, bias=p["c2b"], residual=res, want_stats=True)


class Downsample2D(UrModule):
    """conv3x3 stride 2; ``padding=0`` = the VAE's asymmetric (0,1,0,1) zero pad (autoencoder.py:19)."""

    def __init__(self, channels, use_conv=True, out_channels=None, padding=1, name="conv"):
        super().__init__()
        self.padding, self.out_channels = padding, out_channels or channels
        self.conv = nn.Conv2d(channels, self.out_channels, 3, stride=2, padding=padding)

    def _pack(self):
        w, b = pack_conv(self.conv)
        return dict(w=w, b=b)

    def run(self, x):
        B, H, W, _ = x.shape
        if self.padding == 1:
            ho, wo, taps = (H - 1) // 2 + 1, (W - 1) // 2 + 1, TAPS_3x3
        else:
            ho, wo, taps = (H - 2) // 2 + 1, (W - 2) // 2 + 1, TAPS_3x3_NOPAD
        return ops.conv_gemm(x, self.pk["w"], self.out_channels, taps=taps, stride=2, hout=ho, wout=wo,
                             bias=self.pk["b"], want_stats=True)


class Upsample2D(UrModule):
    """nearest x2 + conv3x3 (base_model.py:202-203, autoencoder.py:60) as four sub-pixel 2x2 convolutions on the
    LOW-resolution input: for output parity (py,px) the 3x3 taps that hit the same source pixel are pre-summed,
    so the up-sampled tensor is never materialised and the work drops to 4/9 of the naive convolution."""

    def __init__(self, channels, use_conv=True, out_channels=None):
        super().__init__()
        self.out_channels = out_channels or channels
        self.conv = nn.Conv2d(channels, self.out_channels, 3, padding=1)

    def _pack(self):
        w = self.conv.weight.detach().float()                       # [Co, Ci, 3, 3]
        sel = {0: ((0,), (1, 2)), 1: ((0, 1), (2,))}                 # parity -> 3x3 taps merged into 2x2 tap 0 / 1
        phases = {}
        for py in (0, 1):
            for px in (0, 1):
                taps, mats = [], []
                for ry in (0, 1):
                    for rx in (0, 1):
                        m = sum(w[:, :, ky, kx] for ky in sel[py][ry] for kx in sel[px][rx])
                        mats.append(m)
                        taps.append((py - 1 + ry, px - 1 + rx))
                wp = torch.stack(mats, 1).reshape(w.shape[0], -1).contiguous().to(torch.bfloat16)   # [Co, 4*Ci]
                phases[(py, px)] = (tuple(taps), wp)
        return dict(phases=phases, b=_f32(self.conv.bias))

    def run(self, x):
        B, H, W, _ = x.shape
        out = torch.empty((B, 2 * H, 2 * W, self.out_channels), device=x.device, dtype=torch.bfloat16)
        # GroupNorm statistics of the up-sampled tensor (its consumer is always a ResnetBlock2D): the four phase GEMMs
        # accumulate into ONE buffer in their epilogues, so no ur_chan_stats pass reads the tensor again (the VAE's
        # 512^2 x 128 outputs cost 126 us each that way).  The epilogue can only do it while an image contributes >= 32
        # rows to an M tile (H * W >= 32); below that the consumer falls back to its own statistics pass.
        stats = ops.new_stats(x.device, B, self.out_channels) if H * W >= 32 else None
        for (py, px), (taps, wp) in self.pk["phases"].items():
            ops.conv_gemm(x, wp, self.out_channels, taps=taps, bias=self.pk["b"], out=out[:, py::2, px::2], stats=stats)
        if stats is not None:
            out._ur_stats = stats
        return out


# ------------------------------------------------------------------------------------------------- attention
class Attention(UrModule):
    """diffusers ``Attention``.  Spatial form (group_norm + residual; Controller / VAE mid block, A.3) via ``run``;
    token form (BasicTransformerBlock attn1 / attn2, A.4) via ``run_tokens``."""

    def __init__(self, query_dim, cross_attention_dim=None, heads=8, dim_head=64, bias=False, out_bias=True,
                 norm_num_groups=None, eps=1e-5, residual_connection=False, **_):
        super().__init__()
        inner = heads * dim_head
        self.heads, self.dim_head, self.inner, self.query_dim = heads, dim_head, inner, query_dim
        self.is_cross = cross_attention_dim is not None
        self.residual_connection = residual_connection
        self.eps, self.groups = eps, norm_num_groups
        self.group_norm = nn.GroupNorm(norm_num_groups, query_dim, eps=eps) if norm_num_groups is not None else None
        kv_dim = cross_attention_dim if self.is_cross else query_dim
        self.to_q = nn.Linear(query_dim, inner, bias=bias)
        self.to_k = nn.Linear(kv_dim, inner, bias=bias)
        self.to_v = nn.Linear(kv_dim, inner, bias=bias)
        self.to_out = nn.ModuleList([nn.Linear(inner, query_dim, bias=out_bias), nn.Dropout(0.0)])
        self._ctx_ref, self._ctx_kv = None, None

    def _reset_cache(self):
        self._pk = None
        self._ctx_ref, self._ctx_kv = None, None

    def _pack(self):
        p = {}
        cat_b = lambda ls: (torch.cat([_f32(l.bias) for l in ls]) if ls[0].bias is not None else None)
        if self.is_cross:
            p["wq"], p["bq"] = pack_conv(self.to_q)
            p["wkv"] = torch.cat([ops.pack_conv_weight(self.to_k.weight.detach()),
                                  ops.pack_conv_weight(self.to_v.weight.detach())], 0).contiguous()
            p["bkv"] = cat_b([self.to_k, self.to_v])
        else:
            p["wqkv"] = torch.cat([ops.pack_conv_weight(l.weight.detach()) for l in (self.to_q, self.to_k, self.to_v)],
                                  0).contiguous()
            p["bqkv"] = cat_b([self.to_q, self.to_k, self.to_v])
        p["wo"], p["bo"] = pack_conv(self.to_out[0])
        if self.group_norm is not None:
            p["gn_g"], p["gn_b"] = _f32(self.group_norm.weight), _f32(self.group_norm.bias)
        return p

    def _kv(self, ctx):
        """K/V of the (constant) encoder states; cached while the same ctx tensor is passed (base_model.py:221)."""
        if self._ctx_ref is not ctx or self._ctx_kv is None:
            kv = ops.conv_gemm(ctx, self.pk["wkv"], 2 * self.inner, bias=self.pk["bkv"])
            self._ctx_ref, self._ctx_kv = ctx, kv
        kv = self._ctx_kv
        return kv[..., : self.inner], kv[..., self.inner:]

    def run_tokens(self, x, ctx=None, residual=None, want_stats=False):
        """x bf16 [B,T,C] -> to_out(attn(x[, ctx])) (+ residual)."""
        p, C = self.pk, self.inner
        if self.is_cross:
            q = ops.conv_gemm(x, p["wq"], C, bias=p["bq"])
            k, v = self._kv(ctx)
        else:
            qkv = ops.conv_gemm(x, p["wqkv"], 3 * C, bias=p["bqkv"])
            q, k, v = qkv[..., :C], qkv[..., C:2 * C], qkv[..., 2 * C:]
        a = ops.attention(q, k, v, self.heads)
        return ops.conv_gemm(a, p["wo"], self.query_dim, bias=p["bo"], residual=residual, want_stats=want_stats)

    def run(self, x):
        """Spatial self-attention block on bf16 NHWC: GN -> qkv -> SDPA -> to_out -> + x."""
        p = self.pk
        B, H, W, C = x.shape
        xn = ops.group_norm(x, self.groups, p["gn_g"], p["gn_b"], self.eps) if self.group_norm is not None else x
        res = x.view(B, H * W, C) if self.residual_connection else None
        y = self.run_tokens(xn.view(B, H * W, C), residual=res, want_stats=True)
        return _view_keep_stats(y, (B, H, W, C))


class GEGLU(nn.Module):
    def __init__(self, dim_in, dim_out):
        super().__init__()
        self.proj = nn.Linear(dim_in, dim_out * 2)


class FeedForward(UrModule):
    """GEGLU feed-forward: net.0.proj (C -> 8C, a*gelu(g) fused in the GEMM epilogue), net.2 (4C -> C)."""

    def __init__(self, dim, mult=4):
        super().__init__()
        self.dim = dim
        self.net = nn.ModuleList([GEGLU(dim, dim * mult), nn.Dropout(0.0), nn.Linear(dim * mult, dim)])

    def _pack(self):
        proj = self.net[0].proj
        n = proj.weight.shape[0]
        bn = ops.pick_bn(n, True)
        w, b = ops.pack_gated_weight(ops.pack_conv_weight(proj.weight.detach()), _f32(proj.bias), bn)
        w2, b2 = pack_conv(self.net[2])
        return dict(w1=w, b1=b, bn=bn, n1=n, w2=w2, b2=b2)

    def run(self, x, residual=None):
        p = self.pk
        h = ops.conv_gemm(x, p["w1"], p["n1"], bias=p["b1"], act=UR_ACT_GEGLU, bn=p["bn"])
        return ops.conv_gemm(h, p["w2"], self.dim, bias=p["b2"], residual=residual)


class BasicTransformerBlock(UrModule):
    def __init__(self, dim, heads, dim_head, cross_attention_dim):
        super().__init__()
        self.norm1 = nn.LayerNorm(dim, eps=1e-5)
        self.attn1 = Attention(dim, None, heads, dim_head, bias=False)
        self.norm2 = nn.LayerNorm(dim, eps=1e-5)
        self.attn2 = Attention(dim, cross_attention_dim, heads, dim_head, bias=False)
        self.norm3 = nn.LayerNorm(dim, eps=1e-5)
        self.ff = FeedForward(dim)

    def _pack(self):
        return {k: (_f32(getattr(self, k).weight), _f32(getattr(self, k).bias), getattr(self, k).eps)
                for k in ("norm1", "norm2", "norm3")}

    def run(self, x, ctx):
        p = self.pk
        x = self.attn1.run_tokens(ops.layernorm(x, *p["norm1"]), residual=x)
        x = self.attn2.run_tokens(ops.layernorm(x, *p["norm2"]), ctx, residual=x)
        return self.ff.run(ops.layernorm(x, *p["norm3"]), residual=x)


class Transformer2DModel(UrModule):
    """Continuous-input transformer, ``use_linear_projection=True`` (A.4); called base_model.py:138,159,191."""

    def __init__(self, num_attention_heads, attention_head_dim, in_channels, cross_attention_dim,
                 norm_num_groups=32, num_layers=1, **_):
        super().__init__()
        inner = num_attention_heads * attention_head_dim
        self.groups, self.in_channels, self.inner = norm_num_groups, in_channels, inner
        self.norm = nn.GroupNorm(norm_num_groups, in_channels, eps=1e-6, affine=True)
        self.proj_in = nn.Linear(in_channels, inner)
        self.transformer_blocks = nn.ModuleList([
            BasicTransformerBlock(inner, num_attention_heads, attention_head_dim, cross_attention_dim)
            for _ in range(num_layers)])
        self.proj_out = nn.Linear(inner, in_channels)

    def _pack(self):
        p = dict(g=_f32(self.norm.weight), b=_f32(self.norm.bias))
        p["wi"], p["bi"] = pack_conv(self.proj_in)
        p["wo"], p["bo"] = pack_conv(self.proj_out)
        return p

    def run(self, x, ctx):
        p = self.pk
        B, H, W, C = x.shape
        t = ops.group_norm(x, self.groups, p["g"], p["b"], self.norm.eps)
        t = ops.conv_gemm(t.view(B, H * W, C), p["wi"], self.inner, bias=p["bi"])
        for blk in self.transformer_blocks:
            t = blk.run(t, ctx)
        y = ops.conv_gemm(t, p["wo"], C, bias=p["bo"], residual=x.view(B, H * W, C), want_stats=True)
        return _view_keep_stats(y, (B, H, W, C))


# ------------------------------------------------------------------------------------------------- UNet blocks
def _resnets(n, cin, cout, temb, eps, groups):
    return nn.ModuleList([ResnetBlock2D(in_channels=cin if i == 0 else cout, out_channels=cout, temb_channels=temb,
                                        eps=eps, groups=groups) for i in range(n)])


class DownBlock2D(nn.Module):
    has_cross_attention = False

    def __init__(self, *, num_layers, in_channels, out_channels, temb_channels, add_downsample, resnet_eps,
                 resnet_groups=32, downsample_padding=1, **_):
        super().__init__()
        self.resnets = _resnets(num_layers, in_channels, out_channels, temb_channels, resnet_eps, resnet_groups)
        self.downsamplers = (nn.ModuleList([Downsample2D(out_channels, True, out_channels, downsample_padding)])
                             if add_downsample else None)

    def run(self, x, temb=None):
        outs = []
        for r in self.resnets:
            x = r.run(x, temb)
            outs.append(x)
        if self.downsamplers is not None:
            x = self.downsamplers[0].run(x)
            outs.append(x)
        return x, outs


class AttnDownBlock2D(nn.Module):
    """resnet + spatial self-attention per layer (Controller, controller.py:101-125)."""
    has_cross_attention = False

    def __init__(self, *, num_layers, in_channels, out_channels, temb_channels, add_downsample, resnet_eps,
                 resnet_groups=32, attention_head_dim=1, downsample_padding=1, **_):
        super().__init__()
        self.resnets = _resnets(num_layers, in_channels, out_channels, temb_channels, resnet_eps, resnet_groups)
        self.attentions = nn.ModuleList([
            Attention(out_channels, heads=out_channels // attention_head_dim, dim_head=attention_head_dim,
                      eps=resnet_eps, norm_num_groups=resnet_groups, residual_connection=True, bias=True)
            for _ in range(num_layers)])
        self.downsamplers = (nn.ModuleList([Downsample2D(out_channels, True, out_channels, downsample_padding)])
                             if add_downsample else None)

    def run(self, x, temb=None):
        outs = []
        for r, a in zip(self.resnets, self.attentions):
            x = a.run(r.run(x, temb))
            outs.append(x)
        if self.downsamplers is not None:
            x = self.downsamplers[0].run(x)
            outs.append(x)
        return x, outs


class CrossAttnDownBlock2D(nn.Module):
    has_cross_attention = True

    def __init__(self, *, num_layers, in_channels, out_channels, temb_channels, add_downsample, resnet_eps,
                 resnet_groups=32, num_attention_heads=1, cross_attention_dim=1024, downsample_padding=1, **_):
        super().__init__()
        self.resnets = _resnets(num_layers, in_channels, out_channels, temb_channels, resnet_eps, resnet_groups)
        self.attentions = nn.ModuleList([
            Transformer2DModel(num_attention_heads, out_channels // num_attention_heads, out_channels,
                               cross_attention_dim, resnet_groups) for _ in range(num_layers)])
        self.downsamplers = (nn.ModuleList([Downsample2D(out_channels, True, out_channels, downsample_padding)])
                             if add_downsample else None)


def get_down_block(down_block_type, **kw):
    """diffusers ``get_down_block`` for the block types the reference instantiates (controller.py:101-125)."""
    kw = {k: v for k, v in kw.items() if v is not None}
    if down_block_type == "DownBlock2D":
        return DownBlock2D(**kw)
    if down_block_type == "AttnDownBlock2D":
        return AttnDownBlock2D(**kw)
    if down_block_type == "CrossAttnDownBlock2D":
        return CrossAttnDownBlock2D(**kw)
    raise ValueError(f"{down_block_type} does not exist.")


class UNetMidBlock2D(nn.Module):
    """res, spatial-attn, res (Controller mid controller.py:133-141; VAE mid block)."""

    def __init__(self, *, in_channels, temb_channels, resnet_eps=1e-6, resnet_groups=32, attention_head_dim=1, **_):
        super().__init__()
        self.resnets = _resnets(2, in_channels, in_channels, temb_channels, resnet_eps, resnet_groups)
        self.attentions = nn.ModuleList([
            Attention(in_channels, heads=in_channels // attention_head_dim, dim_head=attention_head_dim,
                      eps=resnet_eps, norm_num_groups=resnet_groups, residual_connection=True, bias=True)])

    def run(self, x, temb=None):
        x = self.resnets[0].run(x, temb)
        return self.resnets[1].run(self.attentions[0].run(x), temb)


class UNetMidBlock2DCrossAttn(nn.Module):
    has_cross_attention = True

    def __init__(self, *, in_channels, temb_channels, resnet_eps=1e-5, resnet_groups=32, num_attention_heads=1,
                 cross_attention_dim=1024, **_):
        super().__init__()
        self.resnets = _resnets(2, in_channels, in_channels, temb_channels, resnet_eps, resnet_groups)
        self.attentions = nn.ModuleList([
            Transformer2DModel(num_attention_heads, in_channels // num_attention_heads, in_channels,
                               cross_attention_dim, resnet_groups)])


class _UpBlock(nn.Module):
    def __init__(self, *, num_layers, in_channels, out_channels, prev_output_channel, temb_channels, add_upsample,
                 resnet_eps, resnet_groups=32, num_attention_heads=None, cross_attention_dim=1024, **_):
        super().__init__()
        rs = []
        for i in range(num_layers):
            skip = in_channels if i == num_layers - 1 else out_channels
            rin = prev_output_channel if i == 0 else out_channels
            rs.append(ResnetBlock2D(in_channels=rin + skip, out_channels=out_channels, temb_channels=temb_channels,
                                    eps=resnet_eps, groups=resnet_groups))
        self.resnets = nn.ModuleList(rs)
        if num_attention_heads is not None:
            self.attentions = nn.ModuleList([
                Transformer2DModel(num_attention_heads, out_channels // num_attention_heads, out_channels,
                                   cross_attention_dim, resnet_groups) for _ in range(num_layers)])
        self.upsamplers = nn.ModuleList([Upsample2D(out_channels, True, out_channels)]) if add_upsample else None


class UpBlock2D(_UpBlock):
    has_cross_attention = False


class CrossAttnUpBlock2D(_UpBlock):
    has_cross_attention = True


# sd-turbo (SD-2.1 topology) constants -- recalled from the public HF configs; one table.
UNET_CONFIG = dict(in_channels=4, out_channels=4, block_out_channels=(320, 640, 1280, 1280), layers_per_block=2,
                   down_block_types=("CrossAttnDownBlock2D",) * 3 + ("DownBlock2D",),
                   up_block_types=("UpBlock2D",) + ("CrossAttnUpBlock2D",) * 3, cross_attention_dim=1024,
                   num_attention_heads=(5, 10, 20, 20), norm_num_groups=32, norm_eps=1e-5, time_embed_in=320,
                   time_embed_dim=1280)
VAE_CONFIG = dict(in_channels=3, out_channels=3, latent_channels=4, block_out_channels=(128, 256, 512, 512),
                  layers_per_block=2, norm_num_groups=32, resnet_eps=1e-6, scaling_factor=0.18215)


class _Config(dict):
    __getattr__ = dict.__getitem__


class UNet2DConditionModel(UrModule):
    """Container with the sd-turbo UNet topology and diffusers key names (A.5).  Like the reference
    (base_model.py:94-209) the forward is walked by ``ControlledUNet``; only the stem / head live here."""

    def __init__(self, **overrides):
        super().__init__()
        c = dict(UNET_CONFIG)
        c.update(overrides)
        self.config = _Config(c)
        boc, heads, ted = c["block_out_channels"], c["num_attention_heads"], c["time_embed_dim"]
        self.conv_in = nn.Conv2d(c["in_channels"], boc[0], 3, padding=1)
        self.time_proj = Timesteps(c["time_embed_in"], True, 0)
        self.time_embedding = TimestepEmbedding(c["time_embed_in"], ted)
        self.down_blocks = nn.ModuleList()
        out_ch = boc[0]
        for i, t in enumerate(c["down_block_types"]):
            in_ch, out_ch = out_ch, boc[i]
            self.down_blocks.append(get_down_block(
                t, num_layers=c["layers_per_block"], in_channels=in_ch, out_channels=out_ch, temb_channels=ted,
                add_downsample=i != len(boc) - 1, resnet_eps=c["norm_eps"], resnet_groups=c["norm_num_groups"],
                downsample_padding=1, cross_attention_dim=c["cross_attention_dim"],
                num_attention_heads=heads[i] if t.startswith("CrossAttn") else None))
        self.mid_block = UNetMidBlock2DCrossAttn(
            in_channels=boc[-1], temb_channels=ted, resnet_eps=c["norm_eps"], resnet_groups=c["norm_num_groups"],
            num_attention_heads=heads[-1], cross_attention_dim=c["cross_attention_dim"])
        self.up_blocks = nn.ModuleList()
        rev, rheads = list(reversed(boc)), list(reversed(heads))
        out_ch = rev[0]
        for i, t in enumerate(c["up_block_types"]):
            prev, out_ch = out_ch, rev[i]
            in_ch = rev[min(i + 1, len(boc) - 1)]
            cls = CrossAttnUpBlock2D if t == "CrossAttnUpBlock2D" else UpBlock2D
            self.up_blocks.append(cls(
                num_layers=c["layers_per_block"] + 1, in_channels=in_ch, out_channels=out_ch,
                prev_output_channel=prev, temb_channels=ted, add_upsample=i != len(boc) - 1,
                resnet_eps=c["norm_eps"], resnet_groups=c["norm_num_groups"],
                num_attention_heads=rheads[i] if t == "CrossAttnUpBlock2D" else None,
                cross_attention_dim=c["cross_attention_dim"]))
        self.conv_norm_out = nn.GroupNorm(c["norm_num_groups"], boc[0], eps=c["norm_eps"])
        self.conv_act = nn.SiLU()
        self.conv_out = nn.Conv2d(boc[0], c["out_channels"], 3, padding=1)

    @classmethod
    def from_pretrained(cls, model_id=None, subfolder=None, **kw):
        """No weights exist offline: returns the sd-turbo topology with default init."""
        return cls()

    def _pack(self):
        p = dict(g=_f32(self.conv_norm_out.weight), b=_f32(self.conv_norm_out.bias))
        p["w_in"], p["b_in"] = pack_conv(self.conv_in, pad_cin=8)
        p["w_out"], p["b_out"] = pack_conv(self.conv_out, pad_cout=8)
        return p

    def run_conv_in(self, z8):
        return ops.conv_gemm(z8, self.pk["w_in"], self.conv_in.out_channels, taps=TAPS_3x3, bias=self.pk["b_in"],
                             want_stats=True)

    def run_head(self, x):
        """GN -> SiLU -> conv_out; returns fp32 channels-last [B,h,w,8] (channels >= out_channels are zero)."""
        p = self.pk
        h = ops.group_norm(x, self.conv_norm_out.num_groups, p["g"], p["b"], self.conv_norm_out.eps, silu=True)
        return ops.conv_gemm(h, p["w_out"], 8, taps=TAPS_3x3, bias=p["b_out"], out_dtype=torch.float32)


# ------------------------------------------------------------------------------------------------- VAE
class DownEncoderBlock2D(nn.Module):
    def __init__(self, in_channels, out_channels, num_layers, add_downsample, eps, groups):
        super().__init__()
        self.resnets = _resnets(num_layers, in_channels, out_channels, None, eps, groups)
        self.downsamplers = (nn.ModuleList([Downsample2D(out_channels, True, out_channels, 0)])
                             if add_downsample else None)

    def run(self, x):
        for r in self.resnets:
            x = r.run(x)
        if self.downsamplers is not None:
            x = self.downsamplers[0].run(x)
        return x


class UpDecoderBlock2D(nn.Module):
    def __init__(self, in_channels, out_channels, num_layers, add_upsample, eps, groups):
        super().__init__()
        self.resnets = _resnets(num_layers, in_channels, out_channels, None, eps, groups)
        self.upsamplers = nn.ModuleList([Upsample2D(out_channels, True, out_channels)]) if add_upsample else None

    def run(self, x):
        for r in self.resnets:
            x = r.run(x)
        if self.upsamplers is not None:
            x = self.upsamplers[0].run(x)
        return x


class Encoder(UrModule):
    def __init__(self, c):
        super().__init__()
        boc, g, eps = c["block_out_channels"], c["norm_num_groups"], c["resnet_eps"]
        self.conv_in = nn.Conv2d(c["in_channels"], boc[0], 3, padding=1)
        self.down_blocks = nn.ModuleList()
        out_ch = boc[0]
        for i in range(len(boc)):
            in_ch, out_ch = out_ch, boc[i]
            self.down_blocks.append(DownEncoderBlock2D(in_ch, out_ch, c["layers_per_block"], i != len(boc) - 1, eps, g))
        self.mid_block = UNetMidBlock2D(in_channels=boc[-1], temb_channels=None, resnet_eps=eps, resnet_groups=g,
                                        attention_head_dim=boc[-1])
        self.conv_norm_out = nn.GroupNorm(g, boc[-1], eps=1e-6)
        self.conv_act = nn.SiLU()
        self.conv_out = nn.Conv2d(boc[-1], 2 * c["latent_channels"], 3, padding=1)

    def _pack(self):
        p = dict(g=_f32(self.conv_norm_out.weight), b=_f32(self.conv_norm_out.bias))
        p["w_in"], p["b_in"] = pack_conv(self.conv_in, pad_cin=8)
        p["w_out"], p["b_out"] = pack_conv(self.conv_out)
        return p


class Decoder(UrModule):
    def __init__(self, c):
        super().__init__()
        boc, g, eps = c["block_out_channels"], c["norm_num_groups"], c["resnet_eps"]
        self.conv_in = nn.Conv2d(c["latent_channels"], boc[-1], 3, padding=1)
        self.mid_block = UNetMidBlock2D(in_channels=boc[-1], temb_channels=None, resnet_eps=eps, resnet_groups=g,
                                        attention_head_dim=boc[-1])
        self.up_blocks = nn.ModuleList()
        rev = list(reversed(boc))
        out_ch = rev[0]
        for i in range(len(boc)):
            prev, out_ch = out_ch, rev[i]
            self.up_blocks.append(UpDecoderBlock2D(prev, out_ch, c["layers_per_block"] + 1, i != len(boc) - 1, eps, g))
        self.conv_norm_out = nn.GroupNorm(g, boc[0], eps=1e-6)
        self.conv_act = nn.SiLU()
        self.conv_out = nn.Conv2d(boc[0], c["out_channels"], 3, padding=1)

    def _pack(self):
        p = dict(g=_f32(self.conv_norm_out.weight), b=_f32(self.conv_norm_out.bias))
        p["w_in"], p["b_in"] = pack_conv(self.conv_in, pad_cin=8)
        p["w_out"], p["b_out"] = pack_conv(self.conv_out, pad_cout=8)
        return p


class AutoencoderKL(UrModule):
    """sd-turbo VAE container (A.6).  The reference replaces encoder/decoder forward (autoencoder.py:88-90,
    108-110); here ``SkipConnectedAutoEncoder`` drives the children directly."""

    def __init__(self, **overrides):
        super().__init__()
        c = dict(VAE_CONFIG)
        c.update(overrides)
        self.config = _Config(c)
        self.encoder = Encoder(c)
        self.decoder = Decoder(c)
        self.quant_conv = nn.Conv2d(2 * c["latent_channels"], 2 * c["latent_channels"], 1)
        self.post_quant_conv = nn.Conv2d(c["latent_channels"], c["latent_channels"], 1)

    @classmethod
    def from_pretrained(cls, model_id=None, subfolder=None, **kw):
        return cls()

    def _pack(self):
        p = {}
        p["wq"], p["bq"] = pack_conv(self.quant_conv)
        p["wpq"], p["bpq"] = pack_conv(self.post_quant_conv, pad_cin=8, pad_cout=8)
        return p
