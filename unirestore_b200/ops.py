"""Thin tensor-level wrappers over the C-ABI (``_cabi``): torch tensors in, raw pointers out.

Activations are bf16 tensors of logical shape ``[B, H, W, C]`` (channels-last physical layout;
a token matrix ``[B, T, C]`` is treated as ``H = 1, W = T``).  Channel-slice views (``stride(-1) == 1``)
are accepted as inputs and outputs so concatenations / q-k-v splits are never materialised.
Nothing here falls back to PyTorch arithmetic: every op is one launch of a hand-written kernel.
"""
from __future__ import annotations

import ctypes as C
import os

import torch

from . import _cabi
from ._cabi import (UR_ACT_GATE, UR_ACT_GEGLU, UR_ACT_GELU, UR_ACT_NONE, UR_ACT_SILU, UR_DT_BF16, UR_DT_F32,
                    ConvDesc, check)

TAPS_1x1 = ((0, 0),)
TAPS_3x3 = tuple((ky - 1, kx - 1) for ky in range(3) for kx in range(3))          # zero padding 1
TAPS_3x3_NOPAD = tuple((ky, kx) for ky in range(3) for kx in range(3))             # VAE asym pad (0,1,0,1)


def _stream():
    """The current stream of the CURRENT device; every public op runs under ``_on_device`` which makes the device of
    its first tensor argument current, so a model on cuda:1 never launches on cuda:0's stream."""
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _on_device(fn):
    """Run ``fn`` with the device of its first CUDA tensor argument current (a no-op compare on the common path)."""
    import functools

    @functools.wraps(fn)
    def wrapped(*args, **kw):
        for t in args:
            if isinstance(t, torch.Tensor):
                if t.is_cuda:
                    idx = t.device.index
                    if idx not in _cabi._inited:
                        _cabi.ensure_init(idx)
                    if idx != torch.cuda.current_device():
                        with torch.cuda.device(idx):
                            return fn(*args, **kw)
                break
        return fn(*args, **kw)
    return wrapped


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


def _as4(t):
    """[B, T, C] -> [B, 1, T, C] view; [M, C] -> [1, 1, M, C]."""
    if t.dim() == 4:
        return t
    if t.dim() == 3:
        return t.unsqueeze(1)
    if t.dim() == 2:
        return t.unsqueeze(0).unsqueeze(0)
    raise ValueError("expected a 2-4 D tensor, got %s" % (tuple(t.shape),))


def _check_nhwc(t, name):
    if t.dtype != torch.bfloat16 or not t.is_cuda:
        raise ValueError("%s must be a CUDA bf16 tensor" % name)
    B, H, W, Cc = t.shape
    ld = t.stride(2) if W > 1 else (t.stride(1) if H > 1 else (t.stride(0) if B > 1 else Cc))
    ok = (t.stride(3) == 1 or Cc == 1) and (W == 1 or t.stride(2) == ld) and (H == 1 or t.stride(1) == W * ld) \
        and (B == 1 or t.stride(0) == H * W * ld)
    if not ok:
        raise ValueError("%s: unsupported strides %s for shape %s" % (name, t.stride(), tuple(t.shape)))
    return ld


_WS = {}
_ROLE = ["main"]      # which execution context issues the calls: "main" or "ctl" (the Controller's side stream)
_WS_ROLE = [None]     # workspace owner when it differs from the statistics-pool role (the SC-Tuner side stream)


class workspace_role:
    """``with workspace_role("sct"):`` -- kernels issued inside use their own split-K workspace.  Every stream that can
    run concurrently with another one owns a workspace (Controller: role "ctl" via stats_arena_begin; SC-Tuner: "sct")."""

    def __init__(self, name):
        self.name, self.prev = name, None

    def __enter__(self):
        self.prev, _WS_ROLE[0] = _WS_ROLE[0], self.name

    def __exit__(self, *a):
        _WS_ROLE[0] = self.prev


def _workspace(device, nbytes=32 << 20):
    """Caller-owned split-K scratch handed to every ur_conv_gemm call: one per (device, role), since the Controller
    of the next DDIM step and the SC-Tuner adapters run on side streams concurrently with the UNet (stream-ordered
    reuse within a stream)."""
    key = (device, _WS_ROLE[0] or _ROLE[0])
    ws = _WS.get(key)
    if ws is None:
        ws = _WS[key] = torch.empty(nbytes // 4, dtype=torch.float32, device=device)
    return ws


def conv_gemm(x, w, n, *, x2=None, taps=TAPS_1x1, stride=1, hout=None, wout=None, bias=None, rowvec=None,
              chscale=None, residual=None, act=UR_ACT_NONE, alpha=1.0, out=None, out_dtype=torch.bfloat16,
              group_kc=0, group_nc=0, w_batched=False, bn=0, want_stats=False, stats=None):
    """``out = epilogue(alpha * conv(x [cat x2], w))`` -- see ``ur_conv_gemm`` in include/unirestore_b200.h.

    ``w`` is the packed bf16 weight ``[n, ntaps * kc]`` (``[batch, n, k]`` when ``w_batched``).
    ``out`` may be a strided NHWC view (channel slice, or ``full[:, py::2, px::2]`` for sub-pixel phases).
    ``want_stats``: the per-(image, channel) (sum, sumsq) of the bf16 output -- the statistics pass of the GroupNorm
    that consumes it -- are accumulated by the GEMM epilogue and attached to the result as ``._ur_stats``
    (``stats``: accumulate into this existing ``[B, n_out, 2]`` fp64 buffer instead, e.g. the 4 sub-pixel phases).
    """
    squeeze = x.dim()
    x = _as4(x)
    B, H, W, C1 = x.shape
    ld1 = _check_nhwc(x, "x")
    C2, ld2 = 0, 0
    if x2 is not None:
        x2 = _as4(x2)
        if x2.shape[:3] != x.shape[:3]:
            raise ValueError("x2 spatial shape mismatch")
        if C1 % 64:      # the two-source TMA path needs 64-aligned source-1 channels: materialise the concat
            x, x2 = concat_channels(x, x2), None
            C1, ld1 = x.shape[3], x.shape[3]
        else:
            C2, ld2 = x2.shape[3], _check_nhwc(x2, "x2")
    hout = H // stride if hout is None else hout
    wout = W // stride if wout is None else wout
    gated = act in (UR_ACT_GEGLU, UR_ACT_GATE)
    n_out = n // 2 if gated else n
    if out is None:
        out = torch.empty((B, hout, wout, n_out), device=x.device, dtype=out_dtype)
        ret = out if squeeze == 4 else (out.squeeze(1) if squeeze == 3 else out[0, 0])
    else:
        ret = out
        out = _as4(out)
        if tuple(out.shape) != (B, hout, wout, n_out) or (out.stride(3) != 1 and n_out != 1):
            raise ValueError("out shape %s != %s" % (tuple(out.shape), (B, hout, wout, n_out)))
    if w.dtype != torch.bfloat16 or w.stride(-1) != 1:
        raise ValueError("w must be bf16 with unit K stride")
    if (w.dim() == 3) != bool(w_batched):
        raise ValueError("w must be [n, k] (shared) or [batch, n, k] (w_batched)")
    d = ConvDesc()
    d.x1, d.x2 = x.data_ptr(), (x2.data_ptr() if x2 is not None else None)
    d.c1, d.c2, d.ld1, d.ld2 = C1, C2, ld1, ld2
    d.batch, d.hin, d.win = B, H, W
    d.w, d.w_batched, d.n = w.data_ptr(), int(w_batched), n
    d.w_ld = w.stride(-2)
    d.w_bs = w.stride(0) if w_batched else 0
    d.ntaps = len(taps)
    for i, (dy, dx) in enumerate(taps):
        d.tap_dy[i], d.tap_dx[i] = dy, dx
    d.stride, d.group_kc, d.group_nc = stride, group_kc, group_nc
    d.hout, d.wout = hout, wout
    d.out = out.data_ptr()
    d.out_dtype = UR_DT_F32 if out.dtype == torch.float32 else UR_DT_BF16
    d.out_sb, d.out_sy, d.out_sx = out.stride(0), out.stride(1), out.stride(2)
    d.alpha = alpha
    for name, t in (("bias", bias), ("rowvec", rowvec), ("chscale", chscale)):
        if t is not None and (t.dtype != torch.float32 or not t.is_contiguous()):
            raise ValueError("%s must be contiguous fp32" % name)
    d.bias = bias.data_ptr() if bias is not None else None
    if rowvec is not None:
        d.rowvec = rowvec.data_ptr()
        d.rowvec_sb = rowvec.stride(0) if (rowvec.dim() == 2 and rowvec.shape[0] > 1) else 0
    if chscale is not None:
        d.chscale = chscale.data_ptr()
        d.chscale_sb = chscale.stride(0) if (chscale.dim() == 2 and chscale.shape[0] > 1) else 0
    if residual is not None:
        r = _as4(residual)
        if r.dtype != torch.bfloat16 or tuple(r.shape) != tuple(out.shape):
            raise ValueError("residual must be bf16 with the shape of out")
        d.residual = r.data_ptr()
        d.res_sb, d.res_sy, d.res_sx = r.stride(0), r.stride(1), r.stride(2)
    d.act, d.bn = act, bn
    if (want_stats or stats is not None) and out.dtype == torch.bfloat16:
        if stats is None:
            stats = new_stats(x.device, B, n_out)
        d.stats, d.stats_ld, d.stats_off = stats.data_ptr(), n_out, 0
        ret._ur_stats = stats
    ws = _workspace(x.device)
    d.workspace, d.workspace_bytes = ws.data_ptr(), ws.numel() * 4
    check(_cabi.lib().ur_conv_gemm(C.byref(d), _stream()), "ur_conv_gemm")
    return ret


def pack_conv_weight(w: torch.Tensor) -> torch.Tensor:
    """torch conv weight [Cout, Cin/g, kh, kw] (or linear [N, K]) -> bf16 [Cout, kh*kw*Cin/g] (tap-major K)."""
    if w.dim() == 4:
        w = w.permute(0, 2, 3, 1).reshape(w.shape[0], -1)
    return w.contiguous().to(torch.bfloat16)


def pick_bn(n: int, gated: bool = False) -> int:
    return _cabi.lib().ur_conv_gemm_pick_bn(n, int(gated))


def pack_gated_weight(w2d: torch.Tensor, bias: torch.Tensor | None, bn: int):
    """Interleave the two halves of a gated projection (GEGLU / SimpleGate) per N tile so that tile
    columns [0, bn/2) hold `a` and [bn/2, bn) the matching `g` channels."""
    n = w2d.shape[0]
    half, hb = n // 2, bn // 2
    idx = torch.arange(n, device=w2d.device)
    tile, col = idx // bn, idx % bn
    src = torch.where(col < hb, tile * hb + col, half + tile * hb + (col - hb))
    wp = w2d[src].contiguous()
    bp = bias[src].contiguous() if bias is not None else None
    return wp, bp


# ----------------------------------------------------------------------------------------------- helpers
def _lib():
    return _cabi.lib()


def _geom(x):
    """(batch, pixels, channels, ld, img_stride) of an NHWC / [B,T,C] bf16 activation (channel-slice views ok)."""
    x4 = _as4(x)
    ld = _check_nhwc(x4, "x")
    B, H, W, Cc = x4.shape
    return B, H * W, Cc, ld, (x4.stride(0) if B > 1 else H * W * ld)


def _f32(t, name):
    if t is not None and (t.dtype != torch.float32 or not t.is_contiguous() or not t.is_cuda):
        raise ValueError("%s must be a contiguous CUDA fp32 tensor" % name)
    return _ptr(t)


# ----------------------------------------------------------------------------------------------- normalisation
class _StatsArena:
    """Pool of pre-zeroed fp64 statistics buffers: ONE fill per DDIM step instead of one memset node per GroupNorm
    (108 per step).  The used prefix (high-water mark of the previous pass) is what gets zeroed.  One pool per
    (device, role): the Controller runs on a side stream."""

    def __init__(self):
        self.buf, self.off, self.hwm, self.active = None, 0, 0, False


_ARENAS = {}


def _arena_for(device):
    key = (device, _ROLE[0])
    a = _ARENAS.get(key)
    if a is None:
        a = _ARENAS[key] = _StatsArena()
    return a


def stats_arena_begin(device, role="main", capacity=32 << 20):
    """Enter execution context ``role`` (selects the split-K workspace and statistics pool) and zero the pool."""
    _ROLE[0] = role
    a = _arena_for(device)
    if a.buf is None:
        a.buf, a.hwm = torch.empty(capacity // 8, dtype=torch.float64, device=device), 0
    n = a.hwm if a.hwm else a.buf.numel()
    a.buf[:n].zero_()
    a.off, a.active = 0, True


def stats_arena_end(device):
    a = _arena_for(device)
    # the mark never exceeds the pool: requests past the end of the pool keep taking fresh buffers on every pass
    a.hwm, a.active = min(max(a.hwm, a.off), a.buf.numel()), False
    _ROLE[0] = "main"


def new_stats(device, B, channels):
    """Zeroed fp64 ``[B, channels, 2]`` statistics buffer: from the per-step pool when one is active, else a fresh fill."""
    a, n = _arena_for(device), B * channels * 2
    if a.active:
        if a.off + n <= min(a.hwm if a.hwm else a.buf.numel(), a.buf.numel()):
            st = a.buf[a.off:a.off + n].view(B, channels, 2)
            a.off += n
            return st
        a.off += n              # grows the high-water mark for the next pass
    return torch.zeros((B, channels, 2), device=device, dtype=torch.float64)


def chan_stats(x, stats=None, offset=0, total_channels=None, zero=True):
    """fp64 (sum, sumsq) per (image, channel) -> ``[B, total_channels, 2]``."""
    B, P, Cc, ld, ist = _geom(x)
    total = total_channels or Cc
    if stats is None:
        a, n = _arena_for(x.device), B * total * 2
        # the zeroed prefix is [0, hwm) (the whole pool on the first pass)
        if a.active and a.off + n <= min(a.hwm if a.hwm else a.buf.numel(), a.buf.numel()):
            stats, zero = a.buf[a.off:a.off + n].view(B, total, 2), False
            a.off += n
        else:
            if a.active:
                a.off += n          # grows the high-water mark for the next pass
            stats = torch.empty((B, total, 2), device=x.device, dtype=torch.float64)
    check(_lib().ur_chan_stats(_ptr(x), ld, ist, B, P, Cc, _ptr(stats), total, offset, int(zero), _stream()),
          "ur_chan_stats")
    return stats


def norm_apply(x, stats, groups, gamma, beta, eps, silu=False, x2=None, out=None, stats2=None):
    """GroupNorm / InstanceNorm apply on ``cat(x, x2)`` (channel dim); returns a dense bf16 tensor.
    ``stats2`` given: ``stats`` covers the channels of ``x`` and ``stats2`` those of ``x2``."""
    B, P, C1, ld1, is1 = _geom(x)
    C2, ld2, is2 = 0, 0, 0
    if x2 is not None:
        B2, P2, C2, ld2, is2 = _geom(x2)
        if (B2, P2) != (B, P):
            raise ValueError("x2 shape mismatch")
    if out is None:
        out = torch.empty(tuple(x.shape[:-1]) + (C1 + C2,), device=x.device, dtype=torch.bfloat16)
    _, _, Co, ldo, iso = _geom(out)
    check(_lib().ur_norm_apply(_ptr(x), ld1, is1, C1, _ptr(x2), ld2, is2, C2, _ptr(stats), _ptr(stats2), groups, B, P,
                               _f32(gamma, "gamma"), _f32(beta, "beta"), eps, int(silu), _ptr(out), ldo, iso,
                               _stream()), "ur_norm_apply")
    return out


# Tensors up to this size take the one-launch cluster kernel (ur_group_norm).  Measured on B200 (tools/bench_norm.py,
# round 1): with 8-CTA clusters only 64 SMs stream the tensor and the kernel is 20-30 % SLOWER than ur_chan_stats +
# ur_norm_apply on every per-step shape (16-CTA clusters are co-scheduled too sparsely and lose more), so it is off
# by default; set UNIRESTORE_FUSED_GN_MB to route tensors up to that many MiB through it.
FUSED_GN_MAX_BYTES = int(os.environ.get("UNIRESTORE_FUSED_GN_MB", "0")) << 20


def group_norm(x, groups, gamma, beta, eps, silu=False, x2=None):
    """nn.GroupNorm (+SiLU) over ``cat(x, x2)``.

    Default: ``ur_chan_stats`` (per source) + ``ur_norm_apply``.  Tensors up to ``FUSED_GN_MAX_BYTES`` take ONE launch
    of the cluster kernel ``ur_group_norm`` (statistics in distributed shared memory) instead."""
    B, P, C1, ld1, is1 = _geom(x)
    C2 = x2.shape[-1] if x2 is not None else 0
    if B * P * (C1 + C2) * 2 <= FUSED_GN_MAX_BYTES and (C1 + C2) <= 8192:
        ld2, is2 = 0, 0
        if x2 is not None:
            B2, P2, C2, ld2, is2 = _geom(x2)
            if (B2, P2) != (B, P):
                raise ValueError("x2 shape mismatch")
        out = torch.empty(tuple(x.shape[:-1]) + (C1 + C2,), device=x.device, dtype=torch.bfloat16)
        _, _, _, ldo, iso = _geom(out)
        check(_lib().ur_group_norm(_ptr(x), ld1, is1, C1, _ptr(x2), ld2, is2, C2, groups, B, P, _f32(gamma, "gamma"),
                                   _f32(beta, "beta"), eps, int(silu), _ptr(out), ldo, iso, _stream()),
              "ur_group_norm")
        return out
    # statistics: the ones the producing GEMM epilogue accumulated (``._ur_stats``) when present, else a pass here
    st1 = getattr(x, "_ur_stats", None)
    if x2 is None:
        return norm_apply(x, st1 if st1 is not None else chan_stats(x), groups, gamma, beta, eps, silu)
    st2 = getattr(x2, "_ur_stats", None)
    if st1 is None and st2 is None:
        stats = chan_stats(x, total_channels=C1 + C2)
        chan_stats(x2, stats=stats, offset=C1, total_channels=C1 + C2, zero=False)
        return norm_apply(x, stats, groups, gamma, beta, eps, silu, x2)
    return norm_apply(x, st1 if st1 is not None else chan_stats(x), groups, gamma, beta, eps, silu, x2,
                      stats2=st2 if st2 is not None else chan_stats(x2))


def layernorm(x, gamma, beta, eps):
    B, P, Cc, ld, ist = _geom(x)
    if B > 1 and ist != P * ld:
        raise ValueError("layernorm needs a dense token matrix")
    out = torch.empty(x.shape, device=x.device, dtype=torch.bfloat16)
    check(_lib().ur_layernorm(_ptr(x), ld, _ptr(out), Cc, B * P, Cc, _f32(gamma, "gamma"), _f32(beta, "beta"), eps,
                              _stream()), "ur_layernorm")
    return out


def scale_channels_(x, scale):
    B, P, Cc, ld, ist = _geom(x)
    check(_lib().ur_scale_channels(_ptr(x), ld, ist, B, P, Cc, _f32(scale, "scale"), scale.shape[-1], _stream()),
          "ur_scale_channels")
    return x


# ----------------------------------------------------------------------------------------------- attention (unfused)
def softmax_rows(scores, n_valid, n_pad):
    rows = scores.numel() // scores.shape[-1]
    probs = torch.empty(tuple(scores.shape[:-1]) + (n_pad,), device=scores.device, dtype=torch.bfloat16)
    check(_lib().ur_softmax_rows(_ptr(scores), scores.stride(-2), _ptr(probs), n_pad, rows, n_valid, n_pad, _stream()),
          "ur_softmax_rows")
    return probs


def transpose_tokens(x, tokens_pad=None):
    """bf16 [B, T, d] (channel-slice view ok) -> dense [B, d, Tpad] with zero padding."""
    B, T, d = x.shape
    tokens_pad = tokens_pad or ((T + 7) // 8) * 8
    out = torch.empty((B, d, tokens_pad), device=x.device, dtype=torch.bfloat16)
    check(_lib().ur_transpose_tokens(_ptr(x), x.stride(1), x.stride(0), B, T, d, _ptr(out), tokens_pad, _stream()),
          "ur_transpose_tokens")
    return out


def attention(q, k, v, heads, out=None):
    """softmax(q k^T / sqrt(d)) v per head.  q [B,Tq,C], k/v [B or 1,Tk,C] bf16 (channel-slice views ok).

    head_dim 64 / 128 / 512 (the VAE's single 512-wide head): one launch of a fused tcgen05 flash-attention kernel;
    other head dims (reduced test topologies only) take the unfused GEMM -> softmax -> GEMM path."""
    B, Tq, Cc = q.shape
    d = Cc // heads
    if d not in (64, 128, 512):
        return attention_unfused(q, k, v, heads, out)
    Tk = k.shape[1]
    shared = k.shape[0] == 1 and B > 1
    if out is None:
        out = torch.empty((B, Tq, Cc), device=q.device, dtype=torch.bfloat16)
    for t in (q, k, v, out):
        if t.dtype != torch.bfloat16 or t.stride(-1) != 1:
            raise ValueError("attention operands must be bf16 with unit channel stride")
    check(_lib().ur_attention(_ptr(q), q.stride(1), q.stride(0), _ptr(k), k.stride(1), k.stride(0), _ptr(v), v.stride(1),
                              v.stride(0), _ptr(out), out.stride(1), out.stride(0), B, heads, d, Tq, Tk, int(shared),
                              float(d) ** -0.5, _stream()), "ur_attention")
    return out


def attention_unfused(q, k, v, heads, out=None):
    """Per head S = QK^T (ur_conv_gemm, fp32) -> ur_softmax_rows -> O = P V^T (ur_conv_gemm)."""
    B, Tq, Cc = q.shape
    Tk = k.shape[1]
    d = Cc // heads
    shared = k.shape[0] == 1 and B > 1          # one K/V for every image (constant prompt, base_model.py:221)
    if out is None:
        out = torch.empty((B, Tq, Cc), device=q.device, dtype=torch.bfloat16)
    tk_pad = ((Tk + 7) // 8) * 8
    scale = float(d) ** -0.5
    for h in range(heads):
        sl = slice(h * d, (h + 1) * d)
        kh = k[0, :, sl] if shared else k[..., sl]
        s = conv_gemm(q[..., sl], kh, Tk, w_batched=not shared, alpha=scale, out_dtype=torch.float32)
        p = softmax_rows(s, Tk, tk_pad)
        vt = transpose_tokens(v[..., sl], tk_pad)
        conv_gemm(p, vt[0] if shared else vt, d, w_batched=not shared, out=out[..., sl])
    return out


# ----------------------------------------------------------------------------------------------- CFRM / TFA helpers
def dwconv3x3_gate(x, weight9, bias):
    """x bf16 [B,H,W,2c] dense -> (y bf16 [B,H,W,c], stats fp64 [B,c,2] with the per-channel sums of y)."""
    B, H, W, C2 = x.shape
    if not x.is_contiguous():
        raise ValueError("dwconv3x3_gate needs a dense input")
    c = C2 // 2
    y = torch.empty((B, H, W, c), device=x.device, dtype=torch.bfloat16)
    stats = torch.empty((B, c, 2), device=x.device, dtype=torch.float64)
    check(_lib().ur_dwconv3x3_gate(_ptr(x), B, H, W, c, _f32(weight9, "weight"), _f32(bias, "bias"), _ptr(y),
                                   _ptr(stats), _stream()), "ur_dwconv3x3_gate")
    return y, stats


ACT_IDS = {None: 0, "silu": 1, "gelu": 2, "tanh": 3}


def small_linear(x, w, bias=None, groups=1, act_in=None, act_out=None, stats_scale=None):
    """Tiny fp32 linear ``[B,K] -> [B,N]``; ``stats_scale`` set: x is a chan_stats array, input = sum * stats_scale."""
    n, k = w.shape
    if stats_scale is not None:
        B, x_ld, mode, sc = x.shape[0], x.shape[1], 1, float(stats_scale)
    else:
        _f32(x, "x")
        B, x_ld, mode, sc = x.shape[0], x.shape[1], 0, 1.0
    y = torch.empty((B, n), device=w.device, dtype=torch.float32)
    check(_lib().ur_small_linear(_ptr(x), mode, sc, x_ld, _f32(w, "w"), _f32(bias, "bias"), _ptr(y), n, B, n, k,
                                 groups, ACT_IDS[act_in], ACT_IDS[act_out], _stream()), "ur_small_linear")
    return y


def timestep_embedding(timesteps, dim):
    if timesteps.dtype != torch.int64 or not timesteps.is_cuda:
        raise ValueError("timesteps must be a CUDA int64 tensor")
    B = timesteps.numel()
    out = torch.empty((B, dim), device=timesteps.device, dtype=torch.float32)
    check(_lib().ur_timestep_embedding(_ptr(timesteps.contiguous()), B, dim, _ptr(out), _stream()),
          "ur_timestep_embedding")
    return out


def adanaf_scales(stats, pixels, groups, w_intra, b_intra, w_inter, b_inter):
    B, C4 = stats.shape[0], stats.shape[1]
    scale = torch.empty((B, C4), device=stats.device, dtype=torch.float32)
    check(_lib().ur_adanaf_scales(_ptr(stats), pixels, B, C4, groups, _f32(w_intra, "w_intra"),
                                  _f32(b_intra, "b_intra"), _f32(w_inter, "w_inter"), _f32(b_inter, "b_inter"),
                                  _ptr(scale), _stream()), "ur_adanaf_scales")
    return scale


def tfa_gates(stats, pixels, cond, w_out, b_out, w_pt=None, b_pt=None):
    B, T, D = cond.shape
    o = torch.empty((B, D), device=cond.device, dtype=torch.float32)
    nxt = torch.empty((B, T, D // 2), device=cond.device, dtype=torch.float32) if w_pt is not None else None
    check(_lib().ur_tfa_gates(_ptr(stats), pixels, B, T, D, _f32(cond, "cond"), _f32(w_out, "w_out"),
                              _f32(b_out, "b_out"), _f32(w_pt, "w_pt"), _f32(b_pt, "b_pt"), _ptr(o), _ptr(nxt),
                              _stream()), "ur_tfa_gates")
    return o, nxt


# ----------------------------------------------------------------------------------------------- latents / images
def posterior_sample(moments, noise, scaling_factor):
    """moments fp32 [B,h,w,8] channels-last, noise fp32 [B,4,h,w] -> (z fp32 [B,4,h,w], z8 bf16 [B,h,w,8])."""
    B, h, w, _ = moments.shape
    z = torch.empty((B, 4, h, w), device=moments.device, dtype=torch.float32)
    z8 = torch.empty((B, h, w, 8), device=moments.device, dtype=torch.bfloat16)
    check(_lib().ur_posterior_sample(_f32(moments, "moments"), _f32(noise, "noise"), scaling_factor, B, h * w, _ptr(z),
                                     _ptr(z8), _stream()), "ur_posterior_sample")
    return z, z8


def latent_axpby(x, a, y=None, b=0.0, want_out=True, want_nhwc8=False, scale8=1.0):
    B, _, h, w = x.shape
    out = torch.empty_like(x) if want_out else None
    out8 = torch.empty((B, h, w, 8), device=x.device, dtype=torch.bfloat16) if want_nhwc8 else None
    check(_lib().ur_latent_axpby(_f32(x, "x"), a, _f32(y, "y"), b, B, h * w, _ptr(out), _ptr(out8), scale8, _stream()),
          "ur_latent_axpby")
    return out, out8


def ddim_step_(x, eps_nhwc, coefs, clip_sample=False, want_nhwc8=True):
    """In-place DDIM update of fp32 NCHW latents ``x``; ``eps_nhwc`` fp32 [B,h,w,ld>=4]."""
    B, _, h, w = x.shape
    x8 = torch.empty((B, h, w, 8), device=x.device, dtype=torch.bfloat16) if want_nhwc8 else None
    sa, sb, sap, sbp = coefs
    check(_lib().ur_ddim_step(_f32(x, "x"), _f32(eps_nhwc, "eps"), eps_nhwc.shape[-1], sa, sb, sap, sbp,
                              int(clip_sample), B, h * w, _ptr(x8), _stream()), "ur_ddim_step")
    return x8


def image_to_nhwc8(img, a=1.0, b=0.0):
    """fp32 [B,C<=8,H,W] (any strides) -> bf16 [B,H,W,8] = a*img+b, zero channel padding."""
    if img.dtype != torch.float32 or not img.is_cuda:
        raise ValueError("img must be a CUDA fp32 tensor")
    B, Cc, H, W = img.shape
    out = torch.empty((B, H, W, 8), device=img.device, dtype=torch.bfloat16)
    check(_lib().ur_image_to_nhwc8(_ptr(img), img.stride(0), img.stride(1), img.stride(2), img.stride(3), B, Cc, H, W,
                                   a, b, _ptr(out), _stream()), "ur_image_to_nhwc8")
    return out


def nhwc_to_image(src, channels, h=None, w=None, a=1.0, b=0.0, y0=0, x0=0, quantize=False):
    """fp32 channels-last [B,Hs,Ws,ld] -> fp32 NCHW [B,channels,h,w] = a*src[:, y0:y0+h, x0:x0+w]+b;
    ``quantize``: the 8-bit quantisation of eval_image_restoration.py:71 fused into the write-out."""
    B, Hs, Ws, ld = src.shape
    h, w = h or Hs - y0, w or Ws - x0
    out = torch.empty((B, channels, h, w), device=src.device, dtype=torch.float32)
    check(_lib().ur_nhwc_to_image(_f32(src, "src"), ld, Hs, Ws, B, channels, h, w, y0, x0, a, b, int(quantize), _ptr(out),
                                  _stream()), "ur_nhwc_to_image")
    return out


def concat_channels(x1, x2):
    """Dense bf16 ``cat(x1, x2)`` along channels (one copy kernel; channel-slice views accepted as sources)."""
    x1, x2 = _as4(x1), _as4(x2)
    ld1, ld2 = _check_nhwc(x1, "x1"), _check_nhwc(x2, "x2")
    B, H, W, C1 = x1.shape
    C2 = x2.shape[3]
    for t, ld in ((x1, ld1), (x2, ld2)):
        if B > 1 and t.stride(0) != H * W * ld:
            raise ValueError("concat_channels needs a uniform pixel pitch")
    out = torch.empty((B, H, W, C1 + C2), device=x1.device, dtype=torch.bfloat16)
    check(_lib().ur_concat_channels(_ptr(x1), ld1, C1, _ptr(x2), ld2, C2, B * H * W, _ptr(out), _stream()),
          "ur_concat_channels")
    return out


def resize_pad(img, size=None, pad_bottom=0, pad_right=0, quantize=False):
    """``F.pad(F.interpolate(img, size, mode="bicubic", align_corners=False), (0, pad_right, 0, pad_bottom), "reflect")``
    on a CUDA fp32 NCHW image in one kernel (unifie.py:124-134,165-168); ``size=None`` keeps the resolution."""
    if img.dtype != torch.float32 or not img.is_cuda or img.dim() != 4:
        raise ValueError("img must be a CUDA fp32 [B,C,H,W] tensor")
    B, Cc, H, W = img.shape
    hr, wr = (H, W) if size is None else (int(size[0]), int(size[1]))
    out = torch.empty((B, Cc, hr + pad_bottom, wr + pad_right), device=img.device, dtype=torch.float32)
    check(_lib().ur_resize_pad(_ptr(img), img.stride(0), img.stride(1), img.stride(2), img.stride(3), B, Cc, H, W, hr, wr,
                               pad_bottom, pad_right, int(quantize), _ptr(out), _stream()), "ur_resize_pad")
    return out


# every op that launches kernels runs with its first tensor argument's device current (and the library initialised)
for _name in ("conv_gemm", "chan_stats", "norm_apply", "group_norm", "layernorm", "scale_channels_", "softmax_rows",
              "transpose_tokens", "attention", "dwconv3x3_gate", "small_linear", "timestep_embedding", "adanaf_scales",
              "tfa_gates", "posterior_sample", "latent_axpby", "ddim_step_", "image_to_nhwc8", "nhwc_to_image",
              "resize_pad", "concat_channels"):
    globals()[_name] = _on_device(globals()[_name])
del _name
