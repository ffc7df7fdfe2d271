"""Thin tensor-level wrappers over the C-ABI (``_cabi``): torch tensors in, raw pointers out.

Activations are bf16 tensors of logical shape ``[B, H, W, C]`` (channels-last physical layout;
a token matrix ``[B, T, C]`` is treated as ``H = 1, W = T``).  Channel-slice views (``stride(-1) == 1``)
are accepted as inputs and outputs so concatenations / q-k-v splits are never materialised.
Nothing here falls back to PyTorch arithmetic: every op is one launch of a hand-written kernel.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _cabi
from ._cabi import (UR_ACT_GATE, UR_ACT_GEGLU, UR_ACT_GELU, UR_ACT_NONE, UR_ACT_SILU, UR_DT_BF16, UR_DT_F32,
                    ConvDesc, check)

TAPS_1x1 = ((0, 0),)
TAPS_3x3 = tuple((ky - 1, kx - 1) for ky in range(3) for kx in range(3))          # zero padding 1
TAPS_3x3_NOPAD = tuple((ky, kx) for ky in range(3) for kx in range(3))             # VAE asym pad (0,1,0,1)


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


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


def conv_gemm(x, w, n, *, x2=None, taps=TAPS_1x1, stride=1, hout=None, wout=None, bias=None, rowvec=None,
              chscale=None, residual=None, act=UR_ACT_NONE, alpha=1.0, out=None, out_dtype=torch.bfloat16,
              group_kc=0, group_nc=0, w_batched=False, bn=0):
    """``out = epilogue(alpha * conv(x [cat x2], w))`` -- see ``ur_conv_gemm`` in include/unirestore_b200.h.

    ``w`` is the packed bf16 weight ``[n, ntaps * kc]`` (``[batch, n, k]`` when ``w_batched``).
    ``out`` may be a strided NHWC view (channel slice, or ``full[:, py::2, px::2]`` for sub-pixel phases).
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
    if w.dtype != torch.bfloat16 or not w.is_contiguous():
        raise ValueError("w must be contiguous bf16")
    d = ConvDesc()
    d.x1, d.x2 = x.data_ptr(), (x2.data_ptr() if x2 is not None else None)
    d.c1, d.c2, d.ld1, d.ld2 = C1, C2, ld1, ld2
    d.batch, d.hin, d.win = B, H, W
    d.w, d.w_batched, d.n = w.data_ptr(), int(w_batched), n
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
    _cabi.ensure_init(x.device.index or 0)
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
