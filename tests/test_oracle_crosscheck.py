"""Independent cross-checks of oracle leaves whose arithmetic lives in the un-vendored ``diffusers`` (no golden can be
generated offline): the same math through a DIFFERENT implementation that ships with PyTorch.

* ``Attention`` (token form, self and cross) vs ``torch.nn.MultiheadAttention`` with the projections mapped onto
  ``in_proj`` / ``q,k,v_proj_weight`` / ``out_proj`` (1/sqrt(d) scaling, per-head split, softmax over keys).
* ``GEGLU`` / ``FeedForward`` vs an explicit chunk + exact-erf GELU composition.
* ``Timesteps`` sinusoid vs a float64 numpy evaluation of the published formula.
"""
import math

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from oracle import blocks as OB


def test_attention_matches_nn_multihead_attention():
    torch.manual_seed(0)
    C, heads = 64, 4
    att = OB.Attention(C, heads=heads, dim_head=C // heads, bias=True).eval()
    mha = nn.MultiheadAttention(C, heads, bias=True, batch_first=True).eval()
    with torch.no_grad():
        mha.in_proj_weight.copy_(torch.cat([att.to_q.weight, att.to_k.weight, att.to_v.weight]))
        mha.in_proj_bias.copy_(torch.cat([att.to_q.bias, att.to_k.bias, att.to_v.bias]))
        mha.out_proj.weight.copy_(att.to_out[0].weight)
        mha.out_proj.bias.copy_(att.to_out[0].bias)
        x = torch.randn(2, 37, C)
        ref, _ = mha(x, x, x, need_weights=False)
        assert torch.allclose(att(x), ref, atol=2e-6, rtol=1e-5)


def test_cross_attention_matches_nn_multihead_attention():
    torch.manual_seed(1)
    C, Cc, heads = 64, 96, 4
    att = OB.Attention(C, cross_attention_dim=Cc, heads=heads, dim_head=C // heads, bias=False).eval()
    mha = nn.MultiheadAttention(C, heads, bias=True, batch_first=True, kdim=Cc, vdim=Cc).eval()
    with torch.no_grad():
        mha.q_proj_weight.copy_(att.to_q.weight)
        mha.k_proj_weight.copy_(att.to_k.weight)
        mha.v_proj_weight.copy_(att.to_v.weight)
        mha.in_proj_bias.zero_()
        mha.out_proj.weight.copy_(att.to_out[0].weight)
        mha.out_proj.bias.copy_(att.to_out[0].bias)
        x, ctx = torch.randn(3, 20, C), torch.randn(3, 77, Cc)
        ref, _ = mha(x, ctx, ctx, need_weights=False)
        assert torch.allclose(att(x, ctx), ref, atol=2e-6, rtol=1e-5)


def test_geglu_feed_forward_composition():
    torch.manual_seed(2)
    ff = OB.FeedForward(32).eval()
    x = torch.randn(5, 7, 32)
    with torch.no_grad():
        proj = ff.net[0].proj
        a, g = F.linear(x, proj.weight, proj.bias).chunk(2, dim=-1)
        gelu = 0.5 * g * (1.0 + torch.erf(g / math.sqrt(2.0)))                  # exact-erf GELU written out
        ref = F.linear(a * gelu, ff.net[2].weight, ff.net[2].bias)
        assert torch.allclose(ff(x), ref, atol=1e-6, rtol=1e-5)


def test_timesteps_sinusoid_formula():
    ts = torch.tensor([999, 499, 0, 37])
    emb = OB.Timesteps(320, True, 0)(ts).double().numpy()
    i = np.arange(160, dtype=np.float64)
    f = np.exp(-math.log(10000.0) * i / 160.0)
    ref = np.concatenate([np.cos(ts.numpy()[:, None] * f), np.sin(ts.numpy()[:, None] * f)], axis=1)   # cos first
    assert np.abs(emb - ref).max() < 2e-4          # fp32 evaluation of t * f at t ~ 1000
