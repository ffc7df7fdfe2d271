"""bf16-rounding-point oracle: the reference's algorithm with values rounded to bf16 exactly where the CUDA path stores them.

TEST INFRASTRUCTURE (oracle) -- never imported by the product path.

``oracle/blocks.py`` + ``oracle/unirestore.py`` restate the reference in fp32 (the reference's own ``precision: 32``
option, configs/val.yaml:11).  The CUDA path stores every activation as bf16 between kernels (the reference's
``bf16-mixed`` option, val.yaml:12), accumulates in fp32 and keeps statistics / latents / scheduler state in fp32.
Against the fp32 oracle a correct bf16 pipeline differs by ~3e-3..1e-2 rel-L2 per block, which would hide a wrong
epsilon or a dropped bias.  The functions below therefore walk the SAME oracle modules (same parameters, same
state_dict) and apply the same arithmetic, but round to bf16 at the product's kernel boundaries ("identical rounding
points", BASELINE.md section 4 / SURVEY.md section 8c(10)):

  * weights of every tensor-core GEMM / conv are rounded to bf16 (fp32 bias, fp32 accumulate);
  * one rounding per kernel OUTPUT: ``q(acc + bias [+ temb] [act] [* scale] [+ residual])`` -- the fused epilogue
    rounds once, after the residual add; GEGLU / SimpleGate round ``a * act(g)`` once;
  * GroupNorm / LayerNorm / InstanceNorm: fp32 statistics of the (already rounded) input, output rounded once,
    SiLU fused before the rounding;
  * attention: fp32 scores, probabilities rounded to bf16 before ``P V`` (un-normalised in the flash kernel; the row
    sum uses the unrounded fp32 values), output rounded once;
  * tiny fp32 paths stay fp32: time-embedding MLPs and ``time_emb_proj``, NAFBlock SCA, AdaNAF group attention,
    TFA gates, posterior sampling, DDIM / DDPM scalar math, latents.

With rounding switched off (``with exact():``) every function here must reproduce the fp32 oracle module it mirrors to
float rounding -- ``tests/test_oracle_rounded.py`` checks that on the CPU, so this file is pinned to the oracle (and
through it to the reference goldens) and cannot drift from it.  Each function cites the reference file:line of the
module it follows and the product function whose rounding points it mirrors.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

_ON = [True]


class exact:
    """``with exact():`` -- no rounding: the functions reduce to the fp32 oracle arithmetic."""

    def __enter__(self):
        self.prev, _ON[0] = _ON[0], False

    def __exit__(self, *a):
        _ON[0] = self.prev


def q(t):
    """One bf16 store of the CUDA path (round-to-nearest-even), kept in fp32 for the next op."""
    return t.to(torch.bfloat16).to(torch.float32) if _ON[0] else t


# ----------------------------------------------------------------------------------------- leaves
def conv(m, x, weight=None):
    """Tensor-core conv / linear accumulator: bf16 weights, fp32 accumulate, NO bias (added in the epilogue)."""
    w = q(m.weight if weight is None else weight)
    if w.dim() == 2:
        return F.linear(x, w)
    return F.conv2d(x, w, None, m.stride, m.padding, m.dilation, m.groups)


def bias4(m):
    return 0.0 if m.bias is None else m.bias.view(1, -1, 1, 1)


def gn(m, x, silu=False):
    y = F.group_norm(x, m.num_groups, m.weight, m.bias, m.eps)
    return q(F.silu(y) if silu else y)


def ln(m, x):
    return q(F.layer_norm(x, m.normalized_shape, m.weight, m.bias, m.eps))


def sdpa(qh, kh, vh, flash=True):
    """[B, heads, T, d] fp32 (bf16-representable) -> [B, heads, Tq, d]; the rounding points of ur_attention
    (ur_attention.cu: P = exp2(s*scale*log2e - m) rounded to bf16 for P V, row sum from the fp32 values, O / l
    rounded once) or, for head dims without a fused kernel, of the GEMM -> ur_softmax_rows -> GEMM path
    (normalised probabilities rounded)."""
    s = torch.matmul(qh, kh.transpose(-1, -2)) * (qh.shape[-1] ** -0.5)
    p = torch.exp(s - s.amax(-1, keepdim=True))
    l = p.sum(-1, keepdim=True)
    if flash:
        return q(torch.matmul(q(p), vh) / l)
    return q(torch.matmul(q(p / l), vh))


def _heads(t, h):
    b, n, c = t.shape
    return t.view(b, n, h, c // h).transpose(1, 2)


# head dims served by the fused flash kernel (ur_attention); others take GEMM -> ur_softmax_rows -> GEMM
FLASH_HEAD_DIMS = {64, 128, 512}


def _fused_attention(dim_head):
    return dim_head in FLASH_HEAD_DIMS


def attention_tokens(m, x, ctx=None, residual=None):
    """diffusers Attention on [B, T, C] tokens (BasicTransformerBlock attn1 / attn2, base_model.py:138,159,191);
    mirrors sd_blocks.Attention.run_tokens: q/k/v projections rounded, SDPA, ``q(to_out + bias + residual)``."""
    src = x if ctx is None else ctx
    bq = 0.0 if m.to_q.bias is None else m.to_q.bias
    bk = 0.0 if m.to_k.bias is None else m.to_k.bias
    bv = 0.0 if m.to_v.bias is None else m.to_v.bias
    qq, kk, vv = q(conv(m.to_q, x) + bq), q(conv(m.to_k, src) + bk), q(conv(m.to_v, src) + bv)
    a = sdpa(_heads(qq, m.heads), _heads(kk, m.heads), _heads(vv, m.heads), _fused_attention(m.dim_head))
    a = a.transpose(1, 2).reshape(x.shape[0], -1, m.heads * m.dim_head)
    o = conv(m.to_out[0], a) + (0.0 if m.to_out[0].bias is None else m.to_out[0].bias)
    return q(o if residual is None else o + residual)


def attention_spatial(m, x):
    """Spatial self-attention block (Controller AttnDownBlock2D / mid block controller.py:101-141, VAE mid block):
    GN -> tokens -> attention -> + x; mirrors sd_blocks.Attention.run."""
    b, c, h, w = x.shape
    xn = gn(m.group_norm, x) if m.group_norm is not None else x
    tok = lambda t: t.flatten(2).transpose(1, 2)
    y = attention_tokens(m, tok(xn), residual=tok(x) if m.residual_connection else None)
    return y.transpose(1, 2).reshape(b, c, h, w)


def resnet(m, x, temb=None):
    """diffusers ResnetBlock2D (base_model.py:54, controller.py:161-170); mirrors sd_blocks.ResnetBlock2D.run.
    ``x`` may already be the channel concat of the up path (the product never materialises it, same values)."""
    h = gn(m.norm1, x, silu=True)
    t = 0.0
    if temb is not None and m.time_emb_proj is not None:
        t = m.time_emb_proj(F.silu(temb))[:, :, None, None]              # fp32 weights (ur_small_linear)
    h = q(conv(m.conv1, h) + bias4(m.conv1) + t)
    h = gn(m.norm2, h, silu=True)
    res = x if m.conv_shortcut is None else q(conv(m.conv_shortcut, x) + bias4(m.conv_shortcut))
    return q(conv(m.conv2, h) + bias4(m.conv2) + res)


def downsample(m, x):
    if m.padding == 0:
        x = F.pad(x, (0, 1, 0, 1))
    return q(conv(m.conv, x) + bias4(m.conv))


def upsample(m, x):
    """nearest x2 + conv3x3 (base_model.py:202-203, autoencoder.py:60).  The product runs four sub-pixel 2x2
    convolutions whose taps are the fp32 SUMS of the 3x3 taps hitting the same source pixel, rounded to bf16 after
    the sum (sd_blocks.Upsample2D._pack): mirrored here so the weight rounding is identical."""
    w = m.conv.weight
    B, _, H, W = x.shape
    out = x.new_empty(B, w.shape[0], 2 * H, 2 * W)
    sel = {0: ((0,), (1, 2)), 1: ((0, 1), (2,))}
    xp = F.pad(x, (1, 1, 1, 1))
    for py in (0, 1):
        for px in (0, 1):
            acc = 0.0
            for ry in (0, 1):
                for rx in (0, 1):
                    wm = q(sum(w[:, :, ky, kx] for ky in sel[py][ry] for kx in sel[px][rx]))
                    dy, dx = py - 1 + ry, px - 1 + rx
                    src = xp[:, :, 1 + dy:1 + dy + H, 1 + dx:1 + dx + W]
                    acc = acc + torch.einsum("oc,bchw->bohw", wm, src)
            out[:, :, py::2, px::2] = acc + bias4(m.conv)
    return q(out)


def feed_forward(m, x, residual):
    proj, lin2 = m.net[0].proj, m.net[2]
    a, g = (conv(proj, x) + proj.bias).chunk(2, dim=-1)
    h = q(a * F.gelu(g))
    return q(conv(lin2, h) + lin2.bias + residual)


def transformer_block(m, x, ctx):
    x = attention_tokens(m.attn1, ln(m.norm1, x), residual=x)
    x = attention_tokens(m.attn2, ln(m.norm2, x), ctx, residual=x)
    return feed_forward(m.ff, ln(m.norm3, x), x)


def transformer2d(m, x, ctx):
    """diffusers Transformer2DModel, use_linear_projection=True; mirrors sd_blocks.Transformer2DModel.run."""
    b, c, h, w = x.shape
    tok = lambda t: t.flatten(2).transpose(1, 2)
    t = q(conv(m.proj_in, tok(gn(m.norm, x))) + m.proj_in.bias)
    for blk in m.transformer_blocks:
        t = transformer_block(blk, t, ctx)
    y = q(conv(m.proj_out, t) + m.proj_out.bias + tok(x))
    return y.transpose(1, 2).reshape(b, c, h, w)


# ----------------------------------------------------------------------------------------- reference-owned modules
def scedit(m, x, c):
    """CSCEAdapter scedit.py:24-38; mirrors diffuie/scedit.py (three fused-epilogue GEMMs)."""
    s = q(conv(m.proj, c) + bias4(m.proj) + x)
    h = q(F.gelu(conv(m.tuner[0], s) + bias4(m.tuner[0])))
    return q(conv(m.tuner[2], h) + bias4(m.tuner[2]) + s)


def _ln2d(m, x):
    return q(F.layer_norm(x.permute(0, 2, 3, 1), m.normalized_shape, m.weight, m.bias, m.eps).permute(0, 3, 1, 2))


def nafblock(m, x):
    """NAFBlock nafnet_arch.py:28-131; mirrors diffuie/nafnet_arch.py run()."""
    t = q(conv(m.conv1, _ln2d(m.norm1, x)) + bias4(m.conv1))
    d = F.conv2d(t, m.conv2.weight, m.conv2.bias, 1, 1, 1, m.conv2.groups)          # depthwise, fp32 weights
    a, b = d.chunk(2, dim=1)
    g = q(a * b)
    s = F.conv2d(g.mean((2, 3), keepdim=True), m.sca[1].weight, m.sca[1].bias)          # SCA on the pooled vector
    g = q(g * s)
    y = q((conv(m.conv3, g) + bias4(m.conv3)) * m.beta + x)
    a, b = (conv(m.conv4, _ln2d(m.norm2, y)) + bias4(m.conv4)).chunk(2, dim=1)
    u = q(a * b)
    return q((conv(m.conv5, u) + bias4(m.conv5)) * m.gamma + y)


def adanaf(m, x):
    """AdaNAFV2 cfrm.py:12-54; mirrors diffuie/cfrm.py run() (the two group-attention scales are applied in ONE
    pass: GAP(u * intra) = intra * GAP(u))."""
    t = q(conv(m.conv_in, x) + bias4(m.conv_in))
    t = gn(m.group_norm, t)
    u = q(F.gelu(conv(m.group_conv, t) + bias4(m.group_conv)))
    pooled = u.mean((2, 3), keepdim=True)
    intra = m.intra_group_attn[1](pooled)
    inter = m.inter_group_attn[1](pooled * intra)                                       # [B, g, 1, 1]
    scale = intra * inter.repeat_interleave(u.shape[1] // m.groups, dim=1)
    u = q(u * scale)
    y = q(conv(m.pwconv, u) + bias4(m.pwconv) + x)
    return nafblock(m.nafblock, y)


def tfa(m, x, skip, cond):
    """TaskFeatureAdapter taskeditor.py:70-108; mirrors diffuie/taskeditor.py run()."""
    b = skip.shape[0]
    shape = (b, m.prompt_len, m.prompt_dim)
    sn = q(F.instance_norm(skip, eps=m.filter_gate[0].eps))

    def branch(br):
        h = q(F.gelu(conv(br[1], sn) + bias4(br[1])))
        return q(conv(br[3], h) + bias4(br[3])).mean((2, 3)).reshape(shape)

    f = F.softmax(branch(m.filter_gate), dim=-1)
    i = F.softmax(branch(m.info_gate), dim=-1)
    cval = torch.tanh(branch(m.content_trans))
    c2 = f * cond + i * cval
    o = m.out_gate(c2.reshape(b, -1))[:, :, None, None]
    tg = q((conv(m.t_gate1, skip) + bias4(m.t_gate1)) * o)
    skip2 = q(conv(m.t_gate2, tg) + bias4(m.t_gate2) + skip)
    x = q(conv(m.conv_out, torch.cat([x, skip2], 1)) + bias4(m.conv_out) + x)
    return x, (None if m.last_layer else m.prompt_trans(c2))


def controller(m, z, timesteps):
    """Controller.forward controller.py:193-220; mirrors diffuie/controller.py run()."""
    emb = m.time_embedding(m.time_proj(timesteps))                                       # fp32 (ur_small_linear)
    h = q(conv(m.conv_in, q(z)) + bias4(m.conv_in))
    taps = []
    for blk in m.down_blocks:
        outs = []
        attns = getattr(blk, "attentions", None)
        for i, r in enumerate(blk.resnets):
            h = resnet(r, h, emb)
            if attns is not None:
                h = attention_spatial(attns[i], h)
            outs.append(h)
        if blk.downsamplers is not None:
            h = downsample(blk.downsamplers[0], h)
            outs.append(h)
        taps.append(outs[-2])
    mb = m.middle_block
    h = resnet(mb.resnets[0], h, emb)
    taps[-1] = resnet(mb.resnets[1], attention_spatial(mb.attentions[0], h), emb)
    return {t.size(-1): resnet(m.fea_tran[i], t, emb) for i, t in enumerate(taps)}


def controlled_unet(m, zt, control, timesteps):
    """ControlledUNet.forward base_model.py:211-245 (+ :94-209); mirrors diffuie/base_model.py run()."""
    u = m.unet
    ctx = q(m.null_embeds).expand(zt.shape[0], -1, -1)
    emb = u.time_embedding(u.time_proj(timesteps).to(zt.dtype))
    x = q(conv(u.conv_in, q(zt)) + bias4(u.conv_in))
    skips = [x]
    for blk in u.down_blocks:
        attns = blk.attentions if getattr(blk, "has_cross_attention", False) else [None] * len(blk.resnets)
        for r, a in zip(blk.resnets, attns):
            x = resnet(r, x, emb)
            if a is not None:
                x = transformer2d(a, x, ctx)
            skips.append(x)
        if blk.downsamplers is not None:
            x = downsample(blk.downsamplers[0], x)
            skips.append(x)
    x = resnet(u.mid_block.resnets[0], x, emb)
    x = resnet(u.mid_block.resnets[1], transformer2d(u.mid_block.attentions[0], x, ctx), emb)
    for i, ed in enumerate(m.csc_editors):
        skips[i] = scedit(ed, skips[i], control[skips[i].shape[-1]])
    for blk in u.up_blocks:
        attns = blk.attentions if getattr(blk, "has_cross_attention", False) else [None] * len(blk.resnets)
        for r, a in zip(blk.resnets, attns):
            x = resnet(r, torch.cat([x, skips.pop()], 1), emb)
            if a is not None:
                x = transformer2d(a, x, ctx)
        if blk.upsamplers is not None:
            x = upsample(blk.upsamplers[0], x)
    return conv(u.conv_out, gn(u.conv_norm_out, x, silu=True)) + bias4(u.conv_out)    # fp32 eps


def _vae_mid(mb, x):
    x = resnet(mb.resnets[0], x)
    return resnet(mb.resnets[1], attention_spatial(mb.attentions[0], x))


def encode(ae, images, enable_fr=False, noise=None):
    """SkipConnectedAutoEncoder.encode autoencoder.py:132-156 (+ :11-35); mirrors diffuie/autoencoder.py run_encode."""
    enc = ae.vae.encoder
    x = q(conv(enc.conv_in, q(images * 2 - 1)) + bias4(enc.conv_in))
    skips = []
    for i, blk in enumerate(enc.down_blocks[:-1]):
        for r in blk.resnets:
            x = resnet(r, x)
        if blk.downsamplers is not None:
            x = downsample(blk.downsamplers[0], x)
        if enable_fr:
            for sub in enc.fr_blocks[i]:
                x = adanaf(sub, x) if hasattr(sub, "nafblock") else nafblock(sub, x)
        skips.append(x)
    for r in enc.down_blocks[-1].resnets:
        x = resnet(r, x)
    x = _vae_mid(enc.mid_block, x)
    x = q(conv(enc.conv_out, gn(enc.conv_norm_out, x, silu=True)) + bias4(enc.conv_out))
    mom = conv(ae.vae.quant_conv, x) + bias4(ae.vae.quant_conv)                          # fp32 moments
    mean, logvar = mom.chunk(2, dim=1)
    std = torch.exp(0.5 * logvar.clamp(-30.0, 20.0))
    if noise is None:
        noise = torch.randn(mean.shape, dtype=mean.dtype, device=mean.device)
    return (mean + std * noise) * ae.vae.config.scaling_factor, skips


def decode(ae, latents, skips, task):
    """SkipConnectedAutoEncoder.decode autoencoder.py:158-176 (+ :37-72); mirrors run_decode."""
    dec = ae.vae.decoder
    z = q(latents * (1.0 / ae.vae.config.scaling_factor)) if _ON[0] else latents / ae.vae.config.scaling_factor
    z = q(conv(ae.vae.post_quant_conv, z) + bias4(ae.vae.post_quant_conv))
    x = _vae_mid(dec.mid_block, q(conv(dec.conv_in, z) + bias4(dec.conv_in)))
    cond = dec.task_prompts[task].unsqueeze(0).expand(z.shape[0], -1, -1)
    for i, blk in enumerate(dec.up_blocks):
        if i < len(dec.up_blocks) - 1:
            x, cond = tfa(dec.task_editors[i], x, skips[-i - 1], cond)
        for r in blk.resnets:
            x = resnet(r, x)
        if blk.upsamplers is not None:
            x = upsample(blk.upsamplers[0], x)
    y = conv(dec.conv_out, gn(dec.conv_norm_out, x, silu=True)) + bias4(dec.conv_out)
    return (y + 1) / 2 if not _ON[0] else y * 0.5 + 0.5


def predict_z0(m, latents, conditions, timesteps):
    """DiffUIE.predict_z0 unifie.py:91-105."""
    eps = controlled_unet(m.base_model, latents, controller(m.controller, conditions, timesteps), timesteps)
    a = m.ddpm.alphas_cumprod.to(latents.device)[timesteps].view(-1, 1, 1, 1)
    return (latents - (1 - a) ** 0.5 * eps) / a ** 0.5


def forward(m, images, task, noise=None, trace=None):
    """DiffUIE.forward unifie.py:107-169; mirrors diffuie/unifie.py _forward_impl.  ``trace``: list receiving the
    latents after every DDIM step (drift tables)."""
    org_h, org_w = images.shape[-2:]
    h, w = org_h, org_w
    if h < 512 or w < 512:
        s = 512 / min(h, w)
        h, w = round(h * s), round(w * s)
        images = F.interpolate(images, (h, w), mode="bicubic", align_corners=False, antialias=False)
    if h % 64 or w % 64:
        images = F.pad(images, (0, (64 - w % 64) % 64, 0, (64 - h % 64) % 64), mode="reflect")
    n_post, n_diff = noise if noise is not None else (None, None)
    z0, mids = encode(m.ae, images, enable_fr=m.fr_type is not None, noise=n_post)
    if m.control_type:
        t = 999 * torch.ones((len(images),), dtype=torch.long, device=images.device)
        zt, _, _ = m.diffuse(z0, t, n_diff)
        for t in m.scheduler.timesteps:
            ts = t.reshape(-1).to(images.device)
            eps = controlled_unet(m.base_model, zt, controller(m.controller, z0, ts), ts)
            zt = m.scheduler.step(eps, t, zt).prev_sample
            if trace is not None:
                trace.append(zt.clone())
    else:
        zt = z0
    preds = decode(m.ae, zt, mids, task)[..., :h, :w]
    return F.interpolate(preds, (org_h, org_w), mode="bicubic", align_corners=False, antialias=False)
