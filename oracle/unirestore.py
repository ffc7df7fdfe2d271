"""CPU oracle: independent restatement of the reference-owned UniRestore modules.

TEST INFRASTRUCTURE -- never imported by the product path (``unirestore_b200``).

Each class cites the reference file:line it follows.  State-dict key names equal the
reference's so one deterministic weight set
drives the reference files (under shims), this oracle and the CUDA path alike.
Checked against fixtures produced by executing the reference's own files
(oracle/make_golden.py -> tests/golden/, tests/test_oracle_golden.py).
"""
from __future__ import annotations

import os

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import sd_turbo_config as CFG
from .blocks import (AutoencoderKL, LayerNorm2d, ResnetBlock2D, TimestepEmbedding, Timesteps,
                     UNet2DConditionModel, UNetMidBlock2D, get_down_block)
from .schedulers import DDIMScheduler, DDPMScheduler

_ASSET = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "unirestore_b200", "assets",
                      "sd_null_emb.pt")


# ----------------------------------------------------------------------------- SC-Tuner
class CSCEAdapter(nn.Module):
    """scedit.py:24-38 -- ``p = proj(c); out = tuner(x + p) + p + x``."""

    def __init__(self, c_in, c_emb, c_cond):
        super().__init__()
        self.proj = nn.Conv2d(c_cond, c_in, 1)
        self.tuner = nn.Sequential(nn.Conv2d(c_in, c_emb, 1), nn.GELU(), nn.Conv2d(c_emb, c_in, 1))

    def forward(self, x, condition):
        p = self.proj(condition)
        s = x + p
        return self.tuner(s) + s


# ----------------------------------------------------------------------------- CFRM
class NAFBlock(nn.Module):
    """nafnet_arch.py:28-131 (SimpleGate :22-25)."""

    def __init__(self, c):
        super().__init__()
        self.conv1 = nn.Conv2d(c, 2 * c, 1)
        self.conv2 = nn.Conv2d(2 * c, 2 * c, 3, padding=1, groups=2 * c)
        self.conv3 = nn.Conv2d(c, c, 1)
        self.sca = nn.Sequential(nn.AdaptiveAvgPool2d(1), nn.Conv2d(c, c, 1))
        self.conv4 = nn.Conv2d(c, 2 * c, 1)
        self.conv5 = nn.Conv2d(c, c, 1)
        self.norm1 = LayerNorm2d(c)
        self.norm2 = LayerNorm2d(c)
        self.beta = nn.Parameter(torch.zeros(1, c, 1, 1))
        self.gamma = nn.Parameter(torch.zeros(1, c, 1, 1))

    @staticmethod
    def _gate(x):
        a, b = x.chunk(2, dim=1)
        return a * b

    def forward(self, inp):
        x = self._gate(self.conv2(self.conv1(self.norm1(inp))))
        x = self.conv3(x * self.sca(x))
        y = inp + x * self.beta
        x = self.conv5(self._gate(self.conv4(self.norm2(y))))
        return y + x * self.gamma


class AdaNAFV2(nn.Module):
    """cfrm.py:12-54."""

    def __init__(self, c):
        super().__init__()
        g, wide = 16, 4 * c
        self.groups = g
        self.conv_in = nn.Conv2d(c, wide, 1)
        self.group_norm = nn.GroupNorm(g, wide)
        self.group_conv = nn.Conv2d(wide, wide, 3, padding=1, groups=g)
        self.intra_group_attn = nn.Sequential(nn.AdaptiveAvgPool2d(1), nn.Conv2d(wide, wide, 1, groups=g))
        # index 2 of the reference Sequential is a parameter-free Rearrange (cfrm.py:29-33)
        self.inter_group_attn = nn.Sequential(nn.AdaptiveAvgPool2d(1), nn.Conv2d(wide, g, 1))
        self.pwconv = nn.Conv2d(wide, c, 1)
        self.nafblock = NAFBlock(c)

    def forward(self, inp):
        x = F.gelu(self.group_conv(self.group_norm(self.conv_in(inp))))
        x = x * self.intra_group_attn(x)
        iga = self.inter_group_attn(x)                              # [B, g, 1, 1]
        b, ch, h, w = x.shape
        x = (x.view(b, self.groups, ch // self.groups, h, w) * iga[:, :, None]).view(b, ch, h, w)
        return self.nafblock(inp + self.pwconv(x))


# ----------------------------------------------------------------------------- TFA
class TaskFeatureAdapter(nn.Module):
    """taskeditor.py:10-108."""

    def __init__(self, c_out=512, c_skip=256, prompt_len=1, last_layer=False):
        super().__init__()
        d, hid = c_skip, c_skip * prompt_len
        self.prompt_len, self.prompt_dim, self.last_layer = prompt_len, d, last_layer
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

    def forward(self, x, skip, condition):
        b = skip.shape[0]
        shape = (b, self.prompt_len, self.prompt_dim)
        f = F.softmax(self.filter_gate(skip).reshape(shape), dim=-1)
        i = F.softmax(self.info_gate(skip).reshape(shape), dim=-1)
        cval = self.content_trans(skip).reshape(shape)
        cond = f * condition + i * cval
        o = self.out_gate(cond.reshape(b, -1))[:, :, None, None]
        skip = skip + self.t_gate2(o * self.t_gate1(skip))
        x = x + self.conv_out(torch.cat([x, skip], dim=1))
        return x, (None if self.last_layer else self.prompt_trans(cond))


# ----------------------------------------------------------------------------- auto-encoder
class SkipConnectedAutoEncoder(nn.Module):
    """autoencoder.py:74-184 with the patched encoder/decoder forwards of :11-72 inlined."""

    def __init__(self, vae: AutoencoderKL, fr_type=None, tedit=None):
        super().__init__()
        self.vae = vae
        self.tedit_dict = tedit
        if fr_type == "CFRM":
            self.vae.encoder.fr_blocks = nn.ModuleList([
                nn.Sequential(*[NAFBlock(c) for _ in range(n)], AdaNAFV2(c)) for c, n in CFG.CFRM_STACKS])
        elif fr_type is not None:
            raise ValueError("Invalid fr_type")
        if tedit:
            self.task_list, self.tedit_type = tedit["task"], tedit["type"]
            if self.tedit_type != "TFA":
                raise KeyError("%s is not defined in the taskeditor!, please select ['TFA']" % self.tedit_type)
            pl = tedit["prompt_len"]
            self.vae.decoder.task_prompts = nn.ParameterDict(
                {t: nn.Parameter(torch.zeros(pl, 512)) for t in self.task_list})
            self.vae.decoder.task_editors = nn.ModuleList(
                [TaskFeatureAdapter(co, cs, prompt_len=pl, last_layer=last) for co, cs, last in CFG.TFA_SPECS])
        else:
            self.task_list, self.tedit_type = [], None

    def _encoder(self, x, enable_fr):                       # autoencoder.py:11-35
        enc = self.vae.encoder
        x = enc.conv_in(x)
        skips = []
        for i, blk in enumerate(enc.down_blocks[:-1]):
            x = blk(x)
            if enable_fr:
                x = enc.fr_blocks[i](x)
            skips.append(x)
        x = enc.mid_block(enc.down_blocks[-1](x))
        return enc.conv_out(enc.conv_act(enc.conv_norm_out(x))), skips

    def _decoder(self, z, skips, task):                     # autoencoder.py:37-72
        dec = self.vae.decoder
        x = dec.mid_block(dec.conv_in(z))
        cond = dec.task_prompts[task].unsqueeze(0).expand(z.shape[0], -1, -1)
        for i, blk in enumerate(dec.up_blocks[:-1]):
            x, cond = dec.task_editors[i](x, skips[-i - 1], cond)
            x = blk(x)
        x = dec.up_blocks[-1](x)
        return dec.conv_out(dec.conv_act(dec.conv_norm_out(x)))

    def encode(self, images, enable_fr=False, noise=None):  # autoencoder.py:132-156
        h, skips = self._encoder(images * 2 - 1, enable_fr)
        mean, logvar = self.vae.quant_conv(h).chunk(2, dim=1)
        std = torch.exp(0.5 * logvar.clamp(-30.0, 20.0))
        if noise is None:
            noise = torch.randn(mean.shape, dtype=mean.dtype, device=mean.device)
        return (mean + std * noise) * self.vae.config.scaling_factor, skips

    def decode(self, latents, res_samples, task):           # autoencoder.py:158-176
        z = self.vae.post_quant_conv(latents / self.vae.config.scaling_factor)
        return (self._decoder(z, res_samples, task) + 1) / 2

    def forward(self, images, task):                        # autoencoder.py:178-184 (task forced to 'ir')
        z, skips = self.encode(images, enable_fr=True)
        return self.decode(z, skips, "ir")


# ----------------------------------------------------------------------------- controller
class Controller(nn.Module):
    """controller.py:65-220 with ``stablesr_config`` (:29-45)."""

    def __init__(self, in_channels=4, model_channels=256, out_channels=256, num_res_blocks=2,
                 channel_mult=(1, 1, 2, 2), num_heads=4,
                 down_block_types=("AttnDownBlock2D",) * 3 + ("DownBlock2D",), zero_init=True, **_):
        super().__init__()
        ted = model_channels * 4
        self.time_proj = Timesteps(model_channels, True, 0)
        self.time_embedding = TimestepEmbedding(model_channels, ted)
        self.conv_in = nn.Conv2d(in_channels, model_channels, 3, padding=1)
        self.down_blocks = nn.ModuleList()
        widths, ch = [], model_channels
        for i, t in enumerate(down_block_types):
            cin, ch = ch, model_channels * channel_mult[i]
            self.down_blocks.append(get_down_block(
                t, num_layers=num_res_blocks, in_channels=cin, out_channels=ch, temb_channels=ted,
                add_downsample=i != len(channel_mult) - 1, resnet_eps=1e-5, resnet_groups=32,
                downsample_padding=1, attention_head_dim=ch // num_heads))
            widths.append(ch)
        self.middle_block = UNetMidBlock2D(in_channels=ch, temb_channels=ted, resnet_eps=1e-5,
                                           resnet_groups=32, attention_head_dim=ch // num_heads)
        self.fea_tran = nn.ModuleList([
            ResnetBlock2D(in_channels=w, out_channels=out_channels, temb_channels=ted, groups=32, eps=1e-5)
            for w in widths])
        if zero_init:                                       # controller.py:172-185
            from .blocks import Attention
            for m in self.modules():
                if isinstance(m, ResnetBlock2D):
                    nn.init.zeros_(m.conv2.weight), nn.init.zeros_(m.conv2.bias)
                elif isinstance(m, Attention):
                    nn.init.zeros_(m.to_out[0].weight), nn.init.zeros_(m.to_out[0].bias)

    def forward(self, x, timesteps):
        emb = self.time_embedding(self.time_proj(timesteps))
        taps = []
        h = self.conv_in(x)
        for blk in self.down_blocks:
            h, outs = blk(h, emb)
            taps.append(outs[-2])                           # controller.py:205
        taps[-1] = self.middle_block(h, emb)                # controller.py:211
        return {t.size(-1): self.fea_tran[i](t.contiguous(), emb) for i, t in enumerate(taps)}


# ----------------------------------------------------------------------------- UNet
class ControlledUNet(nn.Module):
    """base_model.py:14-245 (``control_type == "scedit"`` only; SPADE is out of scope)."""

    def __init__(self, unet: UNet2DConditionModel, control_type: str, null_embeds=None):
        super().__init__()
        self.unet = unet
        if null_embeds is None:
            null_embeds = torch.load(_ASSET, map_location="cpu")
        self.register_buffer("null_embeds", null_embeds)
        if control_type == "scedit":
            chans = getattr(unet, "sc_chans", CFG.SC_CHANS)
            self.csc_editors = nn.ModuleList([CSCEAdapter(c, c, CFG.SC_COND) for c in chans])
        else:
            raise ValueError(f"control_type '{control_type}' not supported")

    def forward(self, sample, control, timesteps):
        u = self.unet
        ctx = self.null_embeds.expand(sample.shape[0], -1, -1)
        emb = u.time_embedding(u.time_proj(timesteps).to(sample.dtype))
        # encoder (base_model.py:94-162)
        x = u.conv_in(sample)
        skips = [x]
        for blk in u.down_blocks:
            attns = blk.attentions if getattr(blk, "has_cross_attention", False) else [None] * len(blk.resnets)
            for r, a in zip(blk.resnets, attns):
                x = r(x, emb)
                if a is not None:
                    x = a(x, ctx, return_dict=False)[0]
                skips.append(x)
            if blk.downsamplers is not None:
                for d in blk.downsamplers:
                    x = d(x)
                skips.append(x)
        x = u.mid_block.resnets[0](x, emb)
        for a, r in zip(u.mid_block.attentions, u.mid_block.resnets[1:]):
            x = r(a(x, ctx, return_dict=False)[0], emb)
        # SC-Tuner edits the skip tensors only (base_model.py:233-238)
        for i, ed in enumerate(self.csc_editors):
            skips[i] = ed(skips[i], control[skips[i].shape[-1]])
        # decoder (base_model.py:164-209)
        for blk in u.up_blocks:
            attns = blk.attentions if getattr(blk, "has_cross_attention", False) else [None] * len(blk.resnets)
            for r, a in zip(blk.resnets, attns):
                x = r(torch.cat([x, skips.pop()], dim=1), emb)
                if a is not None:
                    x = a(x, ctx, return_dict=False)[0]
            if blk.upsamplers is not None:
                for up in blk.upsamplers:
                    x = up(x)
        return u.conv_out(u.conv_act(u.conv_norm_out(x)))


# ----------------------------------------------------------------------------- assembly
class DiffUIE(nn.Module):
    """unifie.py:22-169 (skipping the leftover FLOPs probe + ``raise`` at :43-53)."""

    def __init__(self, frenc=None, cnet=None, tedit=None, unet=None, vae=None, null_embeds=None):
        super().__init__()
        self.fr_type = frenc["type"] if frenc else None
        self.control_type = cnet["type"] if cnet else None
        self.tedit = tedit if tedit else None
        self.ae = SkipConnectedAutoEncoder(vae or AutoencoderKL(), self.fr_type, self.tedit)
        if self.control_type:
            self.controller = Controller(**CFG.CONTROLLER)
            self.base_model = ControlledUNet(unet or UNet2DConditionModel(), self.control_type, null_embeds)
            self.register_buffer("train_timesteps", torch.tensor([249, 499, 749, 999, 999, 999], dtype=int))
            self.ddpm = DDPMScheduler()
            self.scheduler = DDIMScheduler()
            self.scheduler.set_timesteps(cnet["num_inference_steps"], device=self.train_timesteps.device)

    def diffuse(self, latents, timesteps=None, noise=None):           # unifie.py:77-89
        if timesteps is None:
            idx = torch.randint(0, len(self.train_timesteps), (latents.size(0),), device=latents.device)
            timesteps = self.train_timesteps[idx]
        if noise is None:
            noise = torch.randn_like(latents)
        return self.ddpm.add_noise(latents, noise, timesteps), noise, timesteps

    def predict_z0(self, latents, conditions, timesteps):             # unifie.py:91-105
        eps = self.base_model(latents, self.controller(conditions, timesteps), timesteps)
        a = self.ddpm.alphas_cumprod[timesteps].view(-1, 1, 1, 1)
        return (latents - (1 - a) ** 0.5 * eps) / a ** 0.5

    def forward(self, images, task, noise=None):                      # unifie.py:107-169
        """``noise=(posterior_noise, diffuse_noise)`` injects the two RNG draws for parity."""
        org_h, org_w = images.shape[-2:]
        h, w = org_h, org_w
        if h < 512 or w < 512:
            s = 512 / min(h, w)
            h, w = round(h * s), round(w * s)
            images = F.interpolate(images, (h, w), mode="bicubic", align_corners=False, antialias=False)
        if h % 64 or w % 64:
            images = F.pad(images, (0, (64 - w % 64) % 64, 0, (64 - h % 64) % 64), mode="reflect")
        n_post, n_diff = noise if noise is not None else (None, None)
        z0, mids = self.ae.encode(images, enable_fr=self.fr_type is not None, noise=n_post)
        if self.control_type:
            t = 999 * torch.ones((len(images),), dtype=int, device=images.device)
            zt, _, _ = self.diffuse(z0, t, n_diff)
            for t in self.scheduler.timesteps:
                ts = t.reshape(-1)
                eps = self.base_model(zt, self.controller(z0, ts), ts)
                zt = self.scheduler.step(eps, t, zt).prev_sample
        else:
            zt = z0
        preds = self.ae.decode(zt, mids, task)[..., :h, :w]
        return F.interpolate(preds, (org_h, org_w), mode="bicubic", align_corners=False, antialias=False)
