"""``SkipConnectedAutoEncoder`` -- VAE encode (+CFRM) / decode (+TFA), reference autoencoder.py:74-184 with the
monkey-patched encoder / decoder forwards of :11-72.

encode (autoencoder.py:132-156, :11-35):  x*2-1 -> conv_in -> 3 x {DownEncoderBlock, CFRM fr_block, save skip}
    -> DownEncoderBlock -> mid -> GN/SiLU/conv_out -> quant_conv -> posterior sample * scaling_factor
decode (autoencoder.py:158-176, :37-72):  z/scaling_factor -> post_quant_conv -> conv_in -> mid
    -> 3 x {TFA(x, skip, prompt), UpDecoderBlock} -> UpDecoderBlock -> GN/SiLU/conv_out -> (x+1)/2
Activations stay bf16 channels-last between kernels; latents are fp32 NCHW at the module boundary.
"""
import torch
import torch.nn as nn

from .. import ops
from .cfrm import AdaNAFV2
from .nafnet_arch import NAFBlock
from .sd_blocks import AutoencoderKL
from .taskeditor import TaskFeatureAdapter

CFRM_STACKS = ((128, 1), (256, 1), (512, 9))                      # autoencoder.py:94-96
TFA_SPECS = ((512, 512, False), (512, 256, False), (512, 128, True))   # autoencoder.py:122-126


class UnknownTaskError(KeyError, AttributeError):
    """``task_prompts[task]`` of an unregistered task (autoencoder.py:47): nn.ParameterDict raises AttributeError on the
    reference's torch, KeyError by the mapping protocol -- callers catching either keep working."""


class _Stack(nn.Sequential):
    def run(self, x):
        for m in self:
            x = m.run(x)
        return x


class SkipConnectedAutoEncoder(nn.Module):
    def __init__(self, vae: AutoencoderKL, fr_type=None, tedit=None):
        super().__init__()
        self.vae = vae
        self.tedit_dict = tedit
        boc = vae.config["block_out_channels"]
        if fr_type == "CFRM":
            stacks = CFRM_STACKS if tuple(boc[:3]) == (128, 256, 512) else tuple((c, 1) for c in boc[:3])
            self.vae.encoder.fr_blocks = nn.ModuleList([
                _Stack(*[NAFBlock(c) for _ in range(n)], AdaNAFV2(c)) for c, n in stacks])
        elif fr_type is not None:
            raise ValueError("Invalid fr_type")
        if tedit:
            self.task_list, self.tedit_type = tedit["task"], tedit["type"]
            if self.tedit_type != "TFA":
                raise KeyError("%s is not defined in the taskeditor!, please select ['TFA']" % self.tedit_type)
            pl = tedit["prompt_len"]
            top = boc[-1]
            specs = TFA_SPECS if top == 512 else ((top, boc[2], False), (top, boc[1], False), (top, boc[0], True))
            self.vae.decoder.task_prompts = nn.ParameterDict(
                {t: nn.Parameter(torch.zeros(pl, specs[0][1])) for t in self.task_list})
            self.vae.decoder.task_editors = nn.ModuleList(
                [TaskFeatureAdapter(co, cs, prompt_len=pl, last_layer=last) for co, cs, last in specs])
        else:
            self.task_list, self.tedit_type = [], None

    # ---------------------------------------------------------------------------------- fast paths
    def run_encode(self, images, enable_fr=False, noise=None):
        """images fp32 [B,3,H,W] in [0,1] -> (z fp32 [B,4,h,w], z8 bf16 [B,h,w,8], skips bf16 NHWC)."""
        enc, pk = self.vae.encoder, self.vae.encoder.pk
        x = ops.image_to_nhwc8(images.float(), 2.0, -1.0)                                    # autoencoder.py:151
        x = ops.conv_gemm(x, pk["w_in"], enc.conv_in.out_channels, taps=ops.TAPS_3x3, bias=pk["b_in"], want_stats=True)
        skips = []
        for i, blk in enumerate(enc.down_blocks[:-1]):                                       # autoencoder.py:18-24
            # the reference saves the skip AFTER the down-sampler of block i (block output), CFRM applied first
            x = blk.run(x)
            if enable_fr:
                x = enc.fr_blocks[i].run(x)
            skips.append(x)
        x = enc.mid_block.run(enc.down_blocks[-1].run(x))
        x = ops.group_norm(x, enc.conv_norm_out.num_groups, pk["g"], pk["b"], enc.conv_norm_out.eps, silu=True)
        x = ops.conv_gemm(x, pk["w_out"], enc.conv_out.out_channels, taps=ops.TAPS_3x3, bias=pk["b_out"])
        vp = self.vae.pk
        moments = ops.conv_gemm(x, vp["wq"], 8, bias=vp["bq"], out_dtype=torch.float32)
        B, h, w, _ = moments.shape
        if noise is None:                                                                    # autoencoder.py:152
            noise = torch.randn((B, 4, h, w), device=images.device, dtype=torch.float32)
        z, z8 = ops.posterior_sample(moments, noise.float().contiguous(), float(self.vae.config["scaling_factor"]))
        return z, z8, skips

    def run_decode(self, latents, skips, task, crop_hw=None, quantize=False):
        """latents fp32 [B,4,h,w], skips bf16 NHWC, task key -> fp32 [B,3,H,W] = (decoder + 1) / 2.
        ``quantize``: 8-bit quantisation of the prediction (eval_image_restoration.py:71) fused into the write-out."""
        dec, pk, vp = self.vae.decoder, self.vae.decoder.pk, self.vae.pk
        prompt = None
        if self.tedit_type:
            if task not in dec.task_prompts:                                 # `task_prompts[task]` (autoencoder.py:47)
                raise UnknownTaskError(task)
            prompt = dec.task_prompts[task]
        _, z8 = ops.latent_axpby(latents.float().contiguous(), 1.0, want_out=False, want_nhwc8=True,
                                 scale8=1.0 / float(self.vae.config["scaling_factor"]))          # autoencoder.py:170
        z8 = ops.conv_gemm(z8, vp["wpq"], 8, bias=vp["bpq"])
        x = ops.conv_gemm(z8, pk["w_in"], dec.conv_in.out_channels, taps=ops.TAPS_3x3, bias=pk["b_in"], want_stats=True)
        x = dec.mid_block.run(x)
        B = x.shape[0]
        cond = prompt.detach().float().unsqueeze(0).expand(B, -1, -1).contiguous() if prompt is not None else None
        for i, blk in enumerate(dec.up_blocks[:-1]):                                          # autoencoder.py:49-60
            if self.tedit_type:
                x, cond = dec.task_editors[i].run(x, skips[-i - 1], cond)
            x = blk.run(x)
        x = dec.up_blocks[-1].run(x)
        x = ops.group_norm(x, dec.conv_norm_out.num_groups, pk["g"], pk["b"], dec.conv_norm_out.eps, silu=True)
        y = ops.conv_gemm(x, pk["w_out"], 8, taps=ops.TAPS_3x3, bias=pk["b_out"], out_dtype=torch.float32)
        h, w = crop_hw if crop_hw is not None else (y.shape[1], y.shape[2])
        return ops.nhwc_to_image(y, dec.conv_out.out_channels, h, w, 0.5, 0.5, quantize=quantize)   # autoencoder.py:175

    # ---------------------------------------------------------------------------------- reference API
    def encode(self, images, enable_fr=False, noise=None):
        z, _, skips = self.run_encode(images, enable_fr, noise)
        return z, [s.permute(0, 3, 1, 2) for s in skips]

    def decode(self, latents, res_samples, task):
        skips = [s.permute(0, 2, 3, 1).to(torch.bfloat16).contiguous() for s in res_samples]
        return self.run_decode(latents, skips, task)

    def forward(self, images, task):                            # autoencoder.py:178-184 (task forced to 'ir')
        z, _, skips = self.run_encode(images, enable_fr=True)
        return self.run_decode(z, skips, "ir")
