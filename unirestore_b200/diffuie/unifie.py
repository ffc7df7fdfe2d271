"""``DiffUIE`` -- model assembly and the inference loop, reference unifie.py:22-169.

    resize/pad -> ae.encode(+CFRM) -> noise at t=999 -> N x {Controller, ControlledUNet(+SC-Tuner), DDIM step}
    -> ae.decode(+TFA) -> crop / resize

The leftover FLOPs probe and unconditional ``raise`` at unifie.py:43-53 are not reproduced.  All per-step tensors
stay on the GPU in bf16 channels-last; latents / noise / scheduler state are fp32; timesteps and the DDIM index
math are host-side integers (bit-exact with the reference tables).
"""
import os

import torch
import torch.nn as nn

from .. import ops
from .autoencoder import SkipConnectedAutoEncoder
from .base_model import ControlledUNet
from .controller import Controller, stablesr_config
from .schedulers import DDIMScheduler, DDPMScheduler
from .sd_blocks import AutoencoderKL, TimeCtx, UNet2DConditionModel


class DiffUIE(nn.Module):
    def __init__(self, frenc=None, cnet=None, tedit=None, unet=None, vae=None, null_embeds=None):
        super().__init__()
        self.fr_type = frenc["type"] if frenc else None
        self.control_type = cnet["type"] if cnet else None
        self.tedit = tedit if tedit else None
        self.ae = SkipConnectedAutoEncoder(vae or AutoencoderKL.from_pretrained("stabilityai/sd-turbo", subfolder="vae"),
                                           self.fr_type, self.tedit)
        if self.control_type:
            self.controller = Controller(**stablesr_config)
            self.base_model = ControlledUNet(
                unet or UNet2DConditionModel.from_pretrained("stabilityai/sd-turbo", subfolder="unet"),
                control_type=self.control_type, null_embeds=null_embeds)
            self.register_buffer("train_timesteps", torch.tensor([249, 499, 749, 999, 999, 999], dtype=torch.int64))
            self.ddpm = DDPMScheduler.from_pretrained("stabilityai/sd-turbo", subfolder="scheduler")
            self.scheduler = DDIMScheduler.from_pretrained("stabilityai/sd-turbo", subfolder="scheduler")
            self.scheduler.set_timesteps(cnet["num_inference_steps"], device=self.train_timesteps.device)
        self._ts_dev = {}
        self._side_streams = {}
        self.overlap_controller = os.environ.get("UNIRESTORE_OVERLAP_CONTROLLER", "1") == "1"
        self._keep = []
        # CUDA-graph replay of the whole forward for static shapes (one capture per (shape, task, noise-mode));
        # every kernel behind the C-ABI is capture-safe (no allocation, no synchronisation).
        self.use_cuda_graph = os.environ.get("UNIRESTORE_CUDA_GRAPH", "0") == "1"
        self._graphs = {}
        self._graph_gen = -1
        self.latent_trace = None

    # ---------------------------------------------------------------------------------- training-time helpers
    def diffuse(self, latents, timesteps=None, noise=None):                      # unifie.py:77-89
        if timesteps is None:
            idx = torch.randint(0, len(self.train_timesteps), (latents.size(0),), device=latents.device)
            timesteps = self.train_timesteps.to(latents.device)[idx]
        if noise is None:
            noise = torch.randn_like(latents)
        ts = [int(t) for t in timesteps.reshape(-1).tolist()]
        if len(set(ts)) == 1:
            sa, sb = self.ddpm.noise_coefficients(ts[0])
            out, _ = ops.latent_axpby(latents.float().contiguous(), sa, noise.float().contiguous(), sb)
        else:                                                                       # per-sample timesteps
            out = torch.empty_like(latents, dtype=torch.float32)
            for i, t in enumerate(ts):
                sa, sb = self.ddpm.noise_coefficients(t)
                out[i:i + 1], _ = ops.latent_axpby(latents[i:i + 1].float().contiguous(), sa,
                                                   noise[i:i + 1].float().contiguous(), sb)
        return out, noise, timesteps

    def _time_contexts(self, ts, device):
        """(Controller, UNet) ``TimeCtx`` for the integer timesteps ``ts`` of this forward: one sinusoid + MLP launch
        set per network over all T timesteps (base_model.py:104-106, controller.py:196-197)."""
        key = (tuple(ts), str(device))
        t_dev = self._ts_dev.get(key)
        if t_dev is None:               # host -> device copy of the timestep list: once per (schedule, device)
            t_dev = self._ts_dev[key] = torch.tensor(list(ts), dtype=torch.int64, device=device)
        return TimeCtx(self.controller.time_embed(t_dev)), TimeCtx(self.base_model.time_embed(t_dev))

    def clear_caches(self):
        self._ts_dev = {}
        self._graphs = {}
        for m in self.modules():
            if hasattr(m, "invalidate"):
                m.invalidate()

    def _stream(self, name, device):
        key = (name, device)
        st = self._side_streams.get(key)
        if st is None:
            st = self._side_streams[key] = torch.cuda.Stream(device=device)
        return st

    def _run_controller(self, z0_8, tctx):
        ops.stats_arena_begin(z0_8.device, "ctl")    # one fill instead of a memset node per GroupNorm
        try:
            return self.controller.run(z0_8, tctx)                               # unifie.py:148
        finally:
            ops.stats_arena_end(z0_8.device)

    def _run_unet(self, zt8, control, tctx):
        ops.stats_arena_begin(zt8.device, "main")
        try:
            return self.base_model.run(zt8, control, tctx)                       # unifie.py:149
        finally:
            ops.stats_arena_end(zt8.device)

    def predict_eps(self, zt8, z0_8, t: int):
        ctx_c, ctx_u = self._time_contexts([int(t)], zt8.device)
        self.base_model.begin_forward()
        return self._run_unet(zt8, self._run_controller(z0_8, ctx_c), ctx_u)

    def predict_z0(self, latents, conditions, timesteps):                        # unifie.py:91-105
        """One-step estimate of z0 from noisy latents; ``timesteps`` int64 [1] or [B] (per-sample, as drawn by
        ``diffuse`` during stage-2/3 training, engine_unifie.py:139-147)."""
        tl = [int(t) for t in timesteps.reshape(-1).tolist()]
        B = latents.shape[0]
        if len(tl) not in (1, B):
            raise ValueError("timesteps must have 1 or batch entries")
        zt8, z0_8 = ops.image_to_nhwc8(latents.float()), ops.image_to_nhwc8(conditions.float())
        z = latents.float().clone()
        if len(set(tl)) == 1:
            eps8 = self.predict_eps(zt8, z0_8, tl[0])
            sa, sb = self.ddpm.noise_coefficients(tl[0])
            # x0 = (x - sqrt(1-a) eps) / sqrt(a): the DDIM kernel with a_prev = 1 (sqrt(a_p) = 1, sqrt(1-a_p) = 0)
            ops.ddim_step_(z, eps8, (sa, sb, 1.0, 0.0), want_nhwc8=False)
            return z
        # per-sample timesteps: [B, C] time embeddings -> per-image row vectors in the conv epilogues
        ts = torch.tensor(tl, dtype=torch.int64, device=latents.device)
        control = self.controller.run(z0_8, self.controller.time_embed(ts))
        eps8 = self.base_model.run(zt8, control, self.base_model.time_embed(ts))
        for i, t in enumerate(tl):
            sa, sb = self.ddpm.noise_coefficients(t)
            zi = z[i:i + 1]
            ops.ddim_step_(zi, eps8[i:i + 1], (sa, sb, 1.0, 0.0), want_nhwc8=False)
        return z

    # ---------------------------------------------------------------------------------- inference (hot path)
    def restore_latents(self, z0, z0_8, noise=None):
        """The N-step loop of unifie.py:141-150 on fp32 NCHW latents; returns the denoised latents."""
        if noise is None:
            noise = torch.randn_like(z0)
        sa, sb = self.ddpm.noise_coefficients(999)                               # unifie.py:141-144
        zt, zt8 = ops.latent_axpby(z0, sa, noise.float().contiguous(), sb, want_nhwc8=True)
        # The Controller only depends on (z0, t): the one of step i+1 runs on a side stream concurrently with the UNet
        # of step i and fills the SMs its many small, dependent kernels leave idle (tails / prologues / underfilled
        # grids).  Its outputs are kept alive until the loop ends, so no buffer crosses streams while being recycled.
        ts = [int(t) for t in self.scheduler.timesteps_host]
        ctx_c, ctx_u = self._time_contexts(ts, z0.device)      # all T time embeddings of this forward, two launch sets
        self.base_model.begin_forward()                        # null-prompt K/V are recomputed in this forward
        main = torch.cuda.current_stream(z0.device)
        side = self._stream("ctl", z0.device)
        overlap = self.overlap_controller and len(ts) > 1
        keep = self._keep = []

        def launch_controller(i):
            if not overlap:
                return self._run_controller(z0_8, ctx_c.at(i)), None
            if i == 0:
                side.wait_stream(main)                                            # z0_8 is produced on the main stream
            with torch.cuda.stream(side):
                ctl = self._run_controller(z0_8, ctx_c.at(i))
                ev = torch.cuda.Event()
                ev.record(side)
            keep.append(ctl)
            return ctl, ev

        nxt = launch_controller(0)
        for i, t in enumerate(ts):                                               # unifie.py:146-150
            control, ev = nxt
            if i + 1 < len(ts):
                nxt = launch_controller(i + 1)
            if ev is not None:
                main.wait_event(ev)
            eps8 = self._run_unet(zt8, control, ctx_u.at(i))
            zt8 = ops.ddim_step_(zt, eps8, self.scheduler.step_coefficients(t), bool(self.scheduler.config.clip_sample))
            if self.latent_trace is not None:              # parity runs: latents after every DDIM step (eager only)
                self.latent_trace.append(zt.clone())
        if overlap:
            main.wait_stream(side)
        self._keep = []
        return zt

    @torch.no_grad()
    def forward(self, images, task, noise=None, quantize=False):
        """images fp32 [B,3,H,W] in [0,1] -> restored fp32 [B,3,H,W].  ``noise=(posterior, diffuse)`` injects the two
        RNG draws of the reference (autoencoder.py:152, unifie.py:87) for parity runs.  ``quantize=True`` fuses the
        validate loop's ``pred.mul(255).round_().clamp_(0, 255).div_(255)`` (eval_image_restoration.py:71) into the
        last kernel of the path.  ``images`` may be a strided view (e.g. the centre crop of
        eval_image_restoration.py:113-134, ``center_crop`` below): the first kernel reads it in place."""
        if self.use_cuda_graph and images.is_cuda:
            return self._forward_graphed(images, task, noise, quantize)
        return self._forward_impl(images, task, noise, quantize)

    def _forward_graphed(self, images, task, noise, quantize=False):
        # Captured graphs point at the packed bf16 weights: any UrModule.invalidate() (.to() / .half() / _apply /
        # load_state_dict on this module OR any child, e.g. ``model.controller.load_state_dict`` as in
        # engine_unifie.py:49-126) bumps the weight generation and the graphs recorded before it are dropped.
        from .sd_blocks import weight_generation
        if self._graph_gen != weight_generation():
            self._graphs, self._graph_gen = {}, weight_generation()
        key = (tuple(images.shape), task, noise is not None, str(images.device), bool(quantize))
        g = self._graphs.get(key)
        if g is None:
            g = self._graphs[key] = _GraphedForward(self, images, task, noise, quantize)
        return g(images, noise)

    def _forward_impl(self, images, task, noise=None, quantize=False):
        org_h, org_w = images.shape[-2:]
        h, w = org_h, org_w
        images = images.float()
        resize = h < 512 or w < 512                                              # unifie.py:124-129
        if resize:
            s = 512 / min(h, w)
            h, w = round(h * s), round(w * s)
        pad_r, pad_b = (64 - w % 64) % 64, (64 - h % 64) % 64                     # unifie.py:130-134
        if resize or pad_r or pad_b:
            if not images.is_cuda:
                raise ValueError("DiffUIE.forward needs CUDA tensors (the product path has no CPU fallback)")
            images = ops.resize_pad(images, (h, w) if resize else None, pad_b, pad_r)   # one gather kernel
        n_post, n_diff = noise if noise is not None else (None, None)
        z0, z0_8, mids = self.ae.run_encode(images, enable_fr=self.fr_type is not None, noise=n_post)
        zt = self.restore_latents(z0, z0_8, n_diff) if self.control_type else z0
        back = (h, w) != (org_h, org_w)
        preds = self.ae.run_decode(zt, mids, task, crop_hw=(h, w), quantize=quantize and not back)   # unifie.py:155,164
        if back:                                                                 # unifie.py:165-168
            preds = ops.resize_pad(preds, (org_h, org_w), quantize=quantize)
        return preds


def center_crop(image, upper=(512, 512)):
    """``crop_tensor`` of the validate loop (eval_image_restoration.py:113-134): centre crop to at most ``upper`` as a
    strided VIEW -- ``DiffUIE.forward`` reads it in place (``ur_image_to_nhwc8`` / ``ur_resize_pad`` take strides)."""
    if image.ndim not in (3, 4):
        raise NotImplementedError
    h, w = image.shape[-2:]
    ch, cw = min(h, upper[0]), min(w, upper[1])
    return image[..., h // 2 - ch // 2: h // 2 + ch // 2, w // 2 - cw // 2: w // 2 + cw // 2]


class _GraphedForward:
    """One captured CUDA graph of ``DiffUIE._forward_impl`` for fixed input shapes; inputs are copied into static
    buffers, the graph is replayed, the static output is cloned."""

    def __init__(self, model, images, task, noise, quantize=False):
        self.img = images.detach().float().clone()
        self.noise = tuple(n.detach().float().clone() for n in noise) if noise is not None else None
        side = torch.cuda.Stream(device=images.device)
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):                    # warm-up: weight packing, caches, function attributes
            for _ in range(2):
                model._forward_impl(self.img, task, self.noise, quantize)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        from .. import _cabi
        n0, by0 = _cabi.launch_count, dict(_cabi.launch_counts)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.out = model._forward_impl(self.img, task, self.noise, quantize)
        self.n_launches = _cabi.launch_count - n0          # C-ABI kernel launches recorded in the graph
        self.launches_by_entry = {k: v - by0.get(k, 0) for k, v in _cabi.launch_counts.items() if v - by0.get(k, 0)}

    def __call__(self, images, noise):
        self.img.copy_(images, non_blocking=True)
        if self.noise is not None:
            for dst, src in zip(self.noise, noise):
                dst.copy_(src, non_blocking=True)
        self.graph.replay()
        return self.out.clone()
