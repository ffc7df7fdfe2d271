"""Restatement of the diffusers DDIM / DDPM schedulers as the reference uses them.

TEST INFRASTRUCTURE (oracle).  Call sites: unifie.py:69-75 (construction +
``set_timesteps``), unifie.py:88 (``add_noise``), unifie.py:99 (``alphas_cumprod``),
unifie.py:146-150 (``timesteps`` / ``step(...).prev_sample``).  Algorithm: SURVEY.md
Appendix A.8 (diffusers 0.29, un-vendored -> PARITY UNPINNED; anchored on the
known-answer tables of SURVEY.md section 8c, checked in tests/test_scheduler.py).
"""
from __future__ import annotations

from types import SimpleNamespace

import numpy as np
import torch

from . import sd_turbo_config as CFG


def make_alphas_cumprod(num_train_timesteps=1000, beta_start=0.00085, beta_end=0.012):
    """fp32 cumprod of ``1 - linspace(sqrt(b0), sqrt(b1), T)**2`` ("scaled_linear"), on the host."""
    betas = torch.linspace(beta_start ** 0.5, beta_end ** 0.5, num_train_timesteps, dtype=torch.float32) ** 2
    return torch.cumprod(1.0 - betas, dim=0)


class _Base:
    def __init__(self, **overrides):
        c = dict(CFG.SCHEDULER)
        c.update(overrides)
        self.config = SimpleNamespace(**c)
        self.alphas_cumprod = make_alphas_cumprod(c["num_train_timesteps"], c["beta_start"], c["beta_end"])
        self.final_alpha_cumprod = torch.tensor(1.0) if c["set_alpha_to_one"] else self.alphas_cumprod[0]

    @classmethod
    def from_pretrained(cls, model_id=None, subfolder=None, **kw):
        return cls()


class DDPMScheduler(_Base):
    def add_noise(self, original_samples, noise, timesteps):
        ac = self.alphas_cumprod.to(device=original_samples.device, dtype=original_samples.dtype)
        timesteps = timesteps.to(original_samples.device)
        sa = ac[timesteps] ** 0.5
        sb = (1 - ac[timesteps]) ** 0.5
        sa = sa.flatten()
        sb = sb.flatten()
        while sa.ndim < original_samples.ndim:
            sa, sb = sa.unsqueeze(-1), sb.unsqueeze(-1)
        return sa * original_samples + sb * noise


class DDIMScheduler(_Base):
    def set_timesteps(self, num_inference_steps, device=None):
        T = self.config.num_train_timesteps
        self.num_inference_steps = num_inference_steps
        if self.config.timestep_spacing == "trailing":
            ts = np.round(np.arange(T, 0, -T / num_inference_steps)).astype(np.int64) - 1
        elif self.config.timestep_spacing == "leading":
            ratio = T // num_inference_steps
            ts = (np.arange(0, num_inference_steps) * ratio).round()[::-1].copy().astype(np.int64)
            ts += self.config.steps_offset
        else:
            raise ValueError(self.config.timestep_spacing)
        self.timesteps = torch.from_numpy(ts).to(device)

    def step(self, model_output, timestep, sample, eta=0.0):
        t = int(timestep)
        prev_t = t - self.config.num_train_timesteps // self.num_inference_steps
        a_t = self.alphas_cumprod[t]
        a_p = self.alphas_cumprod[prev_t] if prev_t >= 0 else self.final_alpha_cumprod
        a_t, a_p = a_t.to(sample.device), a_p.to(sample.device)
        beta_t = 1 - a_t
        x0 = (sample - beta_t ** 0.5 * model_output) / a_t ** 0.5
        if self.config.clip_sample:
            x0 = x0.clamp(-1.0, 1.0)
        direction = (1 - a_p) ** 0.5 * model_output       # eta = 0 -> std_dev_t = 0
        prev_sample = a_p ** 0.5 * x0 + direction
        return SimpleNamespace(prev_sample=prev_sample, pred_original_sample=x0)


class EulerDiscreteScheduler(_Base):
    """Imported but never used by the reference (unifie.py:10)."""
