"""DDPM / DDIM schedulers as the reference drives them (unifie.py:69-75,88,99,146-150).

Host-side and exact: the alpha-bar table is built once on the CPU in fp32 exactly as diffusers does
(``cumprod(1 - linspace(sqrt(b0), sqrt(b1), T)**2)``), timesteps are int64 ("trailing" spacing), and the
per-step DDIM coefficients handed to ``ur_ddim_step`` are the fp32 values the reference computes with
``alpha ** 0.5`` / ``(1 - alpha) ** 0.5``.  sd-turbo scheduler constants are recalled
from the public config (clip_sample False, set_alpha_to_one False) and exposed as overrides.
"""
from __future__ import annotations

import numpy as np
import torch

SCHEDULER_CONFIG = dict(num_train_timesteps=1000, beta_start=0.00085, beta_end=0.012, beta_schedule="scaled_linear",
                        prediction_type="epsilon", timestep_spacing="trailing", steps_offset=1, set_alpha_to_one=False,
                        clip_sample=False)


class _Config(dict):
    __getattr__ = dict.__getitem__


def make_alphas_cumprod(num_train_timesteps=1000, beta_start=0.00085, beta_end=0.012):
    betas = torch.linspace(beta_start ** 0.5, beta_end ** 0.5, num_train_timesteps, dtype=torch.float32) ** 2
    return torch.cumprod(1.0 - betas, dim=0)


class _Base:
    def __init__(self, **overrides):
        c = dict(SCHEDULER_CONFIG)
        c.update(overrides)
        self.config = _Config(c)
        if c["beta_schedule"] != "scaled_linear" or c["prediction_type"] != "epsilon":
            raise ValueError("only the sd-turbo scheduler configuration is implemented")
        self.alphas_cumprod = make_alphas_cumprod(c["num_train_timesteps"], c["beta_start"], c["beta_end"])
        self.final_alpha_cumprod = torch.tensor(1.0) if c["set_alpha_to_one"] else self.alphas_cumprod[0]

    @classmethod
    def from_pretrained(cls, model_id=None, subfolder=None, **kw):
        return cls()


class DDPMScheduler(_Base):
    def noise_coefficients(self, t: int):
        """(sqrt(abar_t), sqrt(1 - abar_t)) as python floats of the fp32 values (add_noise, unifie.py:88)."""
        a = self.alphas_cumprod[int(t)]
        return float(a ** 0.5), float((1 - a) ** 0.5)


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
        self.timesteps_host = [int(t) for t in ts]
        self.timesteps = torch.from_numpy(ts).to(device)

    def prev_timestep(self, t: int) -> int:
        return int(t) - self.config.num_train_timesteps // self.num_inference_steps

    def step_coefficients(self, t: int):
        """(sqrt(a_t), sqrt(1-a_t), sqrt(a_prev), sqrt(1-a_prev)), eta = 0 (DDIMScheduler.step, unifie.py:150)."""
        p = self.prev_timestep(t)
        a_t = self.alphas_cumprod[int(t)]
        a_p = self.alphas_cumprod[p] if p >= 0 else self.final_alpha_cumprod
        return (float(a_t ** 0.5), float((1 - a_t) ** 0.5), float(a_p ** 0.5), float((1 - a_p) ** 0.5))
