"""Deterministic synthetic weights keyed by parameter NAME (no checkpoint exists offline).

Every tensor is drawn from its own generator seeded with crc32(key), so the reference files
(under the oracle shims), the CPU oracle and the CUDA modules all receive bit-identical fp32
weights from key names + shapes alone.  Zero-initialised reference tensors (Controller zero
convs controller.py:174-185, NAFBlock beta/gamma nafnet_arch.py:107-108, task prompts
autoencoder.py:117-120) are re-drawn non-zero so those sub-graphs are actually exercised
(SURVEY.md section 8c hazard 2).
"""
from __future__ import annotations

import zlib

import torch


def _draw(key: str, shape, scale: float, shift: float = 0.0) -> torch.Tensor:
    g = torch.Generator(device="cpu")
    g.manual_seed(zlib.crc32(key.encode()) & 0x7FFFFFFF)
    return torch.randn(tuple(shape), generator=g, dtype=torch.float32) * scale + shift


def synth_tensor(key: str, ref: torch.Tensor) -> torch.Tensor:
    leaf = key.rsplit(".", 1)[-1]
    shape = ref.shape
    if not ref.is_floating_point():
        return ref.clone()
    if leaf in ("beta", "gamma"):
        return _draw(key, shape, 0.3)
    if "task_prompts" in key:
        return _draw(key, shape, 0.5)
    if leaf == "weight" and ref.ndim >= 2:
        fan_in = ref[0].numel()
        return _draw(key, shape, 0.8 / fan_in ** 0.5)
    if leaf == "weight":                       # norm scale
        return _draw(key, shape, 0.1, 1.0)
    if leaf == "bias":
        return _draw(key, shape, 0.05)
    return ref.clone()                         # buffers (null_embeds, train_timesteps, ...)


@torch.no_grad()
def deterministic_init_(module: torch.nn.Module, prefix: str = "") -> torch.nn.Module:
    """Overwrite every parameter of ``module`` in place; keys are ``prefix + state_dict key``."""
    for k, p in module.named_parameters():
        p.copy_(synth_tensor(prefix + k, p).to(p.device, p.dtype))
    return module
