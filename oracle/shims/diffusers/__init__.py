"""Shim: names the reference imports from ``diffusers`` (unifie.py:6-12, cfrm.py:5, base_model.py:5)."""
from oracle.blocks import AutoencoderKL, UNet2DConditionModel  # noqa: F401
from oracle.schedulers import DDIMScheduler, DDPMScheduler, EulerDiscreteScheduler  # noqa: F401
