"""Shim for controller.py:4-11."""
from oracle.blocks import (Attention, ResnetBlock2D, Transformer2DModel, UNetMidBlock2D,  # noqa: F401
                           UNetMidBlock2DCrossAttn, get_down_block)
