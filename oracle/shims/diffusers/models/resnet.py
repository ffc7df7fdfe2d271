"""Shim for base_model.py:6."""
from oracle.blocks import ResnetBlock2D  # noqa: F401
