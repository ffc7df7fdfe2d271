"""Shim for nafnet_arch.py:19."""
from oracle.blocks import LayerNorm2d  # noqa: F401
