"""Shim for controller.py:3."""
from oracle.blocks import TimestepEmbedding, Timesteps  # noqa: F401
