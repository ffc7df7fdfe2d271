"""Layout plumbing between the reference-facing NCHW fp32 tensors and the internal bf16 channels-last ones."""
import torch


def to_nhwc(x: torch.Tensor) -> torch.Tensor:
    """[B,C,H,W] any float dtype -> dense bf16 [B,H,W,C] (a free view + cast for channels_last inputs)."""
    return x.permute(0, 2, 3, 1).to(torch.bfloat16).contiguous()


def to_nchw(x: torch.Tensor, dtype=torch.float32) -> torch.Tensor:
    """bf16 [B,H,W,C] -> [B,C,H,W] in ``dtype`` (logical NCHW, channels_last memory)."""
    return x.permute(0, 3, 1, 2).to(dtype)
