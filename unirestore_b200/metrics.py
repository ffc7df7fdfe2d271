"""GPU-side PSNR / SSIM for the validate loop -- mirrors the reference's ``SKPSNR`` / ``SKSSIM``
(src/core/base/eval_image_restoration.py:255-313: ``update(preds, targets)`` / ``compute()``), but the per-image sums are
produced by one CUDA kernel (``ur_image_metrics``) on the restored batch instead of a D2H copy + skimage per image.
Only two doubles per image cross to the host, at ``compute()`` time."""
import math

import torch

from . import ops
from ._cabi import check


def image_metric_sums(preds, targets, quantize=False, data_range=1.0):
    """-> fp64 [B, 2]: (sum of squared error, sum of the SSIM map) per image; preds/targets CUDA fp32 [B,C,H,W]."""
    if preds.shape != targets.shape or preds.dim() != 4:
        raise ValueError("preds / targets must be [B,C,H,W] tensors of the same shape")
    if not preds.is_cuda:
        raise ValueError("image metrics run on the GPU (the product path has no CPU fallback)")
    p, t = preds.float().contiguous(), targets.float().contiguous()
    B, C, H, W = p.shape
    out = torch.empty((B, 2), device=p.device, dtype=torch.float64)
    check(ops._lib().ur_image_metrics(ops._ptr(p), ops._ptr(t), B, C, H, W, int(quantize), float(data_range),
                                      ops._ptr(out), ops._stream()), "ur_image_metrics")
    return out


class _ImageMetric:
    def __init__(self, data_range: float = 1.0, quantize: bool = False):
        self.data_range, self.quantize = data_range, quantize
        self.reset()

    def reset(self):
        self._sums, self.total = [], 0

    def update(self, preds, targets):
        B, C, H, W = preds.shape
        self._sums.append((image_metric_sums(preds, targets, self.quantize, self.data_range), C, H, W))
        self.total += B


class SKPSNR(_ImageMetric):
    """mean over images of 10 log10(data_range^2 / mse) (peak_signal_noise_ratio(target, pred))."""

    def compute(self):
        acc = 0.0
        for s, C, H, W in self._sums:
            for se in s[:, 0].tolist():
                # identical images (mse == 0): skimage's peak_signal_noise_ratio returns inf (division by zero warning)
                acc += 10.0 * math.log10(self.data_range ** 2 / (se / (C * H * W))) if se > 0.0 else float("inf")
        return acc / max(self.total, 1)


class SKSSIM(_ImageMetric):
    """mean over images of structural_similarity(pred, target, channel_axis=0) with skimage defaults."""

    def compute(self):
        acc = 0.0
        for s, C, H, W in self._sums:
            acc += sum(v / (C * (H - 6) * (W - 6)) for v in s[:, 1].tolist())
        return acc / max(self.total, 1)
