"""CPU restatement of the image-quality metrics the reference's validate loop computes right after the hot path
(TEST INFRASTRUCTURE ONLY -- imported by tests/).

Reference: src/core/base/eval_image_restoration.py:71 (8-bit quantisation of the prediction), :255-313 (``SKPSNR`` /
``SKSSIM`` wrap ``skimage.metrics.peak_signal_noise_ratio(target, pred, data_range=1)`` and
``structural_similarity(pred, target, data_range=1, channel_axis=0)``).  scikit-image is an un-vendored dependency
(requirements.txt:20, unpinned) and absent offline, so its published algorithm is restated here with
scipy.ndimage.uniform_filter: **parity unpinned for this leaf** (no skimage golden can be generated in this image).

  PSNR = 10 log10(data_range^2 / mean((t - p)^2))
  SSIM (defaults: win_size 7, uniform window, K1 .01, K2 .03, sample covariance): per channel
      ux, uy, uxx, uyy, uxy = 7x7 box means;  cov_norm = 49 / 48
      vx = cov_norm (uxx - ux^2), vy likewise, vxy = cov_norm (uxy - ux uy)
      S = (2 ux uy + C1)(2 vxy + C2) / ((ux^2 + uy^2 + C1)(vx + vy + C2));   mean of S over the image cropped by 3 pixels
  and the mean over channels.
"""
import numpy as np
from scipy.ndimage import uniform_filter


def quantize8(x):
    """pred.mul(255).round_().clamp_(0, 255).div_(255) (eval_image_restoration.py:71); round half to even like torch."""
    return np.clip(np.rint(np.asarray(x, dtype=np.float32) * np.float32(255.0)), 0, 255).astype(np.float32) / np.float32(255.0)


def psnr(target, pred, data_range=1.0):
    t, p = np.asarray(target, dtype=np.float64), np.asarray(pred, dtype=np.float64)
    return 10.0 * np.log10(data_range ** 2 / np.mean((t - p) ** 2))


def ssim(pred, target, data_range=1.0, win=7):
    """pred / target: [C, H, W]."""
    x, y = np.asarray(pred, dtype=np.float64), np.asarray(target, dtype=np.float64)
    c1, c2 = (0.01 * data_range) ** 2, (0.03 * data_range) ** 2
    npx = win * win
    cov_norm = npx / (npx - 1.0)
    pad = (win - 1) // 2
    vals = []
    for ch in range(x.shape[0]):
        a, b = x[ch], y[ch]
        ux, uy = uniform_filter(a, size=win), uniform_filter(b, size=win)
        uxx, uyy, uxy = uniform_filter(a * a, size=win), uniform_filter(b * b, size=win), uniform_filter(a * b, size=win)
        vx, vy, vxy = cov_norm * (uxx - ux * ux), cov_norm * (uyy - uy * uy), cov_norm * (uxy - ux * uy)
        s = ((2 * ux * uy + c1) * (2 * vxy + c2)) / ((ux * ux + uy * uy + c1) * (vx + vy + c2))
        vals.append(s[pad:-pad, pad:-pad].mean())
    return float(np.mean(vals))
