"""Known answers for the metrics oracle (restated skimage PSNR / SSIM; CPU only)."""
import numpy as np

from oracle import metrics as M


def test_psnr_known_values():
    t = np.zeros((3, 8, 8), dtype=np.float32)
    p = np.full((3, 8, 8), 0.1, dtype=np.float32)
    assert abs(M.psnr(t, p) - 20.0) < 1e-5                 # mse = 0.01 -> 10 log10(1 / 0.01)
    assert abs(M.psnr(t, p, data_range=2.0) - (20.0 + 10 * np.log10(4.0))) < 1e-5


def test_ssim_identities():
    rng = np.random.default_rng(0)
    a = rng.random((3, 32, 40)).astype(np.float32)
    assert abs(M.ssim(a, a) - 1.0) < 1e-12
    b = np.clip(a + 0.1 * rng.standard_normal(a.shape), 0, 1).astype(np.float32)
    s = M.ssim(b, a)
    assert 0.0 < s < 1.0 and abs(s - M.ssim(a, b)) < 1e-12    # symmetric
    # constant images: S = (2 ux uy + C1) / (ux^2 + uy^2 + C1)
    c = np.full((1, 16, 16), 0.5), np.full((1, 16, 16), 0.25)
    assert abs(M.ssim(*c) - (2 * 0.5 * 0.25 + 1e-4) / (0.25 + 0.0625 + 1e-4)) < 1e-12


def test_quantize8_matches_torch_rounding():
    import torch
    x = torch.linspace(-0.1, 1.1, 4001)
    ref = x.mul(255).round_().clamp_(0, 255).div_(255).numpy()
    assert np.array_equal(M.quantize8(x.numpy()), ref)
