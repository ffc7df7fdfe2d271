"""Scheduler known-answer tests: DDIM trailing timesteps (int64, bit-exact), alpha-bar table,
N=1 DDIM affine form, sinusoid constants; product scheduler == oracle scheduler bit for bit."""
import math

import pytest
import torch

from oracle import schedulers as OS
from unirestore_b200.diffuie import schedulers as PS

KAT_TS = {1: [999], 4: [999, 749, 499, 249], 3: [999, 666, 332], 10: list(range(999, 0, -100)),
          20: list(range(999, 0, -50)), 50: list(range(999, 0, -20))}
KAT_ALPHA = {0: 0.99914998, 49: 0.95262527, 249: 0.67543209, 499: 0.27766943, 749: 0.05662345, 999: 0.0046600951}


@pytest.mark.parametrize("n", sorted(KAT_TS))
def test_trailing_timesteps_bit_exact(n):
    for S in (OS.DDIMScheduler, PS.DDIMScheduler):
        s = S()
        s.set_timesteps(n)
        assert s.timesteps.dtype == torch.int64
        assert s.timesteps.tolist() == KAT_TS[n]
    p = PS.DDIMScheduler()
    p.set_timesteps(n)
    assert p.timesteps_host == KAT_TS[n]
    assert p.prev_timestep(KAT_TS[n][-1]) == KAT_TS[n][-1] - 1000 // n


def test_alphas_cumprod_table():
    a_o, a_p = OS.make_alphas_cumprod(), PS.make_alphas_cumprod()
    assert torch.equal(a_o, a_p)
    for t, v in KAT_ALPHA.items():
        assert abs(a_p[t].item() - v) < 5e-8 * max(1.0, v / 1e-2), (t, a_p[t].item())
    assert abs(a_p[999].item() ** 0.5 - 0.068265) < 1e-6 and abs((1 - a_p[999].item()) ** 0.5 - 0.997667) < 1e-6


def test_ddim_single_step_affine_form():
    """N=1: x_prev = 14.642592 x - 14.579279 eps (final alpha = alpha_0, no clipping)."""
    s = PS.DDIMScheduler()
    s.set_timesteps(1)
    sa, sb, sap, sbp = s.step_coefficients(999)
    assert abs(sap / sa - 14.642592) < 2e-4 and abs(sbp - sap * sb / sa + 14.579279) < 2e-4
    o = OS.DDIMScheduler()
    o.set_timesteps(1)
    x, e = torch.randn(2, 4, 8, 8), torch.randn(2, 4, 8, 8)
    ref = o.step(e, torch.tensor(999), x).prev_sample
    x0 = (x - sb * e) / sa
    assert torch.allclose(sap * x0 + sbp * e, ref, rtol=1e-6, atol=1e-6)


@pytest.mark.parametrize("n", [2, 4, 20])
def test_step_coefficients_match_oracle(n):
    o, p = OS.DDIMScheduler(), PS.DDIMScheduler()
    o.set_timesteps(n), p.set_timesteps(n)
    x, e = torch.randn(1, 4, 4, 4), torch.randn(1, 4, 4, 4)
    for t in p.timesteps_host:
        sa, sb, sap, sbp = p.step_coefficients(t)
        ref = o.step(e, torch.tensor(t), x).prev_sample
        got = torch.tensor(sap) * ((x - torch.tensor(sb) * e) / torch.tensor(sa)) + torch.tensor(sbp) * e
        assert torch.equal(got, ref), t


def test_noise_coefficients_match_oracle():
    o, p = OS.DDPMScheduler(), PS.DDPMScheduler()
    x, e = torch.randn(2, 4, 4, 4), torch.randn(2, 4, 4, 4)
    for t in (249, 499, 749, 999):
        sa, sb = p.noise_coefficients(t)
        ref = o.add_noise(x, e, torch.tensor([t, t]))
        assert torch.equal(torch.tensor(sa) * x + torch.tensor(sb) * e, ref)


def test_sinusoid_known_answers():
    from oracle.blocks import Timesteps
    emb = Timesteps(320, True, 0)(torch.tensor([999]))[0]
    for got, want in zip(emb[:3].tolist(), [0.99964982, 0.80267751, -0.27806213]):
        assert abs(got - want) < 2e-4
    for got, want in zip(emb[160:163].tolist(), [-0.02646075, 0.59641331, -0.96056312]):
        assert abs(got - want) < 2e-4
    assert abs(math.cos(999.0) - emb[0].item()) < 2e-4
