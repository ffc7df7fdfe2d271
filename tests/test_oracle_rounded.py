"""CPU: ``oracle/rounded.py`` (the bf16-rounding-point oracle the GPU block tests gate against) with rounding switched
off must reproduce the fp32 oracle modules -- which are pinned to the reference goldens (tests/test_oracle_golden.py)
-- so the rounded restatement cannot drift from the reference algorithm.  With rounding on it must stay within a
bf16-pipeline distance of the fp32 result (sanity of the rounding points, not a parity claim)."""
import pytest
import torch

from oracle import blocks as OB
from oracle import rounded as R
from oracle import unirestore as O
from tests.util import rel_l2
from unirestore_b200.init_utils import deterministic_init_


def rnd(seed, *shape, scale=1.0):
    return torch.randn(*shape, generator=torch.Generator().manual_seed(seed)) * scale


def both(fn_rounded, fn_oracle, tol_exact=2e-6, tol_q=3e-2):
    with torch.no_grad():
        ref = fn_oracle()
        with R.exact():
            ex = fn_rounded()
        rq = fn_rounded()
    pairs = zip(ex, ref, rq) if isinstance(ref, (tuple, list)) else [(ex, ref, rq)]
    for a, b, c in pairs:
        if b is None:
            assert a is None and c is None
            continue
        assert rel_l2(a, b) < tol_exact, rel_l2(a, b)
        e = rel_l2(c, b)
        assert 0 < e < tol_q, e                       # rounding is on (e > 0) and stays at the bf16 level


def mk(mod, prefix=""):
    return deterministic_init_(mod, prefix).eval().requires_grad_(False)


@pytest.mark.parametrize("cin,cout,temb", [(64, 64, True), (64, 128, True), (128, 64, False)])
def test_resnet(cin, cout, temb):
    m = mk(OB.ResnetBlock2D(in_channels=cin, out_channels=cout, temb_channels=96 if temb else None, groups=32, eps=1e-5))
    x, e = R.q(rnd(1, 2, cin, 10, 12)), (rnd(2, 1, 96) if temb else None)
    both(lambda: R.resnet(m, x, e), lambda: m(x, e))


@pytest.mark.parametrize("pad", [1, 0])
def test_resample(pad):
    d = mk(OB.Downsample2D(64, True, 64, pad))
    u = mk(OB.Upsample2D(64, True, 96))
    x = R.q(rnd(3, 2, 64, 8, 10))
    both(lambda: R.downsample(d, x), lambda: d(x))
    both(lambda: R.upsample(u, x), lambda: u(x))


@pytest.mark.parametrize("c,heads", [(128, 2), (512, 1)])
def test_spatial_attention(c, heads):
    m = mk(OB.Attention(c, heads=heads, dim_head=c // heads, eps=1e-5, norm_num_groups=32, residual_connection=True,
                        bias=True))
    x = R.q(rnd(5, 2, c, 6, 6))
    both(lambda: R.attention_spatial(m, x), lambda: m(x))


def test_transformer2d():
    m = mk(OB.Transformer2DModel(5, 64, 320, 1024))
    x, ctx = R.q(rnd(6, 2, 320, 4, 6)), rnd(7, 1, 77, 1024)
    both(lambda: R.transformer2d(m, x, R.q(ctx).expand(2, -1, -1)),
         lambda: m(x, ctx.expand(2, -1, -1), return_dict=False)[0])


def test_scedit_naf_adanaf_tfa():
    s = mk(O.CSCEAdapter(64, 64, 32))
    x, c = R.q(rnd(8, 2, 64, 6, 6)), R.q(rnd(9, 2, 32, 6, 6))
    both(lambda: R.scedit(s, x, c), lambda: s(x, c))
    n = mk(O.NAFBlock(64))
    both(lambda: R.nafblock(n, x), lambda: n(x))
    a = mk(O.AdaNAFV2(64))
    both(lambda: R.adanaf(a, x), lambda: a(x))
    for last in (False, True):
        t = mk(O.TaskFeatureAdapter(64, 32, 1, last))
        sk, cond = R.q(rnd(10, 2, 32, 6, 6, scale=2.0)), rnd(11, 2, 1, 32)
        both(lambda: R.tfa(t, x, sk, cond), lambda: t(x, sk, cond))


def test_controller_and_unet_small():
    ctl = mk(O.Controller(model_channels=64, out_channels=32, num_heads=1), "controller.")
    z, t = rnd(12, 1, 4, 16, 16), torch.tensor([499])
    with torch.no_grad():
        ref = ctl(z, t)
        with R.exact():
            ex = R.controller(ctl, z, t)
        rq = R.controller(ctl, z, t)
    for k in ref:
        assert rel_l2(ex[k], ref[k]) < 2e-6
        assert 0 < rel_l2(rq[k], ref[k]) < 5e-2


@pytest.fixture(scope="module")
def full_model():
    cfg = (dict(type="CFRM"), dict(type="scedit", num_inference_steps=2), dict(type="TFA", prompt_len=1, task=["ir", "seg"]))
    return mk(O.DiffUIE(*cfg))


def test_full_topology_unet_encode_decode(full_model):
    """The assembled sd-turbo-topology model at small spatial sizes: ControlledUNet (+Controller, +SC-Tuner),
    encode (+CFRM) and decode (+TFA) walked by oracle/rounded.py == the fp32 oracle with rounding off."""
    m = full_model
    zt, z0, t = rnd(20, 1, 4, 16, 16), rnd(21, 1, 4, 16, 16), torch.tensor([499])
    with torch.no_grad():
        ref = m.base_model(zt, m.controller(z0, t), t)
        with R.exact():
            ex = R.controlled_unet(m.base_model, zt, R.controller(m.controller, z0, t), t)
            pz = R.predict_z0(m, zt, z0, t)
        assert rel_l2(ex, ref) < 5e-6
        assert rel_l2(pz, m.predict_z0(zt, z0, t)) < 5e-6
        rq = R.controlled_unet(m.base_model, zt, R.controller(m.controller, z0, t), t)
        assert 0 < rel_l2(rq, ref) < 5e-2
        img, noise = torch.rand(1, 3, 64, 64, generator=torch.Generator().manual_seed(3)), rnd(22, 1, 4, 8, 8)
        z_ref, sk_ref = m.ae.encode(img, enable_fr=True, noise=noise)
        with R.exact():
            z_ex, sk_ex = R.encode(m.ae, img, enable_fr=True, noise=noise)
            d_ex = R.decode(m.ae, z_ref, sk_ref, "seg")
        assert rel_l2(z_ex, z_ref) < 5e-6
        for a, b in zip(sk_ex, sk_ref):
            assert rel_l2(a, b) < 5e-6
        assert rel_l2(d_ex, m.ae.decode(z_ref, sk_ref, "seg")) < 5e-6
