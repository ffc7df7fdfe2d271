"""Block-level parity at IDENTICAL ROUNDING POINTS: every CUDA module against
``oracle/rounded.py`` -- the reference's algorithm on the same oracle modules with bf16 rounding exactly where the kernels
store bf16 -- run on the GPU in fp32 (TF32 off).  Gates are <= 2x the measured value (profiles/parity_r2.txt); the
distance to the pure-fp32 oracle is logged next to it (that one measures bf16 storage, not implementation error)."""
import pytest
import torch

from tests.util import _log, assert_close, rel_l2

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
TOL = 2e-3            # identical rounding points: accumulation order + rare rounding flips only


@pytest.fixture(autouse=True, scope="module")
def _fp32_reference_math():
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    yield


def rnd(seed, *shape, scale=1.0):
    return torch.randn(*shape, generator=torch.Generator().manual_seed(seed)) * scale


def pair(make_oracle, make_cuda, prefix=""):
    from unirestore_b200.init_utils import deterministic_init_
    o = deterministic_init_(make_oracle(), prefix).eval().requires_grad_(False)
    m = make_cuda().eval().requires_grad_(False)
    m.load_state_dict(o.state_dict(), strict=True)
    return o.to(DEV), m.to(DEV)


def nhwc(x):
    return x.permute(0, 2, 3, 1).contiguous().to(torch.bfloat16)


def nchw(x):
    return x.permute(0, 3, 1, 2).float()


def bq(x):
    return x.to(torch.bfloat16).float()


def check(got, ref_q, ref32, what, tol=TOL):
    e = assert_close(got, ref_q, tol, what + " vs rounded oracle")
    _log("%-60s rel-L2 %.3e (vs fp32 oracle, informational)" % (what, rel_l2(got.float(), ref32.float())))
    return e


@pytest.mark.parametrize("cin,cout,temb,two,hw", [(64, 64, True, False, (12, 20)), (64, 128, True, False, (12, 20)),
                                                  (128, 64, False, False, (12, 20)), (192, 64, True, True, (12, 20)),
                                                  (320, 320, True, False, (32, 32)), (1280, 1280, True, False, (8, 8)),
                                                  (2560, 1280, True, True, (8, 8)), (960, 640, True, True, (16, 16))])
def test_resnet_block(cin, cout, temb, two, hw):
    from oracle import blocks as OB, rounded as R
    from unirestore_b200.diffuie import sd_blocks as SB
    kw = dict(in_channels=cin, out_channels=cout, temb_channels=96 if temb else None, groups=32, eps=1e-5)
    o, m = pair(lambda: OB.ResnetBlock2D(**kw), lambda: SB.ResnetBlock2D(**kw))
    x = bq(rnd(1, 2, cin, *hw)).to(DEV)
    e = rnd(2, 1, 96).to(DEV) if temb else None
    xb = nhwc(x)
    c1 = (cin * 2) // 3 if cin % 3 == 0 else cin // 2
    y = m.run(xb[..., :c1].contiguous(), e, x2=xb[..., c1:].contiguous()) if two else m.run(xb, e)
    check(nchw(y), R.resnet(o, x, e), o(x, e), "ResnetBlock2D %d->%d @%dx%d" % (cin, cout, *hw))


def test_resample():
    from oracle import blocks as OB, rounded as R
    from unirestore_b200.diffuie import sd_blocks as SB
    x = bq(rnd(3, 2, 64, 16, 24)).to(DEV)
    for pad in (1, 0):
        o, m = pair(lambda: OB.Downsample2D(64, True, 64, pad), lambda: SB.Downsample2D(64, True, 64, pad))
        check(nchw(m.run(nhwc(x))), R.downsample(o, x), o(x), "Downsample2D pad=%d" % pad)
    o, m = pair(lambda: OB.Upsample2D(64, True, 128), lambda: SB.Upsample2D(64, True, 128))
    check(nchw(m.run(nhwc(x))), R.upsample(o, x), o(x), "Upsample2D")


@pytest.mark.parametrize("c,heads,hw", [(256, 4, (8, 8)), (512, 4, (16, 16)), (512, 1, (8, 8)), (512, 1, (32, 32))])
def test_spatial_attention(c, heads, hw):
    from oracle import blocks as OB, rounded as R
    from unirestore_b200.diffuie import sd_blocks as SB
    kw = dict(heads=heads, dim_head=c // heads, eps=1e-5, norm_num_groups=32, residual_connection=True, bias=True)
    o, m = pair(lambda: OB.Attention(c, **kw), lambda: SB.Attention(c, **kw))
    x = bq(rnd(5, 2, c, *hw)).to(DEV)
    check(nchw(m.run(nhwc(x))), R.attention_spatial(o, x), o(x), "Attention c=%d h=%d @%dx%d" % (c, heads, *hw))


@pytest.mark.parametrize("c,heads,hw", [(320, 5, (16, 16)), (640, 10, (8, 8)), (1280, 20, (8, 8))])
def test_transformer2d(c, heads, hw):
    from oracle import blocks as OB, rounded as R
    from unirestore_b200.diffuie import sd_blocks as SB
    o, m = pair(lambda: OB.Transformer2DModel(heads, c // heads, c, 1024),
                lambda: SB.Transformer2DModel(heads, c // heads, c, 1024))
    x = bq(rnd(6, 2, c, *hw)).to(DEV)
    ctx = bq(rnd(7, 1, 77, 1024)).to(DEV)
    y = m.run(nhwc(x), ctx.to(torch.bfloat16).contiguous())
    check(nchw(y), R.transformer2d(o, x, ctx.expand(2, -1, -1)), o(x, ctx.expand(2, -1, -1), return_dict=False)[0],
          "Transformer2DModel c=%d @%dx%d" % (c, *hw), tol=4e-3)       # ~15 rounding points in series


def test_scedit_cfrm_tfa():
    from oracle import rounded as R, unirestore as O
    from unirestore_b200 import diffuie as D
    for c in (320, 1280):
        o, m = pair(lambda: O.CSCEAdapter(c, c, 256), lambda: D.CSCEAdapter(c, c, 256))
        x, cd = bq(rnd(8, 2, c, 16, 16)).to(DEV), bq(rnd(9, 2, 256, 16, 16)).to(DEV)
        check(m(x, cd), R.scedit(o, x, cd), o(x, cd), "CSCEAdapter(%d)" % c)
    for c in (128, 256):
        x = bq(rnd(10, 1, c, 16, 24)).to(DEV)
        o, m = pair(lambda: O.NAFBlock(c), lambda: D.NAFBlock(c))
        check(m(x), R.nafblock(o, x), o(x), "NAFBlock(%d)" % c)
        o, m = pair(lambda: O.AdaNAFV2(c), lambda: D.AdaNAFV2(c))
        check(m(x), R.adanaf(o, x), o(x), "AdaNAFV2(%d)" % c, tol=3e-3)
    for co, cs, last in ((512, 256, False), (512, 128, True)):
        o, m = pair(lambda: O.TaskFeatureAdapter(co, cs, 1, last), lambda: D.TaskFeatureAdapter(co, cs, 1, last))
        x, s, cond = bq(rnd(11, 2, co, 8, 12)).to(DEV), bq(rnd(12, 2, cs, 8, 12, scale=2.0)).to(DEV), rnd(13, 2, 1, cs).to(DEV)
        (yx, yc), (rx, rc), (fx, fc) = m(x, s, cond), R.tfa(o, x, s, cond), o(x, s, cond)
        check(yx, rx, fx, "TFA(%d,%d) x" % (co, cs))
        if not last:
            check(yc, rc, fc, "TFA(%d,%d) cond" % (co, cs))


@pytest.fixture(scope="module")
def models():
    from oracle import unirestore as O
    from unirestore_b200.diffuie import DiffUIE
    from unirestore_b200.init_utils import deterministic_init_
    cfg = (dict(type="CFRM"), dict(type="scedit", num_inference_steps=2), dict(type="TFA", prompt_len=1, task=["ir", "seg"]))
    o = deterministic_init_(O.DiffUIE(*cfg)).eval().requires_grad_(False)
    m = DiffUIE(*cfg).eval().requires_grad_(False)
    m.load_state_dict(o.state_dict(), strict=True)
    return o.to(DEV), m.to(DEV)


def test_controller_unet_encode_decode(models):
    """The assembled networks (sd-turbo topology) at a 32x32 latent / 256x256 image."""
    from oracle import rounded as R
    o, m = models
    z0, zt, t = rnd(20, 2, 4, 32, 32).to(DEV), rnd(21, 2, 4, 32, 32).to(DEV), torch.tensor([499], device=DEV)
    with torch.no_grad():
        ctl, ctl_q, ctl_f = m.controller(z0, t), R.controller(o.controller, z0, t), o.controller(z0, t)
        for k in ctl_q:
            check(ctl[k], ctl_q[k], ctl_f[k], "Controller[%d]" % k, tol=1.3e-2)     # 10..20 blocks in series
        # feed the SAME (rounded-oracle) control tensors to both sides so the UNet comparison starts from equal inputs
        eps = m.base_model(zt, ctl_q, t)
        check(eps, R.controlled_unet(o.base_model, zt, ctl_q, t), o.base_model(zt, ctl_f, t), "ControlledUNet", tol=1.5e-2)
        img = torch.rand(2, 3, 256, 256, generator=torch.Generator().manual_seed(3)).to(DEV)
        noise = rnd(22, 2, 4, 32, 32).to(DEV)
        z, skips = m.ae.encode(img, enable_fr=True, noise=noise)
        zq, sq = R.encode(o.ae, img, enable_fr=True, noise=noise)
        zf, sf = o.ae.encode(img, enable_fr=True, noise=noise)
        check(z, zq, zf, "encode z (+CFRM)", tol=1.5e-2)
        for i in range(3):
            check(skips[i].float(), sq[i], sf[i], "encode skip%d" % i, tol=1.5e-2)
        z2, _ = m.ae.encode(img, enable_fr=False, noise=noise)                       # engine_unifie.py:139 (stage-1 target)
        zq2, _ = R.encode(o.ae, img, enable_fr=False, noise=noise)
        check(z2, zq2, o.ae.encode(img, enable_fr=False, noise=noise)[0], "encode z (enable_fr=False)", tol=1.5e-2)
        for task in ("ir", "seg"):                                                    # double decode engine_unifie.py:220-222
            y = m.ae.decode(zq, sq, task)
            check(y, R.decode(o.ae, zq, sq, task), o.ae.decode(zq, sf, task), "decode[%s] (+TFA)" % task, tol=1.5e-2)
