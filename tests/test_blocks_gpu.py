"""Block-level parity of the CUDA modules (unirestore_b200.diffuie) against the CPU/GPU oracle and against the
golden fixtures produced by executing the reference's own files (oracle/make_golden.py).

The CUDA path computes in bf16 (fp32 accumulate, fp32 statistics) while the oracle / goldens are fp32, so the
gate here is a bf16-pipeline tolerance (rel-L2, written per test); kernel-level parity at identical rounding
points (<= 1e-3) is asserted in test_conv_gemm_gpu.py / test_kernels_gpu.py."""
import os

import pytest
import torch

from tests.util import assert_close, bf16_round

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")
TOL_BLOCK = 7.5e-3    # one block, bf16 activations vs fp32 oracle: <= 2x the measured 2.4e-3..3.8e-3 (profiles/parity_r2.txt);
                      # the gates at identical rounding points (<= 2e-3) live in test_parity_rounded_gpu.py
DEV = "cuda:0"


@pytest.fixture(autouse=True, scope="module")
def _fp32_reference_math():
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    yield


def rnd(seed, *shape, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(*shape, generator=g) * scale


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


@pytest.mark.parametrize("cin,cout,temb,two", [(64, 64, True, False), (64, 128, True, False), (128, 64, False, False),
                                               (192, 64, True, True), (320, 320, True, False)])
def test_resnet_block(cin, cout, temb, two):
    from oracle import blocks as OB
    from unirestore_b200.diffuie import sd_blocks as SB
    kw = dict(in_channels=cin, out_channels=cout, temb_channels=96 if temb else None, groups=32, eps=1e-5)
    o, m = pair(lambda: OB.ResnetBlock2D(**kw), lambda: SB.ResnetBlock2D(**kw))
    x = rnd(1, 2, cin, 12, 20).to(DEV)
    e = rnd(2, 1, 96).to(DEV) if temb else None
    ref = o(x, e)
    xb = nhwc(x)
    if two:
        y = m.run(xb[..., :128].contiguous(), e, x2=xb[..., 128:].contiguous())
    else:
        y = m.run(xb, e)
    assert_close(nchw(y), ref, TOL_BLOCK, "ResnetBlock2D")


@pytest.mark.parametrize("pad", [1, 0])
def test_downsample(pad):
    from oracle import blocks as OB
    from unirestore_b200.diffuie import sd_blocks as SB
    o, m = pair(lambda: OB.Downsample2D(64, True, 64, pad), lambda: SB.Downsample2D(64, True, 64, pad))
    x = rnd(3, 2, 64, 16, 24).to(DEV)
    assert_close(nchw(m.run(nhwc(x))), o(x), TOL_BLOCK, "Downsample2D pad=%d" % pad)


def test_upsample_subpixel():
    from oracle import blocks as OB
    from unirestore_b200.diffuie import sd_blocks as SB
    o, m = pair(lambda: OB.Upsample2D(64, True, 128), lambda: SB.Upsample2D(64, True, 128))
    x = rnd(4, 2, 64, 8, 12).to(DEV)
    assert_close(nchw(m.run(nhwc(x))), o(x), TOL_BLOCK, "Upsample2D")


@pytest.mark.parametrize("B,H,W", [(2, 8, 12), (3, 8, 8), (1, 40, 24), (5, 4, 4)])
def test_upsample_carries_fused_statistics(B, H, W):
    """The four sub-pixel phase GEMMs accumulate the GroupNorm statistics of the up-sampled tensor into one buffer in their
    epilogues (ragged tiles, 1 / 2 images per M tile); below 32 pixels per image the tensor carries none and its
    consumer runs its own statistics pass."""
    from unirestore_b200 import ops
    from unirestore_b200.diffuie import sd_blocks as SB
    torch.manual_seed(7)
    m = SB.Upsample2D(64, True, 128).to(DEV)
    x = rnd(6, B, 64, H, W).to(DEV)
    y = m.run(nhwc(x))
    st = getattr(y, "_ur_stats", None)
    if H * W < 32:
        assert st is None
        return
    ref = ops.chan_stats(y.contiguous())
    torch.cuda.synchronize()
    scale = ref.abs().amax().clamp_min(1e-9)
    # (fp32 partial sums per tile in the epilogue, per pixel lane in the statistics pass: ~1e-7 apart)
    assert ((st - ref).abs() / scale).max().item() < 1e-6, "fused up-sample statistics differ from a statistics pass"


@pytest.mark.parametrize("c,heads,hw", [(256, 4, (8, 8)), (512, 4, (4, 6)), (512, 1, (8, 8)), (64, 1, (2, 2))])
def test_spatial_attention(c, heads, hw):
    from oracle import blocks as OB
    from unirestore_b200.diffuie import sd_blocks as SB
    kw = dict(heads=heads, dim_head=c // heads, eps=1e-5, norm_num_groups=32, residual_connection=True, bias=True)
    o, m = pair(lambda: OB.Attention(c, **kw), lambda: SB.Attention(c, **kw))
    x = rnd(5, 2, c, *hw).to(DEV)
    assert_close(nchw(m.run(nhwc(x))), o(x), TOL_BLOCK, "Attention")


@pytest.mark.parametrize("c,heads,hw", [(320, 5, (8, 8)), (640, 10, (4, 4))])
def test_transformer2d(c, heads, hw):
    from oracle import blocks as OB
    from unirestore_b200.diffuie import sd_blocks as SB
    o, m = pair(lambda: OB.Transformer2DModel(heads, c // heads, c, 1024), lambda: SB.Transformer2DModel(heads, c // heads, c, 1024))
    x = rnd(6, 2, c, *hw).to(DEV)
    ctx = rnd(7, 1, 77, 1024).to(DEV)
    ref = o(x, ctx.expand(2, -1, -1), return_dict=False)[0]
    y = m.run(nhwc(x), ctx.to(torch.bfloat16).contiguous())
    assert_close(nchw(y), ref, TOL_BLOCK, "Transformer2DModel")


def test_scedit_golden():
    from unirestore_b200.diffuie import CSCEAdapter
    from unirestore_b200.init_utils import deterministic_init_
    g = torch.load(os.path.join(GOLD, "scedit.pt"))
    m = deterministic_init_(CSCEAdapter(*g["args"])).eval().to(DEV)
    x, c = rnd(g["x_seed"], *g["x_shape"]).to(DEV), rnd(g["c_seed"], *g["c_shape"]).to(DEV)
    assert_close(m(x, c), g["out"].to(DEV), TOL_BLOCK, "CSCEAdapter vs reference golden")


def test_cfrm_golden():
    from unirestore_b200.diffuie import AdaNAFV2, NAFBlock
    from unirestore_b200.init_utils import deterministic_init_
    g = torch.load(os.path.join(GOLD, "cfrm.pt"))
    x = rnd(g["x_seed"], *g["x_shape"], scale=g["x_scale"]).to(DEV)
    naf = deterministic_init_(NAFBlock(g["c"])).eval().to(DEV)
    assert_close(naf(x), g["naf"].to(DEV), TOL_BLOCK, "NAFBlock vs reference golden")
    ada = deterministic_init_(AdaNAFV2(g["c"])).eval().to(DEV)
    assert_close(ada(x), g["ada"].to(DEV), TOL_BLOCK, "AdaNAFV2 vs reference golden")


@pytest.mark.parametrize("c", [128, 256])
def test_cfrm_wide_vs_oracle(c):
    from oracle import unirestore as O
    from unirestore_b200.diffuie import AdaNAFV2
    o, m = pair(lambda: O.AdaNAFV2(c), lambda: AdaNAFV2(c))
    x = rnd(8, 1, c, 16, 24).to(DEV)
    assert_close(m(x), o(x), TOL_BLOCK, "AdaNAFV2(%d)" % c)


def test_tfa_golden():
    from unirestore_b200.diffuie import TaskFeatureAdapter
    from unirestore_b200.init_utils import deterministic_init_
    for case in torch.load(os.path.join(GOLD, "taskeditor.pt")):
        co, cs, pl, last = case["args"]
        m = deterministic_init_(TaskFeatureAdapter(co, cs, pl, last)).eval().to(DEV)
        s1, s2, s3 = case["seeds"]
        h, w = case["hw"]
        x, s, cond = rnd(s1, 2, co, h, w).to(DEV), rnd(s2, 2, cs, h, w, scale=2.0).to(DEV), rnd(s3, 2, pl, cs).to(DEV)
        yx, yc = m(x, s, cond)
        assert_close(yx, case["out_x"].to(DEV), TOL_BLOCK, "TFA x %s" % (case["args"],))
        if case["out_cond"] is not None:
            assert_close(yc, case["out_cond"].to(DEV), TOL_BLOCK, "TFA cond %s" % (case["args"],))


@pytest.mark.parametrize("c_out,c_skip", [(512, 128), (512, 256)])
def test_tfa_wide_vs_oracle(c_out, c_skip):
    from oracle import unirestore as O
    from unirestore_b200.diffuie import TaskFeatureAdapter
    o, m = pair(lambda: O.TaskFeatureAdapter(c_out, c_skip, 1, False), lambda: TaskFeatureAdapter(c_out, c_skip, 1, False))
    x, s, cond = rnd(9, 2, c_out, 8, 12).to(DEV), rnd(10, 2, c_skip, 8, 12, scale=2.0).to(DEV), rnd(11, 2, 1, c_skip).to(DEV)
    (rx, rc), (yx, yc) = o(x, s, cond), m(x, s, cond)
    assert_close(yx, rx, TOL_BLOCK, "TFA x")
    assert_close(yc, rc, TOL_BLOCK, "TFA cond")


def test_controller_golden():
    from unirestore_b200.diffuie import Controller, stablesr_config
    from unirestore_b200.init_utils import deterministic_init_
    g = torch.load(os.path.join(GOLD, "controller.pt"))
    m = deterministic_init_(Controller(**stablesr_config), "controller.").eval().to(DEV)
    y = m(rnd(g["x_seed"], *g["x_shape"]).to(DEV), torch.tensor([g["t"]], device=DEV))
    assert set(y) == set(g["out"])
    for k in y:
        assert_close(y[k], g["out"][k].to(DEV), 2.2e-2, "Controller[%d] vs reference golden" % k)     # measured 6e-3..1.1e-2


@pytest.mark.parametrize("B,H,W,C1,C2,G,silu", [(2, 16, 16, 320, 0, 32, True), (3, 8, 8, 1280, 1280, 32, True),
                                                 (1, 64, 64, 640, 320, 32, False), (8, 4, 4, 256, 0, 32, True),
                                                 (2, 5, 7, 64, 32, 8, False), (1, 3, 3, 64, 0, 32, True)])
def test_group_norm_cluster_kernel(B, H, W, C1, C2, G, silu):
    """ur_group_norm (cluster / DSMEM kernel) against torch GroupNorm on the same bf16 inputs, incl. the two-source
    concat, slabs that do not divide the pixel count, fewer pixels than CTAs, and against the two-launch path."""
    import torch.nn.functional as F
    from unirestore_b200 import ops
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(5)
    x1 = (torch.randn(B, H, W, C1, generator=g) * 2 + 0.5).to(dev).to(torch.bfloat16)
    x2 = (torch.randn(B, H, W, C2, generator=g) - 0.3).to(dev).to(torch.bfloat16) if C2 else None
    gamma = torch.randn(C1 + C2, generator=g).to(dev)
    beta = torch.randn(C1 + C2, generator=g).to(dev)
    old = ops.FUSED_GN_MAX_BYTES
    ops.FUSED_GN_MAX_BYTES = 1 << 40            # route through the cluster kernel
    try:
        y = ops.group_norm(x1, G, gamma, beta, 1e-5, silu=silu, x2=x2)
    finally:
        ops.FUSED_GN_MAX_BYTES = old
    xc = torch.cat([x1, x2], -1) if C2 else x1
    ref = F.group_norm(xc.float().permute(0, 3, 1, 2), G, gamma, beta, 1e-5)
    ref = (F.silu(ref) if silu else ref).permute(0, 2, 3, 1)
    assert_close(y, bf16_round(ref), 2e-3, "cluster group_norm")
    st = ops.chan_stats(x1, total_channels=C1 + C2)
    if C2:
        ops.chan_stats(x2, stats=st, offset=C1, total_channels=C1 + C2, zero=False)
    y2 = ops.norm_apply(x1, st, G, gamma, beta, 1e-5, silu, x2)
    assert_close(y, y2, 5e-3, "cluster vs two-launch group_norm")


@pytest.mark.parametrize("B,H,W,size,pb,pr", [(2, 256, 256, (512, 512), 0, 0), (1, 300, 200, (768, 512), 0, 0),
                                              (1, 512, 520, None, 0, 56), (1, 200, 333, (512, 853), 0, 43),
                                              (2, 576, 640, (300, 333), 0, 0), (1, 515, 700, None, 61, 4)])
def test_resize_pad_matches_torch(B, H, W, size, pb, pr):
    """ur_resize_pad == F.pad(F.interpolate(bicubic, align_corners=False), reflect) (unifie.py:124-134,165-168)."""
    import torch.nn.functional as F
    from unirestore_b200 import ops
    dev = torch.device("cuda:0")
    x = torch.rand(B, 3, H, W, generator=torch.Generator().manual_seed(3)).to(dev)
    ref = F.interpolate(x, size, mode="bicubic", align_corners=False, antialias=False) if size else x
    if pb or pr:
        ref = F.pad(ref, (0, pr, 0, pb), mode="reflect")
    got = ops.resize_pad(x, size, pb, pr)
    assert got.shape == ref.shape
    assert (got - ref).abs().max().item() < 2e-5
    xt = x.permute(0, 1, 3, 2).contiguous().permute(0, 1, 3, 2)          # non-contiguous (channels / strides) input
    assert (ops.resize_pad(xt, size, pb, pr) - ref).abs().max().item() < 2e-5


@pytest.mark.parametrize("B,H,W,quant", [(2, 64, 80, False), (1, 200, 333, True), (3, 7, 9, False)])
def test_image_metrics_match_oracle(B, H, W, quant):
    """ur_image_metrics (PSNR / SSIM sums) vs the numpy/scipy restatement of the skimage calls of the reference's
    validate loop (eval_image_restoration.py:71,255-313)."""
    import numpy as np
    from oracle import metrics as OM
    from unirestore_b200.metrics import SKPSNR, SKSSIM
    g = torch.Generator().manual_seed(11)
    tgt = torch.rand(B, 3, H, W, generator=g)
    prd = (tgt + 0.05 * torch.randn(B, 3, H, W, generator=g)).clamp(0, 1)
    ps, ss = SKPSNR(quantize=quant), SKSSIM(quantize=quant)
    ps.update(prd.to("cuda:0"), tgt.to("cuda:0"))
    ss.update(prd.to("cuda:0"), tgt.to("cuda:0"))
    pn = OM.quantize8(prd.numpy()) if quant else prd.numpy()
    ref_p = np.mean([OM.psnr(tgt[i].numpy(), pn[i]) for i in range(B)])
    ref_s = np.mean([OM.ssim(pn[i], tgt[i].numpy()) for i in range(B)])
    assert abs(ps.compute() - ref_p) < 1e-6 * abs(ref_p), (ps.compute(), ref_p)
    assert abs(ss.compute() - ref_s) < 1e-8, (ss.compute(), ref_s)
