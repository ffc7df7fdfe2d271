"""ur_conv_gemm (tcgen05 implicit GEMM) against a plain fp32 PyTorch conv/linear on the same
bf16-rounded operands.  Tolerance: rel-L2 <= 1e-3 after rounding the reference to bf16 too
(identical rounding points: bf16 operands, fp32 accumulate, one bf16 rounding of the output)."""
import pytest
import torch
import torch.nn.functional as F

from tests.util import assert_close, bf16_round

pytestmark = pytest.mark.gpu
TOL = 1e-3


def _dev():
    return torch.device("cuda:0")


def _rand(*shape, seed=0, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(*shape, generator=g) * scale).to(_dev())


def _nhwc(x_nchw):
    return x_nchw.permute(0, 2, 3, 1).contiguous().to(torch.bfloat16)


def _ref_out(y_nchw):
    return bf16_round(y_nchw.permute(0, 2, 3, 1))


@pytest.mark.parametrize("M,K,N", [(256, 128, 128), (128, 64, 64), (300, 320, 320), (77, 1024, 640),
                                   (4096, 320, 960), (130, 512, 256), (512, 256, 1024), (64, 1280, 1280),
                                   (200, 192, 72)])
def test_linear(M, K, N):
    from unirestore_b200 import ops
    x, w, b = _rand(M, K, seed=1), _rand(N, K, seed=2, scale=K ** -0.5), _rand(N, seed=3)
    xb, wb = x.to(torch.bfloat16), w.to(torch.bfloat16)
    y = ops.conv_gemm(xb, wb, N, bias=b)
    ref = bf16_round(F.linear(xb.float(), wb.float(), b))
    assert_close(y, ref, TOL, "linear %dx%dx%d" % (M, K, N))


def test_linear_residual_fp32_out_alpha():
    from unirestore_b200 import ops
    M, K, N = 384, 256, 320
    x, w, r = _rand(M, K, seed=4), _rand(N, K, seed=5, scale=K ** -0.5), _rand(M, N, seed=6)
    xb, wb, rb = x.to(torch.bfloat16), w.to(torch.bfloat16), r.to(torch.bfloat16)
    y = ops.conv_gemm(xb, wb, N, residual=rb, alpha=0.125, out_dtype=torch.float32)
    ref = 0.125 * F.linear(xb.float(), wb.float()) + rb.float()
    assert y.dtype == torch.float32
    assert_close(y, ref, 1e-5, "linear alpha+residual fp32")


@pytest.mark.parametrize("act", ["silu", "gelu"])
def test_linear_act_chscale(act):
    from unirestore_b200 import ops
    M, K, N = 2 * 96, 128, 256
    x, w, b = _rand(2, 96, K, seed=7), _rand(N, K, seed=8, scale=K ** -0.5), _rand(N, seed=9)
    cs, rv = _rand(2, N, seed=10), _rand(2, N, seed=11)
    xb, wb = x.to(torch.bfloat16), w.to(torch.bfloat16)
    y = ops.conv_gemm(xb, wb, N, bias=b, rowvec=rv, chscale=cs,
                      act=ops.UR_ACT_SILU if act == "silu" else ops.UR_ACT_GELU)
    pre = F.linear(xb.float(), wb.float(), b) + rv[:, None, :]
    ref = bf16_round((F.silu(pre) if act == "silu" else F.gelu(pre)) * cs[:, None, :])
    assert_close(y, ref, TOL, "linear act=%s chscale" % act)


@pytest.mark.parametrize("B,H,W,Cin,Cout", [(2, 16, 16, 64, 64), (1, 24, 40, 128, 320), (3, 8, 8, 320, 640),
                                            (2, 2, 2, 1280, 1280), (1, 64, 64, 320, 320), (5, 4, 4, 256, 512),
                                            (1, 20, 136, 64, 128)])
def test_conv3x3(B, H, W, Cin, Cout):
    from unirestore_b200 import ops
    x = _rand(B, Cin, H, W, seed=20)
    w = _rand(Cout, Cin, 3, 3, seed=21, scale=(9 * Cin) ** -0.5)
    b = _rand(Cout, seed=22)
    xb, wb = _nhwc(x), w.to(torch.bfloat16)
    y = ops.conv_gemm(xb, ops.pack_conv_weight(wb), Cout, taps=ops.TAPS_3x3, bias=b)
    ref = _ref_out(F.conv2d(xb.float().permute(0, 3, 1, 2), wb.float(), b, padding=1))
    assert_close(y, ref, TOL, "conv3x3 %s" % ((B, H, W, Cin, Cout),))


@pytest.mark.parametrize("asym", [False, True])
@pytest.mark.parametrize("B,H,W,C", [(2, 16, 16, 64), (1, 24, 40, 128), (2, 64, 64, 320), (1, 6, 10, 256)])
def test_conv3x3_stride2(B, H, W, C, asym):
    from unirestore_b200 import ops
    x = _rand(B, C, H, W, seed=30)
    w = _rand(C, C, 3, 3, seed=31, scale=(9 * C) ** -0.5)
    b = _rand(C, seed=32)
    xb, wb = _nhwc(x), w.to(torch.bfloat16)
    xin = xb.float().permute(0, 3, 1, 2)
    if asym:   # diffusers Downsample2D(padding=0): F.pad (0,1,0,1) then stride 2 (VAE encoder, autoencoder.py:19)
        ref = F.conv2d(F.pad(xin, (0, 1, 0, 1)), wb.float(), b, stride=2)
        taps = ops.TAPS_3x3_NOPAD
    else:
        ref = F.conv2d(xin, wb.float(), b, stride=2, padding=1)
        taps = ops.TAPS_3x3
    y = ops.conv_gemm(xb, ops.pack_conv_weight(wb), C, taps=taps, stride=2, hout=ref.shape[2], wout=ref.shape[3], bias=b)
    assert_close(y, _ref_out(ref), TOL, "conv3x3 s2 asym=%s %s" % (asym, (B, H, W, C)))


def test_two_source_concat_and_slices():
    """torch.cat([x, skip], 1) -> 1x1 conv (+residual), reading channel slices and writing a channel slice."""
    from unirestore_b200 import ops
    B, H, W, C1, C2, N = 2, 8, 12, 128, 64, 160
    big1 = _rand(B, H, W, C1 + 64, seed=40).to(torch.bfloat16)
    big2 = _rand(B, H, W, C2 + 8, seed=41).to(torch.bfloat16)
    x1, x2 = big1[..., 64:], big2[..., :C2]
    w = _rand(N, C1 + C2, seed=42, scale=(C1 + C2) ** -0.5).to(torch.bfloat16)
    res = _rand(B, H, W, N, seed=43).to(torch.bfloat16)
    outbig = torch.zeros(B, H, W, N + 32, device=_dev(), dtype=torch.bfloat16)
    ops.conv_gemm(x1, w, N, x2=x2, residual=res, out=outbig[..., 32:])
    ref = bf16_round(F.linear(torch.cat([x1, x2], -1).float(), w.float()) + res.float())
    assert_close(outbig[..., 32:], ref, TOL, "two-source 1x1")
    assert outbig[..., :32].abs().max().item() == 0.0


@pytest.mark.parametrize("act", ["geglu", "gate"])
@pytest.mark.parametrize("K,N2", [(320, 2560), (128, 256), (64, 128)])
def test_gated(act, K, N2):
    from unirestore_b200 import ops
    M = 200
    x, w, b = _rand(M, K, seed=50), _rand(N2, K, seed=51, scale=K ** -0.5), _rand(N2, seed=52)
    xb, wb = x.to(torch.bfloat16), w.to(torch.bfloat16)
    bn = ops.pick_bn(N2, True)
    wp, bp = ops.pack_gated_weight(wb, b, bn)
    y = ops.conv_gemm(xb, wp, N2, bias=bp, act=ops.UR_ACT_GEGLU if act == "geglu" else ops.UR_ACT_GATE, bn=bn)
    a, g = F.linear(xb.float(), wb.float(), b).chunk(2, -1)
    ref = bf16_round(a * F.gelu(g) if act == "geglu" else a * g)
    assert_close(y, ref, TOL, "gated %s" % act)


@pytest.mark.parametrize("C,G", [(256, 4), (512, 4), (128, 2)])
def test_grouped_conv3x3(C, G):
    from unirestore_b200 import ops
    B, H, W = 2, 12, 16
    x = _rand(B, C, H, W, seed=60)
    w = _rand(C, C // G, 3, 3, seed=61, scale=(9 * C // G) ** -0.5)
    b = _rand(C, seed=62)
    xb, wb = _nhwc(x), w.to(torch.bfloat16)
    y = ops.conv_gemm(xb, ops.pack_conv_weight(wb), C, taps=ops.TAPS_3x3, bias=b, group_kc=C // G, group_nc=C // G,
                      bn=min(128, C // G))
    ref = _ref_out(F.conv2d(xb.float().permute(0, 3, 1, 2), wb.float(), b, padding=1, groups=G))
    assert_close(y, ref, TOL, "grouped conv C=%d G=%d" % (C, G))


def test_batched_gemm():
    from unirestore_b200 import ops
    Bt, M, K, N = 3, 200, 512, 256
    a = _rand(Bt, M, K, seed=70).to(torch.bfloat16)
    w = _rand(Bt, N, K, seed=71, scale=K ** -0.5).to(torch.bfloat16)
    y = ops.conv_gemm(a, w, N, w_batched=True, alpha=0.5, out_dtype=torch.float32)
    ref = 0.5 * torch.einsum("bmk,bnk->bmn", a.float(), w.float())
    assert_close(y, ref, 1e-5, "batched gemm")


def test_subpixel_phase_output_view():
    """Strided output view (full[:, py::2, px::2]) used by the nearest-x2 upsample convolution."""
    from unirestore_b200 import ops
    B, H, W, C = 1, 8, 8, 64
    xb = _rand(B, H, W, C, seed=80).to(torch.bfloat16)
    w = _rand(C, C, seed=81, scale=C ** -0.5).to(torch.bfloat16)
    full = torch.zeros(B, 2 * H, 2 * W, C, device=_dev(), dtype=torch.bfloat16)
    ops.conv_gemm(xb, w, C, out=full[:, 1::2, 0::2])
    ref = bf16_round(F.linear(xb.float(), w.float()))
    assert_close(full[:, 1::2, 0::2], ref, TOL, "phase view")
    assert full[:, 0::2].abs().max().item() == 0.0


@pytest.fixture
def pair_mode():
    """Force the persistent kernel's CTA-pair (tcgen05 cta_group::2) mode on / off; restores auto selection."""
    from unirestore_b200 import _cabi

    def setter(mode):
        _cabi.ensure_init(0)
        _cabi.lib().ur_debug_set_gemm_pair_mode(mode)
    yield setter
    _cabi.lib().ur_debug_set_gemm_pair_mode(-1)


@pytest.mark.parametrize("mode", [0, 1])
@pytest.mark.parametrize("bn", [0, 64, 128, 160, 256])
def test_conv3x3_pair_modes(pair_mode, mode, bn):
    """Same conv through single-CTA and CTA-pair tiles, every N-tile width: odd k-block counts (45), an odd number of
    M tiles (ghost pair tile), ragged N tiles, bias + residual epilogue."""
    from unirestore_b200 import ops
    pair_mode(mode)
    for (B, H, W, Cin, Cout) in [(3, 16, 8, 320, 320), (1, 32, 32, 128, 640), (5, 8, 8, 64, 200)]:
        x = _rand(B, Cin, H, W, seed=90)
        w = _rand(Cout, Cin, 3, 3, seed=91, scale=(9 * Cin) ** -0.5)
        b = _rand(Cout, seed=92)
        r = _rand(B, H, W, Cout, seed=93).to(torch.bfloat16)
        xb, wb = _nhwc(x), w.to(torch.bfloat16)
        y = ops.conv_gemm(xb, ops.pack_conv_weight(wb), Cout, taps=ops.TAPS_3x3, bias=b, residual=r, bn=bn)
        ref = bf16_round(F.conv2d(xb.float().permute(0, 3, 1, 2), wb.float(), b, padding=1).permute(0, 2, 3, 1)
                         + r.float())
        assert_close(y, ref, TOL, "conv3x3 pair=%d bn=%d %s" % (mode, bn, (B, H, W, Cin, Cout)))


@pytest.mark.parametrize("mode", [0, 1])
def test_gated_and_linear_pair_modes(pair_mode, mode):
    from unirestore_b200 import ops
    pair_mode(mode)
    M, K, N2 = 1000, 1280, 2560
    x, w, b = _rand(M, K, seed=94), _rand(N2, K, seed=95, scale=K ** -0.5), _rand(N2, seed=96)
    xb, wb = x.to(torch.bfloat16), w.to(torch.bfloat16)
    bn = ops.pick_bn(N2, True)
    wp, bp = ops.pack_gated_weight(wb, b, bn)
    y = ops.conv_gemm(xb, wp, N2, bias=bp, act=ops.UR_ACT_GEGLU, bn=bn)
    a, g = F.linear(xb.float(), wb.float(), b).chunk(2, -1)
    assert_close(y, bf16_round(a * F.gelu(g)), TOL, "geglu pair=%d" % mode)
    for (M, K, N) in [(300, 320, 320), (4096, 64, 960), (129, 1152, 72)]:
        x, w, b = _rand(M, K, seed=97), _rand(N, K, seed=98, scale=K ** -0.5), _rand(N, seed=99)
        xb, wb = x.to(torch.bfloat16), w.to(torch.bfloat16)
        y = ops.conv_gemm(xb, wb, N, bias=b)
        assert_close(y, bf16_round(F.linear(xb.float(), wb.float(), b)), TOL, "linear pair=%d %s" % (mode, (M, K, N)))


@pytest.mark.parametrize("B,H,W,Cin,Cout", [(8, 8, 8, 1280, 1280), (8, 8, 8, 2560, 1280), (3, 8, 8, 512, 512),
                                            (1, 16, 16, 1280, 320)])
def test_conv3x3_split_k(B, H, W, Cin, Cout):
    """Few output tiles + long K: the K range is split over CTAs (one fp32 slab per K slice in the caller's workspace,
    plain stores; the finishing kernel adds the slabs in a fixed order with bias / temb row vector / residual);
    compared with torch and with the unsplit kernel, and run twice: bit-identical (no atomics on the data path)."""
    from unirestore_b200 import _cabi, ops
    x = _rand(B, Cin, H, W, seed=100)
    w = _rand(Cout, Cin, 3, 3, seed=101, scale=(9 * Cin) ** -0.5)
    b, tv = _rand(Cout, seed=102), _rand(1, Cout, seed=103)
    r = _rand(B, H, W, Cout, seed=104).to(torch.bfloat16)
    xb, wb = _nhwc(x), w.to(torch.bfloat16)
    wp = ops.pack_conv_weight(wb)
    y = ops.conv_gemm(xb, wp, Cout, taps=ops.TAPS_3x3, bias=b, rowvec=tv, residual=r)
    ref = bf16_round((F.conv2d(xb.float().permute(0, 3, 1, 2), wb.float(), b, padding=1) + tv[0][None, :, None, None])
                     .permute(0, 2, 3, 1) + r.float())
    assert_close(y, ref, TOL, "conv3x3 split-K %s" % ((B, H, W, Cin, Cout),))
    old = _cabi.lib().ur_debug_set_gemm_splitk(0)
    try:
        y0 = ops.conv_gemm(xb, wp, Cout, taps=ops.TAPS_3x3, bias=b, rowvec=tv, residual=r)
    finally:
        _cabi.lib().ur_debug_set_gemm_splitk(old)
    assert_close(y, y0, 2e-3, "split-K vs unsplit")
    y2 = ops.conv_gemm(xb, wp, Cout, taps=ops.TAPS_3x3, bias=b, rowvec=tv, residual=r)
    assert torch.equal(y, y2), "split-K is not bit-reproducible"


@pytest.mark.parametrize("B,H,W,Cin,Cout,taps", [(2, 16, 16, 64, 320, 9), (1, 24, 40, 128, 320, 9), (3, 8, 8, 320, 640, 9),
                                                 (8, 8, 8, 1280, 1280, 9), (2, 64, 64, 320, 320, 1),
                                                 (5, 4, 4, 256, 200, 9), (1, 20, 136, 64, 128, 9),
                                                 (8, 8, 8, 1280, 1280, 1), (7, 8, 8, 640, 320, 1), (6, 4, 8, 128, 256, 1),
                                                 (3, 8, 4, 192, 320, 9)])
def test_epilogue_group_norm_statistics(B, H, W, Cin, Cout, taps):
    """want_stats: the (sum, sumsq) per (image, channel) accumulated by the GEMM epilogue (tiles holding 1, 2 or 4
    whole images, odd batch counts), by the split-K finishing kernel, or by the fall-back pass (more than 4 images per
    tile) equal a ur_chan_stats pass over the finished output."""
    from unirestore_b200 import ops
    x = _rand(B, Cin, H, W, seed=110)
    k = 3 if taps == 9 else 1
    w = _rand(Cout, Cin, k, k, seed=111, scale=(taps * Cin) ** -0.5)
    b = _rand(Cout, seed=112)
    r = _rand(B, H, W, Cout, seed=113).to(torch.bfloat16)
    xb, wb = _nhwc(x), w.to(torch.bfloat16)
    y = ops.conv_gemm(xb, ops.pack_conv_weight(wb), Cout, taps=ops.TAPS_3x3 if taps == 9 else ops.TAPS_1x1, bias=b,
                      residual=r, want_stats=True)
    got = y._ur_stats
    ref = ops.chan_stats(y)
    torch.cuda.synchronize()
    assert got.shape == ref.shape == (B, Cout, 2)
    scale = ref.abs().amax(dim=(0, 1), keepdim=True).clamp_min(1e-6)
    err = ((got - ref).abs() / scale).max().item()
    assert err < 1e-4, "epilogue statistics differ from ur_chan_stats: %.3e" % err


def test_group_norm_uses_epilogue_statistics():
    """group_norm on a tensor that carries epilogue statistics (one or both concat sources) == the two-pass result."""
    from unirestore_b200 import ops
    B, H, W, C = 2, 16, 16, 320
    xb = _nhwc(_rand(B, 64, H, W, seed=120))
    w = _rand(C, 64, 3, 3, seed=121, scale=(9 * 64) ** -0.5).to(torch.bfloat16)
    g, be = _rand(2 * C, seed=122), _rand(2 * C, seed=123)
    y1 = ops.conv_gemm(xb, ops.pack_conv_weight(w), C, taps=ops.TAPS_3x3, want_stats=True)
    y2 = ops.conv_gemm(xb, ops.pack_conv_weight(w), C, taps=ops.TAPS_3x3)          # no statistics attached
    a = ops.group_norm(y1, 32, g[:C].contiguous(), be[:C].contiguous(), 1e-5, silu=True)
    bq = ops.group_norm(y2, 32, g[:C].contiguous(), be[:C].contiguous(), 1e-5, silu=True)
    assert_close(a, bq, 2e-3, "group_norm with epilogue statistics")
    c1 = ops.group_norm(y1, 32, g, be, 1e-5, x2=y2)
    c2 = ops.group_norm(y2, 32, g, be, 1e-5, x2=y2)
    assert_close(c1, c2, 2e-3, "concat group_norm with mixed statistics sources")


@pytest.mark.parametrize("B,H,W,Cin,Cout,taps", [(2, 64, 64, 320, 320, 1), (3, 16, 16, 640, 640, 9), (8, 8, 8, 1280, 640, 1),
                                                 (2, 20, 24, 128, 200, 9), (1, 1, 300, 320, 960, 1), (5, 4, 8, 128, 256, 1)])
def test_epilogue_store_modes_agree(B, H, W, Cin, Cout, taps):
    """The three store paths of the persistent kernel's epilogue -- st.global (0), one TMA store per 128-row sub-block
    after a group barrier (1), one TMA store per epilogue warp with 32-row boxes and no barrier (2, default) -- produce
    bit-identical outputs and statistics, with bias + residual + fused statistics, ragged tiles and 1/2/4 images per
    tile."""
    from unirestore_b200 import _cabi, ops
    x = _rand(B, Cin, H, W, seed=130)
    k = 3 if taps == 9 else 1
    w = _rand(Cout, Cin, k, k, seed=131, scale=(taps * Cin) ** -0.5)
    b = _rand(Cout, seed=132)
    r = _rand(B, H, W, Cout, seed=133).to(torch.bfloat16)
    xb, wp = _nhwc(x), ops.pack_conv_weight(w.to(torch.bfloat16))
    outs = {}
    for mode in (2, 1, 0):
        old = _cabi.lib().ur_debug_set_gemm_tma_store(mode)
        try:
            y = ops.conv_gemm(xb, wp, Cout, taps=ops.TAPS_3x3 if taps == 9 else ops.TAPS_1x1, bias=b, residual=r,
                              want_stats=True)
            torch.cuda.synchronize()
            outs[mode] = (y.clone(), y._ur_stats.clone())
        finally:
            _cabi.lib().ur_debug_set_gemm_tma_store(old)
    ref = bf16_round(F.conv2d(xb.float().permute(0, 3, 1, 2), w.to(torch.bfloat16).float(), b, padding=k // 2)
                     .permute(0, 2, 3, 1) + r.float())
    assert_close(outs[2][0], ref, TOL, "warp-store epilogue %s" % ((B, H, W, Cin, Cout, taps),))
    for mode in (1, 0):
        assert torch.equal(outs[mode][0], outs[2][0]), "store mode %d output differs from warp mode" % mode
        scale = outs[2][1].abs().amax().clamp_min(1e-6)
        assert ((outs[mode][1] - outs[2][1]).abs() / scale).max().item() < 1e-6
