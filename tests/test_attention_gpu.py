"""ur_attention (fused tcgen05 flash attention) and the unfused GEMM/softmax/GEMM path against fp32
F.scaled_dot_product_attention on the same bf16-rounded q/k/v.  Gate: rel-L2 <= 3e-3 after rounding the reference to
bf16 (P is rounded to bf16 before the second GEMM, as every flash kernel does)."""
import pytest
import torch
import torch.nn.functional as F

from tests.util import assert_close, bf16_round

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _rand(*shape, seed=0, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(*shape, generator=g) * scale).to(DEV).to(torch.bfloat16)


def _ref(q, k, v, heads):
    B, Tq, C = q.shape
    d = C // heads
    sp = lambda t: t.float().view(t.shape[0], -1, heads, d).transpose(1, 2)
    o = F.scaled_dot_product_attention(sp(q), sp(k).expand(B, -1, -1, -1), sp(v).expand(B, -1, -1, -1))
    return bf16_round(o.transpose(1, 2).reshape(B, Tq, C))


@pytest.fixture(params=[(1, 1), (2, 1), (1, 0)], ids=["attention1+kv1", "attention2", "attention1-general"])
def impl(request):
    """The kernels in the tree: attention_kernel (default) with the query-tile-loop kernel for launches whose keys fit
    one tile (attention_kv1_kernel, default), attention2_kernel, and attention_kernel alone."""
    from unirestore_b200 import _cabi
    gen, kv1 = request.param
    old = _cabi.lib().ur_debug_set_attention_impl(gen)
    old_kv1 = _cabi.lib().ur_debug_set_attention_kv1(kv1)
    yield gen
    _cabi.lib().ur_debug_set_attention_impl(old)
    _cabi.lib().ur_debug_set_attention_kv1(old_kv1)


@pytest.mark.parametrize("B,heads,d,Tq,Tk", [(2, 5, 64, 256, 256), (1, 2, 64, 128, 128), (2, 4, 64, 200, 300),
                                             (1, 5, 64, 4096, 4096), (2, 4, 128, 64, 64), (3, 4, 128, 256, 256),
                                             (2, 20, 64, 4, 4), (1, 10, 64, 1024, 1024), (2, 1, 64, 130, 7),
                                             (1, 3, 64, 384, 1000), (2, 2, 64, 129, 129), (1, 2, 128, 300, 200),
                                             (1, 1, 64, 256, 33), (8, 5, 64, 4096, 77), (3, 7, 64, 1000, 128),
                                             (2, 3, 64, 777, 1), (1, 40, 64, 128, 100)])
def test_fused_self_attention_packed_qkv(impl, B, heads, d, Tq, Tk):
    from unirestore_b200 import ops
    torch.backends.cuda.matmul.allow_tf32 = False
    C = heads * d
    if Tq == Tk:     # packed [B,T,3C] buffer, channel-slice views (the layout the QKV GEMM writes)
        qkv = _rand(B, Tq, 3 * C, seed=1, scale=1.5)
        q, k, v = qkv[..., :C], qkv[..., C:2 * C], qkv[..., 2 * C:]
    else:
        q, k, v = _rand(B, Tq, C, seed=2, scale=1.5), _rand(B, Tk, C, seed=3, scale=1.5), _rand(B, Tk, C, seed=4)
    y = ops.attention(q, k, v, heads)
    assert_close(y, _ref(q, k, v, heads), 3e-3, "flash attention %s" % ((B, heads, d, Tq, Tk),))


@pytest.mark.parametrize("B,heads,Tq", [(4, 5, 4096), (2, 10, 1024), (3, 20, 64)])
def test_fused_cross_attention_shared_kv(impl, B, heads, Tq):
    """K/V of the constant 77-token null prompt shared by all images (base_model.py:221)."""
    from unirestore_b200 import ops
    C = heads * 64
    q = _rand(B, Tq, C, seed=5, scale=1.5)
    kv = _rand(1, 77, 2 * C, seed=6, scale=1.5)
    k, v = kv[..., :C], kv[..., C:]
    y = ops.attention(q, k, v, heads)
    assert_close(y, _ref(q, k, v, heads), 3e-3, "flash cross attention %s" % ((B, heads, Tq),))


@pytest.mark.parametrize("B,heads,Tq,Tk", [(2, 1, 96, 96), (1, 1, 1024, 1024), (2, 1, 4096, 4096), (1, 2, 300, 200),
                                           (1, 1, 16384, 16384)])
def test_fused_attention_head_dim_512(B, heads, Tq, Tk):
    """attention512_kernel: the VAE mid-block head (autoencoder.py:32,44) at 512^2 (4 096 tokens) and 1024^2 (16 384
    tokens) latents, ragged token counts and two heads; no score tensor in HBM."""
    from unirestore_b200 import ops
    C = heads * 512
    qkv = _rand(B, max(Tq, Tk), 3 * C, seed=12)
    q, k, v = qkv[:, :Tq, :C], qkv[:, :Tk, C:2 * C], qkv[:, :Tk, 2 * C:]
    y = ops.attention(q, k, v, heads)
    if Tq * Tk > 1 << 26:                      # 16 384^2: check a slice of the queries against the fp32 reference
        idx = torch.arange(0, Tq, 61, device=DEV)
        assert_close(y[:, idx], _ref(q[:, idx], k, v, heads), 3e-3, "flash attention d=512 %s (query slice)" % ((B, heads, Tq, Tk),))
    else:
        assert_close(y, _ref(q, k, v, heads), 3e-3, "flash attention d=512 %s" % ((B, heads, Tq, Tk),))


@pytest.mark.parametrize("B,heads,d,Tq,Tk", [(2, 1, 512, 96, 96), (1, 1, 512, 1024, 1024), (2, 5, 64, 64, 77)])
def test_unfused_attention(B, heads, d, Tq, Tk):
    from unirestore_b200 import ops
    C = heads * d
    q, k, v = _rand(B, Tq, C, seed=7), _rand(B, Tk, C, seed=8), _rand(B, Tk, C, seed=9)
    y = ops.attention_unfused(q, k, v, heads)
    assert_close(y, _ref(q, k, v, heads), 3e-3, "unfused attention %s" % ((B, heads, d, Tq, Tk),))


def test_attention_large_score_range_lazy_rescale():
    """Scores whose running maximum grows by far more than 2^8 from KV tile to KV tile (the lazy-rescaling path of both
    kernels: O and l are rescaled in TMEM) and rows dominated by a single key."""
    from unirestore_b200 import _cabi, ops
    B, heads, d, T = 1, 2, 64, 1024
    g = torch.Generator().manual_seed(21)
    q = torch.randn(B, T, heads * d, generator=g) * 3.0
    k = torch.randn(B, T, heads * d, generator=g)
    k = k * torch.linspace(0.2, 6.0, T).view(1, T, 1)                 # later keys score much higher / lower
    v = torch.randn(B, T, heads * d, generator=g)
    q, k, v = (t.to(DEV).to(torch.bfloat16) for t in (q, k, v))
    ref = _ref(q, k, v, heads)
    for impl in (1, 2):
        old = _cabi.lib().ur_debug_set_attention_impl(impl)
        try:
            y = ops.attention(q, k, v, heads)
        finally:
            _cabi.lib().ur_debug_set_attention_impl(old)
        assert_close(y, ref, 4e-3, "flash attention impl %d, growing score range" % impl)
