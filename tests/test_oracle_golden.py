"""Pins the CPU oracle (oracle/) to the golden fixtures produced by EXECUTING THE REFERENCE'S OWN FILES
(oracle/make_golden.py, run once in the build container against /root/reference).  fp32 vs fp32: max-rel <= 1e-4."""
import os

import pytest
import torch

GOLD = os.path.join(os.path.dirname(__file__), "golden")
TOL = 1e-4


def rnd(seed, *shape, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(*shape, generator=g) * scale


def maxrel(a, b):
    return (a - b).abs().max().item() / max(b.abs().max().item(), 1e-12)


@pytest.fixture(scope="module", autouse=True)
def _threads():
    torch.set_num_threads(os.cpu_count())
    yield


def test_recorded_oracle_vs_reference_report():
    rep = torch.load(os.path.join(GOLD, "oracle_vs_reference_report.pt"))
    assert len(rep) >= 20 and max(rep.values()) < TOL, rep


@torch.no_grad()
def test_scedit():
    from oracle import unirestore as O
    from unirestore_b200.init_utils import deterministic_init_
    g = torch.load(os.path.join(GOLD, "scedit.pt"))
    m = deterministic_init_(O.CSCEAdapter(*g["args"])).eval()
    y = m(rnd(g["x_seed"], *g["x_shape"]), rnd(g["c_seed"], *g["c_shape"]))
    assert maxrel(y, g["out"]) < TOL


@torch.no_grad()
def test_taskeditor():
    from oracle import unirestore as O
    from unirestore_b200.init_utils import deterministic_init_
    for case in torch.load(os.path.join(GOLD, "taskeditor.pt")):
        co, cs, pl, last = case["args"]
        m = deterministic_init_(O.TaskFeatureAdapter(co, cs, pl, last)).eval()
        s1, s2, s3 = case["seeds"]
        h, w = case["hw"]
        yx, yc = m(rnd(s1, 2, co, h, w), rnd(s2, 2, cs, h, w, scale=2.0), rnd(s3, 2, pl, cs))
        assert maxrel(yx, case["out_x"]) < TOL
        if case["out_cond"] is not None:
            assert maxrel(yc, case["out_cond"]) < TOL
        else:
            assert yc is None


@torch.no_grad()
def test_cfrm():
    from oracle import unirestore as O
    from unirestore_b200.init_utils import deterministic_init_
    g = torch.load(os.path.join(GOLD, "cfrm.pt"))
    x = rnd(g["x_seed"], *g["x_shape"], scale=g["x_scale"])
    assert maxrel(deterministic_init_(O.NAFBlock(g["c"])).eval()(x), g["naf"]) < TOL
    assert maxrel(deterministic_init_(O.AdaNAFV2(g["c"])).eval()(x), g["ada"]) < TOL


@torch.no_grad()
def test_controller():
    from oracle import unirestore as O
    from unirestore_b200.init_utils import deterministic_init_
    g = torch.load(os.path.join(GOLD, "controller.pt"))
    m = deterministic_init_(O.Controller(), "controller.").eval()
    y = m(rnd(g["x_seed"], *g["x_shape"]), torch.tensor([g["t"]]))
    assert set(y) == set(g["out"]) == {16, 8, 4, 2}
    for k in y:
        assert maxrel(y[k], g["out"][k]) < TOL, k


@pytest.fixture(scope="module")
def full_oracle():
    from oracle import unirestore as O
    from unirestore_b200.init_utils import deterministic_init_
    m = O.DiffUIE(dict(type="CFRM"), dict(type="scedit", num_inference_steps=2),
                  dict(type="TFA", prompt_len=1, task=["ir", "cls", "seg"])).eval().requires_grad_(False)
    return deterministic_init_(m)


@torch.no_grad()
def test_controlled_unet(full_oracle):
    g = torch.load(os.path.join(GOLD, "base_model.pt"))
    control = torch.load(os.path.join(GOLD, "controller.pt"))["out"]
    y = full_oracle.base_model(rnd(g["zt_seed"], *g["zt_shape"]), control, torch.tensor([g["t"]]))
    assert maxrel(y, g["out"]) < TOL


@torch.no_grad()
def test_autoencoder(full_oracle):
    g = torch.load(os.path.join(GOLD, "autoencoder.pt"))
    img = torch.rand(*g["img_shape"], generator=torch.Generator().manual_seed(g["img_seed"]))
    torch.manual_seed(g["rng_seed"])
    noise = torch.randn(1, 4, g["img_shape"][2] // 8, g["img_shape"][3] // 8)
    z, skips = full_oracle.ae.encode(img, enable_fr=True, noise=noise)
    assert maxrel(z, g["z"]) < TOL
    for s, ref in zip(skips, g["skips"]):
        assert tuple(s.shape) == ref["shape"] and maxrel(s[..., ::4, ::4], ref["sample"]) < TOL
    for task in ("ir", "seg"):
        assert maxrel(full_oracle.ae.decode(z, skips, task), g["decode"][task]) < TOL
    with pytest.raises((KeyError, AttributeError)):      # nn.ParameterDict lookup: AttributeError on torch >= 1.12
        full_oracle.ae.decode(z, skips, "no-such-task")


@torch.no_grad()
@pytest.mark.skipif(os.environ.get("UR_SLOW_TESTS") != "1", reason="~2 min of CPU: set UR_SLOW_TESTS=1")
def test_diffuie_forward(full_oracle):
    g = torch.load(os.path.join(GOLD, "diffuie.pt"))
    img = torch.rand(*g["img_shape"], generator=torch.Generator().manual_seed(g["img_seed"]))
    torch.manual_seed(g["rng_seed"])
    n_post, n_diff = torch.randn(1, 4, 64, 80), torch.randn(1, 4, 64, 80)
    assert full_oracle.scheduler.timesteps.tolist() == g["timesteps"].tolist()
    assert maxrel(full_oracle(img, g["task"], noise=(n_post, n_diff)), g["out"]) < TOL
