"""Host-side logic of the product path that needs no GPU: state_dict contract against the oracle / reference keys,
weight packing (gated interleave, sub-pixel upsample weights, block-diagonal grouped conv), config validation."""
import os

import pytest
import torch
import torch.nn.functional as F

CFG = (dict(type="CFRM"), dict(type="scedit", num_inference_steps=4), dict(type="TFA", prompt_len=1, task=["ir", "cls", "seg"]))


@pytest.fixture(scope="module")
def model():
    from unirestore_b200.diffuie import DiffUIE
    return DiffUIE(*CFG)


def test_state_dict_keys_and_shapes_match_oracle(model):
    from oracle import unirestore as O
    o = O.DiffUIE(*CFG)
    so, sm = o.state_dict(), model.state_dict()
    assert set(so) == set(sm)
    assert all(so[k].shape == sm[k].shape for k in so)
    # checkpoint-surgery prefixes of engine_unifie.py:49-133
    for prefix in ("ae.vae.encoder.fr_blocks.", "controller.", "base_model.csc_editors.", "ae.vae.decoder.task_prompts.",
                   "ae.vae.decoder.task_editors."):
        assert any(k.startswith(prefix) for k in sm), prefix
    assert sum(p.numel() for p in model.base_model.csc_editors.parameters()) == 22147200
    assert [sum(p.numel() for p in t.parameters()) for t in model.ae.vae.decoder.task_editors] == [15602944, 4164480, 1263232]


def test_scheduler_and_buffers(model):
    assert model.scheduler.timesteps.tolist() == [999, 749, 499, 249]
    assert model.train_timesteps.tolist() == [249, 499, 749, 999, 999, 999]
    assert tuple(model.base_model.null_embeds.shape) == (1, 77, 1024)


def test_constructor_errors_match_reference():
    from unirestore_b200.diffuie import ControlledUNet, SkipConnectedAutoEncoder
    from unirestore_b200.diffuie.sd_blocks import AutoencoderKL, UNet2DConditionModel
    small = dict(block_out_channels=(32, 32, 64, 64))
    with pytest.raises(ValueError):
        SkipConnectedAutoEncoder(AutoencoderKL(**small), fr_type="bogus")
    with pytest.raises(KeyError):
        SkipConnectedAutoEncoder(AutoencoderKL(**small), None, dict(type="bogus", task=["ir"], prompt_len=1))
    with pytest.raises(ValueError):
        ControlledUNet(UNet2DConditionModel(block_out_channels=(32, 32, 64, 64), num_attention_heads=(1, 1, 1, 1)),
                       "spade", null_embeds=torch.zeros(1, 77, 1024))


def test_gated_weight_interleave():
    from unirestore_b200.ops import pack_gated_weight
    n, k, bn = 512, 16, 128
    w, b = torch.randn(n, k), torch.randn(n)
    wp, bp = pack_gated_weight(w, b, bn)
    x = torch.randn(5, k)
    full = F.linear(x, wp, bp)
    a, g = F.linear(x, w, b).chunk(2, -1)
    for tile in range(n // bn):
        ta = full[:, tile * bn: tile * bn + bn // 2]
        tg = full[:, tile * bn + bn // 2: (tile + 1) * bn]
        sl = slice(tile * bn // 2, (tile + 1) * bn // 2)
        assert torch.allclose(ta, a[:, sl], atol=1e-6) and torch.allclose(tg, g[:, sl], atol=1e-6)


def test_subpixel_upsample_weights_reproduce_nearest_conv():
    """The four pre-summed 2x2 phase convolutions on the low-res input == conv3x3(nearest_x2(x)) exactly (fp32)."""
    from unirestore_b200.diffuie.sd_blocks import Upsample2D
    up = Upsample2D(8, True, 12)
    torch.nn.init.normal_(up.conv.weight)
    x = torch.randn(2, 8, 5, 7)
    ref = F.conv2d(F.interpolate(x, scale_factor=2.0, mode="nearest"), up.conv.weight, up.conv.bias, padding=1)
    w = up.conv.weight.detach()
    sel = {0: ((0,), (1, 2)), 1: ((0, 1), (2,))}
    out = torch.zeros_like(ref)
    xp = F.pad(x, (1, 1, 1, 1))
    for py in (0, 1):
        for px in (0, 1):
            acc = up.conv.bias.view(1, -1, 1, 1).expand(2, -1, 5, 7).clone()
            for ry in (0, 1):
                for rx in (0, 1):
                    m = sum(w[:, :, ky, kx] for ky in sel[py][ry] for kx in sel[px][rx])
                    dy, dx = py - 1 + ry, px - 1 + rx
                    acc += torch.einsum("oc,bchw->bohw", m, xp[:, :, 1 + dy:1 + dy + 5, 1 + dx:1 + dx + 7])
            out[:, :, py::2, px::2] = acc
    assert torch.allclose(out, ref, atol=1e-4)
    # and the packed phase table has the layout the kernel consumes: 4 taps x Cin per output channel
    pk = up._pack()
    assert set(pk["phases"]) == {(0, 0), (0, 1), (1, 0), (1, 1)}
    taps, wp = pk["phases"][(1, 0)]
    assert taps == ((0, -1), (0, 0), (1, -1), (1, 0)) and tuple(wp.shape) == (12, 32)


def test_block_diagonal_grouped_conv_packing():
    """AdaNAFV2 groups narrower than 64 channels are merged pairwise into block-diagonal 64-channel groups."""
    from unirestore_b200.diffuie import AdaNAFV2
    m = AdaNAFV2(32)                                   # wide = 128, 16 groups of 8 -> merged groups of 64
    torch.nn.init.normal_(m.group_conv.weight)
    pk = m._pack()
    assert pk["kg"] == 64 and tuple(pk["wg"].shape) == (128, 9 * 64)
    wd = pk["wg"].float().view(128, 3, 3, 64).permute(0, 3, 1, 2)         # [Cout, 64, 3, 3] per merged group
    x = torch.randn(1, 128, 6, 6).to(torch.bfloat16).float()
    ref = F.conv2d(x, m.group_conv.weight.to(torch.bfloat16).float(), None, padding=1, groups=16)
    got = F.conv2d(x, wd, None, padding=1, groups=2)
    assert torch.allclose(got, ref, atol=1e-4)


def test_val_yaml_model_kwargs_drive_the_constructor():
    """configs/val.yaml -> model_kwargs -> DiffUIE(**model_kwargs) as engine_unifie.py:38-42 does (the YAML text is
    restated here because /root/reference does not travel)."""
    import yaml
    text = """
    model_kwargs:
      frenc: {type: CFRM, ckpt: $path_to_stage1_ckpt$}
      cnet: {type: scedit, num_inference_steps: 1, ckpt: $path_to_stage1_ckpt$}
      tedit: {type: TFA, prompt_len: 1, task: [ir, cls, seg], ckpt: $path_to_stage2_ckpt$}
    """
    kw = yaml.safe_load(text)["model_kwargs"]
    from unirestore_b200.diffuie import DiffUIE
    m = DiffUIE(**kw)
    assert m.scheduler.timesteps.tolist() == [999] and sorted(m.ae.vae.decoder.task_prompts.keys()) == ["cls", "ir", "seg"]


def test_reference_checkpoint_surgery(model):
    """engine_unifie.py:49-126: CFRM / Controller + SC-Tuner / TFA parameters restored by key prefix from Lightning
    checkpoints; SD-Turbo backbone from diffusers-format state dicts."""
    import torch
    from unirestore_b200.checkpoint import load_reference_checkpoints, load_sd_turbo
    full = {"model." + k: torch.full_like(v, 0.5) if v.is_floating_point() else v.clone()
            for k, v in model.state_dict().items()}
    ckpt = {"state_dict": full}
    rep = load_reference_checkpoints(model, frenc=ckpt, cnet=ckpt, tedit=ckpt)
    assert rep["cnet"]["model.controller."] > 100 and rep["cnet"]["model.base_model.csc_editors."] == 12 * 6
    assert rep["frenc"]["model.ae.vae.encoder.fr_blocks."] > 100
    assert float(model.controller.conv_in.weight.mean()) == 0.5
    assert float(model.base_model.csc_editors[3].proj.bias.mean()) == 0.5
    assert float(model.ae.vae.decoder.task_prompts["seg"].mean()) == 0.5
    assert float(model.base_model.unet.conv_in.weight.mean()) != 0.5          # backbone untouched by the surgery
    broken = {"state_dict": {k: v for k, v in full.items() if "controller.conv_in.bias" not in k}}
    with pytest.raises(RuntimeError):
        load_reference_checkpoints(model, cnet=broken)                         # strict, like the reference
    unet_sd = {k: torch.zeros_like(v) for k, v in model.base_model.unet.state_dict().items()}
    vae_sd = {k: torch.zeros_like(v) for k, v in model.ae.vae.state_dict().items()
              if not k.startswith(("encoder.fr_blocks.", "decoder.task_prompts.", "decoder.task_editors."))}
    load_sd_turbo(model, unet_sd, vae_sd)
    assert float(model.base_model.unet.conv_in.weight.abs().max()) == 0.0
    assert float(model.ae.vae.encoder.fr_blocks[0][0].conv1.weight.mean()) == 0.5   # CFRM kept
    with pytest.raises(RuntimeError):
        load_sd_turbo(model, None, {k: v for k, v in vae_sd.items() if k != "quant_conv.weight"})


def test_bench_stdout_carries_only_the_json_line(tmp_path):
    """bench.claim_stdout(): library writes to fd 1 (NCCL's version banner) end up on stderr, the JSON line on stdout."""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = ("import os, sys, json; sys.path.insert(0, %r); import bench; emit = bench.claim_stdout(); "
            "os.write(1, b'NCCL version 2.28.9\\n'); print('stray print'); emit(json.dumps({'ok': 1}))" % root)
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    assert r.stdout == '{"ok": 1}\n'
    assert "NCCL version" in r.stderr and "stray print" in r.stderr
