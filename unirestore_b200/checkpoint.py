"""Checkpoint ingest for the drop-in ``DiffUIE``.

The reference restores three groups of trained parameters by key prefix from Lightning checkpoints
(``torch.load(path)["state_dict"]``, src/core/engine_unifie.py:49-126):

    frenc["ckpt_path"] -> model.ae.vae.encoder.fr_blocks.*                       (CFRM)
    cnet["ckpt_path"]  -> model.controller.* , model.base_model.csc_editors.*    (Controller + SC-Tuner)
    tedit["ckpt_path"] -> model.ae.vae.decoder.task_prompts.* (strict=False) , model.ae.vae.decoder.task_editors.*

and takes the frozen SD-Turbo UNet / VAE from the diffusers hub.  Attribute paths and parameter names of
``unirestore_b200.diffuie`` are identical, so the same surgery applies; every ``load_state_dict`` drops the packed
bf16 weight caches of the touched modules (``UrModule`` post-hook), the next forward re-packs them.
"""
from __future__ import annotations

import torch

_PREFIXES = {
    "frenc": (("model.ae.vae.encoder.fr_blocks.", lambda m: m.ae.vae.encoder.fr_blocks, True),),
    "cnet": (("model.controller.", lambda m: m.controller, True),
             ("model.base_model.csc_editors.", lambda m: m.base_model.csc_editors, True)),
    "tedit": (("model.ae.vae.decoder.task_prompts.", lambda m: m.ae.vae.decoder.task_prompts, False),
              ("model.ae.vae.decoder.task_editors.", lambda m: m.ae.vae.decoder.task_editors, True)),
}


def _state_dict(ckpt):
    if isinstance(ckpt, (str, bytes)) or hasattr(ckpt, "__fspath__"):
        ckpt = torch.load(ckpt, map_location="cpu")
    return ckpt["state_dict"] if "state_dict" in ckpt else ckpt


def load_reference_checkpoints(model, frenc=None, cnet=None, tedit=None):
    """Apply the reference's prefix surgery; each argument is a checkpoint path, a loaded checkpoint dict or None.
    Returns ``{group: {prefix: number of tensors loaded}}``.  Missing / unexpected keys raise exactly where the
    reference's ``load_state_dict`` calls would (strict, except the task prompts)."""
    report = {}
    for group, ckpt in (("frenc", frenc), ("cnet", cnet), ("tedit", tedit)):
        if ckpt is None:
            continue
        sd = _state_dict(ckpt)
        report[group] = {}
        for prefix, target, strict in _PREFIXES[group]:
            sub = {k[len(prefix):]: v for k, v in sd.items() if k.startswith(prefix)}
            target(model).load_state_dict(sub, strict=strict)
            report[group][prefix] = len(sub)
    if hasattr(model, "clear_caches"):
        model.clear_caches()            # time-embedding / K-V caches and CUDA graphs depend on the weights
    return report


def load_sd_turbo(model, unet_state_dict=None, vae_state_dict=None):
    """Load diffusers-format SD-Turbo weights (``unet/diffusion_pytorch_model`` / ``vae/diffusion_pytorch_model``
    state dicts, e.g. read with safetensors) into the frozen backbone.  The VAE dict does not contain the modules the
    reference adds to it (``encoder.fr_blocks``, ``decoder.task_prompts`` / ``task_editors``), hence strict=False there
    with an explicit check that nothing else is missing."""
    if unet_state_dict is not None:
        model.base_model.unet.load_state_dict(unet_state_dict, strict=True)
    if vae_state_dict is not None:
        res = model.ae.vae.load_state_dict(vae_state_dict, strict=False)
        added = ("encoder.fr_blocks.", "decoder.task_prompts.", "decoder.task_editors.")
        missing = [k for k in res.missing_keys if not k.startswith(added)]
        if missing or res.unexpected_keys:
            raise RuntimeError("VAE state dict mismatch: missing %s unexpected %s" % (missing[:5], res.unexpected_keys[:5]))
    if hasattr(model, "clear_caches"):
        model.clear_caches()
