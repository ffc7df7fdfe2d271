"""CPU oracle for the UniRestore hot path.  TEST INFRASTRUCTURE ONLY.

This package is an independent plain-PyTorch (fp32, CPU) restatement of the
reference's per-step ``Controller -> ControlledUNet(+SC-Tuner)`` forward and the
``VAE-encode(+CFRM)`` / ``VAE-decode(+TFA)`` bookends.  Only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import it, and only as the checker / CPU baseline.
The product path (``unirestore_b200``) never imports it and has no CPU fallback.

Pinning status
--------------
* Reference-owned arithmetic (``scedit.py``, ``taskeditor.py``, ``nafnet_arch.py``,
  ``cfrm.py``) and reference-owned wiring (``controller.py``, ``base_model.py``,
  ``autoencoder.py``, ``unifie.py``) are PINNED: ``oracle/make_golden.py`` executes the
  reference's own files from ``/root/reference`` (under the import shims in
  ``oracle/shims``) on deterministic weights and writes ``tests/golden/*.pt``; the
  oracle is checked against those fixtures in ``tests/test_oracle_golden.py``.
* The leaf arithmetic that lives in the un-vendored third-party dependency
  ``diffusers`` (comment-pinned ``diffusers==0.29.0`` at requirements.txt:14; absent
  from this image and from /root/reference) is restated in ``oracle/blocks.py`` and
  ``oracle/schedulers.py`` from its published algorithm: PARITY UNPINNED for those
  leaves (no golden vector of diffusers itself exists offline); they are anchored on
  the reference's call sites and on the known-answer constants of SURVEY.md section 8c.
"""
