"""The C-ABI library loads on a CPU-only box and exports every symbol include/unirestore_b200.h declares; the ctypes
signature table covers exactly that set; the product path fails loudly when the library is missing."""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "unirestore_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(ur_[a-z0-9_]+)\s*\(", src)))


def test_header_declares_the_expected_surface():
    syms = declared_symbols()
    for must in ("ur_init", "ur_last_error", "ur_conv_gemm", "ur_attention", "ur_chan_stats", "ur_norm_apply",
                 "ur_layernorm", "ur_ddim_step", "ur_posterior_sample"):
        assert must in syms


def test_library_exports_every_declared_symbol():
    from unirestore_b200 import _cabi, build
    build.build()
    lib = _cabi.lib()
    for name in declared_symbols():
        assert hasattr(lib, name), "libunirestore_b200.so does not export %s" % name
    assert sorted(_cabi.SIGNATURES) == declared_symbols()
    assert lib.ur_version() >= 100
    assert isinstance(lib.ur_last_error(), bytes)


def test_conv_desc_matches_header_field_order():
    from unirestore_b200._cabi import ConvDesc
    src = open(os.path.join(ROOT, "include", "unirestore_b200.h")).read()
    body = src[src.index("typedef struct ur_conv_desc {"):src.index("} ur_conv_desc;")]
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    names = []
    for decl in body.split("{", 1)[1].split(";"):
        decl = decl.strip()
        if not decl:
            continue
        decl = re.sub(r"^(const\s+)?(void|float|double|int64_t|int)\s*\*?", "", decl).strip()
        names += [re.sub(r"\[\d+\]|[\*\s]", "", n) for n in decl.split(",")]
    assert names == [f[0] for f in ConvDesc._fields_]


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    from unirestore_b200 import _cabi
    monkeypatch.setattr(_cabi, "_lib", None)
    monkeypatch.setattr(_cabi, "LIB_PATH", str(tmp_path / "nope.so"))
    with pytest.raises(_cabi.UrError):
        _cabi.lib()
