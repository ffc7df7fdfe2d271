"""ctypes binding of the C-ABI declared in ``include/unirestore_b200.h``.

The product path has NO fallback: if the shared library is missing (or a call fails) this module
raises.  ``lib()`` loads ``unirestore_b200/libunirestore_b200.so`` (built in-tree by
``unirestore_b200/build.py``); loading works on a CPU-only box (symbol checks), compute calls need
an sm_100 GPU.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("UNIRESTORE_B200_LIB") or os.path.join(_HERE, "libunirestore_b200.so")   # (override: A/B of two builds)

UR_ACT_NONE, UR_ACT_SILU, UR_ACT_GELU, UR_ACT_GEGLU, UR_ACT_GATE = 0, 1, 2, 3, 4
UR_DT_BF16, UR_DT_F32 = 0, 1


class UrError(RuntimeError):
    pass


class ConvDesc(C.Structure):
    _fields_ = [
        ("x1", C.c_void_p), ("x2", C.c_void_p),
        ("c1", C.c_int), ("c2", C.c_int), ("ld1", C.c_int), ("ld2", C.c_int),
        ("batch", C.c_int), ("hin", C.c_int), ("win", C.c_int),
        ("w", C.c_void_p), ("w_batched", C.c_int), ("w_ld", C.c_int64), ("w_bs", C.c_int64), ("n", C.c_int),
        ("ntaps", C.c_int), ("tap_dy", C.c_int * 9), ("tap_dx", C.c_int * 9), ("stride", C.c_int),
        ("group_kc", C.c_int), ("group_nc", C.c_int),
        ("hout", C.c_int), ("wout", C.c_int),
        ("out", C.c_void_p), ("out_dtype", C.c_int),
        ("out_sb", C.c_int64), ("out_sy", C.c_int64), ("out_sx", C.c_int64),
        ("alpha", C.c_float),
        ("bias", C.c_void_p), ("rowvec", C.c_void_p), ("rowvec_sb", C.c_int64),
        ("chscale", C.c_void_p), ("chscale_sb", C.c_int64),
        ("residual", C.c_void_p), ("res_sb", C.c_int64), ("res_sy", C.c_int64), ("res_sx", C.c_int64),
        ("act", C.c_int), ("bn", C.c_int),
        ("stats", C.c_void_p), ("stats_ld", C.c_int), ("stats_off", C.c_int),
        ("workspace", C.c_void_p), ("workspace_bytes", C.c_int64),
    ]


_P, _I, _I64, _F = C.c_void_p, C.c_int, C.c_int64, C.c_float

# name -> (restype, argtypes); every symbol include/unirestore_b200.h declares
SIGNATURES = {
    "ur_init": (C.c_int, [C.c_int]),
    "ur_last_error": (C.c_char_p, []),
    "ur_version": (C.c_int, []),
    "ur_conv_gemm": (C.c_int, [C.POINTER(ConvDesc), C.c_void_p]),
    "ur_conv_gemm_pick_bn": (C.c_int, [C.c_int, C.c_int]),
    "ur_debug_force_gemm_v1": (C.c_int, [C.c_int]),
    "ur_debug_set_gemm_trace": (C.c_int, [_P]),
    "ur_debug_set_gemm_pair_mode": (C.c_int, [C.c_int]),
    "ur_debug_set_gemm_splitk": (C.c_int, [C.c_int]),
    "ur_debug_set_gemm_tma_store": (C.c_int, [C.c_int]),
    "ur_debug_set_attention_trace": (C.c_int, [_P]),
    "ur_debug_set_attention_impl": (C.c_int, [C.c_int]),
    "ur_debug_set_attention_kv1": (C.c_int, [C.c_int]),
    "ur_debug_set_attention_poly": (C.c_int, [C.c_int]),
    "ur_chan_stats": (C.c_int, [_P, _I64, _I64, _I, _I, _I, _P, _I, _I, _I, _P]),
    "ur_group_norm": (C.c_int, [_P, _I64, _I64, _I, _P, _I64, _I64, _I, _I, _I, _I, _P, _P, _F, _I, _P, _I64, _I64, _P]),
    "ur_group_norm_cluster_size": (C.c_int, []),
    "ur_debug_set_group_norm_cluster": (C.c_int, [C.c_int]),
    "ur_norm_apply": (C.c_int, [_P, _I64, _I64, _I, _P, _I64, _I64, _I, _P, _P, _I, _I, _I, _P, _P, _F, _I, _P, _I64,
                                _I64, _P]),
    "ur_layernorm": (C.c_int, [_P, _I64, _P, _I64, _I64, _I, _P, _P, _F, _P]),
    "ur_scale_channels": (C.c_int, [_P, _I64, _I64, _I, _I, _I, _P, _I, _P]),
    "ur_attention": (C.c_int, [_P, _I64, _I64, _P, _I64, _I64, _P, _I64, _I64, _P, _I64, _I64, _I, _I, _I, _I, _I, _I, _F,
                               _P]),
    "ur_softmax_rows": (C.c_int, [_P, _I64, _P, _I64, _I64, _I, _I, _P]),
    "ur_transpose_tokens": (C.c_int, [_P, _I64, _I64, _I, _I, _I, _P, _I, _P]),
    "ur_dwconv3x3_gate": (C.c_int, [_P, _I, _I, _I, _I, _P, _P, _P, _P, _P]),
    "ur_small_linear": (C.c_int, [_P, _I, _F, _I64, _P, _P, _P, _I64, _I, _I, _I, _I, _I, _I, _P]),
    "ur_timestep_embedding": (C.c_int, [_P, _I, _I, _P, _P]),
    "ur_adanaf_scales": (C.c_int, [_P, _I, _I, _I, _I, _P, _P, _P, _P, _P, _P]),
    "ur_tfa_gates": (C.c_int, [_P, _I, _I, _I, _I, _P, _P, _P, _P, _P, _P, _P, _P]),
    "ur_posterior_sample": (C.c_int, [_P, _P, _F, _I, _I, _P, _P, _P]),
    "ur_latent_axpby": (C.c_int, [_P, _F, _P, _F, _I, _I, _P, _P, _F, _P]),
    "ur_ddim_step": (C.c_int, [_P, _P, _I, _F, _F, _F, _F, _I, _I, _I, _P, _P]),
    "ur_image_to_nhwc8": (C.c_int, [_P, _I64, _I64, _I64, _I64, _I, _I, _I, _I, _F, _F, _P, _P]),
    "ur_image_metrics": (C.c_int, [_P, _P, _I, _I, _I, _I, _I, _F, _P, _P]),
    "ur_resize_pad": (C.c_int, [_P, _I64, _I64, _I64, _I64, _I, _I, _I, _I, _I, _I, _I, _I, _I, _P, _P]),
    "ur_nhwc_to_image": (C.c_int, [_P, _I, _I, _I, _I, _I, _I, _I, _I, _I, _F, _F, _I, _P, _P]),
    "ur_concat_channels": (C.c_int, [_P, _I64, _I, _P, _I64, _I, _I64, _P, _P]),
}

_lib = None
_inited = set()


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise UrError("%s not found: build it with `python -m unirestore_b200.build` "
                          "(the product path has no CPU / PyTorch fallback)" % LIB_PATH)
        _lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(_lib, name)
            fn.restype, fn.argtypes = res, args
    return _lib


launch_count = 0          # C-ABI compute calls issued (each is >= 1 kernel launch); read by bench.py
launch_counts = {}        # the same, per entry point (launches-per-DDIM-step reports)


def check(rc: int, what: str = ""):
    global launch_count
    launch_count += 1
    launch_counts[what] = launch_counts.get(what, 0) + 1
    if rc != 0:
        raise UrError("%s failed (%d): %s" % (what or "unirestore_b200 call", rc,
                                              lib().ur_last_error().decode(errors="replace")))


def ensure_init(device_index: int):
    if device_index not in _inited:
        check(lib().ur_init(device_index), "ur_init")
        _inited.add(device_index)
