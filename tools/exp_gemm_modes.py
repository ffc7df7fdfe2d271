"""Experiment: persistent GEMM, single CTA vs CTA pairs, per N-tile width, on the UNet shapes (CUDA events)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from unirestore_b200 import _cabi, ops  # noqa: E402
from tools.bench_gemm_shapes import SHAPES  # noqa: E402

dev = "cuda:0"
names = sys.argv[1].split(",") if len(sys.argv) > 1 else ["unet_c3_320_64", "unet_c3_640_32", "unet_c3_1280_16",
                                                          "lin_320_2560_4096", "vae_c3_128_512"]
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for name in names:
    B, H, W, Ci, Co, nt = SHAPES[name]
    x = torch.randn(B, H, W, Ci, device=dev).to(torch.bfloat16)
    w = (torch.randn(Co, nt * Ci, device=dev) * (nt * Ci) ** -0.5).to(torch.bfloat16)
    bias = torch.randn(Co, device=dev)
    taps = ops.TAPS_3x3 if nt == 9 else ops.TAPS_1x1
    out = torch.empty(B, H, W, Co, device=dev, dtype=torch.bfloat16)
    fl = 2.0 * B * H * W * Ci * Co * nt
    for pair in (0, 1):
        _cabi.lib().ur_debug_set_gemm_pair_mode(pair)
        for bn in (64, 128, 160, 256):
            try:
                for _ in range(3):
                    ops.conv_gemm(x, w, Co, taps=taps, bias=bias, out=out, bn=bn)
            except Exception as e:  # noqa: BLE001
                print("%-20s pair=%d bn=%3d  skipped (%s)" % (name, pair, bn, str(e)[:60]))
                continue
            ts = []
            for _ in range(7):
                flush.zero_()
                s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                s.record()
                ops.conv_gemm(x, w, Co, taps=taps, bias=bias, out=out, bn=bn)
                e.record()
                torch.cuda.synchronize()
                ts.append(s.elapsed_time(e))
            ts.sort()
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            for _ in range(10):
                ops.conv_gemm(x, w, Co, taps=taps, bias=bias, out=out, bn=bn)
            e.record()
            torch.cuda.synchronize()
            warm = s.elapsed_time(e) / 10
            t = ts[len(ts) // 2]
            print("%-20s pair=%d bn=%3d  cold %8.1f us %7.1f TF/s | back-to-back %8.1f us %7.1f TF/s" % (
                name, pair, bn, t * 1e3, fl / t / 1e9, warm * 1e3, fl / warm / 1e9), flush=True)
_cabi.lib().ur_debug_set_gemm_pair_mode(-1)
