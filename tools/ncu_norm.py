"""Driver for ncu captures of the HBM-bound normalisation kernels on UNet / VAE shapes."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from unirestore_b200 import ops  # noqa: E402

dev = "cuda:0"
SHAPES = [(8, 64, 64, 320, 32), (8, 256, 256, 256, 32)]
if len(sys.argv) > 1:          # B H W C [groups]
    v = [int(t) for t in sys.argv[1:]]
    SHAPES = [(v[0], v[1], v[2], v[3], v[4] if len(v) > 4 else 32)]
for (B, H, W, C, G) in SHAPES:
    x = torch.randn(B, H, W, C, device=dev).to(torch.bfloat16)
    g = torch.ones(C, device=dev)
    b = torch.zeros(C, device=dev)
    for _ in range(3):
        y = ops.group_norm(x, G, g, b, 1e-5, silu=True)
    ln = ops.layernorm(x.view(B, 1, H * W, C), g, b, 1e-5)
torch.cuda.synchronize()
print("done")
