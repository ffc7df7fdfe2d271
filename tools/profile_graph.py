"""In-situ kernel time shares of the benchmarked path: torch.profiler (CUPTI) around CUDA-graph replays of
DiffUIE.forward (B=8, 512x512, 20 DDIM steps by default).  Unlike the ncu launch list (serialised, cold cache) these
are the durations the kernels have INSIDE the graph (warm L2, side streams running) -- sums can exceed the wall span
because of stream overlap.  python tools/profile_graph.py [--batch 8 --size 512 --steps 20] > profiles/..."""
import argparse
import collections
import os
import re
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import CFG, cheap_init_  # noqa: E402
from unirestore_b200.diffuie import DiffUIE  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=8)
ap.add_argument("--size", type=int, default=512)
ap.add_argument("--steps", type=int, default=20)
ap.add_argument("--task", default="ir")
ap.add_argument("--reps", type=int, default=3)
a = ap.parse_args()
dev = "cuda:0"
cfg = (CFG[0], dict(CFG[1], num_inference_steps=a.steps), CFG[2])
m = cheap_init_(DiffUIE(*cfg)).eval().requires_grad_(False).to(dev)
m.use_cuda_graph = True
img = torch.rand(a.batch, 3, a.size, a.size, device=dev)
for _ in range(3):
    m(img, a.task)
torch.cuda.synchronize()
s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
s.record()
for _ in range(a.reps):
    m(img, a.task)
e.record()
torch.cuda.synchronize()
wall = s.elapsed_time(e) / a.reps
from torch.profiler import ProfilerActivity, profile  # noqa: E402
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(a.reps):
        m(img, a.task)
    torch.cuda.synchronize()
agg = collections.defaultdict(lambda: [0.0, 0])
for ev in prof.events():
    if ev.device_type == torch.autograd.DeviceType.CUDA:
        name = re.sub(r"^void ", "", ev.name)
        name = re.sub(r"\(.*$", "", name)
        agg[name][0] += ev.device_time if hasattr(ev, "device_time") else ev.cuda_time
        agg[name][1] += 1
tot = sum(v[0] for v in agg.values())
print("# in-graph kernel durations (torch.profiler / CUPTI), %d replays of DiffUIE.forward B=%d %dx%d %d DDIM steps, task %s"
      % (a.reps, a.batch, a.size, a.size, a.steps, a.task))
print("# wall %.2f ms per forward (CUDA events, no profiler); summed kernel time %.2f ms per forward (streams overlap)"
      % (wall, tot / a.reps / 1e3))
for name, (t, n) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    print("%10.1f us %5.1f%% %6d  avg %7.1f  %s" % (t / a.reps, 100 * t / tot, n // a.reps, t / n, name[:110]))
