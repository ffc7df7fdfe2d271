import sys, time, torch
sys.path.insert(0, '/root/repo')
from unirestore_b200.diffuie import DiffUIE
from unirestore_b200 import _cabi
dev='cuda:0'
m = DiffUIE(dict(type="CFRM"), dict(type="scedit", num_inference_steps=20), dict(type="TFA", prompt_len=1, task=["ir","cls","seg"])).eval().requires_grad_(False)
for n, p in m.named_parameters():
    if p.dim() <= 1 or "beta" in n or "gamma" in n or "task_prompts" in n:
        torch.nn.init.normal_(p, 0.0 if p.dim() > 1 or "bias" in n else 1.0, 0.05)
m = m.to(dev)
for mod in m.modules():
    for n, p in mod.named_parameters(recurse=False):
        if float(p.abs().max()) == 0.0: torch.nn.init.normal_(p, 0.0, 0.02)
img = torch.rand(8,3,512,512, device=dev)
g = torch.Generator(device=dev).manual_seed(1)
noise = (torch.randn(8,4,64,64, device=dev, generator=g), torch.randn(8,4,64,64, device=dev, generator=g))
y0 = m(img, 'ir', noise=noise)
torch.cuda.synchronize()
m.use_cuda_graph = True
t0=time.time(); y1 = m(img, 'ir', noise=noise); torch.cuda.synchronize(); print('capture+first replay %.1fs' % (time.time()-t0))
print('graph vs eager max diff', (y1-y0).abs().max().item(), 'finite', torch.isfinite(y1).all().item())
for _ in range(2): m(img,'ir',noise=noise)
torch.cuda.synchronize()
s,e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
c0=_cabi.launch_count
s.record()
for _ in range(3): y = m(img,'ir',noise=noise)
e.record(); torch.cuda.synchronize()
ms = s.elapsed_time(e)/3
print('graphed forward %.1f ms -> %.2f img/s ; python launches during replay: %d' % (ms, 8/ms*1e3, _cabi.launch_count-c0))
print(torch.cuda.max_memory_allocated()/2**30, 'GiB peak')
