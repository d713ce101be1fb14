"""Runs only the filtered-noise kernels at config-2 shapes (for ncu)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ddsp_pytorch_b200 as ddsp
B, T, NB, bs = 64, 400, 65, 160
g = torch.Generator().manual_seed(0)
mags = torch.randn(B, T, NB, generator=g).cuda()
noise = (torch.rand(B, T, bs, generator=g) * 2 - 1).cuda()
add = torch.randn(B, T * bs, 1, generator=g).cuda()
go = torch.randn(B, T * bs, 1, generator=g).cuda()
ops = torch.ops.ddsp_b200
for _ in range(3):
    y = ops.noise_fwd(mags, noise, add, True, -5.0)
    d = ops.noise_bwd(go, noise, mags, NB, True, -5.0)
torch.cuda.synchronize()
print(float(y.sum()), float(d.sum()))
