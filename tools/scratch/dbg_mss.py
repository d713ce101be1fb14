import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import ddsp_pytorch_b200 as ddsp
from oracle import ddsp_oracle as orc
for B, N, scales in [(3,1000,[64]),(3,1000,[128]),(3,1000,[256]),(1,1000,[64]),(2,1000,[64]),(4,1000,[64]),(3,9001,[64]),(3,9001,[1024]),(3,9001,[512]),(3,9001,[256]),(3,9001,[128]),(3,9000,[1024,512,256,128,64]), (2,64000,[64])]:
    g = torch.Generator().manual_seed(1)
    tgt = 0.1*torch.randn(B,N,generator=g); rec = 0.1*torch.randn(B,N,generator=g)
    r64 = rec.double().requires_grad_(True)
    orc.mss_loss(tgt.double(), r64, scales, 0.75).backward()
    r = rec.cuda().requires_grad_(True)
    ddsp.multiscale_spectral_loss(tgt.cuda(), r, scales, 0.75).backward()
    d = (r.grad.double().cpu()-r64.grad)
    rel = float(d.norm()/r64.grad.norm())
    bad = (d.abs() > 1e-3*r64.grad.abs().max()).nonzero()
    print(B,N,scales,'rel',rel, 'nbad',len(bad), 'first', bad[:3].tolist(), 'last', bad[-3:].tolist())
