"""Runs only the fused multi-scale loss at config-2 shapes (for ncu)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ddsp_pytorch_b200 as ddsp
from ddsp_pytorch_b200.functions import hann_window_like_reference
B, N = 64, 64000
scales = [4096, 2048, 1024, 512, 256, 128]
g = torch.Generator().manual_seed(0)
tgt = (0.1 * torch.randn(B, N, generator=g)).cuda()
rec = (0.1 * torch.randn(B, N, generator=g)).cuda()
win = torch.cat([hann_window_like_reference(s, rec.device) for s in scales])
for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 3):
    loss, d = torch.ops.ddsp_b200.mss_loss_fwd(tgt, rec, scales, 0.75, win, True)
torch.cuda.synchronize()
print(float(loss))
