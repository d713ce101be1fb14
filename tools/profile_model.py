import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ddsp_pytorch_b200 as ddsp
from ddsp_pytorch_b200.models.decoder import DDSPDecoder
from torch.profiler import profile, ProfilerActivity
torch.backends.cuda.matmul.allow_tf32 = "--tf32" in sys.argv
torch.backends.cudnn.allow_tf32 = "--tf32" in sys.argv
torch.manual_seed(0)
B, T, bs, sr = int(os.environ.get("BATCH", 64)), 400, 160, 16000
model = DDSPDecoder(hidden_size=512, n_harmonic=100, n_bands=65, sample_rate=sr, block_size=bs, has_reverb=True).cuda()
model.noise_synth.device_noise = True
opt = torch.optim.Adam(model.parameters(), lr=1e-3)
g = torch.Generator().manual_seed(1)
batch = {"pitch": (torch.rand(B, T, 1, generator=g) * 400 + 100).cuda(), "loudness": torch.randn(B, T, 1, generator=g).cuda(),
         "sig": (0.1 * torch.randn(B, T * bs, generator=g)).cuda()}
def step():
    out = model(batch)
    loss = ddsp.multiscale_spectral_loss(batch["sig"], out["signal"].squeeze(-1), [4096, 2048, 1024, 512, 256, 128], 0.75)
    opt.zero_grad(); loss.backward(); opt.step()
for _ in range(5): step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(5): step()
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=25, max_name_column_width=70))
