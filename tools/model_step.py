"""Full-model training step (control net + synth + loss + backward + Adam) the way train.py:84-130 runs it,
on config-1/2 shapes: (a) drop-in API path (multiscale_fft lists + train.py's loss in torch ops),
(b) fused multiscale_spectral_loss.  The control net is stock torch.nn (cuBLAS / cuDNN)."""
import argparse, json, os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ddsp_pytorch_b200 as ddsp
from ddsp_pytorch_b200.models.decoder import DDSPDecoder

ap = argparse.ArgumentParser(); ap.add_argument("--batch", type=int, default=16); ap.add_argument("--tf32", action="store_true"); ap.add_argument("--graph", action="store_true"); ap.add_argument("--stock-gru", action="store_true"); ap.add_argument("--stock-control-net", action="store_true", help="nn.Linear / nn.LayerNorm / nn.LeakyReLU / nn.GRU (cuBLAS, cuDNN) as in the reference"); args = ap.parse_args()
if args.tf32:
    torch.backends.cuda.matmul.allow_tf32 = True; torch.backends.cudnn.allow_tf32 = True
torch.manual_seed(0)
B, T, bs, sr = args.batch, 400, 160, 16000
model = DDSPDecoder(hidden_size=512, n_harmonic=100, n_bands=65, sample_rate=sr, block_size=bs, has_reverb=True).cuda()
if args.stock_control_net:
    from ddsp_pytorch_b200 import core
    core.to_stock_layers(model)
if args.stock_gru:      # A/B: the cuDNN recurrence the reference runs
    stock = torch.nn.GRU(1024, 512, batch_first=True).cuda(); stock.load_state_dict(model.decoder.gru.state_dict()); model.decoder.gru = stock
opt = torch.optim.Adam(model.parameters(), lr=1e-3)
g = torch.Generator().manual_seed(1)
batch = {"pitch": (torch.rand(B, T, 1, generator=g) * 400 + 100).cuda(), "loudness": torch.randn(B, T, 1, generator=g).cuda(),
         "sig": (0.1 * torch.randn(B, T * bs, generator=g)).cuda(),
         "noise": (torch.rand(B, T, bs, generator=g) * 2 - 1).cuda()}      # pre-drawn (the CPU draw is timed separately)
scales, ov = [4096, 2048, 1024, 512, 256, 128], 0.75

def step(fused):
    out = model(batch)
    rec = out["signal"].squeeze(-1)
    if fused:
        loss = ddsp.multiscale_spectral_loss(batch["sig"], rec, scales, ov)
    else:
        a, b = ddsp.multiscale_fft(batch["sig"], scales, ov), ddsp.multiscale_fft(rec, scales, ov)
        loss = sum((x - y).abs().mean() + (ddsp.safe_log(x) - ddsp.safe_log(y)).abs().mean() for x, y in zip(a, b))
    opt.zero_grad(); loss.backward(); opt.step()
    return loss

def timeit(fn, n=20):
    for _ in range(5): fn()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / n * 1e3

if args.graph:
    # whole training step (control net, synth, fused loss, backward, Adam) captured once and replayed
    opt = torch.optim.Adam(model.parameters(), lr=1e-3, capturable=True)
    side = torch.cuda.Stream(); side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(3): step(True)
    torch.cuda.current_stream().wait_stream(side); torch.cuda.synchronize()
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr):
        static_loss = step(True)
    print(json.dumps({"config": f"batch {B}, tf32={args.tf32}, gru={'cudnn' if args.stock_gru else 'cluster'}", "ms_graph_replay_fused_loss": timeit(gr.replay), "loss": float(static_loss)}))
    sys.exit(0)
res = {"config": f"DDSPDecoder hidden 512, 16 kHz, block 160, H=100, 4 s, batch {B}: full train step incl. control net and Adam, eager",
       "tf32": args.tf32, "gru": "cudnn" if (args.stock_gru or args.stock_control_net) else "cluster", "control_net": "torch.nn (cuBLAS SIMT SGEMM, cuDNN)" if args.stock_control_net else "this repo (split-bf16 tcgen05 GEMM, fused LayerNorm+LeakyReLU, cluster GRU)", "ms_fused_loss": timeit(lambda: step(True)), "ms_list_api_loss": timeit(lambda: step(False))}
with torch.no_grad():
    res["ms_forward_only"] = timeit(lambda: model(batch))
t0 = time.perf_counter(); n = torch.rand(B, T, bs) * 2 - 1; res["ms_cpu_noise_draw"] = (time.perf_counter() - t0) * 1e3
res["samples_per_s_fused"] = B * T * bs / (res["ms_fused_loss"] * 1e-3)
print(json.dumps(res))
