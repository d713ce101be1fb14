"""GRU recurrence alone: cluster-persistent kernel (csrc/gru.cu) against cuDNN's nn.GRU, forward and
forward+backward, at the training shape (B, 400 frames, hidden 512) and the realtime shape (1, 8)."""
import argparse, json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ddsp_pytorch_b200 import core

ap = argparse.ArgumentParser(); ap.add_argument("--shapes", default="64x400,16x400,8x400,1x400,1x8,128x400"); args = ap.parse_args()


def timeit(fn, n=10):
    for _ in range(3): fn()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize(); return a.elapsed_time(b) / n


torch.manual_seed(0)
fast = core.gru(2, 512).cuda()
stock = torch.nn.GRU(1024, 512, batch_first=True).cuda(); stock.load_state_dict(fast.state_dict())
rows = []
for shp in args.shapes.split(","):
    B, T = (int(v) for v in shp.split("x"))
    x = torch.randn(B, T, 1024, device="cuda", requires_grad=True)
    go = torch.randn(B, T, 512, device="cuda")
    row = {"B": B, "T": T}
    for name, m in (("cluster", fast), ("cudnn", stock)):
        with torch.no_grad():
            row[f"{name}_fwd_ms"] = round(timeit(lambda: m(x)), 4)

        def fb():
            m.zero_grad(set_to_none=True); x.grad = None
            m(x)[0].backward(go)
        row[f"{name}_fwd_bwd_ms"] = round(timeit(fb), 4)
    with torch.no_grad():
        row["max_abs_diff"] = float((fast(x)[0] - stock(x)[0]).abs().max())
    rows.append(row); print(json.dumps(row), flush=True)
