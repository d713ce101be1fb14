import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from ddsp_pytorch_b200._lib import get_ops
ops = get_ops()
torch.manual_seed(0)
B, T, H = int(sys.argv[1]), 400, 512
gi = torch.randn(B, T, 3 * H, device="cuda"); w = torch.randn(3 * H, H, device="cuda") * 0.04; b = torch.zeros(3 * H, device="cuda")
def t(fn, n=5):
    for _ in range(2): fn()
    a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); a.record()
    for _ in range(n): fn()
    e.record(); torch.cuda.synchronize(); return a.elapsed_time(e) / n
print("B", B, "probe", os.environ.get("GRU_PROBE"), "fwd ms", round(t(lambda: ops.gru_fwd(gi, w, b, None, True)), 3), "us/step", round(t(lambda: ops.gru_fwd(gi, w, b, None, True)) * 1000 / T, 2))
y, gates = ops.gru_fwd(gi, w, b, None, True)
dy = torch.randn_like(y)
print("bwd ms", round(t(lambda: ops.gru_bwd(dy, None, w, y, None, gates)), 3))
