import os, sys, torch
sys.path.insert(0, os.getcwd())
from ddsp_pytorch_b200._lib import get_ops
ops = get_ops()
M, N, K = 25600, 1536, 1024
a = torch.randn(M, K, device="cuda"); b = torch.randn(N, K, device="cuda")
a_s, b_s = ops.gemm3x_split(a, False), ops.gemm3x_split(b, False)
def t(fn, n=10):
    for _ in range(3): fn()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); s.record()
    for _ in range(n): fn()
    e.record(); torch.cuda.synchronize(); return s.elapsed_time(e) / n
print("probe", os.environ.get("GEMM_PROBE"), "ms", round(t(lambda: ops.gemm3x_mm(a_s, b_s, M, N, K, None, False, False)), 4))
