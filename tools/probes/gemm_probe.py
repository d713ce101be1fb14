"""A/B timing of the GEMM kernel's two halves: run with a build of csrc/gemm3x.cu that carries
two temporary hooks read from GEMM_PROBE -- bit 0: the producer skips the B loads (and expects half the bytes),
bit 1: the issuer skips the MMAs.  Round-1 result at 25600 x 1536 x 1024 on a B200: 0.402 ms (normal), 0.390 ms
(bit 0), 0.230 ms (bit 1), 0.205 ms (both): TMA writes and MMA operand reads serialise on the shared-memory
port.  Without the hooks this script just times the kernel."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
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
