"""Throughput of the split-bf16 tcgen05 GEMM against torch's float32 matmul (SIMT SGEMM) and TF32 matmul on the
control net's shapes (rows = batch 64 x 400 frames)."""
import json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ddsp_pytorch_b200._lib import get_ops
ops = get_ops()


def timeit(fn, n=10):
    for _ in range(3): fn()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize(); return a.elapsed_time(b) / n


for M, N, K in [(25600, 512, 512), (25600, 1536, 1024), (512, 512, 25600), (1536, 1024, 25600), (25600, 101, 512)]:
    a = torch.randn(M, K, device="cuda"); b = torch.randn(N, K, device="cuda") * 0.05
    a_s, b_s = ops.gemm3x_split(a, False), ops.gemm3x_split(b, False)
    row = {"M": M, "N": N, "K": K, "gflop": 2e-9 * M * N * K}
    row["split_a_ms"] = timeit(lambda: ops.gemm3x_split(a, False))
    row["split_a_T_ms"] = timeit(lambda: ops.gemm3x_split(a, True))
    row["gemm3x_ms"] = timeit(lambda: ops.gemm3x_mm(a_s, b_s, M, N, K, None, False, False))
    torch.backends.cuda.matmul.allow_tf32 = False
    row["torch_fp32_ms"] = timeit(lambda: a @ b.t())
    torch.backends.cuda.matmul.allow_tf32 = True
    row["torch_tf32_ms"] = timeit(lambda: a @ b.t())
    torch.backends.cuda.matmul.allow_tf32 = False
    row["gemm3x_tflops_fp32_equiv"] = row["gflop"] / row["gemm3x_ms"]
    row["torch_fp32_tflops"] = row["gflop"] / row["torch_fp32_ms"]
    ref = a.double() @ b.double().t()
    row["err_gemm3x"] = float((ops.gemm3x_mm(a_s, b_s, M, N, K, None, False, False).double() - ref).abs().max() / ref.abs().max())
    row["err_torch_fp32"] = float(((a @ b.t()).double() - ref).abs().max() / ref.abs().max())
    print(json.dumps({k: (round(v, 5) if isinstance(v, float) and v > 1e-3 else v) for k, v in row.items()}), flush=True)
