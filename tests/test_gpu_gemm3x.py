"""The tcgen05 split-bf16 GEMM (csrc/gemm3x.cu) against a float64 product.

It stands in for the SIMT SGEMM torch runs for the reference's float32 nn.Linear layers
(ddsp/core.py:122-129), so the bar is float32 accuracy: error relative to the largest output entry
<= 2e-6 (a float32 dot product of length K accumulates ~sqrt(K) * 6e-8), and never worse than 4x torch's own
float32 matmul on the same operands.
"""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _ops():
    from ddsp_pytorch_b200._lib import get_ops
    return get_ops()


@pytest.mark.parametrize("M,N,K,bias", [(128, 128, 32, False), (128, 128, 512, True), (300, 200, 70, True),
                                         (1, 65, 514, True), (2560, 512, 512, True), (512, 1024, 6400, False),
                                         (101, 512, 3000, True)])
def test_gemm3x_matches_float64(M, N, K, bias):
    ops = _ops()
    g = torch.Generator().manual_seed(M * 7 + N * 3 + K)
    a = torch.randn(M, K, generator=g).cuda()
    b = (torch.randn(N, K, generator=g) * 0.05).cuda()
    bv = torch.randn(N, generator=g).cuda() if bias else None
    ref = a.double() @ b.double().t() + (bv.double() if bias else 0)
    got = ops.gemm3x_mm(ops.gemm3x_split(a, False), ops.gemm3x_split(b, False), M, N, K, bv, False, False)
    assert got.shape == (M, N)
    scale = float(ref.abs().max())
    err = float((got.double() - ref).abs().max()) / scale
    torch.backends.cuda.matmul.allow_tf32 = False
    err_torch = float(((a @ b.t() + (bv if bias else 0)).double() - ref).abs().max()) / scale
    assert err < 2e-6 and err < 4 * err_torch + 1e-7, (err, err_torch)


def test_transposed_split_gives_the_transposed_product():
    ops = _ops()
    g = torch.Generator().manual_seed(5)
    dy = torch.randn(777, 130, generator=g).cuda()          # (M, N)
    x = torch.randn(777, 90, generator=g).cuda()            # (M, K)
    ref = dy.double().t() @ x.double()                      # dW = dy^T x  (N, K)
    got = ops.gemm3x_mm(ops.gemm3x_split(dy, True), ops.gemm3x_split(x, True), 130, 90, 777, None, False, False)
    assert float((got.double() - ref).abs().max()) / float(ref.abs().max()) < 2e-6


@pytest.mark.parametrize("rows,n,k", [(777, 130, 90), (64, 64, 64), (3000, 512, 514), (130, 1536, 1024)])
def test_mn_major_operands_need_no_transposed_copies(rows, n, k):
    """One split per matrix serves all three GEMMs of a layer: dx = dy W reads W MN-major, dW = dy^T x reads dy
    and x MN-major (the contraction index is the row index of the stored matrix)."""
    ops = _ops()
    g = torch.Generator().manual_seed(rows + n)
    dy = torch.randn(rows, n, generator=g).cuda()
    x = torch.randn(rows, k, generator=g).cuda()
    w = (torch.randn(n, k, generator=g) * 0.05).cuda()
    dys, xs, ws = ops.gemm3x_split(dy, False), ops.gemm3x_split(x, False), ops.gemm3x_split(w, False)
    dw = ops.gemm3x_mm(dys, xs, n, k, rows, None, True, True)
    ref = dy.double().t() @ x.double()
    assert float((dw.double() - ref).abs().max()) / float(ref.abs().max()) < 2e-6
    dx = ops.gemm3x_mm(dys, ws, rows, k, n, None, False, True)
    ref = dy.double() @ w.double()
    assert float((dx.double() - ref).abs().max()) / float(ref.abs().max()) < 2e-6
    y = ops.gemm3x_mm(xs, ws, rows, n, k, None, False, False)
    ref = x.double() @ w.double().t()
    assert float((y.double() - ref).abs().max()) / float(ref.abs().max()) < 2e-6


def test_split_parts_are_exact():
    ops = _ops()
    x = torch.randn(50, 37, device="cuda") * 1e3
    s = ops.gemm3x_split(x, False)
    assert s.dtype == torch.bfloat16 and s.shape == (192, 64)          # 3 parts x 64 rows (50 padded with zero rows)
    parts = s.view(3, 64, 64).float()
    assert torch.equal(parts[0, :50, :37] + parts[1, :50, :37] + parts[2, :50, :37], x)   # 3 x 8 mantissa bits: exact
    assert float(parts[:, :, 37:].abs().max()) == 0.0 and float(parts[:, 50:].abs().max()) == 0.0     # padding is zero
    both, both_t = ops.gemm3x_split_both(x)
    st = ops.gemm3x_split(x, True)
    assert torch.equal(both_t, st) and torch.equal(both.view(3, 50, 64), s.view(3, 64, 64)[:, :50])


@pytest.mark.parametrize("rows,fan_in,fan_out", [(4 * 400, 512, 512), (3 * 300, 514, 512), (2000, 512, 101)])
def test_linear_module_matches_float64_forward_and_backward(rows, fan_in, fan_out):
    """core.Linear = nn.Linear's parameters and call; y, dx, dW, db against float64 (grads 1e-3 relative bar
    of the north star, met with three orders of margin)."""
    from ddsp_pytorch_b200 import core
    torch.manual_seed(rows + fan_in)
    lin = core.Linear(fan_in, fan_out).cuda()
    assert set(lin.state_dict()) == {"weight", "bias"}
    x = torch.randn(rows // 100, 100, fan_in, device="cuda", requires_grad=True)
    go = torch.randn(rows // 100, 100, fan_out, device="cuda")
    y = lin(x)
    y.backward(go)
    xr = x.detach().double().requires_grad_(True)
    w, b = lin.weight.detach().double().requires_grad_(True), lin.bias.detach().double().requires_grad_(True)
    yr = torch.nn.functional.linear(xr, w, b)
    yr.backward(go.double())

    def rel(a, r):
        return float((a.double() - r).abs().max() / r.abs().max())
    assert rel(y.detach(), yr.detach()) < 2e-6
    assert rel(x.grad, xr.grad) < 2e-6
    assert rel(lin.weight.grad, w.grad) < 2e-6
    assert rel(lin.bias.grad, b.grad) < 2e-6


@pytest.mark.parametrize("rows,cols", [(1, 8), (100, 101), (25600, 512), (700, 1536), (130, 2500)])
def test_split_with_column_sums(rows, cols):
    ops = _ops()
    x = torch.randn(rows, cols, device="cuda")
    s, colsum = ops.gemm3x_split_colsum(x)
    assert torch.equal(s, ops.gemm3x_split(x, False))
    ref = x.double().sum(0)
    assert float((colsum.double() - ref).abs().max()) <= 1e-5 * max(1.0, float(ref.abs().max()))
    assert torch.equal(colsum, ops.gemm3x_split_colsum(x)[1])        # fixed summation order


def test_c_abi_direct_call_of_the_gemm():
    """split + GEMM through libddsp_b200.so with plain pointers (ctypes), no torch op."""
    import ctypes
    from ddsp_pytorch_b200._lib import core_library
    lib = core_library()
    vp, i64, i32 = ctypes.c_void_p, ctypes.c_int64, ctypes.c_int
    lib.ddsp_b200_gemm3x_ld.restype, lib.ddsp_b200_gemm3x_ld.argtypes = i64, [i64]
    lib.ddsp_b200_gemm3x_split.restype = i32
    lib.ddsp_b200_gemm3x_split.argtypes = [vp, i64, i64, i64, i32, vp, i64, vp]
    lib.ddsp_b200_gemm3x.restype = i32
    lib.ddsp_b200_gemm3x.argtypes = [vp, i64, i64, i32, vp, i64, i64, i32, vp, vp, i64, i32, i32, i32, vp, vp]
    M, N, K = 300, 200, 70
    g = torch.Generator().manual_seed(9)
    x, w, b = torch.randn(M, K, generator=g).cuda(), torch.randn(N, K, generator=g).cuda(), torch.randn(N, generator=g).cuda()
    y = torch.empty(M, N, device="cuda")
    ld = lib.ddsp_b200_gemm3x_ld(K)
    assert ld == 128

    def rp(r):
        return (r + 63) // 64 * 64
    xs = torch.empty(3 * rp(M), ld, dtype=torch.bfloat16, device="cuda")
    ws = torch.empty(3 * rp(N), ld, dtype=torch.bfloat16, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    assert lib.ddsp_b200_gemm3x_split(x.data_ptr(), M, K, K, 0, xs.data_ptr(), rp(M), st) == 0
    assert lib.ddsp_b200_gemm3x_split(w.data_ptr(), N, K, K, 0, ws.data_ptr(), rp(N), st) == 0
    assert lib.ddsp_b200_gemm3x(xs.data_ptr(), rp(M), ld, 0, ws.data_ptr(), rp(N), ld, 0, b.data_ptr(), y.data_ptr(), N,
                                M, N, K, None, st) == 0
    torch.cuda.synchronize()
    ref = x.double() @ w.double().t() + b.double()
    assert float((y.double() - ref).abs().max() / ref.abs().max()) < 2e-6
    assert lib.ddsp_b200_gemm3x(None, rp(M), ld, 0, ws.data_ptr(), rp(N), ld, 0, None, y.data_ptr(), N, M, N, K, None, st) == -1


@pytest.mark.parametrize("rows,fan_in,fan_out", [(2, 1, 512), (2, 512, 101), (25600, 1, 512), (3, 514, 65), (1, 16, 512)])
def test_small_and_narrow_layers_run_on_the_kernel(rows, fan_in, fan_out):
    """No library fallback for CUDA float32 inputs: 2-row calls of the realtime path, the decoder's fan-in-1 first
    layers (decoder.py:18-19) and odd widths all take the split-bf16 GEMM, forward and backward, against float64."""
    from ddsp_pytorch_b200 import core
    torch.manual_seed(rows + fan_in)
    lin = core.Linear(fan_in, fan_out).cuda()
    ref = torch.nn.Linear(fan_in, fan_out).double()
    ref.load_state_dict({k: v.detach().cpu().double() for k, v in lin.state_dict().items()})
    x = torch.randn(rows, fan_in)
    xd = x.cuda().requires_grad_(True)
    y = lin(xd)
    x64 = x.double().requires_grad_(True)
    y64 = ref(x64)
    go = torch.randn(rows, fan_out)
    (y * go.cuda()).sum().backward()
    (y64 * go.double()).sum().backward()
    scale = float(y64.abs().max())
    assert float((y.detach().cpu().double() - y64.detach()).abs().max()) <= 2e-6 * max(1.0, scale)
    for got, want in ((xd.grad, x64.grad), (lin.weight.grad, ref.weight.grad), (lin.bias.grad, ref.bias.grad)):
        assert float((got.cpu().double() - want).norm()) <= 1e-5 * max(1e-30, float(want.norm()))
    assert isinstance(y.grad_fn, torch.autograd.function.BackwardCFunction) and "Linear3x" in type(y.grad_fn).__name__
