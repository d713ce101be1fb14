"""Fused LayerNorm + LeakyReLU (csrc/layernorm.cu) against the two stock torch ops in float64
(ddsp/core.py:122-129).  Bars: forward 1e-5 max abs, gradients 1e-3 relative (north star per-op bar)."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("rows,N", [(1, 512), (37, 128), (1000, 256), (999, 384), (64 * 400, 512)])
def test_ln_lrelu_forward_backward(rows, N):
    from ddsp_pytorch_b200 import core
    torch.manual_seed(rows + N)
    m = core.LayerNormLeakyReLU(N).cuda()
    with torch.no_grad():
        m.weight.uniform_(0.5, 1.5)
        m.bias.normal_(0, 0.3)
    x = (torch.randn(rows, N, device="cuda") * 2 + 0.5).requires_grad_(True)
    go = torch.randn(rows, N, device="cuda")
    y = m(x)
    y.backward(go)
    xr = x.detach().double().requires_grad_(True)
    w, b = m.weight.detach().double().requires_grad_(True), m.bias.detach().double().requires_grad_(True)
    yr = F.leaky_relu(F.layer_norm(xr, (N,), w, b, m.eps), 0.01)
    yr.backward(go.double())

    def rel(a, r):
        return float((a.double() - r).abs().max() / r.abs().max())
    assert float((y.detach().double() - yr.detach()).abs().max()) < 1e-5
    assert rel(x.grad, xr.grad) < 1e-3
    assert rel(m.weight.grad, w.grad) < 1e-3
    assert rel(m.bias.grad, b.grad) < 1e-3


def test_mlp_keeps_reference_state_dict_keys_and_values():
    """core.mlp == Sequential(Linear, LayerNorm, LeakyReLU) x 3 of the reference: same keys, same output."""
    import torch.nn as nn
    from ddsp_pytorch_b200 import core
    torch.manual_seed(1)
    ours = core.mlp(1, 512, 3).cuda()
    layers = []
    for a, b in [(1, 512), (512, 512), (512, 512)]:
        layers += [nn.Linear(a, b), nn.LayerNorm(b), nn.LeakyReLU()]
    ref = nn.Sequential(*layers).cuda()
    assert list(ours.state_dict()) == list(ref.state_dict())
    ref.load_state_dict(ours.state_dict())
    x = torch.randn(8, 400, 1, device="cuda")
    assert float((ours(x) - ref(x)).abs().max()) < 2e-5


def test_unsupported_width_composes_stock_ops():
    from ddsp_pytorch_b200 import core
    m = core.LayerNormLeakyReLU(100).cuda()
    x = torch.randn(5, 100, device="cuda")
    assert torch.allclose(m(x), F.leaky_relu(F.layer_norm(x, (100,), m.weight, m.bias, m.eps), 0.01))
