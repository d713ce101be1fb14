"""Golden vectors for the control net from the UNMODIFIED reference ``GRUDecoder``
(``/root/reference/ddsp/models/decoder.py:9-68``) at the real width (hidden 512), evaluated in float64 on the CPU.

TEST INFRASTRUCTURE; runs only in the build container (see make_golden.py).  The 6 M weights are not stored:
reference and test both build the module after ``torch.manual_seed(SEED)`` (the parameter creation order is
the same, the fixture holds a per-tensor checksum to prove it); stored are the inputs, a sub-sampled slice of the
float64 output, d(output)/d(loudness input) and a few parameter gradients for a closed-form grad_output, and the reference's own float32
deviation.

    PYTHONDONTWRITEBYTECODE=1 python oracle/make_golden_control_net.py
"""
from __future__ import annotations

import torch

from make_golden import import_reference, npz

SEED, B, T = 21, 4, 160            # 640 rows: above the row threshold of the tensor-core GEMM path


def grad_output(b: int, t: int, c: int) -> torch.Tensor:
    """Closed-form grad_output (no RNG, so it need not be stored): the test rebuilds it with the same formula."""
    bi = torch.arange(b, dtype=torch.float64).view(b, 1, 1)
    ti = torch.arange(t, dtype=torch.float64).view(1, t, 1)
    ci = torch.arange(c, dtype=torch.float64).view(1, 1, c)
    return torch.sin(0.37 * bi + 0.011 * ti * (ci % 7 + 1) + 0.05 * ci).float()


def main():
    import_reference()
    from ddsp.models.decoder import GRUDecoder
    torch.set_num_threads(8)
    torch.manual_seed(SEED)
    dec = GRUDecoder(hidden_size=512)
    g = torch.Generator().manual_seed(SEED + 1)
    f0 = torch.rand(B, T, 1, generator=g) * 500 + 100
    loud = torch.randn(B, T, 1, generator=g)
    go = grad_output(B, T, 512)
    sums = torch.stack([p.detach().double().sum() for p in dec.state_dict().values()])
    abss = torch.stack([p.detach().double().abs().sum() for p in dec.state_dict().values()])
    out32 = dec(f0, loud).detach()
    dec64 = dec.double()
    l64 = loud.double().requires_grad_(True)
    out64 = dec64(f0.double(), l64)
    (out64 * go.double()).sum().backward()
    grads = {k: p.grad for k, p in dec64.named_parameters()}
    npz("control_net_gru_decoder", seed=SEED, f0=f0, loudness=loud,
        weight_sums=sums, weight_abs_sums=abss,
        out=out64.detach()[:, ::4, ::4],                     # (4, 40, 128) slice of the (4, 160, 512) output
        out_ref_fp32_max_abs=(out32.double() - out64.detach()).abs().max(),
        d_loudness=l64.grad,
        d_gru_bias_hh=grads["gru.bias_hh_l0"], d_out_mlp_ln_weight=grads["out_mlp.7.weight"],
        d_f0_mlp_w0=grads["f0_mlp.0.weight"], d_gru_weight_hh_slice=grads["gru.weight_hh_l0"][::16, ::16])


if __name__ == "__main__":
    main()
