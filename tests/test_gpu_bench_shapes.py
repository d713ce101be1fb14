"""GPU parity at the BENCHMARKED shapes and through whole chains (round-2 additions, VERDICT r01 "parity holes").

* the long convolution / reverb, forward AND gradients (d x, d noise, d decay, d wet) for a fixed grad_output at
  config 2 (B=64, N=64000, L=16000: transform length 20 x 4096, split partial spectra) and config 4 (N=192000, L=48000: n = 2^18)
* the gradient chain controls -> harmonic + noise -> reverb with a LINEAR loss on the signal (no L1 sign ties), 1e-3
* the whole model's parameter gradients against a float64 run of the unmodified reference (golden fixture)
* the ctypes example of INTEGRATION.md section 3, executed verbatim
Oracle = oracle/ddsp_oracle.py in float64 (pinned to the reference by tests/test_oracle_golden.py).
"""
import os
import re

import numpy as np
import pytest
import torch

from conftest import ROOT, load_golden

pytestmark = pytest.mark.gpu

GRAD_REL = 1e-3


@pytest.fixture(scope="module")
def ddsp():
    assert torch.cuda.is_available(), "gpu tests need a CUDA device"
    import ddsp_pytorch_b200 as pkg
    return pkg


@pytest.fixture(scope="module")
def orc():
    from oracle import ddsp_oracle
    return ddsp_oracle


def rel(got, ref):
    ref = ref.detach().double().cpu()
    return float((got.detach().double().cpu().reshape(ref.shape) - ref).norm() / ref.norm().clamp_min(1e-30))


# ------------------------------------------------------------------------------- K3 at the bench shapes
@pytest.mark.parametrize("B,N,L,sr", [(64, 64000, 16000, 16000), (4, 192000, 48000, 48000), (5, 64000, 16000, 16000),
                                       (3, 70000, 20000, 16000)])
def test_reverb_forward_and_gradients_at_bench_shapes(ddsp, orc, B, N, L, sr):
    """modules.py:21-35 + core.py:169-176 at the shapes bench.py runs (config 2: transform length 20 x 4096; config 4:
    2^18), an odd batch, and a length that takes the 24 x 4096 transform."""
    from ddsp_pytorch_b200.models.modules import Reverb
    torch.manual_seed(B)
    rv = Reverb(L, sr, initial_wet=0.4, initial_decay=3.0)
    g = torch.Generator().manual_seed(N + B)
    x = 0.1 * torch.randn(B, N, 1, generator=g)
    go = torch.randn(B, N, 1, generator=g)
    # float64 oracle
    p64 = [rv.noise.detach().double().requires_grad_(True), rv.decay.detach().double().requires_grad_(True),
           rv.wet.detach().double().requires_grad_(True)]
    x64 = x.double().requires_grad_(True)
    y64 = orc.reverb(x64, p64[0], p64[1], p64[2], rv.t.double())
    (y64 * go.double()).sum().backward()
    # kernels
    rv = rv.cuda()
    xd = x.cuda().requires_grad_(True)
    y = rv(xd)
    (y * go.cuda()).sum().backward()
    peak = float(y64.detach().abs().max())
    err = float((y.detach().double().cpu() - y64.detach()).abs().max())
    assert err <= 1e-4 * max(1.0, peak), f"audio max abs {err:.3e} (peak {peak:.2f})"
    assert rel(xd.grad, x64.grad) <= GRAD_REL
    assert rel(rv.noise.grad, p64[0].grad) <= GRAD_REL
    assert abs(float(rv.decay.grad) - float(p64[1].grad)) <= GRAD_REL * abs(float(p64[1].grad)) + 1e-6
    assert abs(float(rv.wet.grad) - float(p64[2].grad)) <= GRAD_REL * abs(float(p64[2].grad)) + 1e-6


# ------------------------------------------------------------------------------- chain gradients, tie free
@pytest.mark.parametrize("B,frames,L", [(3, 400, 16000), (2, 60, 2000)])
def test_chain_gradients_with_linear_loss(ddsp, orc, B, frames, L):
    """K0 -> K1 / K2 -> mix -> K3 as SynthStep runs them (decoder.py:110-125), loss = sum(signal * go): every gradient
    (amp_raw, dist_raw, mag_raw, reverb noise / decay / wet) within 1e-3 of the float64 oracle.  The spectral loss is
    left out on purpose: its L1 sign ties make end-to-end float32-vs-float64 gradients meaningless."""
    from ddsp_pytorch_b200.hotpath import SynthShapes, SynthStep, synthetic_inputs
    shapes = SynthShapes(batch=B, frames=frames, block_size=160, n_harmonic=100, n_bands=65, sample_rate=16000,
                         reverb_length=L)
    torch.manual_seed(1)
    step = SynthStep(shapes, "cuda")
    with torch.no_grad():
        step.reverb.wet.fill_(0.5)
        step.reverb.decay.fill_(3.5)
    host = synthetic_inputs(shapes, seed=7)
    step.load_inputs(host, non_blocking=False)
    go = torch.randn(B, shapes.samples, 1, generator=torch.Generator().manual_seed(3))
    leaves = step._leaves()
    signal = step.forward(leaves)
    grads = torch.autograd.grad((signal * go.cuda()).sum(), leaves)
    # oracle
    d = {k: v.double() for k, v in host.items()}
    rp = {k: v.detach().double().cpu() for k, v in step.reverb.state_dict().items()}
    l64 = [d["amp_raw"].requires_grad_(True), d["dist_raw"].requires_grad_(True), d["mag_raw"].requires_grad_(True),
           rp["noise"].requires_grad_(True), rp["decay"].requires_grad_(True), rp["wet"].requires_grad_(True)]
    out = orc.synth_chain(l64[0], l64[1], l64[2], d["pitch"], d["noise"], 160, 16000,
                          {"noise": l64[3], "decay": l64[4], "wet": l64[5], "t": rp["t"]})
    g64 = torch.autograd.grad((out["signal"] * go.double()).sum(), l64)
    assert float((signal.detach().double().cpu() - out["signal"].detach()).abs().max()) <= 1e-4
    names = ["amp_raw", "dist_raw", "mag_raw", "reverb.noise", "reverb.decay", "reverb.wet"]
    for name, got, ref in zip(names, grads, g64):
        e = rel(got, ref)
        assert e <= GRAD_REL, f"d {name}: relative error {e:.3e}"


# ------------------------------------------------------------------------------- a14 backward
@pytest.mark.parametrize("name,cls", [("model_decoder", "DDSPDecoder"), ("model_autoencoder", "DDSPAutoencoder")])
def test_model_backward_matches_reference(ddsp, name, cls):
    """Every parameter gradient of the whole model (control net, projections, synth, reverb) for the linear loss
    sum(signal * go), against the float64 run of the unmodified reference stored in tests/golden/<name>_grads.npz."""
    g = load_golden(name)
    gg = load_golden(name + "_grads")
    from ddsp_pytorch_b200.models import decoder, encoder
    ctor = getattr(decoder, cls, None) or getattr(encoder, cls)
    model = ctor(hidden_size=16, n_harmonic=12, n_bands=65, sample_rate=16000, block_size=160, has_reverb=True)
    sd = {k[3:]: torch.as_tensor(v) for k, v in g.items() if k.startswith("sd_")}
    assert abs(sum(float(v.double().abs().sum()) for v in sd.values()) - float(gg["sd_checksum"])) < 1e-6
    model.load_state_dict(sd)
    model.cuda()
    batch = {k[3:]: torch.as_tensor(v).cuda() for k, v in g.items() if k.startswith("in_")}
    torch.manual_seed(int(gg["noise_seed"]))
    out = model(batch)
    assert float((out["signal"].detach().double().cpu() - torch.as_tensor(gg["out_signal"])).abs().max()) <= 1e-4
    (out["signal"] * torch.as_tensor(gg["go"]).float().cuda()).sum().backward()
    params = dict(model.named_parameters())
    assert {"grad_" + k for k in params} == {k for k in gg if k.startswith("grad_")}
    worst = ("", 0.0)
    for k, p in params.items():
        ref = torch.as_tensor(gg["grad_" + k])
        assert p.grad is not None, k
        # parameters whose gradient is rounding noise relative to the model's scale are judged on that scale
        scale = max(float(ref.norm()), 1e-6 * float(torch.as_tensor(gg["go"]).norm()))
        e = float((p.grad.double().cpu() - ref).norm()) / scale
        if e > worst[1]:
            worst = (k, e)
    assert worst[1] <= GRAD_REL, f"d {worst[0]}: relative error {worst[1]:.3e}"


# ------------------------------------------------------------------------------- INTEGRATION.md section 3
def test_integration_md_ctypes_example(ddsp, orc):
    """The ctypes block of INTEGRATION.md section 3 is executed verbatim (only the library path is made absolute) and its
    output compared with the oracle, so the documented argtypes cannot drift from include/ddsp_b200.h again."""
    text = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    sec = text[text.index("## 3. C ABI from any host"):text.index("## 4. libtorch C++ host")]
    code = re.search(r"```python\n(.*?)```", sec, re.S).group(1)
    assert "[vp] * 5 + [i64, i32, i32, i32, ctypes.c_float, vp]" in code, "five pointers precede `rows` (header :115)"
    code = code.replace('"ddsp_pytorch_b200/libddsp_b200.so"', repr(os.path.join(ROOT, "ddsp_pytorch_b200", "libddsp_b200.so")))
    torch.manual_seed(0)
    ns = {}
    exec(compile(code, "INTEGRATION.md#3", "exec"), ns)
    torch.cuda.synchronize()
    mags, noise, out = ns["mags"], ns["noise"], ns["out"]
    ref = orc.filtered_noise(mags.double().cpu(), noise.double().cpu(), ns["bs"])
    assert out.shape == ref.shape
    assert float((out.double().cpu() - ref).abs().max()) <= 1e-6


# ------------------------------------------------------------------------------- device-side noise draw
def test_device_noise_option_matches_oracle_on_the_same_draw(ddsp, orc):
    """FilteredNoise.device_noise (the training harness's default: no CPU draw, no H2D copy of the noise): the module's
    output equals the oracle's modules.py:116-128 on the very tensor the device generator produced for that seed, and the
    draw is uniform(-1, 1) like the reference's."""
    from ddsp_pytorch_b200.models.modules import FilteredNoise
    B, T, bs, NB = 3, 40, 160, 65
    fn = FilteredNoise(block_size=bs, window_size=NB)
    fn.device_noise = True
    mags = torch.rand(B, T, NB, generator=torch.Generator().manual_seed(2)) * 0.5
    torch.cuda.manual_seed(77)
    y = fn(mags.cuda())
    torch.cuda.manual_seed(77)
    noise = torch.rand(B, T, bs, device="cuda", dtype=torch.float32) * 2 - 1
    ref = orc.filtered_noise(mags.double(), noise.double().cpu(), bs)
    assert y.shape == ref.shape
    assert float((y.double().cpu() - ref).abs().max()) <= 1e-6
    big = torch.rand(64, 400, bs, device="cuda") * 2 - 1
    assert abs(float(big.mean())) < 2e-3 and abs(float(big.var()) - 1.0 / 3.0) < 2e-3
    assert float(big.min()) >= -1.0 and float(big.max()) < 1.0


# ------------------------------------------------------------------------------- paths outside the fast ranges
def test_stft_sizes_outside_the_register_fft_range(ddsp, orc):
    """n_fft = 32 and 16 (below the 64..4096 range of the register FFT) take the generic shared-memory transform:
    magnitudes, their gradient, and the fused loss (per-scale launches, not the all-scales kernel) against the float64
    oracle.  A size nothing supports fails loudly."""
    g = torch.Generator().manual_seed(21)
    B, N = 2, 5000
    scales, ov = [32, 16], 0.75
    with pytest.raises(RuntimeError):
        ddsp.multiscale_fft(torch.zeros(1, 40000, device="cuda"), [16384], ov)
    tgt = 0.1 * torch.randn(B, N, generator=g)
    rec = 0.1 * torch.randn(B, N, generator=g)
    r64 = rec.double().requires_grad_(True)
    mags64 = orc.multiscale_fft(r64, scales, ov)
    r = rec.cuda().requires_grad_(True)
    mags = ddsp.multiscale_fft(r, scales, ov)
    for a, b in zip(mags, mags64):
        assert a.shape == b.shape
        assert float((a.detach().double().cpu() - b.detach()).abs().max()) <= 1e-4
    go = [torch.randn(m.shape, generator=g) for m in mags64]
    sum((m * w.double()).sum() for m, w in zip(mags64, go)).backward()
    sum((m * w.cuda()).sum() for m, w in zip(mags, go)).backward()
    assert rel(r.grad, r64.grad) <= GRAD_REL
    r64b = rec.double().requires_grad_(True)
    ref = orc.mss_loss(tgt.double(), r64b, scales, ov)
    ref.backward()
    r32 = rec.clone().requires_grad_(True)
    orc.mss_loss(tgt, r32, scales, ov).backward()
    rb = rec.cuda().requires_grad_(True)
    loss = ddsp.multiscale_spectral_loss(tgt.cuda(), rb, scales, ov)
    loss.backward()
    assert abs(float(loss) - float(ref)) <= 5e-6 * abs(float(ref))
    ref32 = float((r32.grad.double() - r64b.grad).norm() / r64b.grad.norm())
    assert rel(rb.grad, r64b.grad) <= max(GRAD_REL, 1.5 * ref32)


def test_filtered_noise_without_design_table(ddsp, orc):
    """design == NULL in the C ABI selects the first-generation kernels (cosine sums per frame): forward and backward
    through ctypes against the oracle."""
    import ctypes
    from ddsp_pytorch_b200._lib import core_library
    lib = core_library()
    vp, i64, i32 = ctypes.c_void_p, ctypes.c_int64, ctypes.c_int
    B, T, NB, bs = 2, 30, 65, 160
    g = torch.Generator().manual_seed(5)
    mags = torch.rand(B, T, NB, generator=g)
    noise = torch.rand(B, T, bs, generator=g) * 2 - 1
    go = torch.randn(B, T * bs, 1, generator=g)
    m64 = mags.double().requires_grad_(True)
    ref = orc.filtered_noise(m64, noise.double(), bs)
    (ref * go.double()).sum().backward()
    md, nd, god = mags.cuda(), noise.cuda(), go.cuda().contiguous()
    out = torch.empty(B, T * bs, 1, device="cuda")
    dm = torch.empty(B, T, NB, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    lib.ddsp_b200_filtered_noise_fwd.restype = i32
    lib.ddsp_b200_filtered_noise_fwd.argtypes = [vp] * 5 + [i64, i32, i32, i32, ctypes.c_float, vp]
    lib.ddsp_b200_filtered_noise_bwd.restype = i32
    lib.ddsp_b200_filtered_noise_bwd.argtypes = [vp] * 5 + [i64, i32, i32, i32, ctypes.c_float, vp]
    assert lib.ddsp_b200_filtered_noise_fwd(md.data_ptr(), nd.data_ptr(), None, None, out.data_ptr(), B * T, NB, bs, 0, 0.0, st) == 0
    assert lib.ddsp_b200_filtered_noise_bwd(god.data_ptr(), nd.data_ptr(), None, None, dm.data_ptr(), B * T, NB, bs, 0, 0.0, st) == 0
    torch.cuda.synchronize()
    assert float((out.double().cpu() - ref.detach()).abs().max()) <= 1e-6
    assert rel(dm, m64.grad) <= GRAD_REL
