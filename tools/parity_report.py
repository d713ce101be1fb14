"""Prints the measured deviation of every kernel from the float64 oracle (and the float32 reference
port's own deviation) at config-1 shapes, B=2.  Run on a GPU box: python tools/parity_report.py"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ddsp_pytorch_b200 as ddsp  # noqa: E402
from ddsp_pytorch_b200.hotpath import SynthShapes, synthetic_inputs  # noqa: E402
from oracle import ddsp_oracle as orc  # noqa: E402


def mx(a, b):
    return float((a.detach().double().cpu() - b.detach().double().cpu()).abs().max())


def rel(a, b):
    return float((a.double().cpu() - b.double().cpu()).norm() / b.double().cpu().norm())


def main():
    s = SynthShapes(batch=2, frames=400, block_size=160, n_harmonic=100, n_bands=65, sample_rate=16000,
                    reverb_length=16000)
    h = synthetic_inputs(s, seed=0)
    d64 = {k: v.double() for k, v in h.items()}
    dev = {k: v.cuda() for k, v in h.items()}
    rep = {}
    # controls + harmonic
    c64 = orc.harmonic_controls(d64["amp_raw"], d64["dist_raw"], d64["pitch"], s.sample_rate)
    w64 = c64["harmonic_distribution"] * c64["amplitudes"]
    w64.requires_grad_(True)
    y64 = orc.harmonic_synth(orc.upsample(d64["pitch"], s.block_size), orc.upsample(w64, s.block_size), s.sample_rate)
    go = torch.randn(y64.shape, generator=torch.Generator().manual_seed(1), dtype=torch.float64)
    (y64 * go).sum().backward()
    y32 = orc.harmonic_synth(orc.upsample(h["pitch"], s.block_size), orc.upsample(w64.detach().float(), s.block_size), s.sample_rate)
    w = w64.detach().float().cuda().requires_grad_(True)
    y, _ = ddsp.harmonic_synth_frames(dev["pitch"], w, s.block_size, s.sample_rate)
    (y * go.float().cuda()).sum().backward()
    rep["harmonic_audio_max_abs"] = mx(y, y64)
    rep["harmonic_audio_max_abs_reference_fp32"] = mx(y32, y64)
    rep["harmonic_d_weights_rel"] = rel(w.grad, w64.grad)
    # noise
    m64 = orc.noise_controls(d64["mag_raw"])["magnitudes"].requires_grad_(True)
    n64 = orc.filtered_noise(m64, d64["noise"], s.block_size)
    (n64 * go).sum().backward()
    m = m64.detach().float().cuda().requires_grad_(True)
    n = ddsp.filtered_noise(m, dev["noise"])
    (n * go.float().cuda()).sum().backward()
    rep["noise_audio_max_abs"] = mx(n, n64)
    rep["noise_d_mags_rel"] = rel(m.grad, m64.grad)
    # reverb
    from ddsp_pytorch_b200.models.modules import Reverb
    torch.manual_seed(0)
    rv = Reverb(16000, 16000, initial_wet=1.0, initial_decay=4.0)
    x64 = (y64.detach() + n64.detach()).requires_grad_(True)
    p64 = [rv.noise.detach().double().requires_grad_(True), rv.decay.detach().double().requires_grad_(True),
           rv.wet.detach().double().requires_grad_(True)]
    r64 = orc.reverb(x64, p64[0], p64[1], p64[2], rv.t.double())
    (r64 * go).sum().backward()
    rv.cuda()
    x = x64.detach().float().cuda().requires_grad_(True)
    r = rv(x)
    (r * go.float().cuda()).sum().backward()
    rep["reverb_audio_max_abs"] = mx(r, r64)
    rep["reverb_audio_peak"] = float(r64.abs().max())
    rep["reverb_d_x_rel"] = rel(x.grad, x64.grad)
    rep["reverb_d_noise_rel"] = rel(rv.noise.grad, p64[0].grad)
    rep["reverb_d_decay_rel"] = abs(float(rv.decay.grad) - float(p64[1].grad)) / abs(float(p64[1].grad))
    rep["reverb_d_wet_rel"] = abs(float(rv.wet.grad) - float(p64[2].grad)) / abs(float(p64[2].grad))
    # loss
    scales, ov = [4096, 2048, 1024, 512, 256, 128], 0.75
    rec64 = r64.detach().squeeze(-1).requires_grad_(True)
    l64 = orc.mss_loss(d64["target"], rec64, scales, ov)
    l64.backward()
    rec32 = rec64.detach().float().requires_grad_(True)
    l32 = orc.mss_loss(h["target"], rec32, scales, ov)
    l32.backward()
    rec = rec64.detach().float().cuda().requires_grad_(True)
    l = ddsp.multiscale_spectral_loss(dev["target"], rec, scales, ov)
    l.backward()
    rep["mss_loss_rel"] = abs(float(l) - float(l64)) / float(l64)
    rep["mss_loss_rel_reference_fp32"] = abs(float(l32) - float(l64)) / float(l64)
    rep["mss_d_rec_rel"] = rel(rec.grad, rec64.grad)
    rep["mss_d_rec_rel_reference_fp32"] = rel(rec32.grad, rec64.grad)
    mags = ddsp.multiscale_fft(rec.detach(), scales, ov)
    mags64 = orc.multiscale_fft(rec64.detach(), scales, ov)
    rep["stft_mag_max_abs"] = max(mx(a, b) for a, b in zip(mags, mags64))
    # control net: float64 torch.nn layers are the oracle; torch's own float32 layers beside it
    from ddsp_pytorch_b200 import core
    torch.manual_seed(3)
    torch.backends.cuda.matmul.allow_tf32 = False
    B, T, H = 16, 400, 512
    x = torch.randn(B, T, H)
    go2 = torch.randn(B, T, H)
    blk = core.mlp(H, H, 1).cuda()
    gru = core.gru(1, H).cuda()
    ref_blk = torch.nn.Sequential(torch.nn.Linear(H, H), torch.nn.LayerNorm(H), torch.nn.LeakyReLU())
    ref_blk.load_state_dict({k: v.detach().cpu() for k, v in blk.state_dict().items()})
    ref_gru = torch.nn.GRU(H, H, batch_first=True)
    ref_gru.load_state_dict({k: v.detach().cpu() for k, v in gru.state_dict().items()})

    def run(b, g, xx, gg):
        xx = xx.clone().requires_grad_(True)
        b.zero_grad(set_to_none=True)
        g.zero_grad(set_to_none=True)
        mid = b(xx)
        out = g(mid)[0]
        (out * gg).sum().backward()
        return mid.detach(), out.detach(), xx.grad, [p.grad.detach().double().cpu() for p in list(b.parameters()) + list(g.parameters())]

    m64_, o64, gx64, gp64 = run(ref_blk.double(), ref_gru.double(), x.double(), go2.double())
    mk, ok, gxk, gpk = run(blk, gru, x.cuda(), go2.cuda())
    ref_blk.float().cuda(), ref_gru.float().cuda()
    torch.backends.cudnn.allow_tf32 = False
    mt, ot, gxt, gpt = run(ref_blk, ref_gru, x.cuda(), go2.cuda())
    torch.backends.cudnn.allow_tf32 = True
    _, ot32, _, _ = run(ref_blk, ref_gru, x.cuda(), go2.cuda())
    rep["linear_ln_lrelu_max_abs"] = mx(mk, m64_)
    rep["linear_ln_lrelu_max_abs_torch_fp32"] = mx(mt, m64_)
    rep["gru_output_max_abs"] = mx(ok, o64)
    rep["gru_output_max_abs_cudnn_fp32"] = mx(ot, o64)
    rep["gru_output_max_abs_cudnn_default_tf32"] = mx(ot32, o64)
    rep["control_net_d_input_rel"] = rel(gxk, gx64)
    rep["control_net_d_input_rel_torch_fp32"] = rel(gxt, gx64)
    rep["control_net_d_params_rel_max"] = max(rel(a, b) for a, b in zip(gpk, gp64))
    rep["control_net_d_params_rel_max_torch_fp32"] = max(rel(a, b) for a, b in zip(gpt, gp64))
    print(json.dumps(rep, indent=1))


if __name__ == "__main__":
    main()
