"""Generate tests/golden/*.npz by running the UNMODIFIED reference (``/root/reference``).

TEST INFRASTRUCTURE.  Runs only in the build container (the GPU box has no
``/root/reference``); the fixtures it writes are committed so the parity tests travel.
The reference's top-level imports of librosa / crepe / matplotlib / pytorch_lightning
are satisfied with empty stub modules (none is touched by the hot path).

    PYTHONDONTWRITEBYTECODE=1 python oracle/make_golden.py

Every fixture holds the seeded inputs, the reference outputs evaluated in float64,
the reference float32 outputs' deviation from float64 (the "no worse than" clause) and
per-op gradients for a fixed grad_output.
"""
from __future__ import annotations

import os
import sys
import types

import numpy as np
import torch

REF = "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")


def import_reference():
    for name in ["librosa", "crepe", "matplotlib", "matplotlib.pyplot", "pytorch_lightning"]:
        sys.modules[name] = types.ModuleType(name)
    sys.modules["pytorch_lightning"].LightningDataModule = object
    sys.path.insert(0, REF)
    for name in list(sys.modules):
        if name == "ddsp" or name.startswith("ddsp."):
            del sys.modules[name]
    import ddsp  # noqa: the reference package
    assert ddsp.__file__.startswith(REF), ddsp.__file__
    return ddsp


def npz(name, **arrays):
    path = os.path.join(OUT, name + ".npz")
    conv = {}
    for k, v in arrays.items():
        if isinstance(v, torch.Tensor):
            v = v.detach().cpu().numpy()
        conv[k] = np.asarray(v)
    np.savez_compressed(path, **conv)
    print(f"{name}: {os.path.getsize(path) / 1024:.1f} KiB")


def synth_case(ddsp, name, B, T, bs, H, NB, sr, seed, f0_lo=80.0, f0_hi=780.0):
    from ddsp.models.modules import HarmonicSynth, FilteredNoise
    g = torch.Generator().manual_seed(seed)
    amp_raw = torch.randn(B, T, 1, generator=g, dtype=torch.float64)
    dist_raw = torch.randn(B, T, H, generator=g, dtype=torch.float64)
    mag_raw = torch.randn(B, T, NB, generator=g, dtype=torch.float64)
    # f0 is float32-representable so that fp32 kernels see exactly the oracle's input
    f0 = (torch.rand(B, T, 1, generator=g) * (f0_hi - f0_lo) + f0_lo).double()
    noise = (torch.rand(B, T, bs, generator=g) * 2 - 1).double()
    g_harm = torch.randn(B, T * bs, 1, generator=g, dtype=torch.float64)
    g_noise = torch.randn(B, T * bs, 1, generator=g, dtype=torch.float64)

    hs = HarmonicSynth(bs, sr)
    fn = FilteredNoise(bs, NB)
    res = {}
    for dt in (torch.float64, torch.float32):
        a = amp_raw.detach().clone().to(dt).requires_grad_(True)
        d = dist_raw.detach().clone().to(dt).requires_grad_(True)
        m = mag_raw.detach().clone().to(dt).requires_grad_(True)
        f = f0.detach().clone().to(dt).requires_grad_(True)
        ctrl = hs.get_controls(a, d, f)
        amps, dist = ctrl["amplitudes"], ctrl["harmonic_distribution"]
        dist_snapshot = dist.detach().clone()
        harm = hs(amps, dist.clone(), f)          # forward mutates its 2nd argument in place
        nctrl = fn.get_controls(m)
        ir = ddsp.amp_to_impulse_response(nctrl["magnitudes"], bs)
        # FilteredNoise.forward draws its own noise; replay its body with the fixed tensor
        nz = ddsp.fft_convolve(noise.to(dt), ir).contiguous().reshape(B, -1, 1)
        (harm * g_harm.to(dt)).sum().backward()
        (nz * g_noise.to(dt)).sum().backward()
        res[dt] = dict(amps=amps.detach(), dist=dist_snapshot, harm=harm.detach(),
                       mags=nctrl["magnitudes"].detach(), ir=ir.detach(), nz=nz.detach(), d_amp_raw=a.grad, d_dist_raw=d.grad, d_mag_raw=m.grad, d_f0=f.grad)
    r64, r32 = res[torch.float64], res[torch.float32]
    dev = {("dev32_" + k): float((r32[k].double() - r64[k]).abs().max()) for k in r64}
    npz(name, B=B, T=T, bs=bs, H=H, NB=NB, sr=sr,
        amp_raw=amp_raw, dist_raw=dist_raw, mag_raw=mag_raw, f0=f0, noise=noise,
        g_harm=g_harm, g_noise=g_noise, **r64, **dev)


def audio_rate_case(ddsp, name, B, N, H, sr, seed):
    g = torch.Generator().manual_seed(seed)
    f0 = (torch.rand(B, N, 1, generator=g) * 900 + 40).double().requires_grad_(True)
    amps = torch.rand(B, N, H, generator=g, dtype=torch.float64).requires_grad_(True)
    go = torch.randn(B, N, 1, generator=g, dtype=torch.float64)
    y = ddsp.harmonic_synth(f0, amps, sr)
    (y * go).sum().backward()
    y32 = ddsp.harmonic_synth(f0.detach().float(), amps.detach().float(), sr)
    npz(name, sr=sr, f0=f0, amps=amps, go=go, y=y, d_f0=f0.grad, d_amps=amps.grad,
        dev32_y=float((y32.double() - y).abs().max()))


def fftconv_case(ddsp, name, seed):
    g = torch.Generator().manual_seed(seed)
    out = {}
    for tag, ss, ks in [("a", (2, 3, 50), (2, 3, 50)), ("b", (3, 129), (1, 129)), ("c", (1, 1), (1, 1))]:
        s = torch.randn(*ss, generator=g, dtype=torch.float64).requires_grad_(True)
        k = torch.randn(*ks, generator=g, dtype=torch.float64).requires_grad_(True)
        y = ddsp.fft_convolve(s, k)
        go = torch.randn(y.shape, generator=g, dtype=torch.float64)
        (y * go).sum().backward()
        out.update({f"{tag}_s": s, f"{tag}_k": k, f"{tag}_y": y, f"{tag}_go": go,
                    f"{tag}_ds": s.grad, f"{tag}_dk": k.grad})
    npz(name, **out)


def ir_case(ddsp, name, seed):
    g = torch.Generator().manual_seed(seed)
    out = {}
    for tag, nb, ts in [("a", 65, 160), ("b", 65, 512), ("c", 9, 16), ("d", 33, 64)]:
        amp = torch.rand(2, 3, nb, generator=g, dtype=torch.float64).requires_grad_(True)
        ir = ddsp.amp_to_impulse_response(amp, ts)
        go = torch.randn(ir.shape, generator=g, dtype=torch.float64)
        (ir * go).sum().backward()
        out.update({f"{tag}_amp": amp, f"{tag}_ir": ir, f"{tag}_go": go, f"{tag}_damp": amp.grad,
                    f"{tag}_ts": ts})
    npz(name, **out)


def reverb_case(ddsp, name, L, sr, B, N, seed, decay=5.0, wet=0.0):
    from ddsp.models.modules import Reverb
    torch.manual_seed(seed)
    rv = Reverb(L, sr, initial_wet=wet, initial_decay=decay).double()
    g = torch.Generator().manual_seed(seed + 1)
    x = torch.randn(B, N, 1, generator=g, dtype=torch.float64).requires_grad_(True)
    go = torch.randn(B, N, 1, generator=g, dtype=torch.float64)
    y = rv(x)
    (y * go).sum().backward()
    rv32 = Reverb(L, sr, initial_wet=wet, initial_decay=decay)
    rv32.load_state_dict({k: v.float() for k, v in rv.state_dict().items()})
    y32 = rv32(x.detach().float())
    npz(name, L=L, sr=sr, x=x, go=go, noise=rv.noise, decay=rv.decay, wet=rv.wet, t=rv.t,
        impulse=rv.build_impulse(), y=y, d_x=x.grad, d_noise=rv.noise.grad,
        d_decay=rv.decay.grad, d_wet=rv.wet.grad,
        dev32_y=float((y32.double() - y).abs().max()))


def mss_case(ddsp, name, B, N, scales, overlap, seed):
    sys.path.insert(0, REF)
    g = torch.Generator().manual_seed(seed)
    tgt = 0.1 * torch.randn(B, N, generator=g, dtype=torch.float64)
    rec = (0.1 * torch.randn(B, N, generator=g, dtype=torch.float64)).requires_grad_(True)
    st = ddsp.multiscale_fft(tgt, scales, overlap)
    sr_ = ddsp.multiscale_fft(rec, scales, overlap)
    # train.py:70-76 (train.py is a script with side effects; its loss body is replayed here)
    loss = 0
    for sx, sy in zip(st, sr_):
        loss = loss + (sx - sy).abs().mean() + (ddsp.safe_log(sx) - ddsp.safe_log(sy)).abs().mean()
    loss.backward()
    # per-scale magnitude gradient for a fixed grad_output (stft_mag backward fixture)
    rec2 = rec.detach().clone().requires_grad_(True)
    mags = ddsp.multiscale_fft(rec2, scales, overlap)
    gos = [torch.randn(m.shape, generator=g).double() for m in mags]   # float32-representable
    per_scale = []
    for m, go in zip(mags, gos):
        (gr,) = torch.autograd.grad((m * go).sum(), rec2, retain_graph=True)
        per_scale.append(gr)
    arrays = dict(scales=np.array(scales), overlap=overlap, tgt=tgt, rec=rec, loss=loss,
                  d_rec=rec.grad)
    for i, s in enumerate(scales):
        arrays[f"mag_tgt_{s}"] = st[i].float()
        arrays[f"mag_rec_{s}"] = sr_[i].float()
        arrays[f"go_{s}"] = gos[i].float()
        arrays[f"d_rec_{s}"] = per_scale[i]
    # the reference's own float32 deviation of the loss gradient (sign ties)
    rec32 = rec.detach().float().requires_grad_(True)
    l32 = 0
    for sx, sy in zip(ddsp.multiscale_fft(tgt.float(), scales, overlap),
                      ddsp.multiscale_fft(rec32, scales, overlap)):
        l32 = l32 + (sx - sy).abs().mean() + (ddsp.safe_log(sx) - ddsp.safe_log(sy)).abs().mean()
    l32.backward()
    arrays["dev32_loss"] = float(abs(l32.item() - loss.item()))
    arrays["dev32_d_rec_rel"] = float((rec32.grad.double() - rec.grad).norm() / rec.grad.norm())
    npz(name, **arrays)


def model_case(ddsp, name, cls, seed, with_mfcc):
    kw = dict(hidden_size=16, n_harmonic=12, n_bands=65, sample_rate=16000, block_size=160,
              has_reverb=True)
    torch.manual_seed(seed)
    model = cls(**kw)
    with torch.no_grad():
        model.reverb.wet.fill_(0.3)
    B, T = 2, 5
    g = torch.Generator().manual_seed(seed + 1)
    batch = {"pitch": torch.rand(B, T, 1, generator=g) * 500 + 100,
             "loudness": torch.randn(B, T, 1, generator=g)}
    if with_mfcc:
        batch["mfcc"] = torch.randn(B, T, 30, generator=g)
    torch.manual_seed(seed + 2)                      # FilteredNoise draws from the default generator
    out = model({k: v.clone() for k, v in batch.items()})
    arrays = {("in_" + k): v for k, v in batch.items()}
    arrays.update({("sd_" + k): v for k, v in model.state_dict().items()})
    for k in ["signal", "noise", "harmonic_audio"] + (["z"] if with_mfcc else []):
        arrays["out_" + k] = out[k]
    arrays["out_harmonic_distribution"] = out["harmonic_ctrls"]["harmonic_distribution"]
    arrays["out_amplitudes"] = out["harmonic_ctrls"]["amplitudes"]
    arrays["out_magnitudes"] = out["noise_ctrls"]["magnitudes"]
    arrays["noise_seed"] = seed + 2
    npz(name, **arrays)


def model_grad_case(ddsp, name, cls, seed, with_mfcc):
    """Parameter gradients of the whole model (control net + synth + reverb) for a LINEAR loss sum(signal * go):
    no L1 sign ties, so float32 kernels can be held to 1e-3 against this float64 run of the unmodified reference.
    Same seeds as model_case, so the weights equal the ``sd_*`` entries of that fixture (checksum stored)."""
    kw = dict(hidden_size=16, n_harmonic=12, n_bands=65, sample_rate=16000, block_size=160,
              has_reverb=True)
    torch.manual_seed(seed)
    model = cls(**kw)
    with torch.no_grad():
        model.reverb.wet.fill_(0.3)
    checksum = float(sum(v.double().abs().sum() for v in model.state_dict().values()))
    model = model.double()
    B, T = 2, 5
    g = torch.Generator().manual_seed(seed + 1)
    batch = {"pitch": (torch.rand(B, T, 1, generator=g) * 500 + 100).double(),
             "loudness": torch.randn(B, T, 1, generator=g).double()}
    if with_mfcc:
        batch["mfcc"] = torch.randn(B, T, 30, generator=g).double()
    torch.manual_seed(seed + 2)                      # FilteredNoise draws float32 uniforms from the default generator
    out = model({k: v.clone() for k, v in batch.items()})
    go = torch.randn(out["signal"].shape, generator=torch.Generator().manual_seed(seed + 3), dtype=torch.float64)
    (out["signal"] * go).sum().backward()
    arrays = {"go": go, "out_signal": out["signal"], "sd_checksum": checksum, "noise_seed": seed + 2}
    for k, p_ in model.named_parameters():
        arrays["grad_" + k] = p_.grad
    npz(name, **arrays)


def main():
    os.makedirs(OUT, exist_ok=True)
    ddsp = import_reference()
    torch.set_num_threads(1)
    if "--only-model-grads" in sys.argv:             # added in round 2; leaves the other fixtures untouched
        from ddsp.models.decoder import DDSPDecoder
        from ddsp.models.encoder import DDSPAutoencoder
        model_grad_case(ddsp, "model_decoder_grads", DDSPDecoder, seed=11, with_mfcc=False)
        model_grad_case(ddsp, "model_autoencoder_grads", DDSPAutoencoder, seed=12, with_mfcc=True)
        return
    synth_case(ddsp, "synth_c1_small", B=2, T=12, bs=160, H=20, NB=65, sr=16000, seed=1)
    synth_case(ddsp, "synth_c3_buffer", B=1, T=2, bs=512, H=64, NB=65, sr=48000, seed=2)
    synth_case(ddsp, "synth_h100", B=1, T=7, bs=160, H=100, NB=65, sr=16000, seed=3,
               f0_lo=60.0, f0_hi=1200.0)
    audio_rate_case(ddsp, "harmonic_audio_rate", B=2, N=700, H=9, sr=16000, seed=4)
    fftconv_case(ddsp, "fft_convolve", seed=5)
    ir_case(ddsp, "impulse_response", seed=6)
    reverb_case(ddsp, "reverb_pad", L=1000, sr=4000, B=3, N=1600, seed=7, decay=4.0, wet=0.5)
    reverb_case(ddsp, "reverb_crop", L=1000, sr=4000, B=2, N=600, seed=8, decay=5.0, wet=0.0)
    mss_case(ddsp, "mss_full_scales", B=2, N=2400, scales=[4096, 2048, 1024, 512, 256, 128],
             overlap=0.75, seed=9)
    mss_case(ddsp, "mss_small", B=3, N=1000, scales=[256, 128, 64], overlap=0.75, seed=10)
    from ddsp.models.decoder import DDSPDecoder
    from ddsp.models.encoder import DDSPAutoencoder
    model_case(ddsp, "model_decoder", DDSPDecoder, seed=11, with_mfcc=False)
    model_case(ddsp, "model_autoencoder", DDSPAutoencoder, seed=12, with_mfcc=True)
    model_grad_case(ddsp, "model_decoder_grads", DDSPDecoder, seed=11, with_mfcc=False)
    model_grad_case(ddsp, "model_autoencoder_grads", DDSPAutoencoder, seed=12, with_mfcc=True)


if __name__ == "__main__":
    main()
