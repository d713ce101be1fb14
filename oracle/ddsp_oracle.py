"""CPU oracle for the DDSP synthesis hot path  --  TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this module.  The product package
(``ddsp_pytorch_b200``) never imports it and has no CPU fallback.

What it is
----------
A restatement of the algorithm of hugofloresgarcia/ddsp_pytorch's hot path
(``ddsp/core.py``, ``ddsp/models/modules.py``, ``train.py:70-76``) as plain
dtype-generic torch CPU code.  The reference's arithmetic lives in a third-party
dependency, **torch** (``setup.py:18`` asks for ``torch>=1.7.0``, unpinned; this
image has torch 2.11.0): ATen ``cumsum``/``sin``/``upsample_nearest1d``,
``torch.fft`` (pocketfft/MKL on CPU) and ``torch.stft``.  Each function below cites
the reference call site it follows.  Evaluated in float64 it is the parity oracle;
evaluated in float32 on all host threads it is the "port" CPU baseline.

Pinning
-------
The reference ships no tests, golden vectors or fixtures, so
this oracle is pinned against outputs of the reference itself: ``oracle/make_golden.py``
imports the unmodified reference from ``/root/reference`` (build container only),
runs every function on seeded inputs in float64 and stores inputs + outputs +
gradients in ``tests/golden/*.npz``.  ``tests/test_oracle_golden.py`` checks this file
against those fixtures.  ``oracle/closed_form.py`` is a second, structurally
independent numpy restatement (closed-form phase, direct convolution, explicit DFT).
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence

import torch
import torch.nn.functional as F

LN10 = math.log(10.0)


# --------------------------------------------------------------------------- a1
def scale_function(x: torch.Tensor) -> torch.Tensor:
    """ddsp/core.py:77-78  --  2*sigmoid(x)**ln(10) + 1e-7."""
    return 2.0 * torch.sigmoid(x) ** LN10 + 1e-7


# --------------------------------------------------------------------------- a2
def remove_above_nyquist(amplitudes, f0, sample_rate):
    """ddsp/core.py:70-74  --  amp * ((f0*k < sr/2) + 1e-4), k = 1..H.

    The mask is cast with ``.float()`` in the reference, i.e. it is a float32 tensor
    even for float64 inputs; 1 + 1e-4 is therefore rounded to float32 first.
    """
    n_harm = amplitudes.shape[-1]
    k = torch.arange(1, n_harm + 1).to(f0)
    mask = (f0 * k < sample_rate / 2).float() + 1e-4
    return amplitudes * mask


# --------------------------------------------------------------------------- a3
def harmonic_controls(amp_raw, dist_raw, f0, sample_rate) -> Dict[str, torch.Tensor]:
    """ddsp/models/modules.py:44-67 (HarmonicSynth.get_controls)."""
    amp = scale_function(amp_raw)
    dist = remove_above_nyquist(scale_function(dist_raw), f0, sample_rate)
    dist = dist / dist.sum(-1, keepdim=True)
    return {"f0": f0, "harmonic_distribution": dist, "amplitudes": amp}


# --------------------------------------------------------------------------- a4
def upsample(signal, factor: int):
    """ddsp/core.py:64-67  --  nearest-neighbour hold of every frame for ``factor`` samples."""
    x = signal.permute(0, 2, 1)
    x = F.interpolate(x, size=x.shape[-1] * factor)
    return x.permute(0, 2, 1)


# --------------------------------------------------------------------------- a5
def harmonic_synth(f0, amplitudes, sample_rate):
    """ddsp/core.py:136-141  --  audio-rate oscillator bank.

    f0 (B,N,1), amplitudes (B,N,H) -> (B,N,1); phase is the inclusive cumsum of
    2*pi*f0/sr, so the very first sample already carries one phase increment.
    """
    n_harm = amplitudes.shape[-1]
    omega = torch.cumsum(2 * math.pi * f0 / sample_rate, 1)
    k = torch.arange(1, n_harm + 1).to(omega)
    return (torch.sin(omega * k) * amplitudes).sum(-1, keepdim=True)


# --------------------------------------------------------------------------- a6
def harmonic_synth_frames(amplitudes, harmonic_distribution, f0, block_size: int, sample_rate):
    """ddsp/models/modules.py:69-80 (HarmonicSynth.forward), without the in-place write."""
    weights = harmonic_distribution * amplitudes
    return harmonic_synth(upsample(f0, block_size), upsample(weights, block_size), sample_rate)


# --------------------------------------------------------------------------- a7
def amp_to_impulse_response(amp, target_size: int):
    """ddsp/core.py:144-166  --  zero-phase windowed FIR from band magnitudes."""
    spec = torch.complex(amp, torch.zeros_like(amp))
    ir = torch.fft.irfft(spec)
    size = ir.shape[-1]
    ir = torch.roll(ir, size // 2, -1)
    ir = ir * torch.hann_window(size, dtype=ir.dtype, device=ir.device)
    ir = F.pad(ir, (0, int(target_size) - int(size)))
    return torch.roll(ir, -size // 2, -1)


# --------------------------------------------------------------------------- a9
def fft_convolve(signal, kernel):
    """ddsp/core.py:169-176  --  causal linear convolution truncated to the signal length."""
    n = signal.shape[-1]
    s = F.pad(signal, (0, n))
    k = F.pad(kernel, (kernel.shape[-1], 0))
    out = torch.fft.irfft(torch.fft.rfft(s) * torch.fft.rfft(k))
    return out[..., out.shape[-1] // 2:]


# --------------------------------------------------------------------------- a8
def noise_controls(mag_raw, initial_bias: float = -5.0):
    """ddsp/models/modules.py:111-114 (FilteredNoise.get_controls)."""
    return {"magnitudes": scale_function(mag_raw + initial_bias)}


def filtered_noise(magnitudes, noise, block_size: int):
    """ddsp/models/modules.py:116-128 (FilteredNoise.forward) with the uniform(-1,1)
    noise tensor (B,T,block) passed in instead of drawn."""
    ir = amp_to_impulse_response(magnitudes, block_size)
    out = fft_convolve(noise.to(ir), ir).contiguous()
    return out.reshape(out.shape[0], -1, 1)


def draw_noise(batch: int, frames: int, block_size: int) -> torch.Tensor:
    """The draw FilteredNoise.forward does (modules.py:119-123): CPU default generator."""
    return torch.rand(batch, frames, block_size) * 2 - 1


# --------------------------------------------------------------------------- a10
def reverb_impulse(noise_param, decay, wet, t):
    """ddsp/models/modules.py:21-26 (Reverb.build_impulse).  t is (1,L,1) = arange(L)/sr."""
    env = torch.exp(-F.softplus(-decay) * t * 500)
    ir = noise_param * env * torch.sigmoid(wet)
    ir = ir.clone()
    ir[:, 0] = 1
    return ir


def reverb(x, noise_param, decay, wet, t):
    """ddsp/models/modules.py:28-35 (Reverb.forward).  x (B,N,1) -> (B,N,1)."""
    n = x.shape[1]
    length = noise_param.shape[0]
    ir = reverb_impulse(noise_param, decay, wet, t)
    ir = F.pad(ir, (0, 0, 0, n - length))
    return fft_convolve(x.squeeze(-1), ir.squeeze(-1)).unsqueeze(-1)


# --------------------------------------------------------------------------- a11
def multiscale_fft(signal, scales: Sequence[int], overlap: float) -> List[torch.Tensor]:
    """ddsp/core.py:27-41  --  list of |STFT| (B, s/2+1, 1+N//hop), hann, centred, normalised."""
    out = []
    for s in scales:
        spec = torch.stft(
            signal, s, int(s * (1 - overlap)), s,
            torch.hann_window(s).to(signal), True,
            normalized=True, return_complex=True,
        )
        out.append(spec.abs())
    return out


# --------------------------------------------------------------------------- a12
def safe_log(x):
    """ddsp/core.py:10-11."""
    return torch.log(x + 1e-7)


def multiscale_spec_loss(ori_stft: Sequence[torch.Tensor], rec_stft: Sequence[torch.Tensor]):
    """train.py:70-76  --  sum over scales of mean|Sx-Sy| + mean|log Sx - log Sy|."""
    total = 0
    for sx, sy in zip(ori_stft, rec_stft):
        total = total + (sx - sy).abs().mean() + (safe_log(sx) - safe_log(sy)).abs().mean()
    return total


def mss_loss(target, rec, scales, overlap):
    """train.py:92-103  --  the loss as _main_step computes it from two (B,N) signals."""
    return multiscale_spec_loss(multiscale_fft(target, scales, overlap),
                                multiscale_fft(rec, scales, overlap))


# --------------------------------------------------------------------------- a14 (synth part)
def synth_chain(amp_raw, dist_raw, mag_raw, f0, noise, block_size, sample_rate,
                reverb_params: Optional[dict] = None) -> Dict[str, torch.Tensor]:
    """decoder.py:106-125: controls -> harmonic + filtered noise (+ reverb)."""
    hc = harmonic_controls(amp_raw, dist_raw, f0, sample_rate)
    harm = harmonic_synth_frames(hc["amplitudes"], hc["harmonic_distribution"], f0,
                                 block_size, sample_rate)
    nz = filtered_noise(noise_controls(mag_raw)["magnitudes"], noise, block_size)
    sig = harm + nz
    if reverb_params is not None:
        sig = reverb(sig, reverb_params["noise"], reverb_params["decay"],
                     reverb_params["wet"], reverb_params["t"])
    return {"signal": sig, "harmonic_audio": harm, "noise": nz}


def synth_train_step(amp_raw, dist_raw, mag_raw, f0, noise, target, block_size, sample_rate,
                     reverb_params, scales, overlap):
    """One pass of the hot path as train.py:84-129 runs it (without the control net):
    forward synth, multi-scale spectral loss, backward to the synth parameters."""
    leaves = [amp_raw, dist_raw, mag_raw, reverb_params["noise"], reverb_params["decay"],
              reverb_params["wet"]]
    leaves = [x.detach().requires_grad_(True) for x in leaves]
    rp = {"noise": leaves[3], "decay": leaves[4], "wet": leaves[5], "t": reverb_params["t"]}
    out = synth_chain(leaves[0], leaves[1], leaves[2], f0, noise, block_size, sample_rate, rp)
    loss = mss_loss(target, out["signal"].squeeze(-1), scales, overlap)
    grads = torch.autograd.grad(loss, leaves)
    return loss.detach(), grads
