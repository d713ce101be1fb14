"""Second, structurally independent CPU restatement (numpy float64)  --  TEST INFRASTRUCTURE.

Where ``oracle/ddsp_oracle.py`` follows the reference call by call (cumsum, FFTs,
torch.stft), this file writes the same maths in closed form: phase as a frame-rate
prefix sum plus a ramp, convolutions as explicit sums, the FIR design as an explicit
cosine series and the STFT as reflect-pad + frame + DFT.  Agreement of the two (and of
both with the golden fixtures produced by the real reference) is what pins the oracle.
These are also the formulas the CUDA kernels implement.
Only ``tests/`` may import this module.
"""
from __future__ import annotations

import math

import numpy as np

F32_ONE_PLUS_EPS4 = float(np.float32(1.0) + np.float32(1e-4))   # (mask.float() + 1e-4), in band
F32_EPS4 = float(np.float32(0.0) + np.float32(1e-4))             # same, above Nyquist


def scale_function(x):
    """core.py:77-78."""
    return 2.0 * (1.0 / (1.0 + np.exp(-x))) ** math.log(10.0) + 1e-7


def harmonic_controls(amp_raw, dist_raw, f0, sample_rate):
    """modules.py:44-67: scale, Nyquist mask (float32 constants), normalise."""
    h = dist_raw.shape[-1]
    k = np.arange(1, h + 1, dtype=f0.dtype)
    mask = np.where(f0 * k < sample_rate / 2, F32_ONE_PLUS_EPS4, F32_EPS4)
    d = scale_function(dist_raw) * mask
    return scale_function(amp_raw), d / d.sum(-1, keepdims=True)


def harmonic_frames(weights, f0, block_size, sample_rate):
    """modules.py:69-80 + core.py:136-141 in closed form:
    sample n = t*bs + j has phase 2*pi*(bs*sum_{t'<t} f0[t'] + (j+1)*f0[t])/sr."""
    b, t, h = weights.shape
    turns = f0[..., 0] / sample_rate                                  # (B,T) cycles per sample
    start = block_size * (np.cumsum(turns, axis=1) - turns)           # exclusive prefix
    j = np.arange(1, block_size + 1)
    phase = start[:, :, None] + turns[:, :, None] * j                  # (B,T,bs) in turns
    phase = phase - np.floor(phase)
    k = np.arange(1, h + 1)
    s = np.sin(2 * np.pi * phase[..., None] * k)                       # (B,T,bs,H)
    out = (s * weights[:, :, None, :]).sum(-1)
    return out.reshape(b, t * block_size, 1)


def harmonic_audio_rate(f0, amps, sample_rate):
    """core.py:136-141 with the phase kept in turns (mod 1) instead of radians."""
    turns = np.cumsum(f0[..., 0] / sample_rate, axis=1)
    turns = turns - np.floor(turns)
    k = np.arange(1, amps.shape[-1] + 1)
    return (np.sin(2 * np.pi * turns[..., None] * k) * amps).sum(-1, keepdims=True)


def fir_taps(mags):
    """core.py:144-166 as a cosine series.  Returns (causal[..., 0:F/2], far[..., 0:F/2]) with
    F = 2*(NB-1): ``causal[d]`` is IR index d, ``far[e]`` is IR index target-F/2+e (far[0]==0)."""
    nb = mags.shape[-1]
    fsz = 2 * (nb - 1)
    half = fsz // 2
    n = np.arange(fsz)
    k = np.arange(1, nb - 1)
    full = (mags[..., :1] + mags[..., -1:] * np.cos(np.pi * n)
            + 2 * (mags[..., 1:-1, None] * np.cos(2 * np.pi * k[:, None] * n / fsz)).sum(-2)) / fsz
    win = 0.5 - 0.5 * np.cos(2 * np.pi * n / fsz)
    causal = full[..., :half] * win[half:]
    far = full[..., half:] * win[:half]
    return causal, far


def impulse_response(mags, target_size):
    """core.py:144-166 assembled from ``fir_taps`` (target_size >= filter size)."""
    causal, far = fir_taps(mags)
    half = causal.shape[-1]
    ir = np.zeros(mags.shape[:-1] + (target_size,))
    ir[..., :half] = causal
    ir[..., target_size - half:] += far
    return ir


def causal_conv(signal, kernel):
    """core.py:169-176 as the explicit truncated sum out[i] = sum_{j<=i} s[j] k[i-j]."""
    n = signal.shape[-1]
    out = np.zeros(np.broadcast_shapes(signal.shape, kernel.shape))
    for d in range(n):
        out[..., d:] += kernel[..., d:d + 1] * signal[..., :n - d]
    return out


def filtered_noise(mags, noise, block_size):
    """modules.py:116-128."""
    y = causal_conv(noise, impulse_response(mags, block_size))
    return y.reshape(y.shape[0], -1, 1)


def reverb_impulse(noise_param, decay, wet, t):
    """modules.py:21-26; noise_param (L,), t (L,) = the module's ``t`` buffer (a float32
    arange(L)/sr, modules.py:17-19, so it is passed in rather than recomputed); returns (L,)."""
    softplus = np.log1p(np.exp(-decay))
    ir = noise_param * np.exp(-softplus * t * 500) / (1.0 + np.exp(-wet))
    ir[0] = 1.0
    return ir


def reverb(x, noise_param, decay, wet, t):
    """modules.py:28-35; x (B,N) -> (B,N).  IR is zero padded or cropped to N."""
    n = x.shape[-1]
    ir = reverb_impulse(noise_param, decay, wet, t)
    full = np.zeros(n)
    m = min(n, ir.shape[0])
    full[:m] = ir[:m]
    return np.stack([np.convolve(row, full)[:n] for row in x])


def hann_f32(n_fft):
    """core.py:35 builds ``torch.hann_window(s)`` in float32 on the CPU and only then casts it
    ``.to(signal)``: the window carries float32 rounding even in a float64 evaluation."""
    import torch
    return torch.hann_window(n_fft).double().numpy()


def stft_mag(signal, n_fft, hop, window=None):
    """core.py:27-41 for one scale: reflect pad n_fft/2, periodic hann, rfft, / sqrt(n_fft), abs.
    signal (B,N) -> (B, n_fft/2+1, 1+N//hop)."""
    pad = n_fft // 2
    x = np.pad(signal, ((0, 0), (pad, pad)), mode="reflect")
    frames = 1 + signal.shape[-1] // hop
    idx = np.arange(frames)[:, None] * hop + np.arange(n_fft)[None, :]
    win = hann_f32(n_fft) if window is None else window
    assert np.abs(win - (0.5 - 0.5 * np.cos(2 * np.pi * np.arange(n_fft) / n_fft))).max() < 1e-6
    spec = np.fft.rfft(x[:, idx] * win, axis=-1) / math.sqrt(n_fft)
    return np.abs(spec).transpose(0, 2, 1)


def mss_loss(target, rec, scales, overlap):
    """train.py:70-76 over core.py:27-41."""
    total = 0.0
    for s in scales:
        hop = int(s * (1 - overlap))
        a, b = stft_mag(target, s, hop), stft_mag(rec, s, hop)
        total += np.abs(a - b).mean() + np.abs(np.log(a + 1e-7) - np.log(b + 1e-7)).mean()
    return total
