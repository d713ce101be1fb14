"""Synth modules with the reference's classes, constructor arguments, method names and
state_dict keys (ddsp/models/modules.py), on the fused sm_100a kernels.

Only the ``.plot`` helpers (matplotlib) are left out.
"""
from __future__ import annotations

from typing import Dict, Optional

import torch
import torch.nn as nn

from .. import core
from .. import functions as F_


class Reverb(nn.Module):
    """ddsp/models/modules.py:7-35.  Parameters ``noise`` (L,1), ``decay``, ``wet``; buffer ``t``."""

    def __init__(self, length, sample_rate, initial_wet=0, initial_decay=5):
        super().__init__()
        self.length = length
        self.sample_rate = sample_rate
        self.noise = nn.Parameter((torch.rand(length) * 2 - 1).unsqueeze(-1))
        self.decay = nn.Parameter(torch.tensor(float(initial_decay)))
        self.wet = nn.Parameter(torch.tensor(float(initial_wet)))
        t = torch.arange(self.length) / self.sample_rate
        self.register_buffer("t", t.reshape(1, -1, 1))

    def build_impulse(self):
        """modules.py:21-26 -> (1, L, 1)."""
        return F_.ReverbImpulse.apply(self.noise, self.decay, self.wet, self.t)

    def forward(self, x):
        """modules.py:28-35: x (B,N,1) -> (B,N,1).  The impulse is zero-padded to N (or cropped when
        N < L, the reference's negative pad); the convolution kernel takes the shorter of the two."""
        n = x.shape[1]
        impulse = self.build_impulse()
        taps = min(n, self.length)
        kernel = impulse.reshape(1, self.length)[:, :taps]
        y = F_.FFTConvolve.apply(x.squeeze(-1), kernel)
        return y.unsqueeze(-1)


class HarmonicSynth(nn.Module):
    """ddsp/models/modules.py:38-80."""

    def __init__(self, block_size: int, sample_rate: int):
        super().__init__()
        self.block_size = block_size
        self.sample_rate = sample_rate

    def get_controls(self, amplitudes, harmonic_distribution, f0) -> Dict[str, torch.Tensor]:
        """modules.py:44-67: scale_function on both, Nyquist mask, normalise -- one fused launch."""
        amplitudes, harmonic_distribution = F_.HarmonicControls.apply(
            amplitudes, harmonic_distribution, f0, float(self.sample_rate))
        return {"f0": f0, "harmonic_distribution": harmonic_distribution, "amplitudes": amplitudes}

    def forward(self, amplitudes, harmonic_distribution, f0, phase0: Optional[torch.Tensor] = None):
        """modules.py:69-80 (run get_controls first).  Like the reference, ``harmonic_distribution``
        is scaled by ``amplitudes`` IN PLACE, so the caller's controls dict afterwards holds
        distribution x amplitude.  The two upsamples and the (B,N,H) oscillator tensor
        are fused away."""
        harmonic_distribution *= amplitudes
        audio, self._phase_end = core.harmonic_synth_frames(
            f0, harmonic_distribution, self.block_size, self.sample_rate, phase0)
        return audio


    def synthesize(self, param, f0, phase0: Optional[torch.Tensor] = None):
        """Extension: ``get_controls`` + ``forward`` on the projection output ``param`` (B,T,H+1) of decoder.py:106
        (amplitude = column 0, distribution = columns 1..H) in one launch: the controls are computed in the oscillator
        bank's prologue.  Returns (audio, harmonic_ctrls) with the dict the reference holds AFTER forward
        (``harmonic_distribution`` already scaled by the amplitudes, modules.py:73)."""
        fusable = (param.is_cuda and param.dtype == torch.float32 and not f0.requires_grad
                   and core.harmonic_raw_supported(param.shape[-1] - 1, self.block_size))
        if not fusable:
            ctrls = self.get_controls(param[..., :1], param[..., 1:], f0)
            return self.forward(ctrls["amplitudes"], ctrls["harmonic_distribution"], f0, phase0), ctrls
        audio, self._phase_end, amps, weights = core.harmonic_synth_from_raw(
            param, None, f0, self.block_size, self.sample_rate, phase0)
        return audio, {"f0": f0, "harmonic_distribution": weights, "amplitudes": amps}


class FilteredNoise(nn.Module):
    """ddsp/models/modules.py:101-128."""

    def __init__(self, block_size: int, window_size: int, initial_bias: int = -5.0):
        super().__init__()
        self.block_size = block_size
        self.window_size = window_size
        self.initial_bias = initial_bias
        # extension: draw the noise on the device instead (different random stream, no 4-19 ms CPU draw
        # + H2D copy per step at batch 16-64); default keeps the reference's CPU generator semantics
        self.device_noise = False

    def get_controls(self, magnitudes):
        return {"magnitudes": core.scale_function(magnitudes + self.initial_bias)}

    def draw_noise(self, magnitudes):
        """modules.py:119-123: the reference draws uniform(-1,1) with the CPU default generator and
        moves it to the device; doing exactly that keeps 'the same noise tensor' for a given seed."""
        if self.device_noise:
            return torch.rand(magnitudes.shape[0], magnitudes.shape[1], self.block_size,
                              device=magnitudes.device, dtype=magnitudes.dtype) * 2 - 1
        noise = torch.rand(magnitudes.shape[0], magnitudes.shape[1], self.block_size)
        return noise.to(magnitudes) * 2 - 1

    def forward(self, magnitudes, noise: Optional[torch.Tensor] = None):
        """modules.py:116-128.  ``noise`` (extension): pass the (B,T,block) draw explicitly."""
        if noise is None:
            noise = self.draw_noise(magnitudes)
        return core.filtered_noise(magnitudes, noise)
