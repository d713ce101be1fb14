"""Control network + synth wiring with the reference's classes and state_dict keys
(ddsp/models/decoder.py).  The control net's layers (``core.mlp``, ``core.gru``, ``core.Linear``) are subclasses of
the stock torch.nn modules that run on this repo's kernels when the shapes qualify (split-bf16 tcgen05 GEMM, fused
LayerNorm + LeakyReLU, cluster-persistent GRU: DESIGN 3.6); the synth stages they drive are the fused kernels.
"""
from __future__ import annotations

from typing import Dict, Optional

import torch
import torch.nn as nn

from .. import core
from .modules import FilteredNoise, HarmonicSynth, Reverb


class GRUDecoder(nn.Module):
    """ddsp/models/decoder.py:9-68."""

    def __init__(self, hidden_size: int, z_dim: Optional[int] = None):
        super().__init__()
        self.register_buffer("cache_gru", torch.zeros(1, 1, hidden_size))
        n_layers = 3
        self.f0_mlp = core.mlp(in_size=1, hidden_size=hidden_size, n_layers=n_layers)
        self.loudness_mlp = core.mlp(in_size=1, hidden_size=hidden_size, n_layers=n_layers)
        self.add_z = z_dim is not None
        if self.add_z:
            self.z_mlp = core.mlp(z_dim, hidden_size, n_layers)
        self.gru = core.gru(3 if self.add_z else 2, hidden_size)
        self.out_mlp = core.mlp(hidden_size + 2, hidden_size, n_layers)

    def forward(self, f0, loudness, z=None, realtime: bool = False):
        parts = [self.f0_mlp(f0), self.loudness_mlp(loudness)]
        if self.add_z:
            assert z is not None
            parts.append(self.z_mlp(z))
        hidden = torch.cat(parts, -1)
        if realtime:
            gru_out, cache = self.gru(hidden, self.cache_gru)
            self.cache_gru.copy_(cache)
        else:
            gru_out = self.gru(hidden)[0]
        return self.out_mlp(torch.cat([gru_out, f0, loudness], -1))


class _SynthWiring(nn.Module):
    """decoder.py:106-136 / encoder.py:76-103: projections -> controls -> harmonic + noise (+ reverb)."""

    def _init_synth(self, hidden_size, n_harmonic, n_bands, sample_rate, block_size, has_reverb):
        self.register_buffer("sample_rate", torch.tensor(sample_rate))
        self.register_buffer("block_size", torch.tensor(block_size))
        self.harmonic_proj = core.Linear(hidden_size, n_harmonic + 1)
        self.noise_proj = core.Linear(hidden_size, n_bands)
        self.harmonic_synth = HarmonicSynth(block_size=block_size, sample_rate=sample_rate)
        self.noise_synth = FilteredNoise(block_size=block_size, window_size=n_bands)
        self.has_reverb = has_reverb
        self.reverb = Reverb(sample_rate, sample_rate)
        self.register_buffer("phase", torch.zeros(1))

    def _synthesize(self, hidden, f0, loudness, noise=None) -> Dict[str, torch.Tensor]:
        # decoder.py:106-110: projection -> get_controls -> forward; the controls ride in the oscillator bank's launch
        param = self.harmonic_proj(hidden)
        harmonic, harmonic_ctrls = self.harmonic_synth.synthesize(param, f0)

        magnitudes = self.noise_proj(hidden)
        noise_ctrls = self.noise_synth.get_controls(magnitudes)
        noise_audio = self.noise_synth(noise_ctrls["magnitudes"], noise=noise)

        signal = harmonic + noise_audio
        if self.has_reverb:
            signal = self.reverb(signal)
        return {
            "f0": f0,
            "loudness": loudness,
            "signal": signal,
            "noise": noise_audio,
            "harmonic_audio": harmonic,
            "noise_ctrls": noise_ctrls,
            "harmonic_ctrls": harmonic_ctrls,
        }


class DDSPDecoder(_SynthWiring):
    """ddsp/models/decoder.py:70-136: f0 + loudness in, dict of audio and controls out."""

    def __init__(self, hidden_size: int, n_harmonic: int, n_bands: int, sample_rate: int,
                 block_size: int, has_reverb: bool):
        super().__init__()
        self.decoder = GRUDecoder(hidden_size=hidden_size, z_dim=None)
        self._init_synth(hidden_size, n_harmonic, n_bands, sample_rate, block_size, has_reverb)

    def forward(self, batch: dict):
        f0, loudness = batch["pitch"], batch["loudness"]
        hidden = self.decoder(f0, loudness)
        return self._synthesize(hidden, f0, loudness, batch.get("noise"))

    def realtime_forward(self, f0, loudness):
        """decoder.py:138-158, repaired: GRU state is cached in ``decoder.cache_gru``,
        the oscillator phase is carried in the ``phase`` buffer (the reference registers it at
        decoder.py:99 but restarts the phase at 0 every call), reverb is left to the host."""
        hidden = self.decoder(f0, loudness, realtime=True)
        param = self.harmonic_proj(hidden)
        phase0 = self.phase.double().expand(f0.shape[0]).contiguous()
        harmonic, _ = self.harmonic_synth.synthesize(param, f0, phase0=phase0)
        self.phase.copy_(self.harmonic_synth._phase_end[:1].to(self.phase.dtype))
        magnitudes = self.noise_synth.get_controls(self.noise_proj(hidden))["magnitudes"]
        return harmonic + self.noise_synth(magnitudes)
