"""Workload shapes and seeded synthetic inputs of the hot path.

Pure Python + torch CPU: this file imports nothing from the package, so ``bench.py --impl reference`` can load it
by path (``importlib``) without importing ``ddsp_pytorch_b200`` and without mapping the native libraries.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional, Sequence

import torch


@dataclass
class SynthShapes:
    batch: int
    frames: int
    block_size: int
    n_harmonic: int
    n_bands: int
    sample_rate: int
    reverb_length: Optional[int]            # None = no reverb (realtime export path)
    scales: Sequence[int] = (4096, 2048, 1024, 512, 256, 128)
    overlap: float = 0.75

    @property
    def samples(self) -> int:
        return self.frames * self.block_size


def synthetic_inputs(shapes: SynthShapes, seed: int = 0, pitch_lo: float = 36.0, pitch_hi: float = 84.0):
    """SURVEY 8d synthetic inputs, drawn with the CPU generator (seeded), as CPU float32 tensors:
    smooth pitch contours (MIDI note per voice + 5 Hz vibrato), N(0,1) decoder outputs, uniform
    noise draw, 0.1*N(0,1) target audio."""
    s = shapes
    g = torch.Generator().manual_seed(seed)
    midi = torch.rand(s.batch, 1, 1, generator=g) * (pitch_hi - pitch_lo) + pitch_lo
    t = torch.arange(s.frames).view(1, -1, 1) * (s.block_size / s.sample_rate)
    vib = 0.5 * torch.sin(2 * torch.pi * 5.0 * t + 2 * torch.pi * torch.rand(s.batch, 1, 1, generator=g))
    pitch = 440.0 * torch.pow(2.0, (midi + vib - 69.0) / 12.0)
    return {
        "amp_raw": torch.randn(s.batch, s.frames, 1, generator=g),
        "dist_raw": torch.randn(s.batch, s.frames, s.n_harmonic, generator=g),
        "mag_raw": torch.randn(s.batch, s.frames, s.n_bands, generator=g),
        "pitch": pitch.float().contiguous(),
        "noise": torch.rand(s.batch, s.frames, s.block_size, generator=g) * 2 - 1,
        "target": 0.1 * torch.randn(s.batch, s.samples, generator=g),
    }
