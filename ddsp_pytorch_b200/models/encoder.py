"""MFCC-conditioned autoencoder with the reference's classes and state_dict keys
(ddsp/models/encoder.py)."""
from __future__ import annotations

import torch
import torch.nn as nn

from .. import core
from .decoder import GRUDecoder, _SynthWiring


class MFCCEncoder(nn.Module):
    """ddsp/models/encoder.py:10-27: LayerNorm -> GRU -> Linear to z."""

    def __init__(self, sample_rate: int, block_size: int, hidden_size: int, n_mfccs: int,
                 z_dim: int = None):
        super().__init__()
        self.hidden_size = hidden_size
        self.z_dim = z_dim
        self.norm = nn.LayerNorm(n_mfccs)
        self.gru = core.ClusterGRU(n_mfccs, hidden_size, batch_first=True)     # nn.GRU's parameters and keys
        self.proj = core.Linear(hidden_size, z_dim)

    def forward(self, mfccs: torch.Tensor):
        x, _ = self.gru(self.norm(mfccs))
        return self.proj(x)


class DDSPAutoencoder(_SynthWiring):
    """ddsp/models/encoder.py:29-103: adds z = encoder(mfcc) to the decoder inputs and the output."""

    def __init__(self, hidden_size: int, n_harmonic: int, n_bands: int, sample_rate: int,
                 block_size: int, has_reverb: bool):
        super().__init__()
        self.encoder = MFCCEncoder(sample_rate, block_size, hidden_size, n_mfccs=30, z_dim=16)
        self.decoder = GRUDecoder(hidden_size=hidden_size, z_dim=16)
        self._init_synth(hidden_size, n_harmonic, n_bands, sample_rate, block_size, has_reverb)

    def forward(self, batch: dict):
        f0, loudness, mfcc = batch["pitch"], batch["loudness"], batch["mfcc"]
        z = self.encoder(mfcc)
        hidden = self.decoder(f0, loudness, z=z)
        out = self._synthesize(hidden, f0, loudness, batch.get("noise"))
        out["z"] = z
        return out
