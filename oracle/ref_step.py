"""The hot path run by the UNMODIFIED reference code in ``oracle/_ref`` (see ``oracle/build_ref.py``).

TEST / BENCH INFRASTRUCTURE, NOT PRODUCT CODE: used by ``bench.py --impl reference``, by ``bench.py``'s
``cpu_baseline`` leg and by tests that pin the oracle port against the reference.  It imports nothing from the
product package and maps none of its shared libraries.

One step = what ``bench.py``'s B200 arm does, through the reference's own modules and functions:
    decoder.py:110-125   harmonic_synth.get_controls -> harmonic_synth(**ctrls); noise_synth.get_controls ->
                         noise_synth(**ctrls) (draws its own uniform noise, modules.py:119-123); sum; reverb
    train.py:92-103      multiscale_fft of the target and of the reconstruction, multiscale_spec_loss
    train.py:129         loss.backward()  (gradients reach the synth inputs and the reverb parameters)
``multiscale_spec_loss`` lives in the reference's train.py, a script that cannot be imported (it parses
arguments and starts training at import); its six lines (train.py:70-76) are restated below.
"""
from __future__ import annotations

import os
import sys
import types

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")


def available() -> bool:
    return os.path.exists(os.path.join(REF_DIR, "MANIFEST.json"))


def import_reference():
    """``import ddsp`` from oracle/_ref.  librosa / crepe / matplotlib / pytorch_lightning (top-level imports of
    core.py, utils.py, decoder.py, encoder.py that the hot path never calls) and the reference's own data / preprocess
    modules (not taken into oracle/_ref) are satisfied by empty stub modules."""
    if not available():
        raise ImportError("oracle/_ref is missing: run `python oracle/build_ref.py` in the build container")
    for name in ["librosa", "crepe", "matplotlib", "matplotlib.pyplot", "pytorch_lightning"]:
        sys.modules.setdefault(name, types.ModuleType(name))
    for name in list(sys.modules):
        if name == "ddsp" or name.startswith("ddsp."):
            del sys.modules[name]
    sys.modules["ddsp.data"] = types.ModuleType("ddsp.data")
    if REF_DIR not in sys.path:
        sys.path.insert(0, REF_DIR)
    import ddsp  # noqa: the reference package
    assert os.path.dirname(os.path.abspath(ddsp.__file__)).startswith(REF_DIR), ddsp.__file__
    return ddsp


def multiscale_spec_loss(ddsp, ori_stft, rec_stft):
    """train.py:70-76, verbatim semantics."""
    loss = 0
    for s_x, s_y in zip(ori_stft, rec_stft):
        lin_loss = (s_x - s_y).abs().mean()
        log_loss = (ddsp.safe_log(s_x) - ddsp.safe_log(s_y)).abs().mean()
        loss = loss + lin_loss + log_loss
    return loss


class ReferenceStep:
    """The reference's synth modules + loss for one workload (``shapes`` = ddsp_pytorch_b200/shapes.SynthShapes)."""

    def __init__(self, shapes, reverb_state=None, dtype=torch.float32):
        self.ddsp = import_reference()
        from ddsp.models.modules import FilteredNoise, HarmonicSynth, Reverb      # the reference's classes
        s = self.shapes = shapes
        self.harmonic = HarmonicSynth(block_size=s.block_size, sample_rate=s.sample_rate)
        self.noise = FilteredNoise(block_size=s.block_size, window_size=s.n_bands)
        self.reverb = None
        if s.reverb_length is not None:
            self.reverb = Reverb(s.reverb_length, s.sample_rate)
            if reverb_state is not None:
                self.reverb.load_state_dict(reverb_state)
            self.reverb = self.reverb.to(dtype)

    def forward(self, amp_raw, dist_raw, mag_raw, pitch):
        hc = self.harmonic.get_controls(amp_raw, dist_raw, pitch)
        harmonic = self.harmonic(**hc)
        nc = self.noise.get_controls(mag_raw)
        noise = self.noise(**nc)
        signal = harmonic + noise
        if self.reverb is not None:
            signal = self.reverb(signal)
        return signal

    def step(self, host):
        """forward + loss + backward; returns (loss, grads of amp_raw, dist_raw, mag_raw[, reverb noise, decay, wet])"""
        s = self.shapes
        leaves = [host[k].detach().clone().requires_grad_(True) for k in ("amp_raw", "dist_raw", "mag_raw")]
        params = list(self.reverb.parameters()) if self.reverb is not None else []
        for p in params:
            p.grad = None
        signal = self.forward(leaves[0], leaves[1], leaves[2], host["pitch"])
        rec = signal.squeeze(-1)
        sig_stft = self.ddsp.multiscale_fft(host["target"], list(s.scales), s.overlap)
        rec_stft = self.ddsp.multiscale_fft(rec, list(s.scales), s.overlap)
        loss = multiscale_spec_loss(self.ddsp, sig_stft, rec_stft)
        loss.backward()
        return loss.detach(), [x.grad for x in leaves] + [p.grad for p in params]
