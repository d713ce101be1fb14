"""The other BASELINE.json configurations as measurable functions (bench.py reports them as sub-objects of its
JSON line; tools/*.py are thin command-line wrappers).  Every function builds seeded synthetic inputs of the
configuration's shape on the device, warms up, and times with CUDA events (L2 flushed between iterations) or, for
the realtime path, with the host clock around the whole host -> device -> host round trip.

configs[0]  forward_b16     config.yaml defaults, batch 16, forward harmonic + noise + reverb
configs[1]  model_step      the full DDSPDecoder training step the headline metric's hot path is the core of
configs[2]  realtime        scripted export, batch 1, 1024-sample buffers (latency)
configs[3]  bulk_render     48 kHz, 256 harmonics, 1 s reverb, forward, voices per GPU
configs[4]  model_step(autoencoder=True)
"""
from __future__ import annotations

import os
import tempfile
import time
from typing import Dict

import torch

SCALES = [4096, 2048, 1024, 512, 256, 128]
OVERLAP = 0.75


def _flusher(dev):
    buf = torch.empty(256 << 20, dtype=torch.uint8, device=dev)      # > 126 MB of L2
    return lambda: buf.zero_()


def _event_ms(fn, iters, warmup, flush):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        flush()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        b.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    return ts[len(ts) // 2], ts[0]


def model_step(batch: int = 64, autoencoder: bool = False, iters: int = 20, device="cuda") -> Dict:
    """train.py:84-130 on config-1 shapes: control net + synth + fused loss + backward + Adam, float32, the whole
    step captured in one CUDA graph.  The noise draw is an input (drawn on the device outside the timed region)."""
    import ddsp_pytorch_b200 as ddsp
    from .models.decoder import DDSPDecoder
    from .models.encoder import DDSPAutoencoder
    dev = torch.device(device)
    torch.manual_seed(0)
    T, bs, sr = 400, 160, 16000
    cls = DDSPAutoencoder if autoencoder else DDSPDecoder
    model = cls(hidden_size=512, n_harmonic=100, n_bands=65, sample_rate=sr, block_size=bs, has_reverb=True).to(dev)
    opt = torch.optim.Adam(model.parameters(), lr=1e-3, capturable=True)
    g = torch.Generator().manual_seed(1)
    data = {"pitch": (torch.rand(batch, T, 1, generator=g) * 400 + 100).to(dev),
            "loudness": torch.randn(batch, T, 1, generator=g).to(dev),
            "sig": (0.1 * torch.randn(batch, T * bs, generator=g)).to(dev),
            "noise": (torch.rand(batch, T, bs, generator=g) * 2 - 1).to(dev)}
    if autoencoder:
        data["mfcc"] = torch.randn(batch, T, 30, generator=g).to(dev)

    def step():
        out = model(data)
        loss = ddsp.multiscale_spectral_loss(data["sig"], out["signal"].squeeze(-1), SCALES, OVERLAP)
        opt.zero_grad(set_to_none=False)
        loss.backward()
        opt.step()
        return loss

    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(3):
            step()
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        static_loss = step()
    med, best = _event_ms(graph.replay, iters, 3, _flusher(dev))
    n_param = sum(p.numel() for p in model.parameters())
    return {"config": ("configs[4] autoencoder.yaml" if autoencoder else "configs[1] full model") +
            f": {cls.__name__} hidden 512, 16 kHz, block 160, 100 harmonics, 4 s, batch {batch}; control net + synth + "
            "fused loss + backward + Adam, float32, one CUDA graph per step",
            "batch": batch, "ms_per_step": med, "ms_per_step_best": best,
            "samples_per_s": batch * T * bs / (med * 1e-3), "parameters": n_param, "loss": float(static_loss)}


def forward_b16(iters: int = 30, device="cuda") -> Dict:
    """configs[0]: config.yaml defaults, batch 16, forward harmonic + noise + reverb (no control net), graph replay."""
    from .hotpath import SynthShapes, SynthStep, synthetic_inputs
    dev = torch.device(device)
    shapes = SynthShapes(batch=16, frames=400, block_size=160, n_harmonic=100, n_bands=65, sample_rate=16000,
                         reverb_length=16000)
    torch.manual_seed(0)
    step = SynthStep(shapes, dev)
    step.load_inputs({k: v.pin_memory() for k, v in synthetic_inputs(shapes, seed=3).items()})
    step.run_forward()
    step.capture(forward_only=True)
    med, best = _event_ms(lambda: step.replay(True), iters, 3, _flusher(dev))
    return {"config": "configs[0]: 16 kHz, block 160, 100 harmonics, 65 bands, 4 s, batch 16, forward harmonic + noise "
            "+ reverb, CUDA graph replay", "ms_per_step": med, "ms_per_step_best": best,
            "samples_per_s": shapes.batch * shapes.samples / (med * 1e-3)}


def realtime_latency(iters: int = 300, device="cuda") -> Dict:
    """configs[2]: per-buffer latency of the scripted export the way the Pd external calls it
    (ddsp_model.cpp:32-51): 1024 host floats of pitch and loudness -> device -> forward -> host."""
    from .export import export_torchscript
    from .models.decoder import DDSPDecoder
    dev = torch.device(device)
    torch.manual_seed(0)
    model = DDSPDecoder(hidden_size=512, n_harmonic=64, n_bands=65, sample_rate=48000, block_size=512, has_reverb=True)
    path = os.path.join(tempfile.mkdtemp(), "rt.ts")
    export_torchscript(model.to(dev).eval(), path, mean_loudness=-30.0, std_loudness=10.0, realtime=True)
    rt = torch.jit.load(path).to(dev)
    pitch = torch.full((1, 1024, 1), 220.0).pin_memory()
    loud = torch.full((1, 1024, 1), -25.0).pin_memory()
    out_host = torch.empty(1, 1024, 1).pin_memory()

    def call():
        with torch.no_grad():
            y = rt(pitch.to(dev, non_blocking=True), loud.to(dev, non_blocking=True))
            out_host.copy_(y, non_blocking=True)
            torch.cuda.synchronize()

    for _ in range(30):
        call()
    ts = []
    for _ in range(iters):
        t0 = time.perf_counter()
        call()
        ts.append((time.perf_counter() - t0) * 1e3)
    ts.sort()
    budget = 1024 / 48000 * 1e3
    return {"config": "configs[2] realtime export: 48 kHz, block 512, 64 harmonics, batch 1, 1024-sample buffers, "
            "reverb left to the host", "ms_per_buffer_median": ts[len(ts) // 2], "ms_per_buffer_p99": ts[int(len(ts) * 0.99)],
            "budget_ms": budget, "realtime_factor": budget / ts[len(ts) // 2],
            "includes": "H2D of 2x1024 floats, control net, synth kernels, D2H of 1024 floats; host clock"}


class BulkRenderer:
    """configs[3] on one GPU: voices x 4 s at 48 kHz, block 512, 256 harmonics, 65 bands, 1 s reverb IR, forward
    only.  Voices are independent, so a multi-GPU job gives each rank its own slice and nothing crosses GPUs.  Voices
    are rendered in chunks through static buffers; the uniform noise is drawn on the device per chunk."""

    SR, BS, H, NB, T, L = 48000, 512, 256, 65, 375, 48000

    def __init__(self, chunk: int = 128, device="cuda", seed: int = 0):
        from .hotpath import SynthShapes, synthetic_inputs
        from .models.modules import Reverb
        self.dev = torch.device(device)
        self.chunk = chunk
        self.N = self.T * self.BS
        self.shapes = SynthShapes(batch=chunk, frames=self.T, block_size=self.BS, n_harmonic=self.H, n_bands=self.NB,
                                  sample_rate=self.SR, reverb_length=self.L)
        self.host = synthetic_inputs(self.shapes, seed=seed, pitch_lo=28.0, pitch_hi=72.0)      # C4 pitch range
        self.inp = {k: v.to(self.dev) for k, v in self.host.items() if k != "target"}
        torch.manual_seed(0)
        self.impulse = Reverb(self.L, self.SR).to(self.dev).build_impulse().detach().reshape(1, self.L)
        self.noise = torch.empty(chunk, self.T, self.BS, device=self.dev)
        # the learned IR is the same for every voice and chunk: its spectrum is computed once
        self.hspec = torch.ops.ddsp_b200.fftconv_spectrum(self.impulse, self.N)

    def render_chunk(self, draw: bool = True):
        ops, i = torch.ops.ddsp_b200, self.inp
        if draw:
            self.noise.uniform_(-1, 1)
        audio = ops.harmonic_raw_fwd(i["amp_raw"], i["dist_raw"], i["pitch"], self.BS, float(self.SR), None)[0]
        sig = ops.noise_fwd(i["mag_raw"], self.noise, audio, True, -5.0)
        return ops.fftconv_fwd(sig.squeeze(-1), self.impulse, False, self.hspec)[0]

    def render(self, voices: int):
        out = None
        for _ in range(voices // self.chunk):
            out = self.render_chunk()
        return out


def bulk_render(voices: int = 1024, chunk: int = 128, device="cuda") -> Dict:
    r = BulkRenderer(chunk, device)
    r.render(voices)                       # warm-up pass: allocator pools and lazily loaded kernels
    torch.cuda.synchronize()
    times = []
    for _ in range(3):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        r.render(voices)
        b.record()
        torch.cuda.synchronize()
        times.append(a.elapsed_time(b))
    ms = sorted(times)[1]                  # median of three passes
    return {"config": f"configs[3] slice: {voices} voices x 4 s @ 48 kHz, 256 harmonics, reverb 48000 taps, forward, "
            f"1 GPU, chunks of {chunk}", "voices": voices, "total_ms": ms, "samples_per_s": voices * r.N / (ms * 1e-3),
            "harmonic_samples_per_s": voices * r.N * r.H / (ms * 1e-3)}
