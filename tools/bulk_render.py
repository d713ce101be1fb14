"""Config 4 (bulk render): V voices x 4 s at 48 kHz, block 512, 256 harmonics, 65 bands, 1 s reverb IR,
forward only, on ONE GPU (the 8-GPU job shards 8192 voices as 1024 per GPU with no cross-GPU traffic).
Voices are rendered in chunks; noise is drawn on the device.  Reports samples/s, the harmonic bank's
harmonic-samples/s against the FP32 pipe and per-stage times.

    python tools/bulk_render.py [--voices 1024] [--chunk 128]
"""
import argparse, json, os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ddsp_pytorch_b200 as ddsp  # noqa
from ddsp_pytorch_b200.hotpath import SynthShapes, synthetic_inputs

ap = argparse.ArgumentParser()
ap.add_argument("--voices", type=int, default=1024)
ap.add_argument("--chunk", type=int, default=128)
ap.add_argument("--check", action="store_true", help="compare one voice against the float64 oracle")
args = ap.parse_args()
sr, bs, H, NB, T, L = 48000, 512, 256, 65, 375, 48000
N = T * bs
ops = torch.ops.ddsp_b200
dev = torch.device("cuda")
shapes = SynthShapes(batch=args.chunk, frames=T, block_size=bs, n_harmonic=H, n_bands=NB, sample_rate=sr, reverb_length=L)
host = synthetic_inputs(shapes, seed=0, pitch_lo=28.0, pitch_hi=72.0)          # C4 pitch range
inp = {k: v.to(dev) for k, v in host.items()}
torch.manual_seed(0)
from ddsp_pytorch_b200.models.modules import Reverb
rv = Reverb(L, sr, initial_wet=0.0, initial_decay=5.0).to(dev)
impulse = rv.build_impulse().detach().reshape(1, L)

def render(noise):
    amps, dist, w = ops.harmonic_controls_fwd(inp["amp_raw"], inp["dist_raw"], inp["pitch"], float(sr), True)
    audio, _, _, _ = ops.harmonic_fwd(inp["pitch"], w, bs, float(sr), None)
    sig = ops.noise_fwd(inp["mag_raw"], noise, audio, True, -5.0)
    return ops.fftconv_fwd(sig.squeeze(-1), impulse, False)[0]

chunks = args.voices // args.chunk
noise = torch.empty(args.chunk, T, bs, device=dev)
for _ in range(2):
    noise.uniform_(-1, 1); out = render(noise)
torch.cuda.synchronize()
ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
ev[0].record()
for c in range(chunks):
    noise.uniform_(-1, 1)
    out = render(noise)
ev[1].record(); torch.cuda.synchronize()
total_ms = ev[0].elapsed_time(ev[1])

def t(fn, n=5):
    fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n
amps, dist, w = ops.harmonic_controls_fwd(inp["amp_raw"], inp["dist_raw"], inp["pitch"], float(sr), True)
audio = ops.harmonic_fwd(inp["pitch"], w, bs, float(sr), None)[0]
stage = {
    "harmonic_ms": t(lambda: ops.harmonic_fwd(inp["pitch"], w, bs, float(sr), None)),
    "noise_ms": t(lambda: ops.noise_fwd(inp["mag_raw"], noise, audio, True, -5.0)),
    "reverb_ms": t(lambda: ops.fftconv_fwd(audio.squeeze(-1), impulse, False)),
    "rng_ms": t(lambda: noise.uniform_(-1, 1)),
}
hs = args.chunk * N * H
peaks = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json"))) if os.path.exists(
    os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")) else {"sm_max_mhz": 1965.0, "hbm_gbs": 6650.0}
fma_peak = 148 * 128 * peaks["sm_max_mhz"] * 1e6
res = {"config": f"configs[3] bulk render slice: {args.voices} voices x 4 s @48 kHz, H=256, reverb 48000 taps, fwd, 1 GPU",
       "voices": args.voices, "chunk": args.chunk, "total_ms": total_ms,
       "samples_per_s": args.voices * N / (total_ms * 1e-3),
       "seconds_for_1024_voices": total_ms * 1e-3 * 1024 / args.voices,
       "stage_ms_per_chunk": stage,
       "harmonic_samples_per_s": hs / (stage["harmonic_ms"] * 1e-3),
       "harmonic_fp32_roofline_frac_2fma": 2 * hs / (stage["harmonic_ms"] * 1e-3) / fma_peak,
       "noise_hbm_frac": 4 * args.chunk * T * (NB + 3 * bs) / (stage["noise_ms"] * 1e-3) / (peaks["hbm_gbs"] * 1e9),
       "reverb_hbm_frac": 4 * (2 * args.chunk * N + L) / (stage["reverb_ms"] * 1e-3) / (peaks["hbm_gbs"] * 1e9)}
if args.check:
    from oracle import ddsp_oracle as orc
    d = {k: v[:1].double() for k, v in host.items()}
    hc = orc.harmonic_controls(d["amp_raw"], d["dist_raw"], d["pitch"], sr)
    ref = orc.harmonic_synth_frames(hc["amplitudes"], hc["harmonic_distribution"], d["pitch"], bs, sr)
    res["harmonic_max_abs_err_vs_fp64"] = float((audio[:1].double().cpu() - ref).abs().max())
print(json.dumps(res))
