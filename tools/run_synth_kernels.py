"""Runs K0-K3 (controls, harmonic bank, filtered noise, reverb convolution) forward and backward a few times at
config-2 shapes, and the harmonic bank at config-4 shapes, for ncu:

    ncu --set full -k regex:"harmonic_frames|filtered_noise2|cols_|rows_|controls_" ... python tools/run_synth_kernels.py [iters]
"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ddsp_pytorch_b200  # noqa
from ddsp_pytorch_b200.hotpath import SynthShapes, SynthStep, synthetic_inputs
from ddsp_pytorch_b200.workloads import BulkRenderer

iters = int(sys.argv[1]) if len(sys.argv) > 1 else 2
shapes = SynthShapes(batch=64, frames=400, block_size=160, n_harmonic=100, n_bands=65, sample_rate=16000, reverb_length=16000)
torch.manual_seed(0)
step = SynthStep(shapes, "cuda")
step.load_inputs(synthetic_inputs(shapes, seed=1), non_blocking=False)
ops = torch.ops.ddsp_b200
i = step.inputs
for _ in range(iters):
    amps, dist, w = ops.harmonic_controls_fwd(i["amp_raw"], i["dist_raw"], i["pitch"], 16000.0, True)
    audio, _, phi, delta = ops.harmonic_fwd(i["pitch"], w, 160, 16000.0, None)
    g = torch.randn_like(audio)
    ops.harmonic_bwd(g, w, phi, delta, 160, 16000.0, False)
    # the step's own form: controls computed in the bank's prologue / the controls' backward in the backward's epilogue
    a2, _, phi2, delta2, _, _ = ops.harmonic_raw_fwd(i["amp_raw"], i["dist_raw"], i["pitch"], 160, 16000.0, None)
    ops.harmonic_raw_bwd(g, i["amp_raw"], i["dist_raw"], i["pitch"], phi2, delta2, 160, 16000.0)
    ops.noise_fwd(i["mag_raw"], i["noise"], audio, True, -5.0)
    ops.noise_bwd(g, i["noise"], i["mag_raw"], 65, True, -5.0)
    sig2 = audio.squeeze(-1).contiguous()
    imp = step.reverb.build_impulse().detach().reshape(1, -1)
    kept = ops.fftconv_fwd(sig2, imp, True)
    ops.fftconv_bwd(sig2, sig2, imp, kept[1], kept[2], True, True)
torch.cuda.synchronize()
r = BulkRenderer(128, "cuda")           # config 4: 48 kHz, block 512, 256 harmonics
for _ in range(iters):
    r.render_chunk()
torch.cuda.synchronize()
print("ok")
