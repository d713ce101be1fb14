"""Config 3 (realtime export path): per-buffer latency of the scripted model, measured the way the Pd
external calls it (ddsp_model.cpp:32-51): 1024 host floats of pitch and loudness -> device -> forward ->
host.  48 kHz, block 512, 64 harmonics, 65 bands, hidden 512, batch 1, reverb left to the host.
Budget per buffer: 1024/48000 = 21.3 ms."""
import json, os, sys, time, tempfile
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ddsp_pytorch_b200  # noqa
from ddsp_pytorch_b200.models.decoder import DDSPDecoder
from ddsp_pytorch_b200.export import export_torchscript

torch.manual_seed(0)
model = DDSPDecoder(hidden_size=512, n_harmonic=64, n_bands=65, sample_rate=48000, block_size=512, has_reverb=True)
path = os.path.join(tempfile.mkdtemp(), "rt.ts")
export_torchscript(model.cuda().eval(), path, mean_loudness=-30.0, std_loudness=10.0, realtime=True)
rt = torch.jit.load(path).cuda()
pitch = torch.full((1, 1024, 1), 220.0).pin_memory()
loud = torch.full((1, 1024, 1), -25.0).pin_memory()
out_host = torch.empty(1, 1024, 1).pin_memory()
def call():
    with torch.no_grad():
        y = rt(pitch.cuda(non_blocking=True), loud.cuda(non_blocking=True))
        out_host.copy_(y, non_blocking=True)
        torch.cuda.synchronize()
for _ in range(30):
    call()
ts = []
for _ in range(300):
    t0 = time.perf_counter(); call(); ts.append((time.perf_counter() - t0) * 1e3)
ts.sort()
print(json.dumps({"config": "configs[2] realtime export: 48 kHz, block 512, 64 harmonics, batch 1, 1024-sample buffers",
                  "ms_per_buffer_median": ts[len(ts) // 2], "ms_per_buffer_p99": ts[int(len(ts) * 0.99)],
                  "budget_ms": 1024 / 48000 * 1e3, "realtime_factor": (1024 / 48000 * 1e3) / ts[len(ts) // 2],
                  "includes": "H2D of 2x1024 floats, control net (cluster GRU kernel, library GEMMs at this size), synth kernels, D2H of 1024 floats"}))
