"""Probe: what the host->device copy of one step's inputs (49.9 MB, pinned) costs alone and while the GPU computes --
under the whole step's graph, under the fused loss only, under the oscillator bank only -- and what the step costs while
the copy engine is busy.  Explains the spread of bench.py's e2e between boxes.

    python tools/probes/h2d_overlap_probe.py
"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from ddsp_pytorch_b200.hotpath import SynthShapes, SynthStep, synthetic_inputs  # noqa: E402


def main():
    dev = torch.device("cuda")
    shapes = SynthShapes(batch=64, frames=400, block_size=160, n_harmonic=100, n_bands=65, sample_rate=16000,
                         reverb_length=16000, scales=(4096, 2048, 1024, 512, 256, 128), overlap=0.75)
    host = synthetic_inputs(shapes, seed=0)
    torch.manual_seed(0)
    step = SynthStep(shapes, dev)
    step.load_inputs(host, non_blocking=False)
    step.capture()
    block = step.pack_host(host)
    dst = torch.empty_like(step._flat_in)
    cs = torch.cuda.Stream()
    ops = torch.ops.ddsp_b200
    i = step.inputs
    from ddsp_pytorch_b200.functions import hann_windows_like_reference
    win = hann_windows_like_reference(list(shapes.scales), dev)
    sig = torch.randn(64, 64000, device=dev) * 0.1

    def loss_only():
        ops.mss_loss_fwd(i["target"], sig, list(shapes.scales), 0.75, win, True)

    def bank_only():
        ops.harmonic_raw_fwd(i["amp_raw"], i["dist_raw"], i["pitch"], 160, 16000.0, None)

    def measure(load, reps=20, copies=True):
        """mean copy time (events on the copy stream) and mean load time (events on the main stream), run together"""
        torch.cuda.synchronize()
        ce, le = [], []
        for _ in range(reps):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            c, d = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            if load is not None:
                c.record()
                load()
                d.record()
            if copies:
                with torch.cuda.stream(cs):
                    a.record(cs)
                    dst.copy_(block, non_blocking=True)
                    b.record(cs)
            torch.cuda.synchronize()
            if copies:
                ce.append(a.elapsed_time(b))
            if load is not None:
                le.append(c.elapsed_time(d))
        m = lambda v: sorted(v)[len(v) // 2] if v else None
        return {"copy_ms": m(ce), "load_ms": m(le)}

    out = {"copy alone": measure(None),
           "step alone": measure(step.replay, copies=False),
           "copy + step graph": measure(step.replay),
           "copy + fused loss": measure(loss_only),
           "loss alone": measure(loss_only, copies=False),
           "copy + oscillator bank": measure(bank_only),
           "bank alone": measure(bank_only, copies=False)}
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
