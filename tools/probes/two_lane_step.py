"""Probe: the batch-64 step as TWO half-batch steps on two streams inside one CUDA graph (voice lanes), against the
single full-batch step.  Events around graph replays, L2 flushed between iterations.

    python tools/probes/two_lane_step.py [--lanes 2] [--batch 64] [--iters 50]
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from ddsp_pytorch_b200.hotpath import SynthShapes, SynthStep, synthetic_inputs  # noqa: E402


def timed(replay, iters, flush):
    a = [torch.cuda.Event(enable_timing=True) for _ in range(iters)]
    b = [torch.cuda.Event(enable_timing=True) for _ in range(iters)]
    for _ in range(5):
        replay()
    for i in range(iters):
        flush.zero_()
        a[i].record()
        replay()
        b[i].record()
    torch.cuda.synchronize()
    ts = sorted(x.elapsed_time(y) for x, y in zip(a, b))
    return ts[len(ts) // 2], ts[0]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--lanes", type=int, default=2)
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--iters", type=int, default=50)
    args = ap.parse_args()
    dev = torch.device("cuda")
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    base = dict(frames=400, block_size=160, n_harmonic=100, n_bands=65, sample_rate=16000, reverb_length=16000,
                scales=(4096, 2048, 1024, 512, 256, 128), overlap=0.75)
    full_shapes = SynthShapes(batch=args.batch, **base)
    host = synthetic_inputs(full_shapes, seed=0)
    torch.manual_seed(0)
    full = SynthStep(full_shapes, dev)
    full.load_inputs(host, non_blocking=False)
    full.capture()
    out = {"full": dict(zip(("ms_median", "ms_min"), timed(full.replay, args.iters, flush)))}
    per = args.batch // args.lanes
    lanes, streams = [], []
    for i in range(args.lanes):
        st = SynthStep(SynthShapes(batch=per, **base), dev, reverb_state=full.reverb.state_dict())
        st.load_inputs({k: v[i * per:(i + 1) * per].contiguous() for k, v in host.items()}, non_blocking=False)
        lanes.append(st)
        streams.append(torch.cuda.Stream())
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(3):
            for st in lanes:
                st.run()
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        cur = torch.cuda.current_stream()
        for st, s in zip(lanes, streams):
            s.wait_stream(cur)
            with torch.cuda.stream(s):
                st.run()
        for s in streams:
            cur.wait_stream(s)
    out[f"lanes{args.lanes}"] = dict(zip(("ms_median", "ms_min"), timed(g.replay, args.iters, flush)))
    loss_lanes = sum(float(st.loss) for st in lanes) / args.lanes
    out["loss_full"], out["loss_lanes"] = float(full.loss), loss_lanes
    print(json.dumps(out))


if __name__ == "__main__":
    main()
