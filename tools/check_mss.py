"""Fused multi-scale loss (csrc/mss_fused.cu): parity against the float64 oracle at a few shapes, then device time
at the benchmark shape (events, L2 flushed between iterations).  Prints one JSON object.

    python tools/check_mss.py [--batch 64] [--iters 30] [--no-parity]
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ddsp_pytorch_b200 as ddsp  # noqa: E402
from ddsp_pytorch_b200.functions import hann_window_like_reference  # noqa: E402

SCALES = [4096, 2048, 1024, 512, 256, 128]


def parity(B, N, scales, seed):
    from oracle import ddsp_oracle as orc
    g = torch.Generator().manual_seed(seed)
    tgt = 0.1 * torch.randn(B, N, generator=g)
    rec = 0.1 * torch.randn(B, N, generator=g)
    r64 = rec.double().requires_grad_(True)
    ref = orc.mss_loss(tgt.double(), r64, scales, 0.75)
    ref.backward()
    r32 = rec.clone().requires_grad_(True)
    orc.mss_loss(tgt, r32, scales, 0.75).backward()
    r = rec.cuda().requires_grad_(True)
    loss = ddsp.multiscale_spectral_loss(tgt.cuda(), r, scales, 0.75)
    loss.backward()
    torch.cuda.synchronize()
    gref = r64.grad
    return {"B": B, "N": N, "scales": scales,
            "loss_rel": abs(float(loss) - float(ref)) / abs(float(ref)),
            "grad_rel": float((r.grad.double().cpu() - gref).norm() / gref.norm()),
            "grad_rel_reference_fp32": float((r32.grad.double() - gref).norm() / gref.norm()),
            "grad_max_abs": float((r.grad.double().cpu() - gref).abs().max()), "grad_ref_max": float(gref.abs().max())}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--iters", type=int, default=30)
    ap.add_argument("--no-parity", action="store_true")
    args = ap.parse_args()
    out = {"parity": [], "timing": []}
    if not args.no_parity:
        out["parity"].append(parity(2, 64000, SCALES, 1))
        out["parity"].append(parity(3, 9001, [1024, 512, 256, 128, 64], 2))      # odd length, 5 scales, n_fft 64
        out["parity"].append(parity(1, 2500, [4096], 3))                         # pad almost as long as the signal
        out["parity"].append(parity(5, 16000, [2048, 256], 4))
    N = 64000
    flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    for B in sorted({args.batch, 8}):
        g = torch.Generator().manual_seed(0)
        tgt = (0.1 * torch.randn(B, N, generator=g)).cuda()
        rec = (0.1 * torch.randn(B, N, generator=g)).cuda()
        win = torch.cat([hann_window_like_reference(s, rec.device) for s in SCALES])
        for need in (True, False):
            fn = lambda: torch.ops.ddsp_b200.mss_loss_fwd(tgt, rec, SCALES, 0.75, win, need)
            for _ in range(3):
                fn()
            torch.cuda.synchronize()
            ts = []
            for _ in range(args.iters):
                flush_buf.zero_()
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                fn()
                b.record()
                b.synchronize()
                ts.append(a.elapsed_time(b))
            ts.sort()
            out["timing"].append({"B": B, "grad": need, "ms_median": ts[len(ts) // 2], "ms_min": ts[0]})
    # one scale at a time (fused kernel + finalize + combine of that scale only)
    B = args.batch
    g = torch.Generator().manual_seed(0)
    tgt = (0.1 * torch.randn(B, N, generator=g)).cuda()
    rec = (0.1 * torch.randn(B, N, generator=g)).cuda()
    out["per_scale_ms"] = {}
    for s_ in SCALES:
        win = hann_window_like_reference(s_, rec.device)
        for need in (True, False):
            fn = lambda: torch.ops.ddsp_b200.mss_loss_fwd(tgt, rec, [s_], 0.75, win, need)
            for _ in range(3):
                fn()
            torch.cuda.synchronize()
            ts = []
            for _ in range(10):
                flush_buf.zero_()
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                fn()
                b.record()
                b.synchronize()
                ts.append(a.elapsed_time(b))
            ts.sort()
            out["per_scale_ms"][f"{s_}{'_grad' if need else ''}"] = ts[len(ts) // 2]
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
