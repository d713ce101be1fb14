#!/usr/bin/env python
"""Benchmark of the DDSP synthesis hot path on B200 (BASELINE.json metric / configs[1]).

A "step" is one pass of the hot path over one batch of synthetic decoder outputs:
controls -> harmonic bank + filtered noise -> reverb -> multi-scale spectral loss -> backward to the
synth parameters (decoder.py:106-125 + train.py:92-103,129 of the reference, without the control
network).  Workload: 16 kHz, block 160, 100 harmonics, 65 noise bands, 4 s, batch 64 sharded over
the N GPUs (strong scaling), reverb 16000 taps, 6 STFT scales.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
    torchrun --nproc-per-node N bench.py --gpus N ...     (one rank per GPU)

Prints ONE JSON line (rank 0).  `value` = audio samples/s of forward+backward with inputs resident
in HBM (CUDA events, max over ranks); `e2e` = the same through pinned host buffers with the H2D
copies and the loss read-back inside the timed region; `roofline` = the dominant kernel against the
measured peak; `cpu_baseline` = the oracle port (the reference's algorithm as torch CPU float32) on
this box's host cores.  `--impl reference` times only that CPU path.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

METRIC = "DDSP synth audio samples/sec (fwd+bwd)"
K4L_DRAM_TRAFFIC = 197.2e6       # bytes per step, from the round-1 ncu --set full capture (profiles/)
UNIT = "samples/s"
WORKLOAD = dict(sample_rate=16000, block_size=160, n_harmonic=100, n_bands=65, frames=400, batch=64,
                reverb_length=16000, scales=(4096, 2048, 1024, 512, 256, 128), overlap=0.75)
WORKLOAD_NAME = ("configs[1]: synth hot path fwd+bwd with multiscale_fft loss, 16 kHz, block 160, "
                 "100 harmonics, 65 bands, 4 s, batch 64 sharded over N GPUs, reverb 16000 taps")


def parse():
    p = argparse.ArgumentParser()
    p.add_argument("--gpus", type=int, default=1)
    p.add_argument("--steps", type=int, default=100)
    p.add_argument("--warmup", type=int, default=10)
    p.add_argument("--impl", default="b200", choices=["b200", "reference"])
    p.add_argument("--no-graph", action="store_true", help="launch eagerly instead of replaying a CUDA graph")
    p.add_argument("--cpu-batch", type=int, default=8, help="voices in the bounded CPU-baseline sample")
    p.add_argument("--skip-cpu", action="store_true")
    p.add_argument("--global-batch", type=int, default=WORKLOAD["batch"],
                   help="experiments only: the benchmark config is batch 64 (BASELINE.json configs[1])")
    p.add_argument("--skip-kernels", action="store_true")
    return p.parse_args()


# ------------------------------------------------------------------------------------------ CPU arm
def cpu_port_step(shapes, host, reverb_state, orc):
    rp = {k: v for k, v in reverb_state.items()}
    return orc.synth_train_step(host["amp_raw"], host["dist_raw"], host["mag_raw"], host["pitch"],
                                host["noise"], host["target"], shapes.block_size, shapes.sample_rate, rp,
                                list(shapes.scales), shapes.overlap)


def cpu_reverb_state(shapes, seed=0):
    g = torch.Generator().manual_seed(seed)
    L = shapes.reverb_length
    return {"noise": (torch.rand(L, generator=g) * 2 - 1).unsqueeze(-1), "decay": torch.tensor(5.0),
            "wet": torch.tensor(0.0), "t": (torch.arange(L) / shapes.sample_rate).reshape(1, -1, 1)}


def time_cpu_port(batch, steps, warmup):
    """The oracle port (kind "port"): the reference's algorithm as torch CPU float32 on all host threads."""
    from ddsp_pytorch_b200.hotpath import SynthShapes, synthetic_inputs
    from oracle import ddsp_oracle as orc
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    w = dict(WORKLOAD)
    w["batch"] = batch
    shapes = SynthShapes(**w)
    host = synthetic_inputs(shapes, seed=0)
    rs = cpu_reverb_state(shapes)
    for _ in range(warmup):
        cpu_port_step(shapes, host, rs, orc)
    times = []
    for _ in range(steps):
        t0 = time.perf_counter()
        cpu_port_step(shapes, host, rs, orc)
        times.append(time.perf_counter() - t0)
    return shapes, times, cores


def run_reference(args, rank):
    if rank != 0:
        return
    # size the per-step sample so that the whole run stays within a couple of minutes
    batch = 2
    shapes, t1, cores = time_cpu_port(batch, 1, 1)
    budget = 120.0 / max(1, args.steps + args.warmup)
    while batch < args.cpu_batch and t1[0] * 2 <= budget:
        batch *= 2
        t1 = [t1[0] * 2]
    shapes, times, cores = time_cpu_port(batch, args.steps, args.warmup)
    total = sum(times)
    value = batch * shapes.samples * len(times) / total
    sample = f"{batch} of 64 voices per step (same shapes), float32, {cores} threads"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / len(times),
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "config": {"workload": WORKLOAD_NAME, "per_step_sample": sample},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    """Samples SM clock and throttle reasons through NVML while the timed region runs."""

    REASONS = {0x1: "gpu_idle", 0x2: "applications_clocks_setting", 0x4: "sw_power_cap", 0x8: "hw_slowdown",
               0x10: "sync_boost", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown",
               0x80: "hw_power_brake_slowdown", 0x100: "display_clock_setting"}

    def __init__(self, index):
        self.samples, self.reasons, self.max_mhz, self.ok = [], set(), None, False
        self._stop = threading.Event()
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:
            self.ok = False
        self.t = threading.Thread(target=self._loop, daemon=True)

    def _loop(self):
        while not self._stop.is_set():
            try:
                self.samples.append(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM))
                bits = self.nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                for b, name in self.REASONS.items():
                    if bits & b and name != "gpu_idle":
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.005)

    def __enter__(self):
        if self.ok:
            self.t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        if self.ok:
            self.t.join()

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": []}
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2], "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s)}


# ------------------------------------------------------------------------------------------ B200 arm
def event_time(fn, iters, warmup, flush=None):
    """Average device time of fn() over iters, CUDA events on the current stream, L2 flushed between."""
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    total = 0.0
    for _ in range(iters):
        if flush is not None:
            flush()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        b.synchronize()
        total += a.elapsed_time(b)
    return total / iters          # ms


def kernel_table(step, shapes, flush, peaks):
    """Per-kernel device time (events, L2 flushed) and roofline fraction for the stages of the path."""
    import ddsp_pytorch_b200 as ddsp
    ops = torch.ops.ddsp_b200
    i = {k: v.detach() for k, v in step.inputs.items()}
    B, T, bs, H, NB, N = shapes.batch, shapes.frames, shapes.block_size, shapes.n_harmonic, shapes.n_bands, shapes.samples
    sr = float(shapes.sample_rate)
    amps, dist, w = ops.harmonic_controls_fwd(i["amp_raw"], i["dist_raw"], i["pitch"], sr, True)
    audio, _, phi, delta = ops.harmonic_fwd(i["pitch"], w, bs, sr, None)
    g = torch.randn_like(audio)
    sig2 = audio.squeeze(-1).contiguous()
    imp = step.reverb.build_impulse().detach().reshape(1, -1)
    from ddsp_pytorch_b200.functions import hann_window_like_reference
    windows = torch.cat([hann_window_like_reference(s, audio.device) for s in shapes.scales])
    hbm = peaks["hbm_gbs"] * 1e9
    clk = (peaks.get("sm_max_mhz") or 1965.0) * 1e6
    fma_peak = 148 * 128 * clk                      # FP32 FMA lanes / s at max clock
    hs = B * N * H                                  # harmonic-samples
    rows = []

    def add(name, fn, alg_bytes=None, fma_ops=None, note=""):
        ms = event_time(fn, 20, 3, flush)
        r = {"kernel": name, "ms": ms}
        if alg_bytes is not None:
            r.update(bound="hbm", achieved=alg_bytes / (ms * 1e-3) / 1e9, peak=peaks["hbm_gbs"], unit="GB/s")
        if fma_ops is not None:
            r.update(bound="fp32", achieved=fma_ops / (ms * 1e-3) / 1e12, peak=fma_peak / 1e12, unit="TFMA/s")
        r["frac"] = r["achieved"] / r["peak"]
        r["note"] = note
        rows.append(r)

    add("K1 harmonic_frames_fwd", lambda: ops.harmonic_fwd(i["pitch"], w, bs, sr, None), fma_ops=2 * hs,
        note="SURVEY 8d: 2 FMA per harmonic-sample (recurrence + weighted sum) vs 148 SM x 128 lanes x max clock")
    add("K1 harmonic_frames_bwd", lambda: ops.harmonic_bwd(g, w, phi, delta, bs, sr, False), fma_ops=2 * hs)
    add("K2 filtered_noise_fwd", lambda: ops.noise_fwd(i["mag_raw"], i["noise"], audio, True, -5.0), alg_bytes=4 * B * T * (NB + 3 * bs))
    add("K2 filtered_noise_bwd", lambda: ops.noise_bwd(g, i["noise"], i["mag_raw"], NB, True, -5.0), alg_bytes=4 * B * T * (NB + 2 * bs))
    kept = ops.fftconv_fwd(sig2, imp, True)
    add("K3 reverb fftconv_fwd (5 launches)", lambda: ops.fftconv_fwd(sig2, imp, True), alg_bytes=4 * (2 * B * N + imp.numel()))
    add("K3 reverb fftconv_bwd (5 launches, transforms kept by fwd)", lambda: ops.fftconv_bwd(sig2, sig2, imp, kept[1], kept[2], True, True),
        alg_bytes=4 * (3 * B * N + 2 * imp.numel()))
    add("K4L mss_loss fwd+grad (6 scales + finish)",
        lambda: ops.mss_loss_fwd(i["target"], sig2, list(shapes.scales), shapes.overlap, windows, True),
        alg_bytes=4 * 3 * B * N, note="SURVEY 8d: read rec+target, write grad")
    # the same kernel against the FP32 pipe: 5 n log2 n flops per complex FFT, one forward per frame
    # (rec + i*target) and one inverse per frame pair
    import math
    fft_flops = 0.0
    for s_ in shapes.scales:
        hop = int(s_ * (1 - shapes.overlap))
        fft_flops += 1.5 * (1 + N // hop) * 5.0 * s_ * math.log2(s_)
    fft_flops *= B
    k4 = rows[-1]
    k4["fp32"] = {"achieved": fft_flops / (k4["ms"] * 1e-3) / 1e12, "peak": 2 * fma_peak / 1e12, "unit": "TFLOP/s",
                  "frac": fft_flops / (k4["ms"] * 1e-3) / (2 * fma_peak),
                  "note": "FFT butterflies only (7.0 GFLOP/step); the kernel is FP32/latency bound, not HBM bound"}
    add("K0 harmonic_controls_fwd", lambda: ops.harmonic_controls_fwd(i["amp_raw"], i["dist_raw"], i["pitch"], sr, True),
        alg_bytes=4 * B * T * (2 * H + 3))
    return rows


def run_b200(args, rank, world):
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist = None
    if world > 1:
        # keep stdout to the single JSON line: NCCL prints its version banner there at level VERSION
        if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
            os.environ["NCCL_DEBUG"] = "WARN"
        import torch.distributed as dist_
        dist = dist_
        dist.init_process_group("nccl", device_id=dev)

    import ddsp_pytorch_b200  # noqa: F401  (fails loudly without the native libraries)
    from ddsp_pytorch_b200._lib import core_library
    from ddsp_pytorch_b200.hotpath import SynthShapes, SynthStep, synthetic_inputs
    import ctypes
    lib = core_library()
    lib.ddsp_b200_launch_count.restype = ctypes.c_uint64

    w = dict(WORKLOAD)
    w["batch"] = args.global_batch
    assert w["batch"] % world == 0, "the global batch must divide over the GPUs"
    w["batch"] //= world
    shapes = SynthShapes(**w)
    torch.manual_seed(0)
    step = SynthStep(shapes, dev)
    host = {k: v.pin_memory() for k, v in synthetic_inputs(shapes, seed=100 + rank).items()}
    h2d = step.load_inputs(host)
    torch.cuda.synchronize()

    # parameter gradients (reverb.noise/decay/wet) are averaged across ranks: the only collective
    from ddsp_pytorch_b200.distributed import GradBucket
    params = list(step.reverb.parameters())
    n_param = sum(p.numel() for p in params)
    bucket = GradBucket([p.shape for p in params], dev)

    def allreduce_grads():
        if dist is not None:
            bucket.all_reduce_mean(step.grads[3:])

    c0 = lib.ddsp_b200_launch_count()
    step.run()
    launches = int(lib.ddsp_b200_launch_count() - c0)
    use_graph = not args.no_graph
    graph_note = "CUDA graph replay"
    if use_graph:
        try:
            step.capture(forward_only=False)
            step.capture(forward_only=True)
        except Exception as e:                      # stay measurable: fall back to eager launches
            use_graph = False
            graph_note = f"eager (graph capture failed: {type(e).__name__})"
            torch.cuda.synchronize()
    else:
        graph_note = "eager"
    run_step = (lambda: step.replay()) if use_graph else step.run
    run_fwd = (lambda: step.replay(True)) if use_graph else step.run_forward

    flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device=dev)      # > 126 MB of L2
    flush = lambda: flush_buf.zero_()

    def full_step():
        run_step()
        allreduce_grads()

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        flush()
        full_step()
    barrier()
    starts = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    ends = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    with ClockSampler(local) as clocks:
        t_wall = time.perf_counter()
        for k in range(args.steps):
            flush()
            starts[k].record()
            full_step()
            ends[k].record()
        barrier()
        t_wall = time.perf_counter() - t_wall
    total_ms = sum(a.elapsed_time(b) for a, b in zip(starts, ends))
    tt = torch.tensor([total_ms], device=dev, dtype=torch.float64)
    if dist is not None:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    total_ms = float(tt)
    samples_per_step = args.global_batch * shapes.samples           # whole job, all ranks
    value = samples_per_step * args.steps / (total_ms * 1e-3)

    # forward only (the "fwd" half of the metric)
    fwd_ms = event_time(run_fwd, max(10, args.steps // 4), 3, flush)
    tf = torch.tensor([fwd_ms], device=dev, dtype=torch.float64)
    if dist is not None:
        dist.all_reduce(tf, op=dist.ReduceOp.MAX)
    fwd_ms = float(tf)

    # end to end through the public API with HOST buffers: every step copies its inputs from pinned
    # host memory (the copy of step k+1 is queued before step k is waited for, as a prefetching data
    # loader does), runs, and reads the loss back to the host.
    barrier()
    e2e_steps = max(10, args.steps // 2)
    hosts = [host, {k: v.clone().pin_memory() for k, v in host.items()}]

    def e2e_loop(n):
        step.prefetch(hosts[0])
        last = 0.0
        for k in range(n):
            step.step_prefetched()
            allreduce_grads()
            step.prefetch(hosts[(k + 1) & 1])                    # H2D of the next batch, overlapped
            last = float(step.loss.detach())                     # D2H + sync of this step's result
        return last

    e2e_loop(3)
    barrier()
    t0 = time.perf_counter()
    loss_host = e2e_loop(e2e_steps)
    barrier()
    e2e_s = time.perf_counter() - t0
    te = torch.tensor([e2e_s], device=dev, dtype=torch.float64)
    if dist is not None:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = samples_per_step * e2e_steps / float(te)

    line = None
    if rank == 0:
        peaks = {"hbm_gbs": 6650.0, "source": "fallback (B200_PROFILING.md)"}
        pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
        if os.path.exists(pk):
            with open(pk) as f:
                peaks = json.load(f)
            peaks["source"] = "MEASURED_PEAKS.json (burst copy bandwidth)"
        kernels, roof = [], None
        if not args.skip_kernels:
            kernels = kernel_table(step, shapes, flush, peaks)
            top = max(kernels, key=lambda r: r["ms"])
            roof = {"kernel": top["kernel"], "bound": top["bound"] if top["bound"] == "hbm" else "tensor",
                    "achieved": top["achieved"], "peak": top["peak"], "unit": top["unit"], "frac": top["frac"],
                    "traffic": K4L_DRAM_TRAFFIC if top["kernel"].startswith("K4L") else None, "ms": top["ms"],
                    "peak_source": peaks["source"]}
            if top["kernel"].startswith("K4L"):
                roof["traffic_note"] = ("dram__bytes_read+write summed over the 6 per-scale launches of one step, "
                                        "profiles/r01_ncu_mss_scale_reg_full.csv (ncu flushes caches between "
                                        "launches, so each scale re-reads rec+target = 32.8 MB; algorithmic 49 MB)")
            if top["bound"] != "hbm":
                roof["bound_detail"] = "fp32 FMA pipe, not tensor cores (no GEMM on this path)"
            if "fp32" in top:
                roof["fp32"] = top["fp32"]
        cpu = None
        if world == 1 and not args.skip_cpu:
            cshapes, times, cores = time_cpu_port(args.cpu_batch, 3, 1)
            best = min(times)
            cpu = {"value": args.cpu_batch * cshapes.samples / best, "unit": UNIT, "cores": cores, "kind": "port",
                   "sample": f"{args.cpu_batch} of 64 voices, same shapes, float32 torch CPU, best of 3",
                   "ms_per_sample_step": best * 1e3}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": total_ms / args.steps, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD_NAME, "global_batch": args.global_batch, "per_gpu_batch": shapes.batch,
                       "samples_per_voice": shapes.samples, "parallelism": f"voices sharded x{world}, NCCL "
                       f"all-reduce of {n_param} reverb-parameter grads" if world > 1 else "single GPU",
                       "l2": "256 MiB memset between timed steps (outside the event pairs)",
                       "launch": graph_note},
            "fwd": {"value": samples_per_step / (fwd_ms * 1e-3), "unit": UNIT, "ms_per_step": fwd_ms},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d * world, "d2h_bytes_per_step": 4 * world,
                    "ms_per_step": 1e3 * float(te) / e2e_steps, "loss": loss_host,
                    "how": "pinned host -> H2D (copy stream, next batch prefetched during the step) -> D2D into "
                           "the graph inputs -> step -> loss.item(); wall clock between synchronisations"},
            "gpu_launches": launches, "roofline": roof, "kernels": kernels, "cpu_baseline": cpu,
            "clocks": clocks.summary(), "wall_s_timed_loop": t_wall,
        }
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def main():
    args = parse()
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl b200 needs a CUDA device: there is no CPU fallback for the kernels")
    run_b200(args, rank, world)


if __name__ == "__main__":
    main()
