"""Does torch's symmetric memory one-shot all-reduce work between the ranks of this box?  (2+ GPUs, torchrun)
Prints where it hangs (faulthandler) instead of waiting for the outer timeout."""
import faulthandler, os, sys, time
faulthandler.dump_traceback_later(45, exit=True)
import torch, torch.distributed as dist
local = int(os.environ["LOCAL_RANK"]); torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
import torch.distributed._symmetric_memory as symm
print(local, "backend", getattr(symm, "get_backend", lambda d: "?")(dev) if hasattr(symm, "get_backend") else "n/a", flush=True)
t = symm.empty(16002, dtype=torch.float32, device=dev)
print(local, "empty ok", flush=True)
h = symm.rendezvous(t, dist.group.WORLD)
print(local, "rendezvous ok", type(h).__name__, flush=True)
t.fill_(local + 1.0)
out = torch.empty_like(t)
dist.barrier(); torch.cuda.synchronize()
torch.ops.symm_mem.one_shot_all_reduce_out(t, "sum", dist.group.WORLD.group_name, out)
torch.cuda.synchronize()
print(local, "one-shot ok", float(out[0]), flush=True)
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    torch.ops.symm_mem.one_shot_all_reduce_out(t, "sum", dist.group.WORLD.group_name, out)
for _ in range(3):
    g.replay()
torch.cuda.synchronize()
print(local, "graph ok", float(out[0]), flush=True)
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(100):
    g.replay()
b.record(); torch.cuda.synchronize()
print(local, "us per one-shot all-reduce (graph replay)", a.elapsed_time(b) * 10, flush=True)
x = torch.zeros(16002, device=dev)
a.record()
for _ in range(100):
    dist.all_reduce(x)
b.record(); torch.cuda.synchronize()
print(local, "us per NCCL all-reduce (eager)", a.elapsed_time(b) * 10, flush=True)
dist.destroy_process_group()
