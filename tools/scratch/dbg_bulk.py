import sys, os, torch, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import ddsp_pytorch_b200
from ddsp_pytorch_b200.workloads import BulkRenderer
r = BulkRenderer(128, "cuda")
ops = torch.ops.ddsp_b200
i = r.inp
def t(fn, n=5):
    fn(); torch.cuda.synchronize()
    a,b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b)/n
print("raw_fwd", t(lambda: ops.harmonic_raw_fwd(i["amp_raw"], i["dist_raw"], i["pitch"], r.BS, float(r.SR), None)))
def two():
    _, _, w = ops.harmonic_controls_fwd(i["amp_raw"], i["dist_raw"], i["pitch"], float(r.SR), True)
    return ops.harmonic_fwd(i["pitch"], w, r.BS, float(r.SR), None)
print("two-launch", t(two))
print("chunk", t(lambda: r.render_chunk()))
print("chunk nodraw", t(lambda: r.render_chunk(False)))
