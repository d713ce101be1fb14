import torch, time
torch.cuda.init()
sizes=[25600, 2560000, 1664000, 25600, 4096000, 4096000]
host=[torch.randn(n).pin_memory() for n in sizes]
dev=[torch.empty(n, device='cuda') for n in sizes]
big=torch.randn(sum(sizes)).pin_memory(); dbig=torch.empty(sum(sizes), device='cuda')
s=torch.cuda.Stream()
def run(fn, n=20):
    fn(); torch.cuda.synchronize()
    t0=time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize()
    return (time.perf_counter()-t0)/n*1e3
def six():
    with torch.cuda.stream(s):
        for h,d in zip(host,dev): d.copy_(h, non_blocking=True)
def one():
    with torch.cuda.stream(s): dbig.copy_(big, non_blocking=True)
nbytes=sum(sizes)*4
t6=run(six); t1=run(one)
print(f"6 copies: {t6:.3f} ms = {nbytes/t6/1e6:.1f} GB/s ; 1 copy: {t1:.3f} ms = {nbytes/t1/1e6:.1f} GB/s")
dev2=[torch.empty_like(d) for d in dev]
def d2d():
    for a,b in zip(dev2,dev): a.copy_(b, non_blocking=True)
print(f"6 D2D copies: {run(d2d):.3f} ms")
