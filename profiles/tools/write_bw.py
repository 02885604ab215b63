"""Write-only / copy HBM bandwidth of this GPU with plain torch kernels (context for the roofline of a write-dominated kernel)."""
import torch, time
dev='cuda'
def t(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1)/n*1e-3
for mb in (105, 437, 1024, 4096):
    n = mb*1024*1024//8
    a = torch.empty(n, dtype=torch.float64, device=dev); b = torch.empty(n, dtype=torch.float64, device=dev)
    tf = t(lambda: a.fill_(1.5)); tc = t(lambda: b.copy_(a)); 
    bufs=[torch.empty(n, dtype=torch.float64, device=dev) for _ in range(4)] if mb<=437 else None
    if bufs:
        i=[0]
        def rot():
            bufs[i[0]%4].fill_(2.0); i[0]+=1
        tr=t(rot, 40)
    else: tr=float('nan')
    print(f"{mb} MB: fill {n*8/tf/1e9:.0f} GB/s ({tf*1e6:.1f} us)  rotating-fill {n*8/tr/1e9 if tr==tr else 0:.0f} GB/s  copy {2*n*8/tc/1e9:.0f} GB/s")
    del a,b,bufs
