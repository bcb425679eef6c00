import torch, time
N=64; sz=752*480
h=torch.empty((N,sz),dtype=torch.uint8).pin_memory(); d=torch.empty((N,sz),dtype=torch.uint8,device='cuda')
s=torch.cuda.Stream()
for name,fn in [("64x360KB one stream", lambda: [d[i].copy_(h[i],non_blocking=True) for i in range(N)]),
                ("1x23MB", lambda: d.copy_(h,non_blocking=True))]:
    with torch.cuda.stream(s):
        fn(); torch.cuda.synchronize()
        t=time.perf_counter()
        for _ in range(20): fn()
        torch.cuda.synchronize(); dt=(time.perf_counter()-t)/20
    print(name, "%.3f ms  %.1f GB/s"%(dt*1e3, N*sz/dt/1e9))
# D2H 3.3MB
dd=torch.empty(3300000,dtype=torch.uint8,device='cuda'); hh=torch.empty(3300000,dtype=torch.uint8).pin_memory()
hh.copy_(dd,non_blocking=True); torch.cuda.synchronize(); t=time.perf_counter()
for _ in range(20): hh.copy_(dd,non_blocking=True)
torch.cuda.synchronize(); dt=(time.perf_counter()-t)/20; print("D2H 3.3MB %.3f ms"%(dt*1e3))
# multi streams H2D
ss=[torch.cuda.Stream() for _ in range(16)]
def multi():
    for i in range(N):
        with torch.cuda.stream(ss[i%16]): d[i].copy_(h[i],non_blocking=True)
multi(); torch.cuda.synchronize(); t=time.perf_counter()
for _ in range(20): multi()
torch.cuda.synchronize(); dt=(time.perf_counter()-t)/20; print("64x360KB 16 streams %.3f ms %.1f GB/s"%(dt*1e3,N*sz/dt/1e9))
import os; print("cpus", os.cpu_count())
