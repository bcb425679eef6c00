import torch, time
sz=752*480; N=64
h=torch.empty((N,sz),dtype=torch.uint8).pin_memory(); d=torch.empty((N,sz),dtype=torch.uint8,device='cuda')
ss=[torch.cuda.Stream() for _ in range(32)]
def run(fn, name, reps=30):
    fn(); torch.cuda.synchronize()
    e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
    t=time.perf_counter(); e0.record()
    for _ in range(reps): fn()
    for s in ss: torch.cuda.current_stream().wait_stream(s)
    e1.record(); torch.cuda.synchronize(); dt=(time.perf_counter()-t)/reps
    print("%-40s host %.3f ms/batch  dev %.3f ms/batch  %.1f GB/s"%(name, dt*1e3, e0.elapsed_time(e1)/reps, N*sz/dt/1e9))
def per_img():
    for i in range(N):
        with torch.cuda.stream(ss[i%32]): d[i].copy_(h[i],non_blocking=True)
def per_img_1s():
    with torch.cuda.stream(ss[0]):
        for i in range(N): d[i].copy_(h[i],non_blocking=True)
def chunks8():
    for i in range(8):
        with torch.cuda.stream(ss[i]): d[i*8:(i+1)*8].copy_(h[i*8:(i+1)*8],non_blocking=True)
def chunks2():
    for i in range(2):
        with torch.cuda.stream(ss[i]): d[i*32:(i+1)*32].copy_(h[i*32:(i+1)*32],non_blocking=True)
def one():
    with torch.cuda.stream(ss[0]): d.copy_(h,non_blocking=True)
run(per_img,"64 x 360KB on 32 streams")
run(per_img_1s,"64 x 360KB on 1 stream")
run(chunks8,"8 x 2.9MB on 8 streams")
run(chunks2,"2 x 11.5MB on 2 streams")
run(one,"1 x 23MB")
# D2H packets 32 x 105KB
hp=torch.empty((32,104544),dtype=torch.uint8).pin_memory(); dp=torch.empty((32,104544),dtype=torch.uint8,device='cuda')
def d2h():
    for i in range(32):
        with torch.cuda.stream(ss[i]): hp[i].copy_(dp[i],non_blocking=True)
run(d2h,"D2H 32 x 105KB on 32 streams")
def both():
    per_img(); d2h()
run(both,"H2D 64x360KB + D2H 32x105KB")
