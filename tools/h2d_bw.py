import torch, time
dev=torch.device('cuda')
for mb in (16, 64, 512):
    n=mb*1024*1024//4
    h=torch.empty(n,dtype=torch.float32).pin_memory(); d=torch.empty(n,dtype=torch.float32,device=dev)
    for _ in range(3): d.copy_(h,non_blocking=True)
    torch.cuda.synchronize(); t0=time.perf_counter()
    for _ in range(10): d.copy_(h,non_blocking=True)
    torch.cuda.synchronize(); dt=time.perf_counter()-t0
    print(mb,'MB H2D GB/s', 10*mb/1024/dt)
    t0=time.perf_counter()
    for _ in range(10): h.copy_(d,non_blocking=True)
    torch.cuda.synchronize(); dt=time.perf_counter()-t0
    print(mb,'MB D2H GB/s', 10*mb/1024/dt)
