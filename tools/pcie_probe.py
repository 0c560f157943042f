import torch, time
dev=torch.device('cuda',0)
for mb in (64, 199, 1024):
    h=torch.empty(mb*1024*1024,dtype=torch.uint8).pin_memory()
    d=torch.empty_like(h,device=dev)
    for _ in range(2): d.copy_(h,non_blocking=True)
    torch.cuda.synchronize(); t=time.perf_counter()
    for _ in range(5): d.copy_(h,non_blocking=True)
    torch.cuda.synchronize(); dt=(time.perf_counter()-t)/5
    print('H2D',mb,'MB',round(mb*1.048576/1000/dt,2),'GB/s')
    torch.cuda.synchronize(); t=time.perf_counter()
    for _ in range(5): h.copy_(d,non_blocking=True)
    torch.cuda.synchronize(); dt=(time.perf_counter()-t)/5
    print('D2H',mb,'MB',round(mb*1.048576/1000/dt,2),'GB/s')
