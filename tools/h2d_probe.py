import torch,time
for pin in (True,False):
    h=torch.empty(1<<30,dtype=torch.uint8)
    if pin: h=h.pin_memory()
    d=torch.empty(1<<30,dtype=torch.uint8,device='cuda')
    for _ in range(2): d.copy_(h,non_blocking=True); torch.cuda.synchronize()
    t=time.perf_counter()
    for _ in range(3): d.copy_(h,non_blocking=True)
    torch.cuda.synchronize(); dt=(time.perf_counter()-t)/3
    print('H2D pinned=%s %.1f GB/s'%(pin,(1<<30)/dt/1e9))
    t=time.perf_counter()
    for _ in range(3): h.copy_(d,non_blocking=True)
    torch.cuda.synchronize(); dt=(time.perf_counter()-t)/3
    print('D2H pinned=%s %.1f GB/s'%(pin,(1<<30)/dt/1e9))
