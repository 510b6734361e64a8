import torch, time
n=1<<30
h=torch.empty(n,dtype=torch.uint8,pin_memory=True); d=torch.empty(n,dtype=torch.uint8,device='cuda')
for name,f in (('H2D',lambda: d.copy_(h,non_blocking=True)),('D2H',lambda: h.copy_(d,non_blocking=True))):
    f(); torch.cuda.synchronize(); t=time.perf_counter()
    for _ in range(5): f()
    torch.cuda.synchronize(); dt=(time.perf_counter()-t)/5
    print(name, round(n/dt/1e9,1),'GB/s', round(dt*1e3,2),'ms')
h2=torch.empty(n,dtype=torch.uint8,pin_memory=True); d2=torch.empty(n,dtype=torch.uint8,device='cuda')
s1,s2=torch.cuda.Stream(),torch.cuda.Stream()
torch.cuda.synchronize(); t=time.perf_counter()
for _ in range(5):
    with torch.cuda.stream(s1): d.copy_(h,non_blocking=True)
    with torch.cuda.stream(s2): h2.copy_(d2,non_blocking=True)
torch.cuda.synchronize(); dt=(time.perf_counter()-t)/5
print('bidir', round(2*n/dt/1e9,1),'GB/s total', round(dt*1e3,2),'ms')
