import torch, time
n=1559480000
d=torch.empty(n,dtype=torch.uint8,device='cuda'); h=torch.empty(n,dtype=torch.uint8,pin_memory=True)
for _ in range(2): h.copy_(d,non_blocking=True); torch.cuda.synchronize()
t0=time.perf_counter()
for _ in range(5): h.copy_(d,non_blocking=True)
torch.cuda.synchronize(); dt=(time.perf_counter()-t0)/5
print(f"D2H pinned {n/dt/1e9:.1f} GB/s  ({dt*1e3:.2f} ms for 1.56 GB)")
# chunked 64 MiB copies on a side stream
s=torch.cuda.Stream(); c=64<<20
t0=time.perf_counter()
with torch.cuda.stream(s):
    for _ in range(5):
        for o in range(0,n,c): h[o:o+c].copy_(d[o:o+c],non_blocking=True)
s.synchronize(); dt=(time.perf_counter()-t0)/5
print(f"D2H pinned 64MiB chunks {n/dt/1e9:.1f} GB/s ({dt*1e3:.2f} ms)")
