import torch, time
dev=torch.device("cuda")
n=23600000
h=torch.empty(n,dtype=torch.uint8).pin_memory(); h2=torch.empty(n,dtype=torch.uint8).pin_memory()
d=torch.empty(n,dtype=torch.uint8,device=dev); d2=torch.empty(n,dtype=torch.uint8,device=dev)
s1,s2=torch.cuda.Stream(),torch.cuda.Stream()
def t(fn,it=50):
    torch.cuda.synchronize(); t0=time.perf_counter()
    for _ in range(it): fn()
    torch.cuda.synchronize(); return (time.perf_counter()-t0)/it
a=t(lambda: d.copy_(h,non_blocking=True)); b=t(lambda: h2.copy_(d2,non_blocking=True))
def both():
    with torch.cuda.stream(s1): d.copy_(h,non_blocking=True)
    with torch.cuda.stream(s2): h2.copy_(d2,non_blocking=True)
c=t(both)
print("H2D %.1f GB/s  D2H %.1f GB/s  both: %.3f ms per pair -> %.1f GB/s each" % (n/a/1e9, n/b/1e9, c*1e3, n/c/1e9))
