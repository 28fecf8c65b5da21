"""Does the NUMA node of a pinned buffer explain the H2D bandwidth spread seen on the B200 boxes?  For a set of pinned
buffers (allocated with and without dreamgaussian_b200.hostmem.numa_local) print the node of their pages
(/proc/self/numa_maps) next to the measured H2D / D2H bandwidth."""
import os, re, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from dreamgaussian_b200 import hostmem

dev = torch.device("cuda", 0)


def nodes_of(t):
    addr = t.data_ptr()
    best = None
    try:
        for line in open("/proc/self/numa_maps"):
            a = int(line.split()[0], 16)
            if a <= addr and (best is None or a > best[0]):
                best = (a, line)
    except OSError:
        return "?"
    return " ".join(re.findall(r"N\d+=\d+", best[1])) if best else "?"


def bw(fn, nbytes, it=30):
    for _ in range(3): fn()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(it): fn()
    torch.cuda.synchronize(); return nbytes * it / (time.perf_counter() - t0) / 1e9


print("GPU-local CPUs:", sorted(hostmem.gpu_local_cpus(dev))[:4], "... count", len(hostmem.gpu_local_cpus(dev)), " process affinity:", len(os.sched_getaffinity(0)))
print(open("/proc/self/status").read().split("Mems_allowed_list:")[1].split()[0], "= Mems_allowed_list")
for mode in ("default", "numa_local", "default", "numa_local"):
    for mb in (4, 19.2, 64):
        n = int(mb * 1e6)
        if mode == "numa_local":
            h = hostmem.pinned_empty((n,), torch.uint8, dev)
        else:
            h = torch.empty((n,), dtype=torch.uint8).pin_memory(); h.zero_()
        d = torch.empty((n,), dtype=torch.uint8, device=dev)
        up = bw(lambda: d.copy_(h, non_blocking=True), n)
        down = bw(lambda: h.copy_(d, non_blocking=True), n)
        print("%-10s %5.1f MB  H2D %5.1f GB/s  D2H %5.1f GB/s  pages: %s" % (mode, mb, up, down, nodes_of(h)), flush=True)
        del h, d

print("same buffer over time (fresh 19.2 MB allocations, kept alive):")
keep = []
for i in range(8):
    nb = 19_200_000
    hb = torch.empty((nb,), dtype=torch.uint8).pin_memory(); hb.zero_(); keep.append(hb)
    db = torch.empty((nb,), dtype=torch.uint8, device=dev)
    a = bw(lambda: db.copy_(hb, non_blocking=True), nb)
    b = bw(lambda: db.copy_(hb, non_blocking=True), nb)
    print("  buffer %d: H2D %5.1f GB/s, again %5.1f GB/s   pages %s" % (i, a, b, nodes_of(hb)), flush=True)
