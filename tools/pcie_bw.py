"""PCIe copy bandwidth of the box (what bounds bench.py's e2e number): pinned vs write-combined host memory, 1..4
concurrent copy streams, H2D / D2H / both directions."""
import ctypes, time, sys
import numpy as np, torch

dev = torch.device("cuda")
n = 23_600_000
rt = ctypes.CDLL("libcudart.so.12") if True else None


def wc_buffer(nbytes):
    p = ctypes.c_void_p()
    rc = rt.cudaHostAlloc(ctypes.byref(p), ctypes.c_size_t(nbytes), ctypes.c_uint(0x04))     # cudaHostAllocWriteCombined
    assert rc == 0, rc
    arr = np.ctypeslib.as_array((ctypes.c_uint8 * nbytes).from_address(p.value))
    return torch.from_numpy(arr)


def t(fn, it=40):
    for _ in range(3): fn()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(it): fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / it


pinned = torch.empty(n, dtype=torch.uint8).pin_memory()
pinned2 = torch.empty(n, dtype=torch.uint8).pin_memory()
d = torch.empty(n, dtype=torch.uint8, device=dev); d2 = torch.empty(n, dtype=torch.uint8, device=dev)
streams = [torch.cuda.Stream() for _ in range(4)]
print("H2D pinned 1 stream   %.1f GB/s" % (n / t(lambda: d.copy_(pinned, non_blocking=True)) / 1e9))
print("D2H pinned 1 stream   %.1f GB/s" % (n / t(lambda: pinned2.copy_(d2, non_blocking=True)) / 1e9))
for k in (2, 4):
    per = n // k
    def multi():
        for j in range(k):
            with torch.cuda.stream(streams[j]):
                d[j * per:(j + 1) * per].copy_(pinned[j * per:(j + 1) * per], non_blocking=True)
    print("H2D pinned %d streams  %.1f GB/s" % (k, n / t(multi) / 1e9))
try:
    wc = wc_buffer(n)
    print("H2D write-combined    %.1f GB/s" % (n / t(lambda: d.copy_(wc, non_blocking=True)) / 1e9))
    t0 = time.perf_counter(); wc[: n // 4].copy_(pinned[: n // 4]); dt = time.perf_counter() - t0
    print("   (CPU fill of WC memory: %.1f GB/s)" % (n / 4 / dt / 1e9))
except Exception as e:  # noqa: BLE001
    print("write-combined alloc failed:", e)
def both():
    with torch.cuda.stream(streams[0]): d.copy_(pinned, non_blocking=True)
    with torch.cuda.stream(streams[1]): pinned2.copy_(d2, non_blocking=True)
c = t(both)
print("both directions: %.3f ms per pair -> %.1f GB/s each" % (c * 1e3, n / c / 1e9))
for sz in (1 << 20, 4 << 20, 16 << 20, 64 << 20, 256 << 20):
    h = torch.empty(sz, dtype=torch.uint8).pin_memory(); dd = torch.empty(sz, dtype=torch.uint8, device=dev)
    print("H2D %4d MiB  %.1f GB/s" % (sz >> 20, sz / t(lambda: dd.copy_(h, non_blocking=True), 20) / 1e9))
import subprocess
print(subprocess.run(["nvidia-smi", "--query-gpu=pcie.link.gen.current,pcie.link.width.current,pcie.link.gen.max,pcie.link.width.max", "--format=csv"],
                     capture_output=True, text=True).stdout)
print(subprocess.run("numactl -H 2>/dev/null | head -5; nvidia-smi topo -m 2>/dev/null | head -8", shell=True, capture_output=True, text=True).stdout)
