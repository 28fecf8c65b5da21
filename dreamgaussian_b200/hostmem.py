"""NUMA-local pinned host memory for the host<->device legs of a step.

A pinned buffer lives on the NUMA node of the thread that allocates it; when that is not the node the GPU's PCIe root
hangs off, H2D copies cross the socket interconnect.  `numa_local(device)` narrows the calling thread's CPU affinity to the
GPU-local CPUs for the duration of the allocation and restores it afterwards (compute threads — e.g. an OpenMP CPU
baseline — keep every core).  Measured on the shared B200 boxes (tools/pcie_numa.py): the H2D rate of one and the same
pinned buffer swings between 17 and 55 GB/s over seconds whatever its node (D2H stays at ~56 GB/s), so placement is good
practice here, not a fix — bench.py's e2e number inherits that swing."""
import contextlib
import os

import torch


def _parse_cpulist(text):
    cpus = set()
    for part in text.strip().split(","):
        if not part:
            continue
        if "-" in part:
            a, b = part.split("-")
            cpus.update(range(int(a), int(b) + 1))
        else:
            cpus.add(int(part))
    return cpus


def gpu_local_cpus(device):
    """CPUs on the NUMA node of `device`'s PCIe root (sysfs local_cpulist, NVML as a fallback); empty set if unknown."""
    idx = torch.device(device).index
    if idx is None:
        idx = torch.cuda.current_device()
    try:
        p = torch.cuda.get_device_properties(idx)
        bdf = "%04x:%02x:%02x.0" % (p.pci_domain_id, p.pci_bus_id, p.pci_device_id)
        with open("/sys/bus/pci/devices/%s/local_cpulist" % bdf) as f:
            cpus = _parse_cpulist(f.read())
        if cpus:
            return cpus
    except (OSError, AttributeError, ValueError):
        pass
    try:
        import pynvml
        pynvml.nvmlInit()
        p = torch.cuda.get_device_properties(idx)
        h = pynvml.nvmlDeviceGetHandleByPciBusId(("%08x:%02x:%02x.0" % (p.pci_domain_id, p.pci_bus_id, p.pci_device_id)).encode())
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64)
        return {64 * w + b for w, word in enumerate(words) for b in range(64) if (word >> b) & 1}
    except Exception:  # noqa: BLE001  (NVML absent or refused: no hint)
        return set()


@contextlib.contextmanager
def numa_local(device):
    """Within the block the calling thread runs on (and therefore allocates from) the GPU-local NUMA node."""
    if not hasattr(os, "sched_getaffinity"):
        yield False
        return
    before = os.sched_getaffinity(0)
    target = gpu_local_cpus(device) & before
    if not target or target == before:
        yield False
        return
    os.sched_setaffinity(0, target)
    try:
        yield True
    finally:
        os.sched_setaffinity(0, before)


def pinned_like(t, device):
    """Pinned host tensor of t's shape/dtype, allocated and first-touched on the GPU-local NUMA node, holding t's values."""
    with numa_local(device):
        h = torch.empty(t.shape, dtype=t.dtype).pin_memory()
        h.copy_(t)
    return h


def pinned_empty(shape, dtype, device):
    with numa_local(device):
        h = torch.empty(shape, dtype=dtype).pin_memory()
        h.zero_()
    return h
