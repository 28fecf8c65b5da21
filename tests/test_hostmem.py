"""Host-memory helper of the e2e leg (dreamgaussian_b200/hostmem.py): cpulist parsing and that numa_local restores the
calling thread's affinity.  CPU only."""
import os

import pytest

import helpers  # noqa: F401
from dreamgaussian_b200 import hostmem


def test_cpulist_parsing():
    assert hostmem._parse_cpulist("0-3,8,10-11\n") == {0, 1, 2, 3, 8, 10, 11}
    assert hostmem._parse_cpulist("") == set() and hostmem._parse_cpulist("5") == {5}


@pytest.mark.skipif(not hasattr(os, "sched_getaffinity"), reason="needs sched_getaffinity")
def test_numa_local_restores_affinity(monkeypatch):
    before = os.sched_getaffinity(0)
    some = set(sorted(before)[: max(1, len(before) // 2)])
    monkeypatch.setattr(hostmem, "gpu_local_cpus", lambda device: some)
    with hostmem.numa_local("cuda:0") as narrowed:
        inside = os.sched_getaffinity(0)
        assert inside == (some if some != before else before) and narrowed == (some != before)
    assert os.sched_getaffinity(0) == before
    monkeypatch.setattr(hostmem, "gpu_local_cpus", lambda device: set())          # unknown topology: no-op
    with hostmem.numa_local("cuda:0") as narrowed:
        assert narrowed is False and os.sched_getaffinity(0) == before
