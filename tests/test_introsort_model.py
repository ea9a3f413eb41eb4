"""CPU: the host build of the device introsort (flye_b200/csrc/introsort_warp.cuh, 32 lanes run as loops) must
reproduce libstdc++ std::sort's permutation — ties, presorted / reversed / constant inputs, median-of-3 killers that
force the heap-sort fallback — in one-pass mode, in two-level (task) mode and for the sequential variant."""
import os
import subprocess

import pytest

import parity_util as pu

BIN = os.path.join(pu.ROOT, "tests", "cpu_models", "_bin", "introsort_check")


@pytest.mark.parametrize("small", [0, 1024, 64, 17, -1])
def test_warp_introsort_host_build_equals_std_sort(built, small):
    r = subprocess.run([BIN, "120", str(7 + abs(small)), str(small)], stdout=subprocess.PIPE, text=True, timeout=600)
    assert r.returncode == 0, r.stdout
    assert r.stdout.startswith("OK")
    if small >= 0:
        assert "heapsorts=0" not in r.stdout   # the depth-limit fallback was exercised


@pytest.mark.parametrize("variant", [1, 2, 3])
@pytest.mark.parametrize("small", [0, 512, 64])
def test_shared_memory_variants_equal_std_sort(built, small, variant):
    """the shared-memory kernel's variants: rank-table partition (1), + stable leaf pass by ranking with 16- (2) or
    32-element (3) warp-partition threshold"""
    r = subprocess.run([BIN, "100", str(11 + small + variant), str(small), str(variant)], stdout=subprocess.PIPE, text=True, timeout=600)
    assert r.returncode == 0, r.stdout
    assert r.stdout.startswith("OK")


def test_tie_following_emulation_equals_std_sort(built):
    """The hit sort's fast path (overlap.cu: rangeIsTieFree): introsort is only emulated inside the ranges whose sorted ranks
    contain a duplicated key; every other range is copied from the stable-sorted array.  The model in introsort_check (run for
    small = 0) must reproduce std::sort on tie-heavy, few-tie and heap-sort-fallback inputs, and must actually skip ranges.
    The same run checks the model of sortHugeKernel's partition step (stop lists, s by a monotone search, parallel swaps, cut)
    against the literal __move_median_to_first + __unguarded_partition on every random and killer array."""
    r = subprocess.run([BIN, "400", "23", "0"], stdout=subprocess.PIPE, text=True, timeout=600)
    assert r.returncode == 0, r.stdout
    assert r.stdout.startswith("OK")
    skipped = int(r.stdout.split("skipped=")[1].split()[0])
    followed = int(r.stdout.split("followed=")[1].split()[0])
    assert skipped > 1000 and followed > 1000
