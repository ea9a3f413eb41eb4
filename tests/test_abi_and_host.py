"""CPU: the C-ABI library loads and exports every symbol include/flye_b200.h declares (no compute without a GPU),
and the host-side helpers behave like the reference's host code."""
import ctypes
import os
import re

import numpy as np
import pytest

import parity_util as pu
import flye_b200 as fb


def test_library_exports_every_declared_symbol(built):
    header = open(os.path.join(pu.ROOT, "include", "flye_b200.h")).read()
    declared = set(re.findall(r"\b(fg_[a-z_0-9]+)\s*\(", header))
    assert declared == set(fb.SYMBOLS), declared ^ set(fb.SYMBOLS)
    lib = fb.load_lib()
    for sym in declared:
        assert getattr(lib, sym) is not None


def test_no_cpu_fallback(built):
    """Without a CUDA device the engine must refuse to exist (the product path never computes on the CPU)."""
    lib = fb.load_lib()
    ctx = ctypes.c_void_p()
    rc = lib.fg_ctx_create(0, ctypes.byref(ctx))
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = rc == 0
    if not has_gpu:
        assert rc != 0 and not ctx.value
        with pytest.raises(fb.FlyeB200Error):
            fb.Engine(0)
    else:
        assert rc == 0
        lib.fg_ctx_destroy(ctx)


def test_pack_reads_layout():
    """DnaSequence packing: base j at bits 2*(j%32) of word j/32, A0 C1 G2 T3 (sequence.h:54-69,166-173)"""
    rng = np.random.default_rng(3)
    reads = [bytes(rng.choice(list(b"ACGT"), n).astype(np.uint8)) for n in (1, 31, 32, 33, 64, 100, 257)]
    packed, offs, lens = fb.pack_reads(reads)
    code = {65: 0, 67: 1, 71: 2, 84: 3}
    for i, r in enumerate(reads):
        assert lens[i] == len(r) and offs[i + 1] - offs[i] == (len(r) + 31) // 32
        for j, ch in enumerate(r):
            w = int(packed[int(offs[i]) + j // 32])
            assert (w >> (2 * (j % 32))) & 3 == code[ch]
    with pytest.raises(ValueError):
        fb.pack_reads([b"ACGN"])


def test_read_fasta_min_length(tmp_path):
    p = tmp_path / "x.fasta"
    p.write_text(">a\nACGT\nAC\n>b desc\nACG\n>c\n" + "A" * 10 + "\n")
    assert fb.read_fasta(str(p), 0) == [b"ACGTAC", b"ACG", b"A" * 10]
    assert fb.read_fasta(str(p), 6) == [b"A" * 10]          # strictly longer (sequence_container.cpp:102)


def test_cfg_and_estimate_helpers():
    cfg = pu.load_cfg(os.path.join(pu.CFG_DIR, "raw_reads.cfg"))
    assert cfg["maximum_jump"] == 1500 and cfg["use_minimizers"] == 0 and abs(cfg["meta_read_top_kmer_rate"] - np.float32(0.4)) < 1e-9
    a, b = pu.libc_rand_ids(1000), pu.libc_rand_ids(1000)
    assert a == b and len(a) == 1000 and a[0] == 1804289383 % 1000    # glibc rand() after srand(1)
    assert pu.median_f32([3.0, 1.0, 2.0, 4.0]) == 3.0                  # utils.h quantile(vec, 50): index size*50/100
