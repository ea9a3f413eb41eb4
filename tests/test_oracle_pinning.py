"""CPU, build container only: the CPU restatement against the UNMODIFIED reference built by oracle/Makefile on a
freshly simulated read set (byte-identical dumps incl. the rand()-driven divergence threshold)."""
import os

import pytest

import parity_util as pu

needs_ref = pytest.mark.skipif(not (os.path.exists(pu.REF_HARNESS) and os.path.isdir("/root/reference/src")),
                               reason="reference sources not present (GPU box): pinned by tests/golden instead")


@needs_ref
@pytest.mark.parametrize("cfg,k,sim,opts", [
    ("raw_reads.cfg", 15, dict(genome_len=50000, coverage=10, seed=101), ["--dump-index"]),
    ("raw_reads.cfg", 15, dict(genome_len=40000, coverage=10, seed=102), ["--both-strands", "--force-local", "--max-overlaps", "3"]),
    ("hifi.cfg", None, dict(genome_len=30000, coverage=8, mean_len=7000, shape=20, error=0.005, seed=103), ["--dump-index", "--no-estimate"]),
    # partitionBadMappings (the repeat stage's detector): reads at the divergence threshold, so that nearly every primary overlap
    # is rejected and cut into its low-divergence stretches by checkIdyAndTrim (KSW2) — restated in oracle/restate/ksw_restate.cpp
    ("hifi.cfg", None, dict(genome_len=12000, coverage=6, mean_len=4000, shape=20, error=0.010, seed=212), ["--partition-bad"]),
    ("raw_reads.cfg", 15, dict(genome_len=40000, coverage=10, seed=104), ["--partition-bad", "--all-ext", "--both-strands"]),
])
def test_restatement_equals_reference(built, tmp_path, cfg, k, sim, opts):
    tmp = str(tmp_path)
    reads = pu.simulate(os.path.join(tmp, "r.fasta"), **sim)
    ref = pu.run_oracle(reads, os.path.join(pu.CFG_DIR, cfg), os.path.join(tmp, "ref"), k=k, binary=pu.REF_HARNESS, extra=opts)
    res = pu.run_oracle(reads, os.path.join(pu.CFG_DIR, cfg), os.path.join(tmp, "res"), k=k, binary=pu.RESTATE, extra=opts)
    assert ref["overlaps"] == res["overlaps"]
    if "--partition-bad" in opts and cfg == "hifi.cfg":   # the trimming really produced the result: without it (almost) nothing passes
        plain = pu.run_oracle(reads, os.path.join(pu.CFG_DIR, cfg), os.path.join(tmp, "plain"), k=k, binary=pu.REF_HARNESS)
        assert plain["overlaps"] < 10 < ref["overlaps"]
    for ext in ("hist", "index", "ovlp"):
        a, b = os.path.join(tmp, "ref." + ext), os.path.join(tmp, "res." + ext)
        if os.path.exists(a):
            n, sample = pu.diff_files(a, b)
            assert n == 0, (ext, sample[:3])
