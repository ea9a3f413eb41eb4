"""-m gpu: the drop-in boundary.  oracle/harness.cpp — written against the reference's public C++ API — is compiled
unchanged against the host mirror (flye_b200/host: SequenceContainer, VertexIndex, OverlapDetector, OverlapContainer,
OverlapRange over the C ABI) and must produce the same dumps as the build against the unmodified reference."""
import os
import subprocess
import json

import pytest

import parity_util as pu

pytestmark = pytest.mark.gpu
MIRROR = os.path.join(pu.ROOT, "build", "flye_b200_harness")
RAW = os.path.join(pu.CFG_DIR, "raw_reads.cfg")
HIFI = os.path.join(pu.CFG_DIR, "hifi.cfg")


def run_mirror(reads, cfg, out, k=None, extra=(), threads=4):
    if not os.path.exists(MIRROR):
        from flye_b200 import build
        build.build_host_harness()
    cmd = [MIRROR, "--reads", reads, "--cfg", cfg, "--out", out, "--threads", str(threads)] + (["--k", str(k)] if k else []) + list(extra)
    r = subprocess.run(cmd, check=True, stdout=subprocess.PIPE, text=True)
    return json.loads(r.stdout.strip().splitlines()[-1])


@pytest.mark.parametrize("cfg,k,sim,opts,exts", [
    (RAW, 15, dict(genome_len=80000, coverage=12, seed=41), ["--dump-index"], ["hist", "index", "ovlp"]),
    (RAW, 15, dict(genome_len=80000, coverage=12, seed=42), ["--find-all"], ["ovlp"]),        # closure + cluster filter, multisets
    (RAW, 15, dict(genome_len=60000, coverage=12, seed=44), ["--all-ext", "--keep-aln"], ["ovlp"]),   # kmerMatches through the mirror
    (HIFI, None, dict(genome_len=60000, coverage=10, mean_len=8000, shape=20, error=0.005, seed=43), ["--both-strands"], ["ovlp"]),
])
def test_reference_harness_source_runs_on_the_mirror(built, tmp_path, cfg, k, sim, opts, exts):
    tmp = str(tmp_path)
    reads = pu.simulate(os.path.join(tmp, "r.fasta"), **sim)
    ref = pu.run_oracle(reads, cfg, os.path.join(tmp, "ref"), k=k, extra=opts)
    got = run_mirror(reads, cfg, os.path.join(tmp, "gpu"), k=k, extra=opts)
    assert got["reads"] == ref["reads"] and got["overlaps"] == ref["overlaps"]
    for ext in exts:
        n, sample = pu.diff_files(os.path.join(tmp, "ref." + ext), os.path.join(tmp, "gpu." + ext))
        assert n == 0, (ext, sample[:3])


@pytest.mark.parametrize("mode", ["--lazy", "--per-read"])
def test_on_demand_calls_from_many_threads(built, tmp_path, mode):
    """how extender.cpp / chimera.cpp really call the path: one read per call — lazySeqOverlaps (thread safe, cached; the mirror
    fills its cache for all reads on the first miss) or quickSeqOverlaps — from 8 worker threads at once"""
    tmp = str(tmp_path)
    reads = pu.simulate(os.path.join(tmp, "r.fasta"), genome_len=70000, coverage=10, seed=45)
    opts = [mode] + (["--both-strands"] if mode == "--lazy" else [])
    ref = pu.run_oracle(reads, RAW, os.path.join(tmp, "ref"), k=15, threads=8, extra=opts)
    got = run_mirror(reads, RAW, os.path.join(tmp, "gpu"), k=15, extra=opts, threads=8)
    assert got["overlaps"] == ref["overlaps"] and ref["overlaps"] > 100
    n, sample = pu.diff_files(os.path.join(tmp, "ref.ovlp"), os.path.join(tmp, "gpu.ovlp"))
    assert n == 0, sample[:3]


def test_queries_from_a_second_container(built, tmp_path):
    """read-to-graph style: index over container A, queries from container B (OverlapContainer(detector, B)); minimizer index,
    all primary overlaps kept, local overlaps allowed (SMALL_ALN = 100 in the reference's ReadAligner)"""
    tmp = str(tmp_path)
    a = pu.simulate(os.path.join(tmp, "a.fasta"), genome_len=120000, coverage=6, mean_len=12000, shape=20, error=0.002, seed=51)
    b = pu.simulate(os.path.join(tmp, "b.fasta"), genome_len=120000, coverage=8, mean_len=6000, shape=4, error=0.03, seed=51)
    for cfg, k, opts in [(HIFI, None, ["--query-reads", b, "--all-ext", "--both-strands", "--no-estimate", "--min-overlap", "500"]),
                         (RAW, 15, ["--query-reads", b, "--all-ext", "--keep-aln"])]:
        ref = pu.run_oracle(a, cfg, os.path.join(tmp, "ref"), k=k, extra=opts)
        got = run_mirror(a, cfg, os.path.join(tmp, "gpu"), k=k, extra=opts)
        assert got["overlaps"] == ref["overlaps"] and ref["overlaps"] > 100
        n, sample = pu.diff_files(os.path.join(tmp, "ref.ovlp"), os.path.join(tmp, "gpu.ovlp"))
        assert n == 0, sample[:3]
