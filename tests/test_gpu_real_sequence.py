"""-m gpu: tie-heavy input at size.  Reads simulated from the full 500 kb of real E. coli sequence (homopolymers, repeats):
SURVEY §9.1 measured an (extId, curPos) tie in ~60 % of the queries of such read sets, i.e. most queries take the std::sort-exact
path instead of the stable radix sort.  The CUDA path is compared with the unmodified reference binary (oracle/_ref)."""
import gzip
import os
import shutil

import pytest

import parity_util as pu

pytestmark = pytest.mark.gpu
ECOLI = os.path.join(pu.ROOT, "tests", "golden", "ecoli_500kb.fa.gz")


@pytest.mark.skipif(not os.path.exists(pu.REF_HARNESS), reason="reference binary not built")
@pytest.mark.parametrize("cfg,k,sim,exts", [
    ("raw_reads.cfg", 15, dict(coverage=40, mean_len=7500, shape=2, error=0.12, seed=81), ["hist", "ovlp"]),
    ("hifi.cfg", 17, dict(coverage=30, mean_len=15000, shape=20, error=0.005, seed=82), ["ovlp"]),
])
def test_full_ecoli_500kb_against_the_reference(built, engine, tmp_path, cfg, k, sim, exts):
    tmp = str(tmp_path)
    genome = os.path.join(tmp, "ecoli.fa")
    with gzip.open(ECOLI, "rb") as fi, open(genome, "wb") as fo:
        shutil.copyfileobj(fi, fo)
    reads = pu.simulate(os.path.join(tmp, "r.fasta"), genome_len=0, extra=["--genome-fasta", genome], **sim)
    cfg_path = os.path.join(pu.CFG_DIR, cfg)
    ref = pu.run_oracle(reads, cfg_path, os.path.join(tmp, "ref"), k=k, binary=pu.REF_HARNESS)
    eng, info = pu.gpu_pipeline(reads, cfg_path, os.path.join(tmp, "gpu"), k=k, engine=engine)
    assert info["n_overlaps"] == ref["overlaps"] and ref["overlaps"] > 10000
    tied = eng.timings().get("tied_queries", 0)
    assert tied > 0.2 * ref["reads"], "expected a tie-heavy read set (%s of %s queries had ties)" % (tied, ref["reads"])
    for ext in exts:
        n, sample = pu.diff_files(os.path.join(tmp, "ref." + ext), os.path.join(tmp, "gpu." + ext))
        assert n == 0, (ext, sample[:3])
