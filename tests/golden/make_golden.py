#!/usr/bin/env python
"""Generate the golden fixtures under tests/golden/ from the UNMODIFIED reference (oracle/_ref/flye_ref_harness).

Run in the build container (needs /root/reference for the reference binary and, for the real-sequence cases,
flye/tests/data/ecoli_500kb.fasta).  Per case it commits
    <case>.fasta.gz     the reads (ACGT only)
    <case>.hist         k-mer frequency histogram               (solid-k-mer cases)
    <case>.ovlp         ordered getSeqOverlaps vectors, divergence as float bits
    <case>.index.sha256 digest of the full index dump            (the dump itself is MBs)
    <case>.json         the harness arguments that produced them
The reference has no golden vectors of its own for this path, so these ARE the pinned outputs.
"""
import gzip
import hashlib
import json
import os
import shutil
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import parity_util as pu   # noqa: E402

ECOLI = "/root/reference/flye/tests/data/ecoli_500kb.fasta"

CASES = {
    # name: (simulate kwargs, cfg, k, harness options)
    "clr_random_k15": (dict(genome_len=60000, coverage=12, seed=5), "raw_reads.cfg", 15, ["--both-strands"]),
    "clr_ecoli_k15": (dict(genome_len=0, coverage=14, seed=9, extra=["--genome-fasta", "@ECOLI_SLICE@"]), "raw_reads.cfg", 15, []),
    "clr_random_k17_allext": (dict(genome_len=50000, coverage=12, error=0.10, seed=31), "raw_reads.cfg", 17, ["--all-ext"]),
    "hifi_random": (dict(genome_len=40000, coverage=10, mean_len=8000, shape=20, error=0.005, seed=13), "hifi.cfg", None,
                    ["--both-strands", "--no-estimate"]),
    "hifi_ecoli": (dict(genome_len=0, coverage=8, mean_len=8000, shape=20, error=0.005, seed=17, extra=["--genome-fasta", "@ECOLI_SLICE@"]),
                   "hifi.cfg", None, ["--no-estimate"]),
}


def main():
    if not os.path.exists(pu.REF_HARNESS):
        sys.exit("reference harness missing: make -C oracle ref")
    tmp = tempfile.mkdtemp(prefix="golden_")
    # a 50 kb slice of real E. coli sequence (homopolymers / repeats that random genomes lack)
    slice_path = os.path.join(tmp, "ecoli_slice.fasta")
    seq = "".join(l.strip() for l in open(ECOLI) if not l.startswith(">")).upper()
    with open(slice_path, "w") as f:
        f.write(">ecoli_slice\n" + seq[120000:170000] + "\n")
    for name, (sim, cfg, k, opts) in CASES.items():
        sim = dict(sim)
        extra = [slice_path if x == "@ECOLI_SLICE@" else x for x in sim.pop("extra", [])]
        reads = os.path.join(tmp, name + ".fasta")
        pu.simulate(reads, extra=extra, **sim)
        out = os.path.join(tmp, name)
        info = pu.run_oracle(reads, os.path.join(pu.CFG_DIR, cfg), out, k=k, binary=pu.REF_HARNESS, extra=["--dump-index"] + opts)
        with open(reads, "rb") as fi, gzip.GzipFile(os.path.join(HERE, name + ".fasta.gz"), "wb", mtime=0) as fo:
            shutil.copyfileobj(fi, fo)
        for ext in ("hist", "ovlp"):
            if os.path.exists(out + "." + ext):
                shutil.copy(out + "." + ext, os.path.join(HERE, name + "." + ext))
        with open(os.path.join(HERE, name + ".index.sha256"), "w") as f:
            f.write(hashlib.sha256(open(out + ".index", "rb").read()).hexdigest() + "\n")
        with open(os.path.join(HERE, name + ".json"), "w") as f:
            json.dump({"cfg": cfg, "k": k, "options": opts, "reads": info["reads"], "overlaps": info["overlaps"]}, f, indent=1)
        print(name, info["reads"], "reads", info["overlaps"], "overlaps")
    shutil.rmtree(tmp)


K17_PINS = {   # the read sets of tests/test_gpu_parity.py::test_clr_k17_parity and ::test_ont_like_k17_parity
    "clr_k17": dict(genome_len=150000, coverage=18, error=0.10, seed=21),
    "ont_k17": dict(genome_len=300000, coverage=20, mean_len=19000, error=0.10, seed=3),
}


def k17_pins():
    """pins/k17_reference_pins.json: sha256 of the reference's dumps at k = 17 (8 GiB flat counter + 17 G-entry scan per run,
    70-90 s each) on the read sets the GPU k = 17 tests simulate.  Only the digests are committed (the simulator is deterministic)."""
    tmp = tempfile.mkdtemp(prefix="golden_k17_")
    cfg = "raw_reads.cfg"
    out = {"_comment": "sha256 of the dumps (.hist, .index, .ovlp) of the UNMODIFIED reference (oracle/_ref/flye_ref_harness, k = 17) on the "
                       "simulated read sets of tests/test_gpu_parity.py::test_clr_k17_parity and ::test_ont_like_k17_parity (regenerated from "
                       "`sim`); checked against the CPU restatement by tests/test_oracle_golden.py::test_restatement_k17_matches_reference_pins; "
                       "made by tests/golden/make_golden.py --k17-pins"}
    for name, sim in K17_PINS.items():
        reads = pu.simulate(os.path.join(tmp, name + ".fasta"), **sim)
        info = pu.run_oracle(reads, os.path.join(pu.CFG_DIR, cfg), os.path.join(tmp, name), binary=pu.REF_HARNESS, extra=["--dump-index"], threads=4)
        rec = {"sim": sim, "cfg": cfg, "k": 17, "options": ["--dump-index"], "reads": info["reads"], "overlaps": info["overlaps"]}
        for ext in ("hist", "index", "ovlp"):
            rec[ext + "_sha256"] = hashlib.sha256(open(os.path.join(tmp, name + "." + ext), "rb").read()).hexdigest()
        out[name] = rec
        print(name, info["reads"], "reads", info["overlaps"], "overlaps")
    os.makedirs(os.path.join(HERE, "pins"), exist_ok=True)
    with open(os.path.join(HERE, "pins", "k17_reference_pins.json"), "w") as f:
        json.dump(out, f, indent=1)
    shutil.rmtree(tmp)


if __name__ == "__main__":
    if "--k17-pins" in sys.argv:
        k17_pins()
    else:
        main()
