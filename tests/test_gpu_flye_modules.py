"""-m gpu: the reference's OWN `flye-modules assemble` and `flye-modules repeat` (main_assemble.cpp, extender.cpp, chimera.cpp, parameters_estimator.cpp,
consensus_generator.cpp ... unmodified) linked against the host mirror + libflye_b200.so must produce the draft assembly the
unmodified reference build produces.  Both binaries are built in the build container by oracle/build_flye_modules.sh
(oracle/_ref/flye-modules-{ref,b200}); reads are simulated from the real E. coli sequence the reference ships as its test data
(tests/golden/ecoli_500kb.fa.gz).  The extender asks for overlaps one read at a time (lazySeqOverlaps, quickSeqOverlaps): on
the mirror the first miss fills the cache for all reads in one device pass."""
import gzip
import hashlib
import os
import shutil
import subprocess

import pytest

import parity_util as pu

pytestmark = pytest.mark.gpu
REF_BIN = os.path.join(pu.ROOT, "oracle", "_ref", "flye-modules-ref")
B200_BIN = os.path.join(pu.ROOT, "oracle", "_ref", "flye-modules-b200")
ECOLI = os.path.join(pu.ROOT, "tests", "golden", "ecoli_500kb.fa.gz")


def _assemble(binary, reads, cfg, out, threads, k):
    cmd = [binary, "assemble", "--reads", reads, "--out-asm", out, "--config", cfg, "--log", out + ".log", "--threads", str(threads),
           "--min-ovlp", "1000", "--kmer", str(k), "--genome-size", "500000"]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=1500)
    assert r.returncode == 0, r.stdout[-3000:]
    return hashlib.sha256(open(out, "rb").read()).hexdigest(), os.path.getsize(out)


def _canonical(name, data):
    """The node numbers of repeat_graph_dump and of the .gv files come from iterating a std::unordered_set<GraphNode*>
    (repeat_graph.h:445, output_generator.cpp): they depend on heap addresses, not on the graph, so two runs of the SAME binary
    can number the nodes differently.  Edge ids are stable: a node is renamed to the sorted list of its (direction, edge id)
    incidences, and the lines are sorted."""
    import re
    text = data.decode()
    if name == "repeat_graph_dump":
        inc = {}
        for ln in text.splitlines():
            f = ln.split("\t")
            if f[0] == "Edge":
                inc.setdefault(f[2], []).append("o" + f[1]); inc.setdefault(f[3], []).append("i" + f[1])
        sig = {n: "<" + ",".join(sorted(v)) + ">" for n, v in inc.items()}
        out, block = [], None
        for ln in text.splitlines():
            f = ln.split("\t")
            if f[0] == "Edge":
                f[2], f[3] = sig[f[2]], sig[f[3]]
                block = ["\t".join(f)]; out.append(block)
            elif block is not None and ln.startswith("\t"):
                block.append(ln)
            else:
                block = None; out.append([ln])
        return "\n".join(sorted("\n".join(b) for b in out)).encode()
    if name.endswith(".gv"):
        edge = re.compile(r'^"(\d+)" -> "(\d+)" \[label = "id (-?\d+)')
        inc = {}
        for ln in text.splitlines():
            m = edge.match(ln)
            if m:
                inc.setdefault(m.group(1), []).append("o" + m.group(3)); inc.setdefault(m.group(2), []).append("i" + m.group(3))
        sig = lambda n: "<" + ",".join(sorted(inc.get(n, []))) + ">"
        lines = [re.sub(r'^"(\d+)"( -> )?(?:"(\d+)")?', lambda m: sig(m.group(1)) + (m.group(2) or "") + (sig(m.group(3)) if m.group(3) else ""), ln)
                 for ln in text.splitlines()]
        return "\n".join(sorted(lines)).encode()
    return data


def _repeat(binary, disjointigs, reads, cfg, out_dir, threads, k):
    """the repeat stage: RepeatGraph::build = findAllOverlaps on the disjointigs with keepAlignment, nuclAlignment and
    partitionBadMappings (KSW2 trimming) on a minimizer index (repeat_graph.cpp:74-97), then ReadAligner::alignReads = every read
    against the graph edges (read_aligner.cpp:158-292)"""
    os.makedirs(out_dir, exist_ok=True)
    cmd = [binary, "repeat", "--disjointigs", disjointigs, "--reads", reads, "--out-dir", out_dir, "--config", cfg, "--log",
           os.path.join(out_dir, "log"), "--threads", str(threads), "--min-ovlp", "1000", "--kmer", str(k)]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=1500)
    assert r.returncode == 0, r.stdout[-3000:]
    return {f: hashlib.sha256(_canonical(f, open(os.path.join(out_dir, f), "rb").read())).hexdigest()
            for f in ("repeat_graph_edges.fasta", "repeat_graph_dump", "read_alignment_dump", "graph_before_rr.gv", "graph_after_rr.gv")}


@pytest.mark.skipif(not (os.path.exists(REF_BIN) and os.path.exists(B200_BIN)), reason="flye-modules binaries not built (oracle/build_flye_modules.sh)")
@pytest.mark.parametrize("preset,k,sim", [
    ("full_asm_raw_reads.cfg", 15, dict(coverage=25, mean_len=7500, shape=2, error=0.12, seed=71)),
    ("full_asm_hifi.cfg", 17, dict(coverage=20, mean_len=12000, shape=20, error=0.005, seed=72)),
])
def test_flye_modules_assemble_on_the_mirror(built, tmp_path, preset, k, sim):
    tmp = str(tmp_path)
    genome = os.path.join(tmp, "ecoli.fa")
    with gzip.open(ECOLI, "rb") as fi, open(genome, "wb") as fo:
        shutil.copyfileobj(fi, fo)
    reads = pu.simulate(os.path.join(tmp, "r.fasta"), genome_len=0, extra=["--genome-fasta", genome], **sim)
    cfg = os.path.join(pu.CFG_DIR, preset)
    ref = _assemble(REF_BIN, reads, cfg, os.path.join(tmp, "ref.fasta"), 1, k)
    got = _assemble(B200_BIN, reads, cfg, os.path.join(tmp, "b200.fasta"), 1, k)
    assert ref[1] > 100000, "reference produced no assembly"
    assert got == ref, (got, ref)
    # the same with 8 worker threads on the mirror (lazySeqOverlaps / quickSeqOverlaps called concurrently): the extension order
    # is timing dependent in the reference as well, so only completion and a comparable amount of sequence are required
    many = _assemble(B200_BIN, reads, cfg, os.path.join(tmp, "b200_t8.fasta"), 8, k)
    assert 0.7 * ref[1] < many[1] < 1.4 * ref[1], (many, ref)
    # the next stage on the same disjointigs: repeat graph + read-to-graph alignment, both builds, byte-identical outputs
    rep_ref = _repeat(REF_BIN, os.path.join(tmp, "ref.fasta"), reads, cfg, os.path.join(tmp, "rep_ref"), 1, k)
    rep_got = _repeat(B200_BIN, os.path.join(tmp, "ref.fasta"), reads, cfg, os.path.join(tmp, "rep_b200"), 1, k)
    if rep_got != rep_ref and os.path.isdir(os.path.join(pu.ROOT, "gpurun_out")):   # keep the evidence of a mismatch
        for f in rep_ref:
            if rep_got[f] != rep_ref[f]:
                for side in ("rep_ref", "rep_b200"):
                    shutil.copy(os.path.join(tmp, side, f), os.path.join(pu.ROOT, "gpurun_out", "mismatch_%s_%s" % (side, f)))
                    shutil.copy(os.path.join(tmp, side, "log"), os.path.join(pu.ROOT, "gpurun_out", "mismatch_%s_log" % side))
    assert rep_got == rep_ref, {f: (rep_got[f] == rep_ref[f]) for f in rep_ref}
