"""CPU (build container only — needs the reference sources): the reference's own callers of the hot path compile against the
host mirror.  INTEGRATION.md §1: the mirror headers REPLACE src/sequence/{sequence,sequence_container,kmer,vertex_index,overlap}.h
(a quoted #include "../sequence/overlap.h" resolves relative to the including file, so putting the mirror on -I is not enough);
src/common stays the reference's.  The shadow tree below is exactly that replacement, made of symlinks."""
import os
import subprocess
from concurrent.futures import ThreadPoolExecutor

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
CALLERS = ["assemble/extender.cpp", "assemble/chimera.cpp", "assemble/main_assemble.cpp", "assemble/parameters_estimator.cpp",
           "repeat_graph/read_aligner.cpp", "repeat_graph/repeat_graph.cpp", "repeat_graph/main_repeat.cpp",
           "repeat_graph/repeat_resolver.cpp", "contigger/contig_extender.cpp", "sequence/consensus_generator.cpp"]


@pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "src", "sequence")), reason="reference sources not present")
def test_reference_callers_compile_against_the_mirror(tmp_path):
    src = os.path.join(str(tmp_path), "src")
    for d in ("assemble", "repeat_graph", "common", "contigger", "polishing", "sequence"):
        os.makedirs(os.path.join(src, d))
        for f in os.listdir(os.path.join(REF, "src", d)):
            os.symlink(os.path.join(REF, "src", d, f), os.path.join(src, d, f))
    mirror = os.path.join(ROOT, "flye_b200", "host", "sequence")
    for f in os.listdir(mirror):
        if f.endswith(".h"):
            dst = os.path.join(src, "sequence", f)
            os.remove(dst)
            os.symlink(os.path.join(mirror, f), dst)
    inc = ["-I" + os.path.join(REF, "lib", d) for d in ("libcuckoo", "interval_tree", "lemon", "minimap2")] + ["-I" + os.path.join(ROOT, "include")]

    def check(f):
        r = subprocess.run(["g++", "-std=c++17", "-fsyntax-only", "-w", "-include", "cstdint"] + inc + [os.path.join(src, f)],
                           stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        return f, r.returncode, r.stdout[-1500:]
    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as ex:
        results = list(ex.map(check, CALLERS))
    bad = [(f, out) for f, rc, out in results if rc != 0]
    assert not bad, bad
    # the headers that won the include race are the mirror's (g++ -H prints every header it opens)
    r = subprocess.run(["g++", "-std=c++17", "-fsyntax-only", "-w", "-H", "-include", "cstdint"] + inc + [os.path.join(src, "assemble/extender.cpp")],
                       stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    opened = [l.strip(". ").strip() for l in r.stdout.splitlines() if "sequence/overlap.h" in l or "sequence/vertex_index.h" in l]
    assert opened and all(os.path.realpath(p).startswith(os.path.join(ROOT, "flye_b200", "host")) for p in opened), opened
