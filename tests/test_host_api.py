"""CPU: the value types of the drop-in boundary in the host mirror against the reference's.
oracle/host_api_check.cpp is compiled on the unmodified reference headers (oracle/_ref/host_api_ref, built where /root/reference
exists) and on flye_b200/host (build/flye_b200_host_api); both must print the same bytes and write the same FASTA files:
DnaSequence views, Kmer arithmetic (reverse complement, standard form, hash, rolling appends), IterKmers over whole sequences
and windows of either strand, yieldMinimizers for four window sizes (incl. low-complexity sequences with equal hashes inside a
window), every OverlapRange method incl. project() with and without kmerMatches, SequenceContainer::writeFasta."""
import filecmp
import os
import subprocess

import pytest

import parity_util as pu

REF_BIN = os.path.join(pu.ROOT, "oracle", "_ref", "host_api_ref")
MIRROR_BIN = os.path.join(pu.ROOT, "build", "flye_b200_host_api")


@pytest.mark.skipif(not os.path.exists(REF_BIN), reason="oracle/_ref/host_api_ref is built where the reference sources are")
def test_value_types_match_the_reference(built, tmp_path):
    from flye_b200 import build
    build.build_host_harness()
    outs = {}
    for name, exe in (("ref", REF_BIN), ("mirror", MIRROR_BIN)):
        r = subprocess.run([exe, os.path.join(str(tmp_path), name)], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=300)
        assert r.returncode == 0, r.stderr[-2000:]
        outs[name] = r.stdout
    assert outs["ref"].count("minimizers w=") >= 80 and outs["ref"].count("project ") >= 70
    if outs["ref"] != outs["mirror"]:
        a, b = outs["ref"].splitlines(), outs["mirror"].splitlines()
        first = next((i for i, (x, y) in enumerate(zip(a, b)) if x != y), min(len(a), len(b)))
        raise AssertionError("line %d: reference %r, mirror %r" % (first, a[first:first + 1], b[first:first + 1]))
    for ext in (".fasta", ".pos.fasta"):
        fa, fb = os.path.join(str(tmp_path), "ref" + ext), os.path.join(str(tmp_path), "mirror" + ext)
        assert os.path.getsize(fa) > 1000 and filecmp.cmp(fa, fb, shallow=False), ext
