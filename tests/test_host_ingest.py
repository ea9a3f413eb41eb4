"""CPU: read ingestion and the OverlapRange text format of the host mirror against the reference.
oracle/ingest_check.cpp is compiled twice — on the unmodified reference's SequenceContainer / OverlapRange (oracle/_ref/ingest_ref,
built where /root/reference exists and shipped as a binary) and on the mirror in flye_b200/host (build/flye_b200_ingest) — and both
must print the same bytes: FASTA and FASTQ, plain and gzip, CRLF line ends, blank lines, lower case, non-ACGT letters (replaced
through the same rand() sequence), a missing final newline, ids / reverse-complement ids / names, the global-position round trip,
N50, the parser's error texts with their line numbers, duplicated ids, and dump -> load -> dump of OverlapRange records."""
import gzip
import os
import random
import subprocess

import pytest

import parity_util as pu

REF_BIN = os.path.join(pu.ROOT, "oracle", "_ref", "ingest_ref")
MIRROR_BIN = os.path.join(pu.ROOT, "build", "flye_b200_ingest")


def _inputs(tmp):
    rng = random.Random(3)

    def seq(n):
        return "".join(rng.choice("ACGT") for _ in range(n))
    recs = [("read_%d some description" % i, seq(rng.randint(30, 400))) for i in range(12)]
    recs[3] = ("lower_case", recs[3][1].lower())
    recs[5] = ("with_N", recs[5][1][:50] + "NNRYK" + recs[5][1][50:])
    fa = "".join(">%s\n%s\n" % (h, "\n".join(s[i:i + 60] for i in range(0, len(s), 60))) for h, s in recs)
    fq = "".join("@%s\n%s\n+\n%s\n" % (h, s, "I" * len(s)) for h, s in recs)
    lines = fa.split("\n")
    files = {
        "a.fasta": fa, "a.fastq": fq, "b.fa": fa, "b.fq": fq,
        "bad.fastq": fq.replace("+\n", "-\n", 1), "bad.fasta": "ACGT\n" + fa, "crlf.fasta": fa.replace("\n", "\r\n"),
        "empty.fasta": "", "x.txt": fa, "blank.fasta": "\n\n" + "\n".join(lines[:7]) + "\n\n" + "\n".join(lines[7:]),
        "blank_bad.fasta": "\n\n" + "\n".join(lines[:7]) + "\n\n>only_header\n>next\nACGT\n", "emptyhdr.fasta": ">\nACGT\n",
        "crlf.fastq": fq.replace("\n", "\r\n"),
        "blank.fastq": fq + "\n\n\n\n" + fq.replace("read_", "again_").replace("lower_case", "lc2").replace("with_N", "wn2"),
        "bad2.fastq": fq + "\n" + fq, "nonl.fasta": fa.rstrip("\n"), "dup.fasta": fa + fa,
    }
    paths = []
    for name, text in files.items():
        p = os.path.join(tmp, name)
        with open(p, "w", newline="") as f:
            f.write(text)
        paths.append(p)
    for name, text in (("c.fa.gz", fa), ("c.fq.gz", fq), ("c.fasta.gz", fa)):
        p = os.path.join(tmp, name)
        with gzip.open(p, "wt", newline="") as f:
            f.write(text)
        paths.append(p)
    paths.append(os.path.join(tmp, "missing.fasta"))
    return paths


@pytest.mark.skipif(not os.path.exists(REF_BIN), reason="oracle/_ref/ingest_ref is built where the reference sources are")
def test_ingestion_and_dump_load_match_the_reference(built, tmp_path):
    from flye_b200 import build
    build.build_host_harness()
    paths = _inputs(str(tmp_path))
    outs = []
    for exe in (REF_BIN, MIRROR_BIN):
        for min_len in ("0", "120"):
            r = subprocess.run([exe, min_len] + paths, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, cwd=str(tmp_path), timeout=300)
            assert r.returncode == 0, r.stderr[-2000:]
            outs.append(r.stdout)
    ref, mirror = outs[0] + outs[1], outs[2] + outs[3]
    assert ref.count("FILE ") == 2 * len(paths) and ref.count(" ERROR ") >= 14 and ref.count("REDUMP identical") >= 20
    assert "REDUMP DIFFERENT" not in ref
    if ref != mirror:
        a, b = ref.splitlines(), mirror.splitlines()
        first = next((i for i, (x, y) in enumerate(zip(a, b)) if x != y), min(len(a), len(b)))
        raise AssertionError("line %d: reference %r, mirror %r" % (first, a[first:first + 1], b[first:first + 1]))
