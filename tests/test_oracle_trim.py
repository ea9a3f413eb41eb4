"""CPU: the oracle for the alignment-trimming branch of getSeqOverlaps (overlap.cpp:475-485) — the restatement of
getAlignmentCigarKsw / checkIdyAndTrim (oracle/restate/ksw_restate.cpp: minimap2's banded ksw_extz2_sse restated at the level of its
byte arrays, the traceback, the interval search with libstdc++'s std::sort permutation) against the unmodified reference.
oracle/trim_check.cpp is one driver with two builds; their outputs — every CIGAR and every trimmed overlap with its divergence
bits, on generated pairs with block-wise divergence, homopolymer runs, either strand, band doublings — must be identical.
Where the reference binary is absent the restatement is held to the committed digests of the reference's output."""
import hashlib
import json
import os
import subprocess

import pytest

import parity_util as pu

REF_BIN = os.path.join(pu.ROOT, "oracle", "_ref", "trim_ref")
RESTATE_BIN = os.path.join(pu.ROOT, "oracle", "_ref", "trim_restate")
CORE_BIN = os.path.join(pu.ROOT, "oracle", "_ref", "trim_core")
MIRROR_BIN = os.path.join(pu.ROOT, "build", "flye_b200_trim_mirror")
PINS = json.load(open(os.path.join(pu.ROOT, "tests", "golden", "pins", "trim_reference_pins.json")))


def _run(exe, cases, seed):
    r = subprocess.run([exe, str(cases), str(seed)], stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    return r.stdout


@pytest.mark.parametrize("run", PINS["runs"], ids=lambda r: "seed%d" % r["seed"])
def test_trim_restatement_matches_reference_digests(built, run):
    out = _run(RESTATE_BIN, run["cases"], run["seed"])
    assert out.count(b"    cur [") == run["pieces"] > 100 and out.count(b"\n") == run["lines"]
    assert hashlib.sha256(out).hexdigest() == run["sha256"]


@pytest.mark.skipif(not os.path.exists(REF_BIN), reason="oracle/_ref/trim_ref is built where the reference sources are")
def test_trim_restatement_equals_reference_on_fresh_cases(built):
    seed = 20000 + os.getpid() % 10000          # a new seed every run: the two builds must agree on any input
    ref, res = _run(REF_BIN, 500, seed), _run(RESTATE_BIN, 500, seed)
    if ref != res:
        a, b = ref.splitlines(), res.splitlines()
        first = next((i for i, (x, y) in enumerate(zip(a, b)) if x != y), min(len(a), len(b)))
        raise AssertionError("seed %d line %d: reference %r, restatement %r" % (seed, first, a[first:first + 1], b[first:first + 1]))
    lens = [abs(int(l.split()[3].rstrip(b"+-")) - int(l.split()[5].rstrip(b"+-"))) for l in ref.splitlines() if l.startswith(b"case")]
    assert sum(d > 64 for d in lens) > 50       # the first band (64) was too narrow: the doubling loop ran


def cigar_lines(out):
    return [l for l in out.splitlines() if l.startswith(b"case") or l.startswith(b"  cigar")]


def test_device_routine_host_build_gives_the_reference_cigars(built):
    """flye_b200/csrc/ksw_core.cuh — the scalar routine kswCigarKernel runs, one thread per alignment — compiled for the host
    (oracle/_ref/trim_core, scratch pre-filled with garbage): every CIGAR equals the restatement's, which is pinned to the
    reference's above (and compared with it directly where the reference binary exists)."""
    for cases, seed in ((400, 1), (300, 77)):
        core, res = _run(CORE_BIN, cases, seed), _run(RESTATE_BIN, cases, seed)
        assert cigar_lines(core) == cigar_lines(res) and len(cigar_lines(core)) == 2 * cases
        if os.path.exists(REF_BIN):
            assert cigar_lines(core) == cigar_lines(_run(REF_BIN, cases, seed))


@pytest.mark.parametrize("run", PINS["runs"][:2], ids=lambda r: "seed%d" % r["seed"])
def test_mirror_trimming_tail_matches_reference_digests(built, run):
    """What the host mirror runs around the device alignment with FLYE_B200_DEVICE_KSW=1 (flye_b200/host/sequence/overlap.h:
    hpcRange + trimByCigar — homopolymer compression, =/X split, interval search with std::sort, coordinate mapping), fed with
    the CIGARs of the routine's host build: CIGAR lines and trimmed overlaps with the digest of the reference's checkIdyAndTrim."""
    if not os.path.exists(MIRROR_BIN):
        from flye_b200 import build
        build.build_host_harness()
    out = _run(MIRROR_BIN, run["cases"], run["seed"])
    assert hashlib.sha256(out).hexdigest() == run["sha256"]
