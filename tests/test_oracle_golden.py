"""CPU: the oracle (CPU restatement, oracle/restate) against the golden vectors generated from the unmodified
reference (tests/golden/make_golden.py).  This pins the restatement wherever /root/reference is absent."""
import gzip
import hashlib
import json
import os
import shutil

import pytest

import parity_util as pu

GOLD = os.path.join(pu.ROOT, "tests", "golden")
CASES = sorted(f[:-5] for f in os.listdir(GOLD) if f.endswith(".json"))


def gunzip_case(name, tmp):
    dst = os.path.join(tmp, name + ".fasta")
    with gzip.open(os.path.join(GOLD, name + ".fasta.gz"), "rb") as fi, open(dst, "wb") as fo:
        shutil.copyfileobj(fi, fo)
    return dst


@pytest.mark.parametrize("name", CASES)
def test_restatement_matches_golden(built, tmp_path, name):
    meta = json.load(open(os.path.join(GOLD, name + ".json")))
    reads = gunzip_case(name, str(tmp_path))
    out = os.path.join(str(tmp_path), "res")
    info = pu.run_oracle(reads, os.path.join(pu.CFG_DIR, meta["cfg"]), out, k=meta["k"], binary=pu.RESTATE,
                         extra=["--dump-index"] + meta["options"])
    assert info["reads"] == meta["reads"] and info["overlaps"] == meta["overlaps"]
    for ext in ("hist", "ovlp"):
        gold = os.path.join(GOLD, name + "." + ext)
        if os.path.exists(gold):
            n, sample = pu.diff_files(gold, out + "." + ext)
            assert n == 0, (ext, sample[:3])
    digest = hashlib.sha256(open(out + ".index", "rb").read()).hexdigest()
    assert digest == open(os.path.join(GOLD, name + ".index.sha256")).read().strip()


PINS = json.load(open(os.path.join(GOLD, "pins", "k17_reference_pins.json")))


@pytest.mark.parametrize("name", [n for n in PINS if not n.startswith("_")])
def test_restatement_k17_matches_reference_pins(built, tmp_path, name):
    """k = 17 (64-bit keys): the reference's flat counter takes 8 GiB and a 17 G-entry scan, so the GPU k = 17 tests compare the
    device path with the restatement; here the restatement's dumps of exactly those read sets must have the sha256 of the
    unmodified reference's (tests/golden/pins/k17_reference_pins.json, made in the build container)."""
    pin = PINS[name]
    reads = pu.simulate(os.path.join(str(tmp_path), "r.fasta"), **pin["sim"])
    out = os.path.join(str(tmp_path), "res")
    info = pu.run_oracle(reads, os.path.join(pu.CFG_DIR, pin["cfg"]), out, binary=pu.RESTATE, extra=pin["options"])
    assert info["reads"] == pin["reads"] and info["overlaps"] == pin["overlaps"] and info["k"] == 17
    for ext in ("hist", "index", "ovlp"):
        assert hashlib.sha256(open(out + "." + ext, "rb").read()).hexdigest() == pin[ext + "_sha256"], ext
