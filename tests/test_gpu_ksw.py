"""GPU: the device build of the banded KSW2 alignment behind checkIdyAndTrim (flye_b200/csrc/ksw.cu,
fg_align_cigar_batch) against the oracle: build/flye_b200_trim_device is oracle/trim_check.cpp linked to the library, a separate
process with its own context; its CIGAR lines must equal the restatement's (pinned to the unmodified reference by
tests/test_oracle_trim.py) on the same generated cases, including the pairs whose band has to be doubled.
(First run on a B200: profiles/r2_ksw_device_240_cases.txt — 240 of 240 CIGARs identical.)"""
import os
import subprocess

import pytest

import parity_util as pu

DEVICE_BIN = os.path.join(pu.ROOT, "build", "flye_b200_trim_device")
RESTATE_BIN = os.path.join(pu.ROOT, "oracle", "_ref", "trim_restate")


def _cigar_lines(out):
    return [l for l in out.splitlines() if l.startswith(b"case") or l.startswith(b"  cigar")]


@pytest.mark.gpu
def test_device_ksw_cigars_match_the_oracle(built):
    if not os.path.exists(DEVICE_BIN):
        from flye_b200 import build
        build.build_host_harness()
    cases, seed = 240, 5
    dev = subprocess.run([DEVICE_BIN, str(cases), str(seed)], stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=600)
    assert dev.returncode == 0, dev.stderr[-2000:]
    res = subprocess.run([RESTATE_BIN, str(cases), str(seed)], stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=600)
    assert res.returncode == 0
    a, b = _cigar_lines(dev.stdout), _cigar_lines(res.stdout)
    assert len(a) == 2 * cases and a == b
