"""-m gpu: the CUDA path against the committed golden vectors of the unmodified reference (no oracle binary needed)."""
import hashlib
import json
import os

import pytest

import parity_util as pu
from test_oracle_golden import CASES, GOLD, gunzip_case

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", CASES)
def test_cuda_path_matches_golden(engine, tmp_path, name):
    meta = json.load(open(os.path.join(GOLD, name + ".json")))
    reads = gunzip_case(name, str(tmp_path))
    out = os.path.join(str(tmp_path), "gpu")
    opts = meta["options"]
    mo = int(opts[opts.index("--max-overlaps") + 1]) if "--max-overlaps" in opts else 0
    _, info = pu.gpu_pipeline(reads, os.path.join(pu.CFG_DIR, meta["cfg"]), out, k=meta["k"], dump_index=True,
                              both_strands="--both-strands" in opts, force_local="--force-local" in opts, max_overlaps=mo,
                              all_ext="--all-ext" in opts, estimate="--no-estimate" not in opts, engine=engine)
    assert info["reads"] == meta["reads"] and info["n_overlaps"] == meta["overlaps"]
    for ext in ("hist", "ovlp"):
        gold = os.path.join(GOLD, name + "." + ext)
        if os.path.exists(gold):
            n, sample = pu.diff_files(gold, out + "." + ext)
            assert n == 0, (ext, sample[:3])
    digest = hashlib.sha256(open(out + ".index", "rb").read()).hexdigest()
    assert digest == open(os.path.join(GOLD, name + ".index.sha256")).read().strip()
