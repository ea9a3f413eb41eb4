#!/usr/bin/env python
"""Assembles tests/golden/full_scale/<workload>.json from a finished run of the UNMODIFIED reference (build container):

    oracle/_ref/flye_ref_harness --reads R --cfg tests/cfg/<cfg> --k K --threads T [--chunk N] --out P  > P.json
    tools/_bin/ovlpdigest P.ovlp > P.digest.json
    python tools/make_full_scale_golden.py <workload> P

(recipe and meaning of the fields: tests/golden/full_scale/README.md).  The read set itself is not stored: the simulator is
deterministic and bench.py regenerates it from the parameters recorded under "sim"."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    workload, prefix = sys.argv[1], sys.argv[2]
    import bench
    wl = bench.WORKLOADS[workload]
    ref = json.loads(open(prefix + ".json").read().strip().splitlines()[-1])
    dig = json.loads(open(prefix + ".digest.json").read())
    assert dig["queries"] == ref["queries"] and dig["overlaps"] == ref["overlaps"], (dig["queries"], ref["queries"], dig["overlaps"], ref["overlaps"])
    out = {"workload": wl["name"], "sim": wl["sim"], "cfg": wl["cfg"], "k": wl["k"],
           "made_by": "oracle/_ref/flye_ref_harness (the unmodified reference) in the build container + tools/_bin/ovlpdigest; see tests/golden/full_scale/README.md",
           "reference": ref, "hist": open(prefix + ".hist").read() if os.path.exists(prefix + ".hist") else None,
           "queries": dig["queries"], "overlaps": dig["overlaps"], "chunk_queries": dig["chunk_queries"], "final": dig["final"], "chunks": dig["chunks"]}
    path = os.path.join(ROOT, "tests", "golden", "full_scale", workload + ".json")
    with open(path, "w") as f:
        json.dump(out, f)
    print(path, out["queries"], out["overlaps"], out["final"])


if __name__ == "__main__":
    main()
