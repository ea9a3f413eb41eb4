#!/usr/bin/env python
"""One-off check at BASELINE.json's full sizes: the CUDA path against the unmodified reference on configs[0] / configs[1]
(ordered overlap vectors incl. divergence bits, k-mer histogram).  Writes a summary JSON to stdout.
Usage: python tools/full_scale_parity.py {clr|hifi}"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import parity_util as pu   # noqa: E402
import bench               # noqa: E402

w = sys.argv[1]
wl = bench.WORKLOADS[w]
reads = bench.make_reads(wl, 1.0)
cfg = os.path.join(pu.CFG_DIR, wl["cfg"])
t0 = time.time()
REF, GPU = "/tmp/full_ref_" + w, "/tmp/full_gpu_" + w
if os.path.exists(REF + ".json"):   # a second variant of the CUDA path in the same session: reuse the reference run
    ref = json.load(open(REF + ".json"))
else:
    ref = pu.run_oracle(reads, cfg, REF, k=wl["k"])
    json.dump(ref, open(REF + ".json", "w"))
t_ref = time.time() - t0
t0 = time.time()
_, info = pu.gpu_pipeline(reads, cfg, GPU, k=wl["k"])
t_gpu = time.time() - t0
out = {"workload": wl["name"], "reads": ref["reads"], "reference_overlaps": ref["overlaps"], "gpu_overlaps": info["n_overlaps"],
       "reference": {kk: ref[kk] for kk in ("t_count", "t_index", "t_estimate", "t_overlaps", "threads")},
       "reference_wall_s": round(t_ref, 1), "gpu_wall_s_incl_python_dumps": round(t_gpu, 1),
       "env": {kk: v for kk, v in os.environ.items() if kk.startswith("FG_")}}
for ext in ("hist", "ovlp"):
    a, b = REF + "." + ext, GPU + "." + ext
    if os.path.exists(a):
        n, sample = pu.diff_files(a, b)
        out["diff_lines_" + ext] = n
        if n:
            out["sample_" + ext] = sample[:3]
print(json.dumps(out))
