#!/usr/bin/env python
"""Fills the @PLACEHOLDERS@ of DESIGN.md / README.md / profiles/README.md from the final bench lines (profiles/r1_bench_*.json)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def line(path):
    return json.loads(open(path).read().strip().splitlines()[-1])


def table(d):
    ph = d["phases_ms"]
    skip = {"arena_mallocs", "arena_gib", "tied_queries", "presorted_pairs", "raw_overlaps", "host_total", "host_subbatches", "host_chunk_prep"}
    rows = [(k, v) for k, v in ph.items() if k not in skip and not (k == "edit" and "host_results" in ph and False)]
    tot = d["ms_per_step"]
    out = ["| phase (kernels) | ms | share of the pass |", "|---|---|---|"]
    names = {"select": "select (`minimizerRegKernel` / `selectKernel`)", "expand": "expand (`expandKernel<2>`)", "lookup": "lookup (`queryLookupKernel` + scans)",
             "hit_sort_radix": "hit sort, tie-free queries (`segRadixSortKernel`)", "hit_sort_top": "hit sort, queries with ties: re-expansion, tie prefix, `sortHuge/Level/TailKernel`",
             "hit_sort_small": "hit sort, queries with ties: `sortSmallKernel`", "group": "target groups + prefilters (`pairFilterKernel`, CUB select)",
             "chain_prep": "pair prep + run detection (`pairPrepKernel`)", "chain_runs": "`chainRunsKernel` (re-sorted pairs only)", "chain_order": "pair order (CUB sort of run counts)",
             "chain_dp": "`chainRunDpKernel`", "chain_fill": "`chainFillKernel`", "chain_walk": "`chainWalkKernel<true>`", "edit": "edit distance (`wfaKernel`)",
             "host_results": "gather + D2H of the overlaps (host wall clock, includes `edit`)", "host_epilogue": "host epilogue (glibc logf divergence, thresholds, compaction)",
             "chain_extsort_top": "extPos re-sort of non-monotone pairs (exact)", "chain_extsort_small": "… shared-memory part", "chain_ordsort_top": "score sort of non-presorted pairs (exact)",
             "chain_ordsort_small": "… shared-memory part"}
    for k, v in sorted(rows, key=lambda kv: -kv[1]):
        if k == "host_results":
            v = v - ph.get("edit", 0.0)
            k2 = "gather + D2H of the overlaps (host wall clock)"
        else:
            k2 = names.get(k, k)
        out.append("| %s | %.2f | %.1f %% |" % (k2, v, 100 * v / tot))
    acc = sum(v for k, v in rows) - ph.get("edit", 0.0) * (1 if "host_results" in ph else 0)
    out.append("| python / ctypes glue, host gaps between phases | %.2f | %.1f %% |" % (tot - acc, 100 * (tot - acc) / tot))
    return "\n".join(out)


h, c = line(os.path.join(ROOT, "profiles", "r1_bench_hifi_1gpu_final.json")), line(os.path.join(ROOT, "profiles", "r1_bench_clr_1gpu_final.json"))
rep = {"@HIFI_MS@": "%.0f" % h["ms_per_step"], "@HIFI_RS@": "%.1f k" % (h["value"] / 1e3), "@CLR_MS@": "%.0f" % c["ms_per_step"],
       "@CLR_RS@": "%.0f k" % (c["value"] / 1e3), "@HIFI_TABLE@": table(h), "@CLR_TABLE@": table(c)}
for f in ["DESIGN.md", "README.md", os.path.join("profiles", "README.md")]:
    p = os.path.join(ROOT, f)
    s = open(p).read()
    for k, v in rep.items():
        s = s.replace(k, v)
    open(p, "w").write(s)
print(rep["@HIFI_MS@"], rep["@HIFI_RS@"], rep["@CLR_MS@"], rep["@CLR_RS@"])
