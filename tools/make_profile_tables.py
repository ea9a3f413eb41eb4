#!/usr/bin/env python
"""Markdown tables for profiles/README.md from a bench.py line: `python tools/make_profile_tables.py profiles/r2_bench_clr_1gpu.json`.
Prints (1) the phase table of the one-lane pass (every kernel alone on the device) with its share of that pass and
(2) the roofline table (`roofline_kernels`)."""
import json
import sys

NAMES = {"count_clear": "clear counters + bitmap", "count": "`denseCountKernel` / `denseCountListKernel`", "count_hist": "`denseHistKernel`",
         "count_partition": "`ownerTotals/ownerScatterKernel`", "count_exchange": "all-to-all of the slots (NCCL)", "count_merge": "histogram all-reduce + all-gather of the 16-bit counters",
         "select": "`selectKernel` / `minimizerRegKernel` (+ tandem, finalize)", "emit": "`emitCount/emitWriteKernel`", "index_exchange": "owner partition + all-to-all of the entries (NCCL)",
         "index_sort": "radix sort of the entries by key (CUB)", "index_table": "RLE, classify, table insert, presence bitmap (+ all-gathers)",
         "lookup": "`queryLookupKernel` + scans", "expand": "`expandKernel<2>`", "hit_sort_radix": "`segTileSortKernel`", "epilogue_dev": "`divergenceKernel`, flag compaction, `gatherKeptKernel`",
         "hit_sort_top": "queries with ties: re-expansion, tie prefix, `sortHuge/Level/TailKernel`", "hit_sort_small": "queries with ties: `sortSmallKernel`",
         "group": "group starts, candidates, `pairFilterKernel`", "chain_prep": "`pairPrepKernel`", "chain_extsort_top": "extPos re-sort of non-monotone pairs (exact)",
         "chain_extsort_small": "... shared-memory part", "chain_runs": "`chainRunsKernel` (re-sorted pairs)", "chain_order": "pair order (CUB sort of run counts)",
         "chain_dp": "`chainRunDpKernel`", "chain_fill": "`chainFillKernel`", "chain_ordsort_top": "score sort of non-presorted pairs (exact)",
         "chain_ordsort_small": "... shared-memory part", "chain_walk": "`chainWalkKernel<true>`", "edit": "`hpcReadsKernel` + `wfaKernel`"}
SKIP_PREFIX = ("host_", "prep_", "arena_")
SKIP = {"raw_overlaps", "gathered_overlaps", "tied_queries", "presorted_pairs"}


def main():
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    ph = d.get("phases_ms_one_lane") or d["phases_ms"]
    tot = d.get("roofline_pass", {}).get("ms_per_step", d["ms_per_step"])
    rows = [(k, v) for k, v in ph.items() if not k.startswith(SKIP_PREFIX) and k not in SKIP]
    print("%s — %d GPU(s): %.1f ms per pass (%.0f reads/s), e2e %.1f ms; one-lane pass %.1f ms\n" %
          (d["config"]["workload"].split(":")[0], d["n_gpus"], d["ms_per_step"], d["value"], d["e2e"]["ms_per_step"] if d.get("e2e") else float("nan"), tot))
    print("| phase | ms | share of the one-lane pass |\n|---|---|---|")
    for k, v in sorted(rows, key=lambda kv: -kv[1]):
        print("| %s (`%s`) | %.2f | %.1f %% |" % (NAMES.get(k, k), k, v, 100 * v / tot))
    dev = sum(v for _, v in rows)
    print("| **device phases together** | %.2f | %.1f %% |" % (dev, 100 * dev / tot))
    print("| host: results D2H (`host_results` incl. waits), refilter, glue | %.2f | %.1f %% |" % (tot - dev, 100 * (tot - dev) / tot))
    print("\n| kernel | bound | ms / launch | achieved | peak | frac |\n|---|---|---|---|---|---|")
    for r in d["roofline_kernels"]:
        print("| `%s` | %s | %.2f | %.0f %s | %.0f | %.3f |" % (r["kernel"], r["bound"], r["ms_per_launch"], r["achieved"], r["unit"], r["peak"], r["frac"]))


if __name__ == "__main__":
    main()
