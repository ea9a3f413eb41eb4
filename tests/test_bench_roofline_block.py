"""CPU: the roofline block of bench.py executed on the numbers of a committed bench line (profiles/r2_bench_clr_1gpu.json) — the
block cannot run here end to end (no GPU), so the test lifts it out of main() between its two marker lines and executes exactly
that source with the line's phases and work counters: it must run, reproduce the fractions of the HBM-bound kernels of the line,
and report chainRunDpKernel in the reference's units (tests/golden/full_scale/clr.json: dp_cells_literal) next to the executed
ones."""
import argparse
import json
import os
import textwrap

import numpy as np

import parity_util as pu


def test_roofline_block_runs_on_a_committed_line():
    src = open(os.path.join(pu.ROOT, "bench.py")).read()
    a = src.index("    total_reads = n_reads   # all ranks together")
    b = src.index("    lib_phases = (")
    block = textwrap.dedent(src[a:b])
    line = json.loads(open(os.path.join(pu.ROOT, "profiles", "r2_bench_clr_1gpu.json")).read().strip().splitlines()[-1])
    import bench
    n_reads, n_bases = line["config"]["reads"], line["config"]["bases"]
    lens = np.full(n_reads, n_bases // n_reads, dtype=np.uint32)
    lens[: n_bases - int(lens.sum())] += 1
    ns = {
        "np": np, "os": os, "json": json, "ROOT": pu.ROOT, "measured_peak": bench.measured_peak,
        "n_reads": n_reads, "ms_resident": line["ms_per_step"], "packed": np.zeros(8, np.uint64), "woff": np.zeros(2, np.uint64), "lens": lens,
        "queries": np.arange(0, 2 * n_reads, 2, dtype=np.uint32), "est_ids": list(range(1000)),
        "resident_phases": dict(line["phases_ms"]), "n_raw": [int(line["phases_ms"]["raw_overlaps"])], "world": 1,
        "stats": {"n_hits": line["work"]["kmer_hits"], "n_dp_cells": line["work"]["dp_cells"]}, "n_ovl": line["work"]["overlaps"],
        "k": line["config"]["kmer"], "lo": 0, "hi": n_reads, "ovl_len_sample": np.full(100, 5000.0), "edit_sample": [None, 0.0],
        "args": argparse.Namespace(workload="clr", scale=1.0), "gp": bench.golden_path("clr"),
        "roof_phases": dict(line["phases_ms_one_lane"]), "roof_calls": {p: 1 for p in line["phases_ms_one_lane"]},
        "int_peak": next(r["peak"] for r in line["roofline_kernels"] if r["bound"] == "int"), "cfg": {"use_minimizers": 0},
    }
    exec(compile(block, "bench.py:roofline-block", "exec"), ns)
    got = {r["phase"]: r for r in ns["roofline_kernels"]}
    for r in line["roofline_kernels"]:
        if r["bound"] == "hbm":   # the byte formulas are unchanged: same fraction as the committed line
            assert abs(got[r["phase"]]["frac"] - r["frac"]) < 1e-3 * r["frac"], r["phase"]
    literal = json.load(open(bench.golden_path("clr")))["dp_cells_literal"]
    dp = got["chain_dp"]
    assert ns["dp_cells_literal"] == literal == 3410233385
    assert abs(dp["algorithmic_ops_per_launch"] - 15.0 * literal) < 1.0 and 0.4 < dp["frac"] < 0.6
    assert abs(dp["frac_executed"] - next(r["frac"] for r in line["roofline_kernels"] if r["phase"] == "chain_dp")) < 1e-3
    assert ns["roofline"]["kernel"] == "chainRunDpKernel" and ns["roofline"]["traffic"] is not None
