#!/usr/bin/env python
"""Summaries of ncu outputs for profiles/: (1) per-kernel totals of a `--metrics gpu__time_duration.sum` launch list,
(2) the key counters of a `--set full` report (needs `ncu` on PATH to read the .ncu-rep)."""
import collections
import csv
import subprocess
import sys

KEY = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
       "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
       "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
       "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio", "l1tex__t_sector_hit_rate.pct",
       "lts__t_sector_hit_rate.pct", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
       "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
       "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
       "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
       "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
       "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
       "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active"]


def launches(path):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    agg = collections.defaultdict(lambda: [0, 0.0])
    for row in csv.DictReader(lines):
        name = row["Kernel Name"].split("(")[0]
        try:
            v = float(row["Metric Value"].replace(",", ""))
        except ValueError:
            continue
        unit = row["Metric Unit"]
        ms = v / 1e6 if unit.startswith("n") else v / 1e3 if unit.startswith("u") else v
        agg[name][0] += 1
        agg[name][1] += ms
    tot = sum(v[1] for v in agg.values())
    print("total kernel time %.3f ms over %d launches" % (tot, sum(v[0] for v in agg.values())))
    print("%-64s %6s %10s %7s" % ("kernel", "n", "ms", "share"))
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("%-64s %6d %10.3f %6.1f%%" % (k[:64], v[0], v[1], 100 * v[1] / tot))


def full(path):
    """path: a .ncu-rep, or the CSV that `ncu -i rep --page raw --csv` printed on the GPU box (the report itself stays there:
    gpurun only brings back 64 MiB)"""
    if path.endswith(".csv"):
        raw = open(path).read()
    else:
        raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], stdout=subprocess.PIPE, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = [(k, hdr.index(k)) for k in KEY if k in hdr]
    ni = hdr.index("Kernel Name")
    seen = collections.Counter()
    for r in rows[2:]:
        name = r[ni].split("(")[0]
        seen[name] += 1
        if seen[name] > (int(sys.argv[3]) if len(sys.argv) > 3 else 2):
            continue
        print("--- %s (launch %d)" % (name, seen[name]))
        for k, i in idx:
            print("    %-82s %s %s" % (k, r[i], units[i]))


if __name__ == "__main__":
    (launches if sys.argv[1] == "launches" else full)(sys.argv[2])
