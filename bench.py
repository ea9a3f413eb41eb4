#!/usr/bin/env python
"""bench.py — reads/sec through k-mer index + overlap detection (BASELINE.json metric).

One step = one pass of the hot path over the whole synthetic read set:
    [H2D of the packed reads]  ->  k-mer counting (solid presets)  ->  VertexIndex build  ->
    estimateOverlaperParameters (1000 random queries)  ->  getSeqOverlaps for every forward read  -> [D2H overlaps]
`value`  = reads / step time with the reads already resident in HBM (upload outside the timed region);
`e2e`    = the same through the C-ABI calls with HOST buffers: every step re-uploads the packed reads from
           pinned host memory and brings every overlap record back.
Workloads (BASELINE.json configs):
    clr     : configs[0]  4.6 Mb genome, 50x CLR-like reads (12 % error, mean 7.5 kb), solid k-mers k=15       (default: the
              configuration BASELINE.md / SURVEY §8 quote, and the full north-star path: count + solid index + overlaps)
    hifi    : configs[1]  4.6 Mb genome, 30x HiFi-like reads (0.5 % error, ~15 kb), minimizers w=10 k=17, edlib+HPC
    ont     : configs[2]  140 Mb genome, 30x ONT-like reads (10 % error, mean 19 kb), solid k-mers k=17
    meta    : configs[3]  20 genomes, 65 Mb in total, log-distributed coverage, 10 % error, mean 9 kb, k=17
    hifi250 : configs[4]  250 Mb genome, 30x HiFi-like reads
After the timed steps the ordered overlap vectors of the last step are digested (tools/ovlpdigest.cpp: sha256 per query,
folded in query order over all ranks) and compared with the digest of the UNMODIFIED REFERENCE's dump committed under
tests/golden/full_scale/; a mismatch makes the process exit with status 3 — so a green run at N ranks is also the N-rank
parity proof.
--impl reference times the reference's own CPU implementation (oracle/_ref/flye_ref_harness, the unmodified
reference sources; the CPU restatement if that binary is absent) on all host cores.
"""
import argparse
import ctypes
import hashlib
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

WORKLOADS = {
    "clr": dict(name="configs[0]: simulated 4.6 Mb genome, 50x CLR-like reads (12% error, mean 7.5 kb), asm_raw_reads settings, k=15",
                sim=dict(genome_len=4600000, coverage=50, mean_len=7500, shape=2, error=0.12, seed=1), cfg="raw_reads.cfg", k=15,
                cpu_queries=None),
    "hifi": dict(name="configs[1]: simulated 4.6 Mb genome, 30x HiFi-like reads (0.5% error, ~15 kb), asm_hifi settings "
                      "(k=17 minimizers w=10, HPC edit-distance divergence)",
                 sim=dict(genome_len=4600000, coverage=30, mean_len=15000, shape=20, error=0.005, seed=2), cfg="hifi.cfg", k=17,
                 cpu_queries=None),
    "ont": dict(name="configs[2]: simulated 140 Mb genome, 30x ONT-like reads (10% error, mean 19 kb), asm_raw_reads settings, k=17",
                sim=dict(genome_len=140000000, coverage=30, mean_len=19000, shape=2, error=0.10, seed=3), cfg="raw_reads.cfg", k=17,
                cpu_queries=2000),
    "meta": dict(name="configs[3]: simulated metagenome, 20 genomes of 65 Mb in total, coverage 4*2^(i/2.5) capped at 400, 10% error, "
                      "mean 9 kb, asm_raw_reads settings, k=17 (buildIndexUnevenCoverage is the index path of every solid preset)",
                 sim=dict(genome_len=65000000, coverage=0, mean_len=9000, shape=2, error=0.10, seed=4, meta=20), cfg="raw_reads.cfg", k=17,
                 cpu_queries=2000),
    "hifi250": dict(name="configs[4]: simulated 250 Mb genome, 30x HiFi-like reads (0.5% error, ~15 kb), asm_hifi settings",
                    sim=dict(genome_len=250000000, coverage=30, mean_len=15000, shape=20, error=0.005, seed=5), cfg="hifi.cfg", k=17,
                    cpu_queries=2000),
}
MIN_OVERLAP = 1000
METRIC = "reads/sec through k-mer index + overlap detection"


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def make_reads(wl, scale):
    import parity_util as pu
    from flye_b200 import build
    build.build_tools()
    sim = dict(wl["sim"])
    sim["genome_len"] = max(50000, int(sim["genome_len"] * scale))
    tag = "_".join("%s%s" % (k[0], v) for k, v in sorted(sim.items()))
    path = os.path.join("/tmp", "flye_b200_bench_%s.fasta" % tag)
    if not os.path.exists(path):
        tmp = "%s.%d.tmp" % (path, os.getpid())
        meta = sim.pop("meta", 0)
        if meta:
            sim.pop("coverage")
            total = sim.pop("genome_len")
            pu.simulate(tmp, genome_len=total, extra=["--meta", str(meta), "--meta-total", str(total)], **sim)
        else:
            pu.simulate(tmp, **sim)
        os.replace(tmp, path)
    return path


def golden_path(workload):
    return os.path.join(ROOT, "tests", "golden", "full_scale", workload + ".json")


_digest_lib = None


def digest_lib():
    global _digest_lib
    if _digest_lib is None:
        _digest_lib = ctypes.CDLL(os.path.join(ROOT, "tools", "_bin", "libovlpdigest.so"))
    return _digest_lib


def query_digests(query_ids, offsets, ov):
    """sha256 of every query's block of the ordered overlap dump (oracle/harness.cpp format) -> (n, 32) uint8"""
    import numpy as np
    q = np.ascontiguousarray(query_ids, dtype=np.uint32)
    offs = np.ascontiguousarray(offsets, dtype=np.uint64)
    ov = np.ascontiguousarray(ov)
    out = np.zeros((len(q), 32), dtype=np.uint8)
    digest_lib().fg_digest_queries(q.ctypes.data_as(ctypes.c_void_p), offs.ctypes.data_as(ctypes.c_void_p), ov.ctypes.data_as(ctypes.c_void_p),
                                   ctypes.c_uint64(len(q)), out.ctypes.data_as(ctypes.c_void_p))
    return out


def fold_digests(dq, chunk=1000):
    chunks = [hashlib.sha256(dq[a:a + chunk].tobytes()).digest() for a in range(0, len(dq), chunk)]
    return hashlib.sha256(b"".join(chunks)).hexdigest(), [c.hex() for c in chunks]


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md clocks line)."""

    def __init__(self, index=0):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
            "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q, "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def mark(self):
        """start of the timed region: only the samples from here on are reported (nvidia-smi itself is started before the warm-up
        steps, so that its start-up — NVML initialisation holds the driver for tens of ms — does not land in a timed step)"""
        self.first = len(self.rows)

    def stop(self):
        if self.proc:
            self.proc.terminate()
        self.rows = self.rows[getattr(self, "first", 0):] or self.rows
        sm = sorted(int(float(r[0])) for r in self.rows if r and r[0].replace(".", "").isdigit())
        mx = [int(float(r[1])) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            for name, v in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons),
                "samples": len(sm)}


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def run_reference(args, wl, reads_path, budget_s=150.0):
    """The reference's own CPU path on all host cores.  cpu_queries = None: the whole workload (the unmodified reference
    needs ~17 s for configs[1] and ~31 s for configs[0] on the 16-core GPU box); a number Q: full count + index over all reads +
    the estimate pass + the first Q forward reads as overlap queries, overlap time scaled to all reads.  At most `budget_s`
    seconds of steps are run after the warm-up one; the number of steps actually timed is returned."""
    import parity_util as pu
    binary, kind = pu.oracle_binary()
    cores = os.cpu_count() or 1
    cfg = os.path.join(pu.CFG_DIR, wl["cfg"])
    q = wl["cpu_queries"]
    vals = []
    n_reads = n_bases = None
    warm = min(args.warmup, 1)          # a CPU run has nothing to warm beyond the page cache
    t_begin = None
    for it in range(warm + args.steps):
        if kind == "port" and it > 0:
            break
        if it == warm:
            t_begin = time.time()
        if it > warm and time.time() - t_begin + (vals[-1] if vals else 0) > budget_s:   # keep the whole arm within a few minutes
            break
        r = pu.run_oracle(reads_path, cfg, "/tmp/flye_b200_bench_ref", k=wl["k"], threads=cores, binary=binary,
                          extra=["--max-queries", str(q)] if q is not None else [])
        n_reads, n_bases = r["reads"], r.get("bases")
        nq = max(1, r["queries"])
        t = r["t_count"] + r["t_index"] + r["t_estimate"] + r["t_overlaps"] * (n_reads / nq)
        log("reference step %d: count %.2fs index %.2fs estimate %.2fs overlaps(%d queries) %.2fs -> %.1fs per pass" %
            (it, r["t_count"], r["t_index"], r["t_estimate"], nq, r["t_overlaps"], t))
        if it >= warm or kind == "port":
            vals.append(t)
    t = sum(vals) / len(vals)
    value = n_reads / t
    if q is None:
        sample = "the whole workload: k-mer counting / index over all %d reads + estimateOverlaperParameters + getSeqOverlaps of " \
                 "every forward read (no extrapolation); %d timed passes" % (n_reads, len(vals))
    else:
        sample = "count + index over all %d reads + estimateOverlaperParameters + getSeqOverlaps of the first %d forward reads; overlap " \
                 "time scaled by reads/queries; %d timed passes" % (n_reads, q, len(vals))
    return value, t, n_reads, n_bases, len(vals), dict(value=value, unit="reads/s", cores=cores, kind=kind, sample=sample)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default=os.environ.get("FLYE_B200_WORKLOAD", "clr"), choices=sorted(WORKLOADS))
    ap.add_argument("--scale", type=float, default=float(os.environ.get("FLYE_B200_SCALE", "1.0")), help="genome scale (tests)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true", help="skip the second timed region (host buffers re-uploaded every step)")
    ap.add_argument("--write-golden", action="store_true", help="(development) print the digest record instead of comparing")
    args = ap.parse_args()
    wl = WORKLOADS[args.workload]
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    config = {"workload": wl["name"], "reads": None, "bases": None, "kmer": wl["k"], "min_overlap": MIN_OVERLAP,
              "cache": "inputs (packed reads >= 34 MB, index and hit arrays of several GB) exceed the 126 MB L2",
              "scale": args.scale, "partition": "reads sharded across ranks, index replicated" if world > 1 else "single GPU"}

    if args.impl == "reference":
        if rank != 0:
            return
        reads_path = make_reads(wl, args.scale)
        value, t, n_reads, n_bases, steps_run, cb = run_reference(args, wl, reads_path)
        config["reads"], config["bases"] = n_reads, n_bases
        config["partition"] = "host CPU threads (the reference has no device path)"
        print(json.dumps({"impl": "reference", "metric": METRIC, "value": value,
                          "unit": "reads/s", "n_gpus": args.gpus, "steps": steps_run, "warmup": min(args.warmup, 1), "ms_per_step": t * 1e3,
                          "steps_requested": args.steps, "warmup_requested": args.warmup,
                          "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "int64", "data": "synthetic",
                          "config": config, "cpu_baseline": cb,
                          "e2e": {"value": value, "unit": "reads/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))
        return

    import numpy as np
    import torch
    import flye_b200 as fb
    import parity_util as pu

    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    torch.cuda.set_device(local_rank)
    if world > 1:   # one rank simulates, the others read the file
        if rank == 0:
            make_reads(wl, args.scale)
        dist.barrier()
    reads_path = make_reads(wl, args.scale)
    cfg = pu.load_cfg(os.path.join(pu.CFG_DIR, wl["cfg"]))
    k = wl["k"]
    reads = fb.read_fasta(reads_path, MIN_OVERLAP)
    packed, woff, lens = fb.pack_reads(reads)
    n_reads, n_bases = len(reads), int(lens.sum())
    config["reads"], config["bases"] = n_reads, n_bases
    del reads
    # host buffers in pinned memory (what the e2e arm uploads every step)
    packed_t = torch.from_numpy(packed.view(np.int64)).pin_memory()
    packed_pin = packed_t.numpy().view(np.uint64)
    eng = fb.Engine(local_rank)
    stream = torch.cuda.ExternalStream(eng.stream_ptr(), device=torch.device("cuda", local_rank))
    shard = (0, n_reads)
    if world > 1:
        from flye_b200 import parallel
        eng.comm_init(world, rank, parallel.broadcast_unique_id(fb.Engine, dist, device=torch.device("cuda", local_rank)))
        shard = parallel.shard_reads(lens, world)[rank]
    common = dict(max_jump=int(cfg["maximum_jump"]), min_overlap=MIN_OVERLAP, max_overhang=int(cfg["maximum_overhang"]),
                  only_max_ext=True, nucl_alignment=bool(cfg["reads_base_alignment"]), use_hpc=bool(cfg["hpc_scoring_on"]))
    est_ids = pu.libc_rand_ids(2 * n_reads)
    # this rank's queries = the forward reads of its shard; counting / index emission work on the shard too and are
    # merged over NCCL inside the library (index replicated on every rank)
    lo, hi = shard[0], shard[0] + shard[1]
    queries = np.arange(2 * lo, 2 * hi, 2, dtype=np.uint32)
    phase_ms, phase_calls, ovl_stats, n_ovl, wall = {}, {}, {}, 0, {}
    n_raw = [0]
    ovl_len_sample = np.zeros(0)
    edit_sample = [None, 0.0]
    last = {}

    def step(upload):
        nonlocal phase_ms, phase_calls, ovl_stats, n_ovl, wall, ovl_len_sample
        phase_ms, phase_calls, wall = {}, {}, {}
        t_prev = [time.perf_counter()]

        def lap(name):
            now = time.perf_counter()
            wall[name] = wall.get(name, 0.0) + (now - t_prev[0]) * 1e3
            t_prev[0] = now

        def grab():
            t = eng.timings()
            for name, ms in t.items():
                phase_ms[name] = phase_ms.get(name, 0.0) + ms
                phase_calls[name] = phase_calls.get(name, 0) + eng.last_calls[name]
        if upload:
            eng.upload_packed(packed_pin, woff, lens)
            lap("upload")
        if world > 1:
            eng.set_shard(shard[0], shard[1])
        if int(cfg["use_minimizers"]):
            eng.build_index_minimizers(k, 1, int(cfg["minimizer_window"]), cfg["repeat_kmer_rate"])
        else:
            eng.count_kmers(k)
            grab()
            lap("count")
            eng.build_index_solid(2, cfg["meta_read_top_kmer_rate"], int(cfg["meta_read_filter_kmer_freq"]), cfg["repeat_kmer_rate"],
                                  float(int(cfg["assemble_kmer_sample"])))
        grab()
        lap("index")
        # estimateOverlaperParameters' 1000 random queries and the all-reads pass share ONE device batch: the device
        # work does not depend on the divergence threshold (it only gates the final host-side filter, overlap.cpp:470),
        # so the threshold is computed from the first 1000 result vectors and applied to the rest on the host.
        # multi-GPU: the 1000 estimate queries are dealt round-robin to the ranks; the per-query divergences are
        # all-gathered (a few KB) so that every rank derives the same threshold
        my_est = np.asarray(est_ids[rank::world], dtype=np.uint32)
        all_q = np.concatenate([my_est, queries])
        # Presets with an ABSOLUTE threshold (asm_hifi: assemble_divergence_relative = 0) know it before the estimate: the main
        # queries carry it (per-query thresholds), so the library can stop the edit-distance work of an overlap as soon as
        # it cannot pass; the estimate queries stay unfiltered (1.0) as in overlap.cpp:744-790.
        relative = bool(cfg["assemble_divergence_relative"])
        qthr = None
        if not relative:
            qthr = np.concatenate([np.full(len(my_est), 1.0, np.float32), np.full(len(queries), np.float32(cfg["assemble_ovlp_divergence"]), np.float32)])
        offs, ov, ovl_stats = eng.overlaps(all_q, max_divergence=1.0, copy=False, query_max_divergence=qthr, **common)
        first = int(offs[len(my_est)])
        rng_est = ov["cur_end"][:first] - ov["cur_begin"][:first]
        div_all = ov["seq_divergence"]
        # divergence of the longest overlap of every estimate query (first one on ties, as np.argmax / overlap.cpp:768-783)
        divs = []
        if first:
            o_est = np.asarray(offs[:len(my_est) + 1], dtype=np.int64)
            starts = o_est[:-1][np.diff(o_est) > 0]
            big = first + 1
            keyv = rng_est.astype(np.int64) * big + (big - 1 - np.arange(first, dtype=np.int64))
            best = np.maximum.reduceat(keyv, starts)
            divs = div_all[big - 1 - (best % big)].tolist()
        if world > 1:
            cap = (len(est_ids) + world - 1) // world
            buf = torch.full((cap + 1,), float("nan"), dtype=torch.float32, device="cuda")
            buf[0] = len(divs)
            if divs:
                buf[1:1 + len(divs)] = torch.tensor(np.asarray(divs, dtype=np.float32), device="cuda")
            gathered = [torch.empty_like(buf) for _ in range(world)]
            dist.all_gather(gathered, buf)
            divs = []
            for g in gathered:
                g = g.cpu().numpy()
                divs.extend(g[1:1 + int(g[0])].tolist())
        mean = pu.median_f32(divs) if divs else np.float32(0.5)
        max_div = np.float32((mean if relative else np.float32(0.0)) + np.float32(cfg["assemble_ovlp_divergence"]))
        lap("overlaps")
        grab()
        n_raw[0] = int(offs[-1])
        if relative:   # setDivergenceThreshold (overlap.cpp:817-827): the library re-applies the threshold to the main queries
            offs, ov = eng.refilter(len(my_est), float(max_div), len(all_q))
            div_all = ov["seq_divergence"]
            lap("refilter")
        n_ovl = int(offs[-1]) - int(offs[len(my_est)])
        ovl_len_sample = (ov["cur_end"][first:first + 20000] - ov["cur_begin"][first:first + 20000]).astype(np.float64)
        if common["nucl_alignment"] and len(ov) > first:
            ed, al = ov["edit_distance"][first:first + 20000].astype(np.float64), ov["aln_len"][first:first + 20000].astype(np.float64)
            edit_sample[0] = float(np.mean(ed))
            edit_sample[1] = float(np.mean(al * (np.ceil((2 * ed + 1) / 64.0) + 1)))   # Myers word updates of a band of 2d+1 (edlib's regime)
        lap("filter")
        last["offs"], last["ov"], last["n_est"] = offs, ov, len(my_est)
        return n_ovl

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    step_detail = []   # API wall clock and host-side phase times of every timed step of this rank
    step_walls = []   # host wall clock of every single step of this rank, one list per timed region
    per_rank = []   # ms per step of every rank, one list per timed region (resident, e2e, roofline pass)

    def timed(upload, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        walls = []
        for _ in range(steps):
            t0 = time.perf_counter()
            step(upload)
            walls.append(round((time.perf_counter() - t0) * 1e3, 2))
            step_detail.append({"wall": {p: round(v, 2) for p, v in wall.items()},
                                "host": {p: round(v, 2) for p, v in phase_ms.items() if p.startswith(("host_", "prep_"))}})
        step_walls.append(walls)
        e1.record(stream)
        barrier()
        ms = e0.elapsed_time(e1) / steps
        per_rank.append([round(ms, 3)])
        if world > 1:
            t = torch.tensor([ms], device="cuda")
            g = [torch.empty_like(t) for _ in range(world)]
            dist.all_gather(g, t)
            per_rank[-1] = [round(float(x.item()), 3) for x in g]
            ms = max(per_rank[-1])
        return ms

    eng.upload_packed(packed_pin, woff, lens)
    if world > 1:
        eng.set_shard(shard[0], shard[1])
    sampler = ClockSampler(local_rank)
    sampler.start()
    for _ in range(args.warmup):
        step(False)
    sampler.mark()
    launches0 = eng.launches()
    ms_resident = timed(False, args.steps)
    launches = (eng.launches() - launches0) // max(1, args.steps)
    resident_phases, resident_calls, stats, resident_wall = dict(phase_ms), dict(phase_calls), dict(ovl_stats), dict(wall)
    ms_e2e, e2e_wall, e2e_phases = None, {}, {}
    if not args.no_e2e:
        ms_e2e = timed(True, args.steps)
        e2e_wall = dict(wall)
        e2e_phases = dict(phase_ms)
    # roofline pass: the same resident step with ONE lane, so that every kernel runs alone on the device and its CUDA-event
    # duration is its own (with two lanes the kernels of two sub-batches share the SMs and each event interval also covers the
    # other lane's work).  `value` / `e2e` above are NOT taken from this pass.
    prev_lanes = os.environ.get("FG_LANES")
    os.environ["FG_LANES"] = "1"
    roof_steps = max(1, min(args.steps, 3))
    step(False)
    ms_roof = timed(False, roof_steps)
    roof_phases, roof_calls = dict(phase_ms), dict(phase_calls)
    if prev_lanes is None:
        del os.environ["FG_LANES"]
    else:
        os.environ["FG_LANES"] = prev_lanes
    clocks = sampler.stop()
    int_peak = eng.int_peak() if hasattr(eng, "int_peak") else None

    # ---- parity: digest of the ordered overlap vectors of the last step, all ranks, against the reference's digest ----
    n_est = last["n_est"]
    my_offs = np.asarray(last["offs"][n_est:], dtype=np.uint64)
    dq = query_digests(queries, my_offs - my_offs[0], last["ov"][int(my_offs[0]):int(my_offs[-1])])
    if world > 1:
        counts = [None] * world
        dist.all_gather_object(counts, len(dq))
        mx = max(counts)
        pad = torch.zeros((mx, 32), dtype=torch.uint8, device="cuda")
        if len(dq):
            pad[:len(dq)] = torch.from_numpy(dq).cuda()
        gathered = [torch.empty_like(pad) for _ in range(world)]
        dist.all_gather(gathered, pad)
        dq = np.concatenate([g[:c].cpu().numpy() for g, c in zip(gathered, counts)])
        t = torch.tensor([n_ovl, n_raw[0]], device="cuda", dtype=torch.int64)
        dist.all_reduce(t)
        n_ovl, n_raw[0] = int(t[0].item()), int(t[1].item())
    final, chunks = fold_digests(dq)
    parity = {"digest": final, "queries": int(len(dq)), "overlaps": int(n_ovl), "golden": None, "match": None,
              "what": "sha256 per query of the ordered overlap dump (all integer fields + divergence bits), folded in query order over all ranks"}
    gp = golden_path(args.workload)
    if args.scale == 1.0 and os.path.exists(gp):
        g = json.load(open(gp))
        parity["golden"] = g["final"]
        parity["golden_source"] = "tests/golden/full_scale/%s.json (dump of the unmodified reference, %d overlaps)" % (args.workload, g["overlaps"])
        parity["match"] = bool(g["final"] == final and g["overlaps"] == n_ovl)
        if not parity["match"]:
            parity["first_bad_chunk"] = next((i for i, (a, b) in enumerate(zip(chunks, g["chunks"])) if a != b), min(len(chunks), len(g["chunks"])))
    if args.write_golden and rank == 0:
        print(json.dumps({"queries": len(dq), "overlaps": n_ovl, "chunk_queries": 1000, "final": final, "chunks": chunks}))

    total_reads = n_reads   # all ranks together process every forward read exactly once
    value = total_reads / (ms_resident / 1e3)
    h2d = int(packed.nbytes + woff.nbytes + lens.nbytes + 4 * (len(queries) + len(est_ids)))
    raw_records = int(resident_phases.get("raw_overlaps", n_raw[0])) * world   # records the library copied back (before the threshold)
    d2h = int(72 * max(raw_records, n_raw[0]) + 8 * (len(queries) + len(est_ids) + 1))

    # roofline (algorithmic bytes: SURVEY.md §8d stage formulas, assigned to kernels in DESIGN.md §4).  Only phases that
    # consist of ONE launch of one of our kernels are listed; their durations are CUDA-event times on the library's stream(s).
    M, O = stats.get("n_hits", 0), n_ovl // max(world, 1)
    w = 4 if 2 * k <= 32 else 8
    shard_bases = int(lens[lo:hi].sum())
    n_k = shard_bases - k * (hi - lo)
    raw_ovl = int(resident_phases.get("gathered_overlaps", resident_phases.get("raw_overlaps", n_raw[0])))   # records the edit-distance kernel aligned (before the divergence test)
    mean_ovl_len = float(np.mean(np.maximum(ovl_len_sample, 1))) if len(ovl_len_sample) else 0.0
    k8 = 12.0 * M + 40.0 * O                        # K8 as a whole; its four kernels share it evenly
    alg = {"expand": 20.0 * M,                       # K6, per-hit part: 8 B index entry read + 12 B match written
           "lookup": (w + 8.0) * n_k,                # K6, per-position part
           "hit_sort_radix": 24.0 * M,               # K7: 12 B record read + 12 B record written, once
           "chain_prep": k8 / 4, "chain_dp": k8 / 4, "chain_fill": k8 / 4, "chain_walk": k8 / 4,
           "edit": raw_ovl * 2.0 * mean_ovl_len / 4.0,   # K9: (len_q + len_t) / 4 bytes per overlap
           "select": (w + 5.0) * n_k,                # K4
           "count": 0.25 * shard_bases + w * n_k + 2.0 * w * n_k}   # K2 + the counting part of K3 (3 w N_k without the (w+4) D of the table)
    names = {"expand": "expandKernel", "lookup": "queryLookupKernel", "hit_sort_radix": {"0": "segRadixSortKernel", "1": "segRadixSortClusterKernel", "2": "segTileSortKernel"}.get(os.environ.get("FG_SEG_SORT", "2"), "segTileSortKernel"),
             "chain_prep": "pairPrepKernel",
             "chain_dp": "chainRunDpKernel", "chain_fill": "chainFillKernel", "chain_walk": "chainWalkKernel", "edit": "wfaKernel",
             "select": "minimizerRegKernel" if int(cfg["use_minimizers"]) else "selectKernel", "count": "denseCountKernel"}
    # integer-bound kernels: algorithmic integer operations (~15 int ops per DP cell of the reference's scan; Myers
    # word update ~17 int ops per 64 cells of edlib's band) against the IMAD/LOP3/SHF issue rate measured in this run
    # chain_dp: the reference's scan visits `dp_cells_literal` predecessors on this read set (counted by the CPU restatement, recorded
    # next to the reference's digest); the kernel reaches the same scores visiting `n_dp_cells` run heads.  Like the edit distance,
    # the kernel is reported in units of the reference's algorithm; `work` has both counts.
    dp_cells_literal = None
    if args.scale == 1.0 and world == 1 and os.path.exists(gp):
        dp_cells_literal = json.load(open(gp)).get("dp_cells_literal")
    int_alg = {"chain_dp": 15.0 * (dp_cells_literal or stats.get("n_dp_cells", 0)), "edit": 17.0 * raw_ovl * edit_sample[1]}
    notes = {"edit": "integer issue: O(ND) wavefronts instead of edlib's banded bit-vectors; algorithmic ops = the reference's Myers word updates",
             "chain_dp": ("integer issue: scan over run heads; algorithmic ops = 15 per predecessor the reference's scan visits on this read set "
                          "(tests/golden/full_scale: dp_cells_literal)") if dp_cells_literal else
                         "integer issue: scan over run heads; algorithmic ops = 15 per run-head predecessor the kernel itself evaluates (no literal count recorded for this run)",
             "chain_walk": "latency (pointer chasing in shared memory)"}
    traffic_per_hit = {}
    try:
        with open(os.path.join(ROOT, "profiles", "traffic_per_hit.json")) as f:
            traffic_per_hit.update(json.load(f).get(args.workload, {}))
    except Exception:
        pass
    kernel_phases = {p: ms for p, ms in roof_phases.items() if p in alg and alg[p] > 0 and roof_calls.get(p, 1) >= 1}
    peak, peak_src = measured_peak()
    roofline, roofline_kernels = None, []
    for ph in sorted(kernel_phases, key=kernel_phases.get, reverse=True):
        n_l = max(1, roof_calls.get(ph, 1))
        ms_l = kernel_phases[ph] / n_l
        if ph in int_alg and int_peak:
            achieved = int_alg[ph] / n_l / (ms_l / 1e3) / 1e9
            entry = {"bound": "int", "kernel": names[ph], "phase": ph, "achieved": achieved, "peak": int_peak, "unit": "Gop/s",
                     "frac": achieved / int_peak, "traffic": traffic_per_hit[ph] * M / n_l if ph in traffic_per_hit else None,
                     "peak_source": "measured in this run (fg_debug_int_peak: IMAD/LOP3/SHF mix, all SMs)", "launches": n_l,
                     "ms_per_launch": ms_l, "algorithmic_ops_per_launch": int_alg[ph] / n_l, "note": notes.get(ph)}
            if ph == "chain_dp" and dp_cells_literal:
                # the same kernel in units of what it executes itself (run-head predecessors); `frac` above is in the reference's
                # units and exceeds 1 where the run compression skips more work than the pipe could have done
                ex = 15.0 * stats.get("n_dp_cells", 0) / n_l / (ms_l / 1e3) / 1e9
                entry["achieved_executed"], entry["frac_executed"] = ex, ex / int_peak
            roofline_kernels.append(entry)
            continue
        achieved = alg[ph] / n_l / (ms_l / 1e3) / 1e9
        roofline_kernels.append({"bound": "hbm", "kernel": names[ph], "phase": ph, "achieved": achieved, "peak": peak, "unit": "GB/s",
                                 "frac": achieved / peak, "traffic": traffic_per_hit[ph] * M / n_l if ph in traffic_per_hit else None,
                                 "peak_source": peak_src, "launches": n_l, "ms_per_launch": ms_l, "algorithmic_bytes_per_launch": alg[ph] / n_l,
                                 "note": notes.get(ph)})
    if roofline_kernels:
        roofline = dict(roofline_kernels[0])
        roofline["traffic_source"] = "bytes per hit measured with ncu --set full (profiles/), scaled to this launch's hits"
    lib_phases = ("index_sort", "hit_sort_radix_lib", "chain_order")
    line = {"metric": METRIC, "value": value, "unit": "reads/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_resident, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "int64", "data": "synthetic", "config": config, "clocks": clocks,
            "e2e": None if ms_e2e is None else {"value": total_reads / (ms_e2e / 1e3), "unit": "reads/s", "ms_per_step": ms_e2e,
                                                 "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
            "gpu_launches": int(launches), "roofline": roofline, "roofline_kernels": roofline_kernels, "parity": parity,
            "library_ms": {"cub_radix_sort_of_index_entries": round(resident_phases.get("index_sort", 0.0), 3),
                           "cub_radix_sort_of_hits": round(resident_phases.get("hit_sort_radix_lib", 0.0), 3),
                           "cub_radix_sort_of_pair_sizes": round(resident_phases.get("chain_order", 0.0), 3),
                           "note": "cub::DeviceScan / DeviceSelect calls (prefix sums, compaction of flags) run inside the phases lookup, group, emit, index_table"},
            "roofline_pass": {"lanes": 1, "steps": roof_steps, "ms_per_step": ms_roof,
                              "note": "kernel durations of `roofline` / `roofline_kernels` / `phases_ms_one_lane` are CUDA-event times of a pass with one "
                                      "lane (each kernel alone on the device); `value`, `e2e` and `phases_ms` come from the two-lane passes"},
            "phases_ms": {p: round(v, 3) for p, v in resident_phases.items()},
            "phases_ms_one_lane": {p: round(v, 3) for p, v in roof_phases.items()},
            "phases_ms_e2e": {p: round(v, 3) for p, v in e2e_phases.items()}, "ms_per_step_of_every_rank": per_rank, "step_wall_ms_rank0": step_walls, "step_detail_rank0": step_detail,
            "api_wall_ms": {p: round(v, 2) for p, v in resident_wall.items()}, "e2e_api_wall_ms": {p: round(v, 2) for p, v in e2e_wall.items()},
            "work": {"kmer_hits": int(M), "target_groups": int(stats.get("n_pairs", 0)), "dp_pairs": int(stats.get("n_dp_pairs", 0)),
                     "dp_cells": int(stats.get("n_dp_cells", 0)), "dp_cells_literal": dp_cells_literal, "overlaps": int(n_ovl),
                     "mean_edit_distance_first_20k_overlaps": edit_sample[0],
                     "queries_with_hit_ties": int(resident_phases.get("tied_queries", 0)),
                     "pairs_score_order_presorted": int(resident_phases.get("presorted_pairs", 0))}}
    _ = lib_phases
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            _, _, _, _, _, cb = run_reference(argparse.Namespace(warmup=0, steps=1), wl, reads_path)
            line["cpu_baseline"] = cb
        except Exception as e:   # the baseline is a reported number, never a reason to lose the GPU line
            line["cpu_baseline"] = {"value": None, "unit": "reads/s", "cores": os.cpu_count(), "kind": "unavailable", "sample": str(e)}
    if rank == 0 and not args.write_golden:
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    if parity["match"] is False:
        log("PARITY FAILURE: digest %s != golden %s (first bad chunk of 1000 queries: %s)" % (final, parity["golden"], parity.get("first_bad_chunk")))
        sys.exit(3)


if __name__ == "__main__":
    main()
