"""Shared helpers of the parity tests: run the oracle, run the CUDA path through the C ABI, compare dumps."""
import ctypes
import json
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

REF_HARNESS = os.path.join(ROOT, "oracle", "_ref", "flye_ref_harness")
RESTATE = os.path.join(ROOT, "oracle", "_ref", "flye_restate")
SIMREADS = os.path.join(ROOT, "tools", "_bin", "simreads")
CFG_DIR = os.path.join(ROOT, "tests", "cfg")


def load_cfg(path):
    kv = {}
    with open(path) as f:
        for line in f:
            line = line.strip()
            if not line or line.startswith("#") or "=" not in line:
                continue
            k, v = line.split("=", 1)
            kv[k.strip()] = float(np.float32(float(v.strip())))   # Config stores floats (config.h:69)
    return kv


def simulate(out, genome_len=200000, coverage=20, mean_len=7500, shape=2, error=0.12, seed=1, extra=()):
    cmd = [SIMREADS, "--out", out, "--genome-len", str(genome_len), "--coverage", str(coverage), "--mean-len", str(mean_len),
           "--shape", str(shape), "--error", str(error), "--seed", str(seed)] + list(extra)
    subprocess.run(cmd, check=True, stderr=subprocess.DEVNULL)
    return out


def oracle_binary():
    """The unmodified reference when it was built here (kind 'reference'), else the CPU restatement ('port')."""
    if os.path.exists(REF_HARNESS):
        return REF_HARNESS, "reference"
    return RESTATE, "port"


def run_oracle(reads, cfg, out_prefix, k=None, threads=None, binary=None, extra=()):
    binary = binary or oracle_binary()[0]
    threads = threads or os.cpu_count() or 1
    cmd = [binary, "--reads", reads, "--cfg", cfg, "--out", out_prefix, "--threads", str(threads)]
    if k:
        cmd += ["--k", str(k)]
    cmd += list(extra)
    r = subprocess.run(cmd, check=True, stdout=subprocess.PIPE, text=True)
    return json.loads(r.stdout.strip().splitlines()[-1])


def libc_rand_ids(n_ids, count=1000):
    """The ids OverlapContainer::estimateOverlaperParameters draws: rand() % size with glibc's default seed
    (overlap.cpp:752-756)."""
    libc = ctypes.CDLL("libc.so.6")
    libc.srand(1)
    return [libc.rand() % n_ids for _ in range(count)]


def median_f32(values):
    """utils.h:31-51 quantile(vec, 50) on floats"""
    v = np.sort(np.asarray(values, dtype=np.float32))
    return v[min(len(v) * 50 // 100, len(v) - 1)]


def gpu_pipeline(reads_path, cfg_path, out_prefix, k=None, min_overlap=1000, dump_index=False, both_strands=False,
                 max_overlaps=0, force_local=False, all_ext=False, estimate=True, engine=None, max_queries=None, min_read_len=None,
                 keep_aln=False):
    """Mirror of oracle/harness.cpp on the CUDA path; writes the same dump files.  Returns (engine, info)."""
    import flye_b200 as fb
    cfg = load_cfg(cfg_path)
    k = k or int(cfg["kmer_size"])
    reads = fb.read_fasta(reads_path, min_overlap if min_read_len is None else min_read_len)
    eng = engine or fb.Engine(0)
    eng.upload_ascii(reads)
    info = {"reads": len(reads)}
    if int(cfg["use_minimizers"]):
        st = eng.build_index_minimizers(k, 1, int(cfg["minimizer_window"]), cfg["repeat_kmer_rate"])
    else:
        info["distinct"] = eng.count_kmers(k)
        fb.dump_hist(eng.kmer_hist(), out_prefix + ".hist")
        info["t_count"] = eng.timings()
        st = eng.build_index_solid(2, cfg["meta_read_top_kmer_rate"], int(cfg["meta_read_filter_kmer_freq"]),
                                   cfg["repeat_kmer_rate"], float(int(cfg["assemble_kmer_sample"])))
    info["t_index"] = eng.timings()
    info["stats"] = {f: getattr(st, f) for f, _ in st._fields_}
    if dump_index:
        fb.dump_index(eng, st, out_prefix + ".index")
    common = dict(max_jump=int(cfg["maximum_jump"]), min_overlap=min_overlap, max_overhang=int(cfg["maximum_overhang"]),
                  only_max_ext=not all_ext, nucl_alignment=bool(cfg["reads_base_alignment"]), use_hpc=bool(cfg["hpc_scoring_on"]))
    max_div = np.float32(1.0)
    if estimate:
        ids = libc_rand_ids(2 * len(reads))
        offs, ov, _ = eng.overlaps(ids, max_divergence=1.0, **common)
        divs = []
        for i in range(len(ids)):
            a, b = int(offs[i]), int(offs[i + 1])
            if b > a:
                rng = ov["cur_end"][a:b] - ov["cur_begin"][a:b]
                divs.append(ov["seq_divergence"][a + int(np.argmax(rng))])   # first maximum (strict >, overlap.cpp:768)
        mean = median_f32(divs) if divs else np.float32(0.5)
        rel = bool(cfg["assemble_divergence_relative"])
        max_div = np.float32((mean if rel else np.float32(0.0)) + np.float32(cfg["assemble_ovlp_divergence"]))
    info["max_divergence"] = float(max_div)
    step = 1 if both_strands else 2
    queries = list(range(0, 2 * len(reads), step))
    if max_queries is not None:
        queries = queries[:max_queries]
    offs, ov, stats = eng.overlaps(queries, max_divergence=float(max_div), max_overlaps=max_overlaps, force_local=force_local,
                                   keep_alignment=keep_aln, **common)
    info["t_overlaps"] = eng.timings()
    info["ovl_stats"] = stats
    info["n_overlaps"] = int(offs[-1])
    fb.dump_overlaps(queries, offs, ov, out_prefix + ".ovlp", aln_pairs=stats.pop("aln_pairs", None))
    return eng, info


def diff_files(a, b, max_show=8):
    """Returns (n_differing_lines, sample) comparing two text dumps line by line."""
    with open(a) as fa, open(b) as fb_:
        la, lb = fa.read().splitlines(), fb_.read().splitlines()
    n = abs(len(la) - len(lb))
    sample = []
    for i, (x, y) in enumerate(zip(la, lb)):
        if x != y:
            n += 1
            if len(sample) < max_show:
                sample.append((i + 1, x, y))
    if len(la) != len(lb):
        sample.append(("len", len(la), len(lb)))
    return n, sample
