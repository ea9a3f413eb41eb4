"""-m gpu: the CUDA path, called through the C ABI, against the oracle on the same seeded inputs.
Bar: bit-exact for every integer field, the k-mer histogram, the whole index, and the divergence float bits."""
import ctypes
import os

import numpy as np
import pytest

import parity_util as pu

pytestmark = pytest.mark.gpu
RAW = os.path.join(pu.CFG_DIR, "raw_reads.cfg")
HIFI = os.path.join(pu.CFG_DIR, "hifi.cfg")


def _std_sort(keys, vals, seg):
    lib = ctypes.CDLL(os.path.join(pu.ROOT, "tests", "cpu_models", "_bin", "libstdsort.so"))
    k, v = keys.copy(), vals.copy()
    lib.std_sort_segments(k.ctypes.data_as(ctypes.c_void_p), v.ctypes.data_as(ctypes.c_void_p),
                          seg.ctypes.data_as(ctypes.c_void_p), ctypes.c_uint32(len(seg) - 1))
    return k, v


def test_device_introsort_matches_std_sort(engine):
    """The warp introsort must reproduce libstdc++ std::sort's permutation, ties included."""
    rng = np.random.default_rng(5)
    sizes = [0, 1, 2, 16, 17, 18, 31, 32, 33, 34, 47, 48, 49, 50, 63, 64, 65, 66, 96, 97, 127, 128, 129, 1000, 4097]
    sizes += list(rng.integers(1, 300, 300)) + list(rng.integers(300, 6000, 60)) + [70000, 150001]
    # ranges above 8192 elements take the CTA-wide partition (sortHugeKernel): every key pattern at several sizes
    sizes += [8193, 9000, 12345, 16384, 20001, 33333, 40000, 50000, 65537, 8200, 100000, 262145, 30000, 11111]
    keys, seg = [], [0]
    for i, n in enumerate(sizes):
        mode = i % 6
        if mode == 0:
            k = rng.integers(0, max(2, n // 3 + 1), n)
        elif mode == 1:
            k = rng.integers(0, 4, n)
        elif mode == 2:
            k = np.arange(n) // 2
        elif mode == 3:
            k = (n - np.arange(n)) // 3
        elif mode == 4:
            k = np.full(n, 9)
        else:
            k = rng.integers(0, 1 << 40, n)
        keys.append(k.astype(np.uint64))
        seg.append(seg[-1] + n)
    keys = np.concatenate(keys)
    vals = np.arange(len(keys), dtype=np.uint32)
    seg = np.array(seg, dtype=np.uint64)
    gk, gv = engine.debug_warp_sort(keys, vals, seg)
    sk, sv = _std_sort(keys, vals, seg)
    assert np.array_equal(gk, sk)
    bad = np.nonzero(gv != sv)[0]
    assert len(bad) == 0, "first mismatch at %d (segment %d)" % (bad[0], np.searchsorted(seg, bad[0], side="right") - 1)


def _compare(tmp, name, exts):
    res = {}
    for e in exts:
        n, sample = pu.diff_files(os.path.join(tmp, "ref." + e), os.path.join(tmp, "gpu." + e))
        res[e] = n
        if n:
            print("%s .%s: %d differing lines; first: %s" % (name, e, n, sample[:4]))
    return res


def test_clr_small_full_parity(engine, tmp_path):
    tmp = str(tmp_path)
    reads = pu.simulate(os.path.join(tmp, "r.fasta"), genome_len=200000, coverage=20, seed=7)
    ref = pu.run_oracle(reads, RAW, os.path.join(tmp, "ref"), k=15, extra=["--dump-index"])
    _, info = pu.gpu_pipeline(reads, RAW, os.path.join(tmp, "gpu"), k=15, dump_index=True, engine=engine)
    print(info)
    res = _compare(tmp, "clr_small", ["hist", "index", "ovlp"])
    assert ref["overlaps"] > 1000
    assert res == {"hist": 0, "index": 0, "ovlp": 0}


@pytest.mark.parametrize("env", [{"FG_HIT_RADIX": "0"}, {"FG_HIT_RADIX": "2"}, {"FG_DP_MODE": "0"}, {"FG_DP_MODE": "1"},
                                 {"FG_HIT_RADIX": "0", "FG_DP_MODE": "0"}, {"FG_HIT_RADIX": "0", "FG_SORT_HUGE": "0"},
                                 {"FG_SEG_SORT": "1", "FG_SRS_CLUSTER": "1"}, {"FG_SEG_SORT": "1", "FG_SRS_CLUSTER": "2"},
                                 {"FG_SEG_SORT": "1", "FG_SRS_CLUSTER": "4", "FG_SRS_MIN_PASSES": "3"}, {"FG_SEG_SORT": "1", "FG_SRS_CLUSTER": "8"},
                                 {"FG_SRS_MIN_PASSES": "3"}, {"FG_DEVICE_EPILOGUE": "0"}, {"FG_SRS_DEEP": "0"},
                                 {"FG_SRS_DEEP": "0", "FG_SRS_MIN_PASSES": "3"}, {"FG_SEG_SORT": "2"}, {"FG_SEG_SORT": "2", "FG_SRS_MIN_PASSES": "3"},
                                 {"FG_SEG_SORT": "2", "FG_SEG_OCC": "4"}, {"FG_SEG_SORT": "0"}])
def test_kernel_variants_agree_with_oracle(engine, tmp_path, monkeypatch, env):
    """The alternative device paths stay exact: introsort emulation for every query (no radix fast path; with and without the
    CTA-wide partition of long ranges), the library radix sort instead of the segmented one, the thread-block-cluster version of the
    segmented sort with every cluster size, three radix passes, the segmented sort with the shallow prefetch, the tile version of
    the segmented sort (staged in shared memory, coalesced copy-out) and the register-scatter version, match-by-match DP with and without
    pruned look-back (the default is the run-compressed DP + run-wise chain walk), the host epilogue (divergence with the host's
    logf and per-query replay on host threads; the default computes both on the device).  CLR with all primary overlaps +
    kmerMatches, and HiFi."""
    for name, value in env.items():
        monkeypatch.setenv(name, value)
    tmp = str(tmp_path)
    reads = pu.simulate(os.path.join(tmp, "r.fasta"), genome_len=150000, coverage=15, seed=11)
    pu.run_oracle(reads, RAW, os.path.join(tmp, "ref"), k=15, extra=["--all-ext", "--keep-aln"])
    pu.gpu_pipeline(reads, RAW, os.path.join(tmp, "gpu"), k=15, engine=engine, all_ext=True, keep_aln=True)
    assert _compare(tmp, "variants clr %s" % env, ["ovlp"]) == {"ovlp": 0}
    reads = pu.simulate(os.path.join(tmp, "h.fasta"), genome_len=150000, coverage=15, mean_len=12000, shape=20, error=0.005, seed=12)
    pu.run_oracle(reads, HIFI, os.path.join(tmp, "ref"))
    pu.gpu_pipeline(reads, HIFI, os.path.join(tmp, "gpu"), engine=engine)
    assert _compare(tmp, "variants hifi %s" % env, ["ovlp"]) == {"ovlp": 0}


def test_clr_options(engine, tmp_path):
    """either strand as query, forceLocal, maxOverlaps, all primary overlaps per target"""
    tmp = str(tmp_path)
    reads = pu.simulate(os.path.join(tmp, "r.fasta"), genome_len=120000, coverage=15, seed=3)
    for opts, kw in [(["--both-strands", "--force-local", "--max-overlaps", "5"], dict(both_strands=True, force_local=True, max_overlaps=5)),
                     (["--all-ext"], dict(all_ext=True)),
                     (["--all-ext", "--keep-aln"], dict(all_ext=True, keep_aln=True)),
                     (["--keep-aln", "--both-strands"], dict(keep_aln=True, both_strands=True))]:
        pu.run_oracle(reads, RAW, os.path.join(tmp, "ref"), k=15, extra=opts)
        pu.gpu_pipeline(reads, RAW, os.path.join(tmp, "gpu"), k=15, engine=engine, **kw)
        res = _compare(tmp, "clr_opts %s" % opts, ["ovlp"])
        assert res == {"ovlp": 0}


def test_clr_k17_parity(engine, tmp_path):
    """k = 17 takes the 64-bit key path.  The oracle here is the restatement (the reference's k = 17 flat counter needs 8 GiB and a
    17 G-entry scan); its dumps of exactly this read set are pinned to the unmodified reference's by sha256
    (tests/golden/pins/k17_reference_pins.json, checked by tests/test_oracle_golden.py)."""
    tmp = str(tmp_path)
    reads = pu.simulate(os.path.join(tmp, "r.fasta"), genome_len=150000, coverage=18, error=0.10, seed=21)
    pu.run_oracle(reads, RAW, os.path.join(tmp, "ref"), binary=pu.RESTATE, extra=["--dump-index"])  # the reference's k=17 flat counter needs 8 GiB + a 17 G-entry scan
    pu.gpu_pipeline(reads, RAW, os.path.join(tmp, "gpu"), dump_index=True, engine=engine)
    res = _compare(tmp, "clr_k17", ["hist", "index", "ovlp"])
    assert res == {"hist": 0, "index": 0, "ovlp": 0}


def _edit_distance_np(a, b):
    prev = np.arange(len(b) + 1)
    for i in range(1, len(a) + 1):
        cur = np.empty_like(prev)
        cur[0] = i
        sub = prev[:-1] + (b != a[i - 1])
        best = np.minimum(sub, prev[1:] + 1)
        # left dependency: cur[j] = min(best[j-1], cur[j-1]+1) -> prefix scan
        cur[1:] = best
        cur = np.minimum.accumulate(cur - np.arange(len(cur))) + np.arange(len(cur))
        prev = cur
    return int(prev[-1])


def test_device_edit_distance_exact(engine):
    """wavefront edit distance == textbook DP (global, unit costs), incl. empty / unequal / unrelated inputs"""
    rng = np.random.default_rng(11)
    cases = [(0, 0, 0.0), (0, 7, 0.0), (9, 0, 0.0), (1, 1, 0.5), (50, 50, 0.0), (300, 300, 0.02), (1500, 1500, 0.01),
             (1200, 900, 0.3), (700, 700, 1.0), (2000, 2000, 0.1), (33, 64, 0.2), (31, 32, 0.1), (64, 64, 0.0),
             (4000, 4100, 0.002), (1025, 1024, 0.0)]
    for n, m, err in cases:
        a = rng.integers(0, 4, n).astype(np.uint8)
        if err >= 1.0:
            b = rng.integers(0, 4, m).astype(np.uint8)
        else:
            b = []
            for c in a[:m] if m <= n else a:
                u = rng.random()
                if u < err / 3:
                    continue
                if u < 2 * err / 3:
                    b.append(rng.integers(0, 4)); b.append(c); continue
                b.append((c + 1) % 4 if u < err else c)
            b = np.array(b, dtype=np.uint8) if len(b) else np.zeros(0, np.uint8)
        want = _edit_distance_np(a, b)
        assert engine.debug_edit_distance(a, b) == want, (n, m, err)
        # strand views: d(rc(a), rc(b)) computed through the reverse-complement fetch path equals d(a', b') of the
        # explicitly reverse-complemented strings
        ra, rb = (3 - a[::-1]).astype(np.uint8), (3 - b[::-1]).astype(np.uint8)
        assert engine.debug_edit_distance_rc(ra, True, rb, True) == want, (n, m, err, "rc/rc")
        assert engine.debug_edit_distance_rc(ra, True, b, False) == want, (n, m, err, "rc/fwd")
        assert engine.debug_edit_distance_rc(a, False, rb, True) == want, (n, m, err, "fwd/rc")


def test_device_edit_distance_long_and_bounded(engine, monkeypatch):
    """Distances beyond WF_DMAX (254) continue in global memory; FG_DEBUG_ED_LIMIT bounds the distance the way the divergence
    threshold does in wfaKernel (exact below the limit, '>= limit' otherwise; the wavefronts are pruned to the diamond the limit
    allows)."""
    rng = np.random.default_rng(3)
    for t in range(5):
        n = int(rng.integers(2500, 5000))
        a = rng.integers(0, 4, n).astype(np.uint8)
        err = [0.03, 0.07, 0.05, 0.02, 0.09][t]
        b = []
        for c in a:
            u = rng.random()
            if u < err / 3:
                continue
            if u < 2 * err / 3:
                b.append(rng.integers(0, 4)); b.append(c); continue
            b.append((c + 1) % 4 if u < err else c)
        b = np.array(b, dtype=np.uint8)
        d = _edit_distance_np(a, b)
        for limit in (None, d + 1, d + 40, 3 * d, d, max(1, d - 10), 255, 256):
            if limit is None:
                monkeypatch.delenv("FG_DEBUG_ED_LIMIT", raising=False)
            else:
                monkeypatch.setenv("FG_DEBUG_ED_LIMIT", str(limit))
            got = engine.debug_edit_distance(a, b)
            if limit is None or d < limit:
                assert got == d, (n, len(b), d, limit, got)
            else:
                assert got >= limit, (n, len(b), d, limit, got)
    monkeypatch.delenv("FG_DEBUG_ED_LIMIT", raising=False)


def test_hifi_small_full_parity(engine, tmp_path):
    """minimizer index + homopolymer-compressed edit-distance divergence (BASELINE configs 2/5 code path)"""
    tmp = str(tmp_path)
    reads = pu.simulate(os.path.join(tmp, "r.fasta"), genome_len=150000, coverage=15, mean_len=8000, shape=20, error=0.005, seed=11)
    ref = pu.run_oracle(reads, HIFI, os.path.join(tmp, "ref"), extra=["--dump-index", "--both-strands"])
    _, info = pu.gpu_pipeline(reads, HIFI, os.path.join(tmp, "gpu"), dump_index=True, both_strands=True, engine=engine)
    print(info)
    res = _compare(tmp, "hifi_small", ["index", "ovlp"])
    assert ref["overlaps"] > 1000
    assert res == {"index": 0, "ovlp": 0}


def test_clr_quarter_scale_parity(engine, tmp_path):
    """BASELINE config 1 at a quarter of the genome (1.15 Mb, 50x, 12 % error): ~7.6 k reads, ~110 M k-mer hits,
    several sub-batches; every overlap field identical to the unmodified reference."""
    tmp = str(tmp_path)
    reads = pu.simulate(os.path.join(tmp, "r.fasta"), genome_len=1150000, coverage=50, seed=1)
    ref = pu.run_oracle(reads, RAW, os.path.join(tmp, "ref"), k=15)
    os.environ["FG_HIT_BUDGET"] = str(24 << 20)     # force several sub-batches ...
    os.environ["FG_CHUNK_SLOTS"] = str(20 << 20)    # ... inside several query chunks
    try:
        _, info = pu.gpu_pipeline(reads, RAW, os.path.join(tmp, "gpu"), k=15, engine=engine)
    finally:
        del os.environ["FG_HIT_BUDGET"]
        del os.environ["FG_CHUNK_SLOTS"]
    print(info)
    res = _compare(tmp, "clr_quarter", ["hist", "ovlp"])
    assert ref["overlaps"] > 100000
    assert res == {"hist": 0, "ovlp": 0}


def test_hifi_1mb_parity(engine, tmp_path):
    """BASELINE config 2 code path at 1 Mb / 30x / 15 kb reads against the unmodified reference (edlib)"""
    tmp = str(tmp_path)
    reads = pu.simulate(os.path.join(tmp, "r.fasta"), genome_len=1000000, coverage=30, mean_len=15000, shape=20, error=0.005, seed=2)
    ref = pu.run_oracle(reads, HIFI, os.path.join(tmp, "ref"))
    _, info = pu.gpu_pipeline(reads, HIFI, os.path.join(tmp, "gpu"), engine=engine)
    print(info)
    res = _compare(tmp, "hifi_1mb", ["ovlp"])
    assert ref["overlaps"] > 20000
    assert res == {"ovlp": 0}


def test_metagenome_uneven_coverage_parity(engine, tmp_path):
    """BASELINE config 4 code path at small scale: several genomes with log-distributed coverage (4x .. 16x), k = 15;
    the reference uses buildIndexUnevenCoverage for every solid-k-mer run"""
    tmp = str(tmp_path)
    reads = pu.simulate(os.path.join(tmp, "r.fasta"), coverage=0, mean_len=9000, error=0.10, seed=4,
                        extra=["--meta", "6", "--meta-total", "600000"])
    ref = pu.run_oracle(reads, RAW, os.path.join(tmp, "ref"), k=15, extra=["--dump-index"])
    _, info = pu.gpu_pipeline(reads, RAW, os.path.join(tmp, "gpu"), k=15, dump_index=True, engine=engine)
    res = _compare(tmp, "meta", ["hist", "index", "ovlp"])
    assert ref["overlaps"] > 2000
    assert res == {"hist": 0, "index": 0, "ovlp": 0}


def test_ont_like_k17_parity(engine, tmp_path):
    """BASELINE config 3 code path at small scale: ONT-like reads (10 % error, mean 19 kb), cfg default k = 17 (restatement as the
    oracle, pinned to the reference on this read set: tests/golden/pins/k17_reference_pins.json)"""
    tmp = str(tmp_path)
    reads = pu.simulate(os.path.join(tmp, "r.fasta"), genome_len=300000, coverage=20, mean_len=19000, error=0.10, seed=3)
    ref = pu.run_oracle(reads, RAW, os.path.join(tmp, "ref"), binary=pu.RESTATE, extra=["--dump-index"])
    _, info = pu.gpu_pipeline(reads, RAW, os.path.join(tmp, "gpu"), dump_index=True, engine=engine)
    res = _compare(tmp, "ont_k17", ["hist", "index", "ovlp"])
    assert ref["overlaps"] > 2000
    assert res == {"hist": 0, "index": 0, "ovlp": 0}
