"""-m gpu: edge cases the reference handles implicitly — reads shorter than k, identical reads (every k-mer hit is a
tie for std::sort), tandem repeats (tandem filter + repetitive k-mers), low-complexity sequence (minimizer tie quirk,
homopolymer compression), empty / duplicated / reverse-strand query lists."""
import os

import numpy as np
import pytest

import parity_util as pu

pytestmark = pytest.mark.gpu
RAW = os.path.join(pu.CFG_DIR, "raw_reads.cfg")
HIFI = os.path.join(pu.CFG_DIR, "hifi.cfg")


def _rand(rng, n):
    return "".join(rng.choice(list("ACGT"), n))


def _mutate(rng, s, rate):
    out = []
    for c in s:
        u = rng.random()
        if u < rate / 3:
            continue
        if u < 2 * rate / 3:
            out.append(rng.choice(list("ACGT")))
        out.append(rng.choice(list("ACGT")) if u > 1 - rate / 3 else c)
    return "".join(out)


def _rc(s):
    return s[::-1].translate(str.maketrans("ACGT", "TGCA"))


def _write(path, seqs):
    with open(path, "w") as f:
        for i, s in enumerate(seqs):
            f.write(">e%d\n%s\n" % (i, s))


def _check(tmp, reads, cfg, k, engine, exts, ref_opts=(), **kw):
    pu.run_oracle(reads, cfg, os.path.join(tmp, "ref"), k=k, extra=["--dump-index", "--min-read-len", "0"] + list(ref_opts))
    pu.gpu_pipeline(reads, cfg, os.path.join(tmp, "gpu"), k=k, dump_index=True, engine=engine, min_read_len=0, **kw)
    for ext in exts:
        n, sample = pu.diff_files(os.path.join(tmp, "ref." + ext), os.path.join(tmp, "gpu." + ext))
        assert n == 0, (ext, sample[:3])


def test_identical_and_short_reads(engine, tmp_path):
    rng = np.random.default_rng(1)
    base = _rand(rng, 4000)
    seqs = [base] * 4 + [_rc(base)] * 2 + [base[500:3500], "ACGTACGTAC", "A" * 15, _rand(rng, 16), _rand(rng, 17), _rand(rng, 40)]
    seqs += [_mutate(rng, base, 0.08) for _ in range(6)]
    reads = os.path.join(str(tmp_path), "r.fasta")
    _write(reads, seqs)
    _check(str(tmp_path), reads, RAW, 15, engine, ["hist", "index", "ovlp"], ref_opts=["--both-strands", "--no-estimate"],
           both_strands=True, estimate=False)


def test_tandem_repeats_and_repetitive_kmers(engine, tmp_path):
    rng = np.random.default_rng(2)
    unit = _rand(rng, 23)
    genome = _rand(rng, 3000) + unit * 260 + _rand(rng, 3000) + unit * 150 + _rand(rng, 2500)
    seqs = []
    for _ in range(60):
        a = rng.integers(0, len(genome) - 5000)
        s = _mutate(rng, genome[a:a + rng.integers(3000, 5000)], 0.06)
        seqs.append(s if rng.random() < 0.5 else _rc(s))
    reads = os.path.join(str(tmp_path), "r.fasta")
    _write(reads, seqs)
    _check(str(tmp_path), reads, RAW, 15, engine, ["hist", "index", "ovlp"], ref_opts=["--no-estimate"], estimate=False)


def test_low_complexity_hifi(engine, tmp_path):
    rng = np.random.default_rng(3)
    genome = _rand(rng, 2500) + "A" * 400 + _rand(rng, 2000) + "AC" * 300 + _rand(rng, 2500) + "GGGTTT" * 120 + _rand(rng, 2000)
    seqs = []
    for _ in range(50):
        a = rng.integers(0, len(genome) - 4500)
        s = _mutate(rng, genome[a:a + rng.integers(3000, 4500)], 0.006)
        seqs.append(s if rng.random() < 0.5 else _rc(s))
    reads = os.path.join(str(tmp_path), "r.fasta")
    _write(reads, seqs)
    _check(str(tmp_path), reads, HIFI, None, engine, ["index", "ovlp"], ref_opts=["--both-strands", "--no-estimate"],
           both_strands=True, estimate=False)


def test_query_list_shapes(engine, tmp_path):
    """empty list, duplicates, reverse-strand ids, a read too short to have k-mers"""
    import flye_b200 as fb
    rng = np.random.default_rng(4)
    base = _rand(rng, 6000)
    seqs = [_mutate(rng, base[a:a + 3500], 0.05) for a in (0, 800, 1600, 2400)] + ["ACGT"]
    engine.upload_ascii([s.encode() for s in seqs])
    engine.count_kmers(15)
    engine.build_index_solid()
    offs, ov, _ = engine.overlaps([])
    assert len(offs) == 1 and offs[0] == 0 and len(ov) == 0
    q = [0, 0, 3, 8, 9, 2, 0]
    offs, ov, _ = engine.overlaps(q)
    per = [ov[int(offs[i]):int(offs[i + 1])] for i in range(len(q))]
    assert len(per[0]) > 0
    assert all(np.array_equal(per[0], per[j]) for j in (1, 6))          # the same query gives the same vector
    assert len(per[3]) == 0 and len(per[4]) == 0                        # read 4 ("ACGT") has no k-mers, either strand
    assert (per[2]["cur_id"] == 3).all() and (per[5]["cur_id"] == 2).all()
    with pytest.raises(fb.FlyeB200Error):
        engine.overlaps([10])                                           # id out of range -> FG_ERR_ARG


def test_per_query_divergence_thresholds(engine, tmp_path):
    """fg_overlap_params.query_max_divergence: every query is filtered with its own threshold (estimate queries at 1.0 next to
    thresholded ones in one batch), and the threshold-bounded edit distance never changes a kept record: the overlaps kept at
    0.01 are exactly the records of the unfiltered run whose divergence is below 0.01."""
    import flye_b200 as fb
    cfg = pu.load_cfg(os.path.join(pu.CFG_DIR, "hifi.cfg"))
    reads_path = pu.simulate(os.path.join(str(tmp_path), "h.fasta"), genome_len=120000, coverage=15, mean_len=12000, shape=20, error=0.01, seed=21)
    reads = fb.read_fasta(reads_path, 1000)
    engine.upload_ascii(reads)
    engine.build_index_minimizers(17, 1, int(cfg["minimizer_window"]), cfg["repeat_kmer_rate"])
    common = dict(max_jump=int(cfg["maximum_jump"]), min_overlap=1000, max_overhang=int(cfg["maximum_overhang"]), only_max_ext=True,
                  nucl_alignment=True, use_hpc=True)
    q = np.arange(0, min(2 * len(reads), 120), dtype=np.uint32)
    thr = np.where(np.arange(len(q)) % 3 == 0, 1.0, 0.01).astype(np.float32)
    off_a, ov_a, _ = engine.overlaps(q, max_divergence=1.0, **common)
    off_b, ov_b, _ = engine.overlaps(q, max_divergence=0.01, **common)
    off_m, ov_m, _ = engine.overlaps(q, max_divergence=0.5, query_max_divergence=thr, **common)
    fields = ["cur_id", "cur_begin", "cur_end", "ext_id", "ext_begin", "ext_end", "score", "seq_divergence", "edit_distance", "aln_len"]
    n_low = n_all = 0
    for i in range(len(q)):
        a = ov_a[int(off_a[i]):int(off_a[i + 1])]
        b = ov_b[int(off_b[i]):int(off_b[i + 1])]
        m = ov_m[int(off_m[i]):int(off_m[i + 1])]
        want_b = a[a["seq_divergence"] < np.float32(0.01)]
        for f in fields:
            assert np.array_equal(b[f], want_b[f]), (i, f)
            assert np.array_equal(m[f], (a if thr[i] == 1.0 else want_b)[f]), (i, f)
        n_low += len(want_b); n_all += len(a)
    assert 0 < n_low < n_all      # the threshold really removes overlaps on this input (1 % error per read)


def test_refilter_equals_thresholded_run(engine, tmp_path):
    """fg_overlaps_refilter (setDivergenceThreshold after the estimate): re-filtering the unthresholded result of the queries
    from `first` on gives exactly what a thresholded run returns for them; the queries before `first` keep everything."""
    import flye_b200 as fb
    cfg = pu.load_cfg(os.path.join(pu.CFG_DIR, "raw_reads.cfg"))
    reads_path = pu.simulate(os.path.join(str(tmp_path), "c.fasta"), genome_len=100000, coverage=15, seed=31)
    reads = fb.read_fasta(reads_path, 1000)
    engine.upload_ascii(reads)
    engine.count_kmers(15)
    engine.build_index_solid(2, cfg["meta_read_top_kmer_rate"], int(cfg["meta_read_filter_kmer_freq"]), cfg["repeat_kmer_rate"],
                             float(int(cfg["assemble_kmer_sample"])))
    common = dict(max_jump=int(cfg["maximum_jump"]), min_overlap=1000, max_overhang=int(cfg["maximum_overhang"]), only_max_ext=True)
    q = np.arange(0, min(2 * len(reads), 200), dtype=np.uint32)
    off_a, ov_a, _ = engine.overlaps(q, max_divergence=1.0, **common)
    thr = float(np.median(ov_a["seq_divergence"]))           # removes about half of the records
    off_b, ov_b, _ = engine.overlaps(q, max_divergence=thr, **common)
    engine.overlaps(q, max_divergence=1.0, copy=False, **common)
    first = 37
    off_r, ov_r = engine.refilter(first, thr, len(q), copy=True)
    assert 0 < len(ov_b) < len(ov_a)
    for i in range(len(q)):
        want = ov_a[int(off_a[i]):int(off_a[i + 1])] if i < first else ov_b[int(off_b[i]):int(off_b[i + 1])]
        got = ov_r[int(off_r[i]):int(off_r[i + 1])]
        assert np.array_equal(got, want), i
    off_r2, ov_r2 = engine.refilter(0, thr, len(q), copy=True)     # a second refilter works on the refreshed result
    assert np.array_equal(off_r2, off_b) and np.array_equal(ov_r2, ov_b)
    # only tightening is possible: the records a batch (or an earlier refilter) dropped are gone
    with pytest.raises(fb.FlyeB200Error, match="looser"):
        engine.refilter(0, 2 * thr, len(q))
    engine.overlaps(q, max_divergence=1.0, max_overlaps=3, copy=False, **common)
    with pytest.raises(fb.FlyeB200Error, match="max_overlaps"):
        engine.refilter(0, thr, len(q))


def test_device_epilogue_equals_host_epilogue(engine, tmp_path, monkeypatch):
    """The default epilogue (divergenceKernel: seqDivergence with glibc's logf evaluated on the device, divergence test, per-query
    counts, compaction on the device) returns byte-identical records and offsets to the host epilogue (FG_DEVICE_EPILOGUE=0:
    recordDivergences with the host's logf + sliceEpilogue), for k-mer divergences (CLR) and edit-distance divergences (HiFi),
    with a threshold that removes records, with per-query thresholds, and with forced sub-batching."""
    import flye_b200 as fb
    tmp = str(tmp_path)
    n_cmp = 0
    for preset in ("raw_reads.cfg", "hifi.cfg"):
        cfg = pu.load_cfg(os.path.join(pu.CFG_DIR, preset))
        hifi = preset == "hifi.cfg"
        sim = dict(genome_len=120000, coverage=15, mean_len=12000, shape=20, error=0.01, seed=41) if hifi else dict(genome_len=120000, coverage=15, seed=42)
        reads = fb.read_fasta(pu.simulate(os.path.join(tmp, preset + ".fasta"), **sim), 1000)
        engine.upload_ascii(reads)
        if hifi:
            engine.build_index_minimizers(17, 1, int(cfg["minimizer_window"]), cfg["repeat_kmer_rate"])
        else:
            engine.count_kmers(15)
            engine.build_index_solid(2, cfg["meta_read_top_kmer_rate"], int(cfg["meta_read_filter_kmer_freq"]), cfg["repeat_kmer_rate"],
                                     float(int(cfg["assemble_kmer_sample"])))
        common = dict(max_jump=int(cfg["maximum_jump"]), min_overlap=1000, max_overhang=int(cfg["maximum_overhang"]), only_max_ext=True,
                      nucl_alignment=hifi, use_hpc=hifi)
        q = np.arange(0, 2 * len(reads), dtype=np.uint32)
        monkeypatch.setenv("FG_DEVICE_EPILOGUE", "0")
        _, ov_all, _ = engine.overlaps(q, max_divergence=1.0, **common)
        thr = float(np.median(ov_all["seq_divergence"]))          # removes about half of the records
        per_q = np.where(np.arange(len(q)) % 4 == 0, 1.0, thr).astype(np.float32)
        for kw in (dict(max_divergence=1.0), dict(max_divergence=thr), dict(max_divergence=0.5, query_max_divergence=per_q)):
            for budget in (None, "20000"):
                if budget:
                    monkeypatch.setenv("FG_HIT_BUDGET", budget)
                else:
                    monkeypatch.delenv("FG_HIT_BUDGET", raising=False)
                monkeypatch.setenv("FG_DEVICE_EPILOGUE", "0")
                off_h, ov_h, _ = engine.overlaps(q, **kw, **common)
                assert "epilogue_dev" not in engine.timings()
                monkeypatch.setenv("FG_DEVICE_EPILOGUE", "1")
                off_d, ov_d, _ = engine.overlaps(q, **kw, **common)
                assert "epilogue_dev" in engine.timings()        # the device path really ran (no silent fallback)
                assert np.array_equal(off_h, off_d), (preset, kw.keys(), budget)
                assert ov_h.tobytes() == ov_d.tobytes(), (preset, kw.keys(), budget)
                n_cmp += len(ov_d)
        assert 0 < len(ov_h) < len(ov_all)
    assert n_cmp > 10000
