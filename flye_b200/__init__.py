"""flye_b200 — B200-native read-overlap hot path of Flye (k-mer counting, VertexIndex, getSeqOverlaps).

This Python module is only the ctypes binding of the C ABI in include/flye_b200.h, used by tests/ and
bench.py.  The product is flye_b200/libflye_b200.so (hand-written sm_100a kernels + C ABI) and the C++
mirror of the reference classes in flye_b200/host/.  There is no CPU fallback: importing works without a
GPU (so the symbol table can be checked), creating an Engine does not.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libflye_b200.so")

FG_OK = 0
ERR_NAMES = {-1: "FG_ERR_CUDA", -2: "FG_ERR_ARG", -3: "FG_ERR_KMER_SIZE", -4: "FG_ERR_TOO_FREQ",
             -5: "FG_ERR_OVERFLOW", -6: "FG_ERR_NCCL", -7: "FG_ERR_INTERNAL"}

# every symbol include/flye_b200.h declares
SYMBOLS = ["fg_ctx_create", "fg_ctx_destroy", "fg_last_error", "fg_stream", "fg_kernel_launches", "fg_last_timings",
           "fg_reads_upload", "fg_reads_upload_ascii", "fg_queries_upload", "fg_count_kmers", "fg_kmer_hist", "fg_kmer_freq",
           "fg_build_index_solid", "fg_build_index_minimizers", "fg_index_clear", "fg_index_lookup",
           "fg_index_positions", "fg_index_export", "fg_overlaps_batch", "fg_overlaps_refilter", "fg_overlaps_closure", "fg_comm_unique_id", "fg_comm_init",
           "fg_comm_set_shard", "fg_debug_int_peak", "fg_debug_warp_sort", "fg_debug_edit_distance", "fg_debug_edit_distance_rc", "fg_align_cigar_batch"]


class IndexStats(C.Structure):
    _fields_ = [("n_keys", C.c_uint64), ("n_entries", C.c_uint64), ("n_repetitive", C.c_uint64),
                ("repetitive_frequency", C.c_uint64), ("mean_frequency", C.c_float), ("sample_rate", C.c_float)]


class OverlapParams(C.Structure):
    _fields_ = [("max_jump", C.c_int32), ("min_overlap", C.c_int32), ("max_overhang", C.c_int32),
                ("max_overlaps", C.c_int32), ("force_local", C.c_int32), ("keep_alignment", C.c_int32),
                ("only_max_ext", C.c_int32), ("nucl_alignment", C.c_int32), ("use_hpc", C.c_int32),
                ("max_divergence", C.c_float), ("query_set", C.c_int32), ("query_max_divergence", C.POINTER(C.c_float)),
                ("keep_rejected", C.c_int32), ("pad_", C.c_int32)]


OVERLAP_DTYPE = np.dtype([("cur_id", "<u4"), ("cur_begin", "<i4"), ("cur_end", "<i4"), ("cur_len", "<i4"),
                          ("ext_id", "<u4"), ("ext_begin", "<i4"), ("ext_end", "<i4"), ("ext_len", "<i4"),
                          ("score", "<i4"), ("seq_divergence", "<f4"), ("chain_length", "<i4"),
                          ("filtered_positions", "<i4"), ("edit_distance", "<i4"), ("aln_len", "<i4"),
                          ("aln_first", "<u8"), ("aln_count", "<u4"), ("reserved", "<u4")], align=True)
assert OVERLAP_DTYPE.itemsize == 72


class OverlapResult(C.Structure):
    _fields_ = [("n_queries", C.c_uint32), ("offsets", C.POINTER(C.c_uint64)), ("overlaps", C.c_void_p),
                ("aln_pairs", C.POINTER(C.c_int32)), ("n_aln_pairs", C.c_uint64), ("n_hits", C.c_uint64),
                ("n_pairs", C.c_uint64), ("n_dp_pairs", C.c_uint64), ("n_dp_cells", C.c_uint64)]


_lib = None


def load_lib():
    """dlopen the in-tree library; raises if it has not been built (no fallback of any kind)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError("flye_b200/libflye_b200.so is missing: run `python flye_b200/build.py` "
                           "(there is no CPU or PyTorch fallback for this path)")
    lib = C.CDLL(LIB_PATH)
    vp, u64p, u32p, i32p, u8p = C.c_void_p, C.POINTER(C.c_uint64), C.POINTER(C.c_uint32), C.POINTER(C.c_int32), C.POINTER(C.c_uint8)
    lib.fg_ctx_create.argtypes = [C.c_int, C.POINTER(vp)]
    lib.fg_ctx_destroy.argtypes = [vp]
    lib.fg_ctx_destroy.restype = None
    lib.fg_last_error.argtypes = [vp]
    lib.fg_last_error.restype = C.c_char_p
    lib.fg_stream.argtypes = [vp]
    lib.fg_stream.restype = C.c_void_p
    lib.fg_kernel_launches.argtypes = [vp]
    lib.fg_kernel_launches.restype = C.c_uint64
    lib.fg_last_timings.argtypes = [vp, C.POINTER(C.c_char_p), C.POINTER(C.c_float), C.POINTER(C.c_int), C.c_int]
    lib.fg_reads_upload.argtypes = [vp, u64p, u64p, u32p, C.c_uint32]
    lib.fg_reads_upload_ascii.argtypes = [vp, C.c_char_p, u64p, C.c_uint32]
    lib.fg_queries_upload.argtypes = [vp, u64p, u64p, u32p, C.c_uint32]
    lib.fg_count_kmers.argtypes = [vp, C.c_int, u64p]
    lib.fg_kmer_hist.argtypes = [vp, u64p, u64p, u64p]
    lib.fg_kmer_freq.argtypes = [vp, u64p, C.c_uint32, u32p]
    lib.fg_build_index_solid.argtypes = [vp, C.c_int, C.c_float, C.c_int, C.c_float, C.c_float, C.POINTER(IndexStats)]
    lib.fg_build_index_minimizers.argtypes = [vp, C.c_int, C.c_int, C.c_int, C.c_float, C.POINTER(IndexStats)]
    lib.fg_index_clear.argtypes = [vp]
    lib.fg_index_lookup.argtypes = [vp, u64p, C.c_uint32, u8p, u32p, u64p, u8p]
    lib.fg_index_positions.argtypes = [vp, C.c_uint64, C.c_uint32, C.c_int, u32p, i32p]
    lib.fg_index_export.argtypes = [vp, u64p, u8p, u64p, u32p, u64p, u32p, i32p, u64p]
    lib.fg_overlaps_batch.argtypes = [vp, u32p, C.c_uint32, C.POINTER(OverlapParams), C.POINTER(OverlapResult)]
    lib.fg_overlaps_refilter.argtypes = [vp, C.c_uint32, C.c_float, C.POINTER(OverlapResult)]
    lib.fg_overlaps_closure.argtypes = [vp, C.c_void_p, C.c_uint64, C.c_uint32, C.c_int32, C.POINTER(OverlapResult)]
    lib.fg_comm_unique_id.argtypes = [u8p]
    lib.fg_comm_init.argtypes = [vp, C.c_int, C.c_int, u8p]
    lib.fg_comm_set_shard.argtypes = [vp, C.c_uint32, C.c_uint32]
    lib.fg_debug_int_peak.argtypes = [vp, C.POINTER(C.c_double)]
    lib.fg_debug_warp_sort.argtypes = [vp, u64p, u32p, u64p, C.c_uint32]
    lib.fg_debug_edit_distance.argtypes = [vp, u8p, C.c_int, u8p, C.c_int, C.POINTER(C.c_int)]
    lib.fg_debug_edit_distance_rc.argtypes = [vp, u8p, C.c_int, C.c_int, u8p, C.c_int, C.c_int, C.POINTER(C.c_int)]
    _lib = lib
    return lib


class FlyeB200Error(RuntimeError):
    pass


def _ptr(a, ctype):
    return a.ctypes.data_as(C.POINTER(ctype))


_CODE = np.full(256, 255, dtype=np.uint8)
for _c, _v in zip("ACGTacgt", [0, 1, 2, 3, 0, 1, 2, 3]):
    _CODE[ord(_c)] = _v


def read_fasta(path, min_read_length=0):
    """FASTA -> list of ASCII byte strings; keeps reads strictly longer than min_read_length
    (reference: sequence_container.cpp:100-106)."""
    reads, cur = [], []
    opener = open
    if path.endswith(".gz"):
        import gzip
        opener = gzip.open
    with opener(path, "rb") as f:
        have = False
        for line in f:
            line = line.strip()
            if not line:
                continue
            if line.startswith(b">"):
                if have:
                    s = b"".join(cur)
                    if len(s) > min_read_length:
                        reads.append(s)
                cur, have = [], True
            else:
                cur.append(line)
        if have:
            s = b"".join(cur)
            if len(s) > min_read_length:
                reads.append(s)
    return reads


def pack_reads(reads):
    """list of ASCII reads -> (packed uint64 words, word offsets, lengths) in DnaSequence's layout
    (sequence.h:54-69: base j at bits 2*(j%32) of word j/32)."""
    lengths = np.array([len(r) for r in reads], dtype=np.uint32)
    nwords = (lengths.astype(np.uint64) + 31) // 32
    offs = np.zeros(len(reads) + 1, dtype=np.uint64)
    np.cumsum(nwords, out=offs[1:])
    packed = np.zeros(int(offs[-1]) + 1, dtype=np.uint64)
    shifts = (np.arange(32, dtype=np.uint64) * 2)
    for i, r in enumerate(reads):
        codes = _CODE[np.frombuffer(r, dtype=np.uint8)]
        if (codes > 3).any():
            raise ValueError("non-ACGT letter in read %d" % i)
        pad = (-len(codes)) % 32
        c = np.concatenate([codes, np.zeros(pad, np.uint8)]).astype(np.uint64).reshape(-1, 32)
        packed[int(offs[i]):int(offs[i + 1])] = (c << shifts).sum(axis=1, dtype=np.uint64)
    return packed, offs, lengths


class Engine:
    """One device context (fg_ctx)."""

    def __init__(self, device=0):
        self.lib = load_lib()
        self.ctx = C.c_void_p()
        rc = self.lib.fg_ctx_create(device, C.byref(self.ctx))
        if rc != FG_OK:
            raise FlyeB200Error("fg_ctx_create(%d) failed: %s — a CUDA device is required, there is no CPU path"
                                % (device, ERR_NAMES.get(rc, rc)))
        self.n_reads = 0
        self.lengths = None
        self.k = 0

    def close(self):
        if self.ctx:
            self.lib.fg_ctx_destroy(self.ctx)
            self.ctx = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc != FG_OK:
            raise FlyeB200Error("%s: %s" % (ERR_NAMES.get(rc, rc), self.lib.fg_last_error(self.ctx).decode()))

    # ---- reads ----
    def upload_packed(self, packed, word_offsets, lengths):
        packed = np.ascontiguousarray(packed, dtype=np.uint64)
        word_offsets = np.ascontiguousarray(word_offsets, dtype=np.uint64)
        lengths = np.ascontiguousarray(lengths, dtype=np.uint32)
        self._check(self.lib.fg_reads_upload(self.ctx, _ptr(packed, C.c_uint64), _ptr(word_offsets, C.c_uint64),
                                             _ptr(lengths, C.c_uint32), len(lengths)))
        self.n_reads, self.lengths = len(lengths), lengths.copy()

    def upload_queries(self, reads):
        """second sequence set (list of ASCII reads) for query_set=1 calls"""
        packed, offs, lens = pack_reads(reads)
        self._check(self.lib.fg_queries_upload(self.ctx, _ptr(packed, C.c_uint64), _ptr(offs, C.c_uint64), _ptr(lens, C.c_uint32), len(lens)))

    def upload_ascii(self, reads):
        buf = b"".join(reads)
        offs = np.zeros(len(reads) + 1, dtype=np.uint64)
        np.cumsum([len(r) for r in reads], out=offs[1:])
        self._check(self.lib.fg_reads_upload_ascii(self.ctx, buf, _ptr(offs, C.c_uint64), len(reads)))
        self.n_reads = len(reads)
        self.lengths = np.array([len(r) for r in reads], dtype=np.uint32)

    # ---- counting ----
    def count_kmers(self, k):
        n = C.c_uint64()
        self._check(self.lib.fg_count_kmers(self.ctx, k, C.byref(n)))
        self.k = k
        return n.value

    def kmer_hist(self):
        n = C.c_uint64()
        self._check(self.lib.fg_kmer_hist(self.ctx, None, None, C.byref(n)))
        f = np.zeros(n.value, dtype=np.uint64)
        c = np.zeros(n.value, dtype=np.uint64)
        if n.value:
            self._check(self.lib.fg_kmer_hist(self.ctx, _ptr(f, C.c_uint64), _ptr(c, C.c_uint64), C.byref(n)))
        return dict(zip(f.tolist(), c.tolist()))

    # ---- index ----
    def build_index_solid(self, min_freq=2, select_rate=0.40, tandem_freq=100, repeat_rate=100.0, sample_rate=1.0):
        st = IndexStats()
        self._check(self.lib.fg_build_index_solid(self.ctx, min_freq, select_rate, tandem_freq, repeat_rate, sample_rate, C.byref(st)))
        return st

    def build_index_minimizers(self, k, min_cov=1, window=10, repeat_rate=100.0):
        st = IndexStats()
        self._check(self.lib.fg_build_index_minimizers(self.ctx, k, min_cov, window, repeat_rate, C.byref(st)))
        self.k = k
        return st

    def export_index(self):
        nk, ne = C.c_uint64(), C.c_uint64()
        self._check(self.lib.fg_index_export(self.ctx, None, None, None, None, C.byref(nk), None, None, C.byref(ne)))
        keys = np.zeros(nk.value, np.uint64); rep = np.zeros(nk.value, np.uint8)
        first = np.zeros(nk.value, np.uint64); size = np.zeros(nk.value, np.uint32)
        ids = np.zeros(ne.value, np.uint32); pos = np.zeros(ne.value, np.int32)
        self._check(self.lib.fg_index_export(self.ctx, _ptr(keys, C.c_uint64), _ptr(rep, C.c_uint8), _ptr(first, C.c_uint64),
                                             _ptr(size, C.c_uint32), C.byref(nk), _ptr(ids, C.c_uint32), _ptr(pos, C.c_int32),
                                             C.byref(ne)))
        return keys, rep, first, size, ids, pos

    # ---- overlaps ----
    def overlaps(self, query_ids, max_jump=1500, min_overlap=1000, max_overhang=1500, max_overlaps=0, force_local=False,
                 keep_alignment=False, only_max_ext=True, nucl_alignment=False, use_hpc=False, max_divergence=1.0, copy=True,
                 query_set=0, query_max_divergence=None, keep_rejected=False):
        """copy=False returns views into library-owned memory, valid until the next overlaps() call.
        query_max_divergence: optional per-query thresholds (float32, one per query id) overriding max_divergence."""
        q = np.ascontiguousarray(query_ids, dtype=np.uint32)
        qthr = None
        if query_max_divergence is not None:
            qthr = np.ascontiguousarray(query_max_divergence, dtype=np.float32)
            if len(qthr) != len(q):
                raise ValueError("query_max_divergence needs one threshold per query")
        p = OverlapParams(max_jump, min_overlap, max_overhang, max_overlaps, int(force_local), int(keep_alignment),
                          int(only_max_ext), int(nucl_alignment), int(use_hpc), max_divergence, int(query_set),
                          _ptr(qthr, C.c_float) if qthr is not None else None, int(keep_rejected), 0)
        res = OverlapResult()
        self._check(self.lib.fg_overlaps_batch(self.ctx, _ptr(q, C.c_uint32), len(q), C.byref(p), C.byref(res)))
        offsets = np.ctypeslib.as_array(res.offsets, shape=(len(q) + 1,))
        if copy:
            offsets = offsets.copy()
        n = int(offsets[-1])
        if n:
            buf = (C.c_char * (n * OVERLAP_DTYPE.itemsize)).from_address(res.overlaps)
            ov = np.frombuffer(buf, dtype=OVERLAP_DTYPE, count=n)
            if copy:
                ov = ov.copy()
        else:
            ov = np.zeros(0, dtype=OVERLAP_DTYPE)
        stats = dict(n_hits=res.n_hits, n_pairs=res.n_pairs, n_dp_pairs=res.n_dp_pairs, n_dp_cells=res.n_dp_cells)
        if keep_alignment:
            npts = int(res.n_aln_pairs)
            stats["aln_pairs"] = (np.ctypeslib.as_array(res.aln_pairs, shape=(npts, 2)).copy() if npts else np.zeros((0, 2), np.int32))
        return offsets, ov, stats

    def refilter(self, first_query, max_divergence, n_queries, copy=False):
        """fg_overlaps_refilter: drop the overlaps with seq_divergence >= max_divergence of the queries from first_query on
        (last overlaps() result); returns (offsets, overlaps) like overlaps()."""
        res = OverlapResult()
        self._check(self.lib.fg_overlaps_refilter(self.ctx, C.c_uint32(first_query), C.c_float(max_divergence), C.byref(res)))
        offsets = np.ctypeslib.as_array(res.offsets, shape=(n_queries + 1,))
        n = int(offsets[-1])
        if n:
            buf = (C.c_char * (n * OVERLAP_DTYPE.itemsize)).from_address(res.overlaps)
            ov = np.frombuffer(buf, dtype=OVERLAP_DTYPE, count=n)
        else:
            ov = np.zeros(0, dtype=OVERLAP_DTYPE)
        return (offsets.copy(), ov.copy()) if copy else (offsets, ov)

    def closure(self, records, n_seqs, max_ends_diff):
        """fg_overlaps_closure: symmetric closure + cluster filter of forward-sequence overlap records -> (offsets, records)"""
        recs = np.ascontiguousarray(records, dtype=OVERLAP_DTYPE)
        res = OverlapResult()
        self._check(self.lib.fg_overlaps_closure(self.ctx, recs.ctypes.data_as(C.c_void_p), len(recs), n_seqs, max_ends_diff, C.byref(res)))
        offsets = np.ctypeslib.as_array(res.offsets, shape=(n_seqs + 1,)).copy()
        n = int(offsets[-1])
        if n:
            buf = (C.c_char * (n * OVERLAP_DTYPE.itemsize)).from_address(res.overlaps)
            ov = np.frombuffer(buf, dtype=OVERLAP_DTYPE, count=n).copy()
        else:
            ov = np.zeros(0, dtype=OVERLAP_DTYPE)
        return offsets, ov

    # ---- multi-GPU ----
    @staticmethod
    def comm_unique_id():
        lib = load_lib()
        buf = (C.c_uint8 * 128)()
        rc = lib.fg_comm_unique_id(buf)
        if rc != FG_OK:
            raise FlyeB200Error("fg_comm_unique_id: %s (is libnccl available?)" % ERR_NAMES.get(rc, rc))
        return bytes(buf)

    def comm_init(self, n_ranks, rank, unique_id):
        buf = (C.c_uint8 * 128).from_buffer_copy(unique_id)
        self._check(self.lib.fg_comm_init(self.ctx, n_ranks, rank, buf))

    def set_shard(self, first_read, n_reads):
        self._check(self.lib.fg_comm_set_shard(self.ctx, first_read, n_reads))

    def debug_warp_sort(self, keys, vals, seg_offsets):
        keys = np.ascontiguousarray(keys, dtype=np.uint64).copy()
        vals = np.ascontiguousarray(vals, dtype=np.uint32).copy()
        seg = np.ascontiguousarray(seg_offsets, dtype=np.uint64)
        self._check(self.lib.fg_debug_warp_sort(self.ctx, _ptr(keys, C.c_uint64), _ptr(vals, C.c_uint32), _ptr(seg, C.c_uint64), len(seg) - 1))
        return keys, vals

    def debug_edit_distance(self, a, b):
        a = np.ascontiguousarray(a, dtype=np.uint8)
        b = np.ascontiguousarray(b, dtype=np.uint8)
        d = C.c_int()
        self._check(self.lib.fg_debug_edit_distance(self.ctx, _ptr(a, C.c_uint8), len(a), _ptr(b, C.c_uint8), len(b), C.byref(d)))
        return d.value

    def debug_edit_distance_rc(self, a, rc_a, b, rc_b):
        a = np.ascontiguousarray(a, dtype=np.uint8)
        b = np.ascontiguousarray(b, dtype=np.uint8)
        d = C.c_int()
        self._check(self.lib.fg_debug_edit_distance_rc(self.ctx, _ptr(a, C.c_uint8), len(a), int(rc_a), _ptr(b, C.c_uint8), len(b), int(rc_b),
                                                       C.byref(d)))
        return d.value

    def int_peak(self):
        """measured integer-issue ceiling of the device in Gop/s (IMAD / LOP3 / SHF mix)"""
        g = C.c_double()
        self._check(self.lib.fg_debug_int_peak(self.ctx, C.byref(g)))
        return g.value

    def timings(self):
        names = (C.c_char_p * 64)()
        ms = (C.c_float * 64)()
        calls = (C.c_int * 64)()
        n = self.lib.fg_last_timings(self.ctx, names, ms, calls, 64)
        self.last_calls = {names[i].decode(): calls[i] for i in range(n)}
        return {names[i].decode(): ms[i] for i in range(n)}

    def stream_ptr(self):
        return int(self.lib.fg_stream(self.ctx) or 0)

    def launches(self):
        return int(self.lib.fg_kernel_launches(self.ctx))


# ---- text dumps in the oracle's formats (oracle/harness.cpp) --------------------------------------------
def dump_hist(hist, path):
    with open(path, "w") as f:
        for k in sorted(hist):
            f.write("%d %d\n" % (k, hist[k]))


def dump_index(engine, stats, path):
    keys, rep, first, size, ids, pos = engine.export_index()
    with open(path, "w") as f:
        f.write("sampleRate %08x\n" % np.float32(stats.sample_rate).view(np.uint32))
        for i in range(len(keys)):
            if not rep[i] and not size[i]:
                continue
            a, n = int(first[i]), int(size[i])
            f.write("%x %d %d" % (int(keys[i]), int(rep[i]), n))
            if n:
                f.write(" " + " ".join("%d:%d" % (ids[a + j], pos[a + j]) for j in range(n)))
            f.write("\n")


def dump_overlaps(query_ids, offsets, ov, path, aln_pairs=None):
    bits = ov["seq_divergence"].view(np.uint32)
    with open(path, "w") as f:
        for i, q in enumerate(query_ids):
            a, b = int(offsets[i]), int(offsets[i + 1])
            f.write("# %d %d\n" % (q, b - a))
            for j in range(a, b):
                o = ov[j]
                f.write("%d %d %d %d %d %d %d %d %d %08x\n" % (o["cur_id"], o["cur_begin"], o["cur_end"], o["cur_len"], o["ext_id"],
                                                              o["ext_begin"], o["ext_end"], o["ext_len"], o["score"], bits[j]))
                if aln_pairs is not None:
                    a0, n = int(o["aln_first"]), int(o["aln_count"])
                    f.write("  aln %d %s\n" % (n, " ".join("%d,%d" % (c, e) for c, e in aln_pairs[a0:a0 + n])))
