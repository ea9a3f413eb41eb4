"""-m gpu: fg_overlaps_closure (ensureTransitivity(false) + filterOverlaps() on the device, overlap.cpp:576-627, 681-741)
against a plain Python restatement of the same two functions on random records; lists are compared as multisets (the
reference leaves the order among equal curBegin open)."""
import numpy as np
import pytest

import flye_b200 as fb

pytestmark = pytest.mark.gpu
K = 15


def _variant(r, v):
    cur = [r["cur_id"], r["cur_begin"], r["cur_end"], r["cur_len"]]
    ext = [r["ext_id"], r["ext_begin"], r["ext_end"], r["ext_len"]]
    if v & 1:   # complement (overlap.h:118-147)
        cur = [cur[0] ^ 1, cur[3] - cur[2] - 1, cur[3] - cur[1] - 1, cur[3]]
        ext = [ext[0] ^ 1, ext[3] - ext[2] - 1, ext[3] - ext[1] - 1, ext[3]]
    if v & 2:   # reverse (overlap.h:95-116)
        cur, ext = ext, cur
    return tuple(int(x) for x in cur + ext + [r["score"]])


def _closure_model(recs, n_seqs):
    lists = [[] for _ in range(n_seqs)]
    for r in recs:
        for v in range(4):
            t = _variant(r, v)
            lists[t[0]].append(t)
    out = []
    for s, ovs in enumerate(lists):
        parent = list(range(len(ovs)))

        def find(x):
            while parent[x] != x:
                parent[x] = parent[parent[x]]
                x = parent[x]
            return x
        for i, a in enumerate(ovs):      # overlap.cpp:700-717
            for j, b in enumerate(ovs):
                if a[4] != b[4]:
                    continue
                cur_diff = (a[2] - a[1]) - (min(a[2], b[2]) - max(a[1], b[1]))
                ext_diff = (a[6] - a[5]) - (min(a[6], b[6]) - max(a[5], b[5]))
                if cur_diff < K and ext_diff < K:
                    parent[find(i)] = find(j)
        best = {}
        for i, a in enumerate(ovs):      # best score of every cluster
            c = find(i)
            if c not in best or a[8] > best[c][8]:
                best[c] = a
        out.append(sorted(best.values()))
    return out


def test_closure_and_cluster_filter_match_the_restatement(engine):
    rng = np.random.default_rng(7)
    n_reads = 30
    lens = rng.integers(5000, 20000, n_reads)
    recs = []
    score = 1000
    for _ in range(400):
        a, b = rng.integers(0, n_reads, 2)
        if a == b:
            continue
        cb = int(rng.integers(0, lens[a] - 3000)); ce = cb + int(rng.integers(1500, 2900))
        eb = int(rng.integers(0, lens[b] - 3000)); ee = eb + int(rng.integers(1500, 2900))
        ext_id = 2 * int(b) + int(rng.integers(0, 2))
        for _dup in range(int(rng.integers(1, 4))):   # near-identical copies (ends within k), different scores
            j = rng.integers(-6, 7, 4)
            score += 1
            recs.append((2 * int(a), cb + j[0], ce + j[1], int(lens[a]), ext_id, eb + j[2], ee + j[3], int(lens[b]), score))
    arr = np.zeros(len(recs), dtype=fb.OVERLAP_DTYPE)
    for i, r in enumerate(recs):
        (arr[i]["cur_id"], arr[i]["cur_begin"], arr[i]["cur_end"], arr[i]["cur_len"], arr[i]["ext_id"], arr[i]["ext_begin"], arr[i]["ext_end"],
         arr[i]["ext_len"], arr[i]["score"]) = r
    offs, ov = engine.closure(arr, 2 * n_reads, K)
    want = _closure_model(arr, 2 * n_reads)
    assert int(offs[-1]) == sum(len(w) for w in want)
    for s in range(2 * n_reads):
        got = ov[int(offs[s]):int(offs[s + 1])]
        tup = sorted(tuple(int(g[f]) for f in ("cur_id", "cur_begin", "cur_end", "cur_len", "ext_id", "ext_begin", "ext_end", "ext_len", "score")) for g in got)
        assert tup == want[s], s
        assert list(got["cur_begin"]) == sorted(got["cur_begin"])          # sorted by curBegin (overlap.cpp:736-738)
        for g in got:                                                         # every record names its input record and variant
            assert _variant(arr[int(g["reserved"]) >> 2], int(g["reserved"]) & 3)[:8] == tuple(int(g[f]) for f in (
                "cur_id", "cur_begin", "cur_end", "cur_len", "ext_id", "ext_begin", "ext_end", "ext_len"))
