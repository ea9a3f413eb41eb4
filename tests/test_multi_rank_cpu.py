"""CPU, world_size 2 over gloo: the host logic of the multi-GPU layout — shard computation, unique-id exchange through
the host program's process group, and the merge of per-rank results into id order."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from flye_b200 import parallel   # noqa: E402


def test_shard_reads_tiles_and_balances():
    rng = np.random.default_rng(0)
    lens = rng.integers(1000, 40000, 5000)
    for world in (1, 2, 3, 4, 8):
        shards = parallel.shard_reads(lens, world)
        assert shards[0][0] == 0 and sum(c for _, c in shards) == len(lens)
        for (f0, c0), (f1, _) in zip(shards, shards[1:]):
            assert f0 + c0 == f1                                   # contiguous, ascending in rank order
        bases = [int(lens[f:f + c].sum()) for f, c in shards]
        assert max(bases) - min(bases) <= 2 * lens.max()           # balanced by bases
    assert parallel.shard_reads([5], 4) == [(0, 0), (0, 0), (0, 0), (0, 1)] or sum(c for _, c in parallel.shard_reads([5], 4)) == 1


class _FakeEngine:
    @staticmethod
    def comm_unique_id():
        return bytes(range(128))


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        lens = np.random.default_rng(1).integers(1000, 20000, 301)
        first, count = parallel.shard_reads(lens, world)[rank]
        uid = parallel.broadcast_unique_id(_FakeEngine, dist)
        assert uid == bytes(range(128))
        # fake per-rank results: query q (forward id 2*read) has (read % 3) overlaps tagged with the read index
        qids = np.arange(2 * first, 2 * (first + count), 2, dtype=np.uint32)
        counts = (qids // 2) % 3
        offs = np.concatenate([[0], np.cumsum(counts)]).astype(np.uint64)
        ov = np.repeat(qids // 2, counts).astype(np.int64)
        gathered = [None] * world
        dist.all_gather_object(gathered, (qids, offs, ov))
        ids, moffs, mov = parallel.merge_rank_results(gathered)
        assert np.array_equal(ids, np.arange(0, 2 * len(lens), 2))          # id order, every read exactly once
        for i, q in enumerate(ids):
            seg = mov[int(moffs[i]):int(moffs[i + 1])]
            assert len(seg) == (q // 2) % 3 and (seg == q // 2).all()
        if rank == 0:
            out.put("ok")
    finally:
        dist.destroy_process_group()


def test_two_ranks_gloo():
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert out.get(timeout=5) == "ok"
