"""Multi-GPU host logic: one process per GPU (torch.distributed for the plumbing), reads partitioned into contiguous,
base-balanced shards in rank order, index replicated over NCCL inside the library (csrc/comm.cu)."""
import numpy as np


def shard_reads(lengths, world):
    """Contiguous shards [first, first+count) per rank, balanced by bases.  Shards are ascending in rank order — the
    library relies on that when it concatenates the ranks' index entries (global-position order)."""
    lengths = np.asarray(lengths, dtype=np.int64)
    n = len(lengths)
    csum = np.concatenate([[0], np.cumsum(lengths)])
    total = int(csum[-1])
    bounds = [0]
    for r in range(1, world):
        target = total * r // world
        b = int(np.searchsorted(csum, target, side="left"))
        b = min(max(b, bounds[-1]), n)
        bounds.append(b)
    bounds.append(n)
    return [(bounds[r], bounds[r + 1] - bounds[r]) for r in range(world)]


def broadcast_unique_id(engine_cls, dist, device=None):
    """rank 0 draws the ncclUniqueId from the library, every rank receives the 128 bytes through the host program's own
    process group (any backend)."""
    import torch
    buf = torch.zeros(128, dtype=torch.uint8, device=device) if device is not None else torch.zeros(128, dtype=torch.uint8)
    if dist.get_rank() == 0:
        raw = engine_cls.comm_unique_id()
        buf.copy_(torch.frombuffer(bytearray(raw), dtype=torch.uint8))
    dist.broadcast(buf, src=0)
    return bytes(buf.cpu().numpy().tobytes())


def merge_rank_results(per_rank):
    """per_rank: list (in rank order) of (query_ids, offsets, overlaps) -> one result in query-id order.  Shards are
    ascending, so concatenation in rank order is id order."""
    ids = np.concatenate([p[0] for p in per_rank])
    counts = np.concatenate([np.diff(p[1]) for p in per_rank])
    offs = np.concatenate([[0], np.cumsum(counts)]).astype(np.uint64)
    ov = np.concatenate([p[2] for p in per_rank])
    return ids, offs, ov
