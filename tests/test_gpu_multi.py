"""-m gpu, needs >= 2 GPUs (skipped otherwise): two ranks, reads partitioned, counts / index merged over NCCL — the
replicated index and the concatenated overlaps must be byte-identical to the single-GPU dumps (and so to the oracle)."""
import json
import os
import subprocess
import sys

import pytest

import parity_util as pu

pytestmark = pytest.mark.gpu

WORKER = r'''
import os, sys, json
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.environ["FG_ROOT"]); sys.path.insert(0, os.path.join(os.environ["FG_ROOT"], "tests"))
import flye_b200 as fb, parity_util as pu
from flye_b200 import parallel
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank)
dist.init_process_group("nccl", device_id=torch.device("cuda", rank))
reads_path, cfg_path, out, k = sys.argv[1], sys.argv[2], sys.argv[3], int(sys.argv[4])
cfg = pu.load_cfg(cfg_path)
reads = fb.read_fasta(reads_path, 1000)
eng = fb.Engine(rank)
eng.upload_ascii(reads)
eng.comm_init(world, rank, parallel.broadcast_unique_id(fb.Engine, dist, device=torch.device("cuda", rank)))
first, count = parallel.shard_reads(eng.lengths, world)[rank]
eng.set_shard(first, count)
if int(cfg["use_minimizers"]):
    st = eng.build_index_minimizers(k, 1, int(cfg["minimizer_window"]), cfg["repeat_kmer_rate"])
else:
    eng.count_kmers(k)
    if rank == 0: fb.dump_hist(eng.kmer_hist(), out + ".hist")
    st = eng.build_index_solid(2, cfg["meta_read_top_kmer_rate"], int(cfg["meta_read_filter_kmer_freq"]), cfg["repeat_kmer_rate"],
                               float(int(cfg["assemble_kmer_sample"])))
fb.dump_index(eng, st, out + ".index.rank%d" % rank)
q = np.arange(2 * first, 2 * (first + count), 2, dtype=np.uint32)
offs, ov, _ = eng.overlaps(q, max_jump=int(cfg["maximum_jump"]), min_overlap=1000, max_overhang=int(cfg["maximum_overhang"]),
                           nucl_alignment=bool(cfg["reads_base_alignment"]), use_hpc=bool(cfg["hpc_scoring_on"]), max_divergence=1.0)
gathered = [None] * world
dist.all_gather_object(gathered, (q, offs, ov))
if rank == 0:
    ids, moffs, mov = parallel.merge_rank_results(gathered)
    fb.dump_overlaps(ids.tolist(), moffs, mov, out + ".ovlp")
dist.destroy_process_group()
'''


def _index_diff(a, b):
    """key-level summary of two index dumps (diagnostics of a failure)"""
    da = {l.split(" ", 1)[0]: l for l in open(a).read().splitlines()[1:]}
    db = {l.split(" ", 1)[0]: l for l in open(b).read().splitlines()[1:]}
    only_a, only_b = sorted(set(da) - set(db)), sorted(set(db) - set(da))
    changed = [k for k in da if k in db and da[k] != db[k]]
    print("INDEXDIFF", a, b, "first entries differ?", open(a).readline(), open(b).readline())
    out = {"keys_a": len(da), "keys_b": len(db), "only_a": len(only_a), "only_b": len(only_b), "changed": len(changed),
            "ex_only_a": [da[k] for k in only_a[:3]], "ex_only_b": [db[k] for k in only_b[:3]], "ex_changed": [(da[k], db[k]) for k in changed[:3]]}
    print("INDEXDIFF", json.dumps(out))
    return out["keys_a"], out["keys_b"], out["changed"]


def _n_gpus():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


@pytest.mark.skipif(_n_gpus() < 2, reason="needs two GPUs")
@pytest.mark.parametrize("cfg,k,sim", [
    ("raw_reads.cfg", 15, dict(genome_len=150000, coverage=15, seed=61)),
    ("hifi.cfg", 17, dict(genome_len=100000, coverage=12, mean_len=8000, shape=20, error=0.005, seed=62)),
    # solid k-mers with k = 17 (configs[2] / [3]): 2^33 classes, exactly 2^32 counter slots per rank, 64-bit keys in the index
    ("raw_reads.cfg", 17, dict(genome_len=120000, coverage=14, error=0.10, seed=63)),
])
def test_two_gpu_parity(built, engine, tmp_path, cfg, k, sim):
    tmp = str(tmp_path)
    reads = pu.simulate(os.path.join(tmp, "r.fasta"), **sim)
    cfg_path = os.path.join(pu.CFG_DIR, cfg)
    # single-GPU reference dumps (themselves checked against the oracle elsewhere)
    pu.gpu_pipeline(reads, cfg_path, os.path.join(tmp, "one"), k=k, dump_index=True, estimate=False, engine=engine)
    worker = os.path.join(tmp, "worker.py")
    open(worker, "w").write(WORKER)
    env = dict(os.environ, FG_ROOT=pu.ROOT)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                        "--master-port", "29611", worker, reads, cfg_path, os.path.join(tmp, "two"), str(k)],
                       env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:]
    for a, b in ([("one.hist", "two.hist")] if cfg == "raw_reads.cfg" else []) + \
                [("one.index", "two.index.rank0"), ("one.index", "two.index.rank1"), ("one.ovlp", "two.ovlp")]:
        n, sample = pu.diff_files(os.path.join(tmp, a), os.path.join(tmp, b))
        assert n == 0, (a, b, sample[:3], _index_diff(os.path.join(tmp, a), os.path.join(tmp, b)) if "index" in a else None)
