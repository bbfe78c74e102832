"""N>1 path on CPU: world_size-2 (and 3) gloo process groups exercise the per-rank top-k gather + merge that bench.py and
multi-process deployments use; the per-rank shard results are produced by the oracle (no GPU in this test)."""
import os
import socket

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

from cudasw4_b200 import dbformat, synth
from cudasw4_b200.distributed import gather_topk, merge_topk, shard_of


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, k, scores, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        ids = np.array([i for i in range(len(scores)) if shard_of(i, world) == rank], dtype=np.int64)
        local = scores[ids]
        order = np.lexsort((ids, -local))[:k]  # what the engine returns for this shard: score desc, id asc
        merged = gather_topk(local[order].tolist(), ids[order].tolist(), k)
        q.put((rank, merged))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_topk_merge_matches_global(oracle, world):
    rng = np.random.default_rng(4)
    seqs = [synth.random_residues(rng, int(n)) for n in rng.integers(5, 120, 1500)]
    seqs += [seqs[3].copy() for _ in range(30)]  # ties across shards
    db = dbformat.from_sequences(seqs)
    query = synth.random_residues(rng, 90)
    scores = oracle.scan(62, query, db, -11, -1)
    k = 20
    s, i = oracle.topk(scores, k)
    expected = list(zip(s.tolist(), i.tolist()))
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, k, scores, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, merged in results:
        assert merged == expected, rank


def test_merge_and_shard_rule():
    assert merge_topk([([9, 5], [7, 1]), ([9, 9], [2, 300])], 3) == [(9, 2), (9, 7), (9, 300)]
    assert merge_topk([([], [])], 5) == []
    assert [shard_of(i, 2) for i in (0, 255, 256, 511, 512)] == [0, 0, 1, 1, 0]
    world = 5
    counts = np.bincount([shard_of(i, world) for i in range(10_000)], minlength=world)
    assert counts.max() - counts.min() <= 256
