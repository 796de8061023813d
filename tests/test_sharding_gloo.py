"""N>1 host logic on CPU: world_size-2 gloo run of the sharding / ordered merge / max-over-ranks
helpers bench.py and the multi-GPU driver use.  The per-shard "mapping" is done by the C oracle
(test stand-in for the device)."""
import os
import socket
import sys

import numpy as np
import torch.multiprocessing as mp

import goldenio
import refio


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _worker(rank, world, port, q):
    import torch.distributed as dist
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from walt_b200 import sharding
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    hdr, _ = goldenio.genome()
    z = goldenio.load("se_ct.npz")
    reads = z["reads"][:999]
    lo, hi = sharding.shard_range(len(reads), world, rank)
    local = refio.oracle_se_map(hdr, goldenio.se_pair(False), reads[lo:hi])
    merged = sharding.gather_in_order(local, len(reads), dist, refio.BEST_DT)
    tmax = sharding.max_over_ranks([1.0 + rank, 5.0 - rank], dist)
    dist.barrier()
    if rank == 0:
        want = z["best_m6_b5000"][:999]
        ok = all(np.array_equal(merged[f], want[f]) for f in ("genome_pos", "times", "mismatch", "strand"))
        q.put((ok, tmax))
    dist.destroy_process_group()


def test_two_rank_sharding_and_merge():
    from walt_b200 import sharding
    for n in (0, 1, 7, 1000, 10_000_001):
        for g in (1, 2, 3, 8):
            cuts = [sharding.shard_range(n, g, r) for r in range(g)]
            assert cuts[0][0] == 0 and cuts[-1][1] == n
            assert all(cuts[i][1] == cuts[i + 1][0] for i in range(g - 1))
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    ok, tmax = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert ok
    assert tmax == [2.0, 5.0]
