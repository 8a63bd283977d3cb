"""Multi-GPU path without GPUs: ZMW-range sharding over world_size-2 `gloo` ranks.

ZMWs are independent ("ccs scales linear in the number of ... ZMWs", docs/faq/performance.md:85-86;
`--chunk i/N` + merge, docs/faq/parallelize.md:7-28), so there is no data-path collective: every rank
processes its own index range and the merged output must equal the single-process output byte for byte.
The per-rank work here is the CPU oracle (no GPU in this container); the sharding / max-over-ranks timing /
gather logic is the same code path bench.py uses under torchrun.
"""
import os
import socket
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def shard_range(n_total, rank, world):
    """contiguous ZMW index range of `rank` (shard g = [g*Z/G, (g+1)*Z/G))"""
    return (n_total * rank) // world, (n_total * (rank + 1)) // world


def _consensus_of(indices):
    import oracle_lib as O
    from ccs_b200 import sim
    model = O.synthetic_model()
    cfg = sim.get_config(2, insert_mean=400, insert_sd=20)
    out = []
    for i in indices:
        z = sim.simulate_zmw(model, cfg, i)
        r = O.ccs_zmw(model, z.snr, [z.read(k) for k in range(z.n_reads)], z.cx)
        out.append((i, r["status"], bytes(r["seq"]), bytes(r["qv"])))
    return out


def _worker(rank, world, port, n_total, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = shard_range(n_total, rank, world)
    mine = _consensus_of(range(lo, hi))
    # max-over-ranks timing + ZMW count reduction, as bench.py does
    t = torch.tensor([float(rank + 1)], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    n = torch.tensor([float(len(mine))], dtype=torch.float64)
    dist.all_reduce(n, op=dist.ReduceOp.SUM)
    gathered = [None] * world
    dist.all_gather_object(gathered, mine)
    if rank == 0:
        q.put((float(t.item()), float(n.item()), [x for part in gathered for x in part]))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharding_equals_single_process():
    n_total, world = 6, 2
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_total, q)) for r in range(world)]
    for p in procs:
        p.start()
    tmax, nsum, merged = q.get(timeout=300)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert tmax == 2.0 and nsum == n_total
    single = _consensus_of(range(n_total))
    assert merged == single                       # byte-identical, in ZMW order (ordered merge)
    assert shard_range(10, 0, 3) == (0, 3) and shard_range(10, 2, 3) == (6, 10)


def test_bench_step_sharding_tiles_the_index_space():
    """bench.py's own range arithmetic (product-side sharding of the synthetic ZMW index space under torchrun): the
    (step, rank) ranges are disjoint and tile [0, steps * world * zmws)."""
    import bench
    for world, zmws, steps in ((1, 600, 5), (2, 1000, 4), (8, 600, 3)):
        seen = []
        for step in range(steps):
            for rank in range(world):
                f = bench.shard_first_index(step, rank, world, zmws)
                seen.append((f, f + zmws))
        seen.sort()
        assert seen[0][0] == 0 and seen[-1][1] == steps * world * zmws
        assert all(seen[k][1] == seen[k + 1][0] for k in range(len(seen) - 1))
