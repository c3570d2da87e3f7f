"""world_size-2 gloo test of the N>1 host logic: the batch is cut into contiguous shards, each
rank generates and solves only its own shard (here with the CPU oracle standing in for the
kernel — there is no GPU in the build container), and the results are gathered on rank 0 with
no collective in the data path other than that final gather.  The sharded result must be
bit-identical to the unsharded one."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, batch, n, p, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import psd_b200
    import psd_rng
    from oracle import oracle as O
    lo, hi = psd_b200.shard_bounds(batch, world, rank)
    A = psd_rng.gen_uniform(1234, n, p, hi - lo, first_b=lo)
    _, _, lam, info, _ = O.rpschur_batched(A, wantT=False, wantZ=False, nthreads=1)
    mine = torch.from_numpy(np.concatenate([lam.view(np.float64).reshape(hi - lo, -1),
                                            info.reshape(-1, 1).astype(np.float64)], axis=1))
    sizes = [psd_b200.shard_bounds(batch, world, r) for r in range(world)]
    bufs = [torch.empty((b - a, mine.shape[1]), dtype=torch.float64) for a, b in sizes]
    dist.all_gather(bufs, mine) if len({b - a for a, b in sizes}) == 1 else dist.all_gather_object(
        objs := [None] * world, mine)
    if len({b - a for a, b in sizes}) != 1:
        bufs = objs
    if rank == 0:
        q.put(torch.cat(bufs).numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharding_matches_single_process(oracle):
    import psd_rng
    batch, n, p = 13, 8, 3   # odd batch: ragged shards
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, batch, n, p, q)) for r in range(2)]
    for pr in procs:
        pr.start()
    got = q.get(timeout=120)
    for pr in procs:
        pr.join(timeout=60)
        assert pr.exitcode == 0
    A = psd_rng.gen_uniform(1234, n, p, batch)
    _, _, lam, info, _ = oracle.rpschur_batched(A, wantT=False, wantZ=False, nthreads=1)
    ref = np.concatenate([lam.view(np.float64).reshape(batch, -1), info.reshape(-1, 1).astype(np.float64)], axis=1)
    assert got.shape == ref.shape
    assert got.tobytes() == ref.tobytes()


def test_shard_bounds_cover_batch():
    sys.path.insert(0, ROOT)
    import psd_b200
    for batch in (0, 1, 7, 100000):
        for world in (1, 2, 3, 8):
            edges = [psd_b200.shard_bounds(batch, world, r) for r in range(world)]
            assert edges[0][0] == 0 and edges[-1][1] == batch
            for (a, b), (c, d) in zip(edges, edges[1:]):
                assert b == c and a <= b
