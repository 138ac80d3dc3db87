"""world_size-2 (and 3) gloo runs of the multi-GPU driver logic on CPU: the all-gather layout,
shard ownership and the rank-0 merge.  The per-shard evaluation is injected (the CPU oracle
restricted to the shard's rows) -- on GPUs the same driver calls K2 (tests/test_prefilter_gpu.py
checks that K2's shards partition the pair list the same way)."""
import os
import sys

import numpy as np
import pytest

from conftest import ROOT
from util import random_family_table


def _worker(rank, world, port, n, s, seed, out_path):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    import torch
    import torch.distributed as dist
    import oracle
    from galah_b200 import distributed as gd
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(seed)
    table, counts = random_family_table(n, s, rng)
    n_local = n // world
    lt = torch.from_numpy(table[rank * n_local:(rank + 1) * n_local].view(np.int64).copy())
    lc = torch.from_numpy(counts[rank * n_local:(rank + 1) * n_local].view(np.int32).copy())

    def shard_fn(t, c, k, min_ani, shard, n_shards):
        tt = t.numpy().view(np.uint64)
        cc = c.numpy().view(np.uint32)
        assert np.array_equal(tt, table) and np.array_equal(cc, counts), "all-gather layout"
        rows = gd.rows_of_shard(len(tt), shard, n_shards)
        parts = [oracle.prefilter(tt, cc, k, min_ani, row_begin=int(r), row_end=int(r) + 1) for r in rows]
        return np.concatenate(parts) if parts else np.zeros(0, oracle.PAIR_DTYPE)

    merged = gd.prefilter_sharded(lt, lc, dist, shard_fn, 21, 0.9)
    if rank == 0:
        np.save(out_path, merged)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_prefilter_over_gloo(world, tmp_path):
    import torch.multiprocessing as mp
    import oracle
    n, s, seed = 66 * world, 100, 5
    out = str(tmp_path / "merged.npy")
    port = 29500 + (os.getpid() % 2000) + world
    mp.spawn(_worker, args=(world, port, n, s, seed, out), nprocs=world, join=True)
    merged = np.load(out)
    table, counts = random_family_table(n, s, np.random.default_rng(seed))
    exp = oracle.prefilter(table, counts, 21, 0.9)
    assert len(merged) == len(exp) and len(exp) > 0
    for f in ("i", "j", "common", "total"):
        assert np.array_equal(merged[f], exp[f])
    assert np.array_equal(merged["ani"].view(np.uint32), exp["ani"].view(np.uint32))


def test_shard_ownership_is_a_balanced_partition():
    from galah_b200 import distributed as gd
    n = 10_000
    for g in (1, 2, 4, 8):
        rows = [gd.rows_of_shard(n, r, g) for r in range(g)]
        assert sorted(np.concatenate(rows).tolist()) == list(range(n))
        area = [gd.pairs_of_shard(n, r, g) for r in range(g)]
        assert sum(area) == n * (n - 1) // 2
        assert max(area) / (sum(area) / g) < 1.01  # boustrophedon 64-row blocks balance the triangle
        assert all(int(gd.owner_of_row(int(r[0]), g)) == x for x, r in enumerate(rows))


def test_ring_ownership_partitions_the_block_pairs():
    """The ring exchange's symmetric ownership: every block pair {b1 <= b2} is joined by exactly one
    rank, that rank built one of the two lists, round k only needs the lists of peer rank - k, and
    the shares are balanced."""
    from galah_b200.distributed import ring_round_items
    for world, nbp in ((2, 3), (3, 4), (4, 5), (8, 30), (8, 1)):
        nb = world * nbp
        seen = np.zeros((nb, nb), np.int32)
        shares = []
        for r in range(world):
            rounds = ring_round_items(r, world, nbp)
            assert len(rounds) == world
            total = 0
            for k, items in enumerate(rounds):
                peer = (r - k) % world
                for rb, cb in items:
                    assert rb <= cb
                    owners = {int(rb) // nbp, int(cb) // nbp}
                    assert r in owners and owners <= {r, peer}
                    seen[rb, cb] += 1
                total += len(items)
            # round 0 starts with the diagonal pairs of the rank's own slice
            assert all(a == b for a, b in rounds[0][:nbp])
            shares.append(total)
        iu = np.triu_indices(nb)
        assert np.all(seen[iu] == 1) and seen.sum() == len(iu[0])
        assert max(shares) - min(shares) <= nbp * nbp // 2 + nbp
