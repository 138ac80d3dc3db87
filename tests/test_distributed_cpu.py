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


def _stub_ani(q, r):
    """Deterministic, orientation-dependent stage-2 ANI of (query, reference) genome ids: values around the threshold."""
    q = np.asarray(q, np.uint64); r = np.asarray(r, np.uint64)
    h = (q * np.uint64(0x9E3779B97F4A7C15) + r * np.uint64(0xC2B2AE3D27D4EB4F)) >> np.uint64(40)
    return (90.0 + (h % np.uint64(1000)).astype(np.float32) / np.float32(100.0)).astype(np.float32)


def _family_hits(n, seed):
    import galah_b200 as gb
    rng = np.random.default_rng(seed)
    pairs = [(a, b) for base in range(0, n, 11) for a in range(base, min(n, base + 11)) for b in range(a + 1, min(n, base + 11))
             if rng.uniform() < 0.7]
    pairs += [(int(a), int(b)) for a, b in rng.integers(0, n, size=(n // 8, 2)) if a < b]  # a few links across families
    pairs = sorted(set(pairs))
    hits = np.zeros(len(pairs), gb.PAIR_DTYPE)
    hits["i"], hits["j"] = [p[0] for p in pairs], [p[1] for p in pairs]
    hits["ani"] = rng.uniform(0.9, 1.0, len(pairs)).astype(np.float32)
    return hits


def _wave_worker(rank, world, port, n, seed, out_dir):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    import torch
    import torch.distributed as dist
    import galah_b200 as gb
    from galah_b200 import distributed as gd
    dist.init_process_group("gloo", rank=rank, world_size=world)
    hits = _family_hits(n, seed)
    n_local = n // world
    seen = []

    def evaluate_mine(q, r):
        assert np.all(gd.route_hits(q, n_local, world) == rank), "a request reached a rank that does not own its query"
        seen.append(np.stack([q, r], axis=1))
        return _stub_ani(q, r)

    stats = {}
    clusters, info = gd.cluster_in_waves_replicated(gb, dist, torch.device("cpu"), hits, n, n_local, 95.0, evaluate_mine, stats)
    mine = np.concatenate(seen) if seen else np.zeros((0, 2), np.int64)
    np.savez(os.path.join(out_dir, f"rank{rank}.npz"), members=clusters.members, offsets=clusters.offsets, mine=mine,
             asked=stats["asked"], n_mine=stats["mine"], waves=info["ani_waves"], calls=info["ani_calls"])
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_replicated_wave_engine_over_gloo(world, tmp_path):
    """Stage 2 of the sharded pipeline: every rank runs the wave engine on the same hit list, evaluates the requests
    whose query genome it owns (a stub ANI here; K3 on the GPUs) and the values of a wave are all-gathered.  Every rank
    ends with the clusters of the serial engine; every request was evaluated exactly once, by the owner of its query;
    the number of evaluations is the reference's calculate_ani count."""
    import torch.multiprocessing as mp
    import galah_b200 as gb
    n, seed = 66 * world, 9
    port = 29600 + (os.getpid() % 2000) + world
    mp.spawn(_wave_worker, args=(world, port, n, seed, str(tmp_path)), nprocs=world, join=True)
    hits = _family_hits(n, seed)
    want, winfo = gb.cluster_from_distances(n, hits, 95.0, lambda rep, g: float(_stub_ani([rep], [g])[0]))
    evaluated = []
    for rank in range(world):
        z = np.load(str(tmp_path / f"rank{rank}.npz"))
        got = gb.ClusterList(z["members"], z["offsets"])
        assert got == want, f"rank {rank}"
        assert int(z["calls"]) == winfo["ani_calls"] == int(z["asked"]) and int(z["waves"]) >= 2
        evaluated.append(z["mine"])
        assert len(z["mine"]) == int(z["n_mine"])
    allp = np.concatenate(evaluated)
    assert len(allp) == winfo["ani_calls"] and len({(int(a), int(b)) for a, b in allp}) == len(allp)
    assert all(len(e) > 0 for e in evaluated)
