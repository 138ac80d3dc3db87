"""Multi-GPU driver for the stage-1 path: one process per GPU, torch.distributed for the plumbing.

The path shards by rows of the pair grid: rank r sketches genomes
[r*N/G, (r+1)*N/G) (K1, no collective), ONE all-gather assembles the N x s sketch table on every
rank (the only exchange step on the path; NCCL over NVLink on GPUs, gloo in the CPU tests), and
rank r evaluates the row blocks of GALAH_B200_ROW_BLOCK rows that the boustrophedon rule
0,1,..,G-1,G-1,..,0,0,1,.. assigns to it (so the triangular pair area is balanced).  Pair lists are gathered to rank 0 and merged by (i, j), which
is the iteration order of the reference's SortedPairGenomeDistanceCache.
"""
import numpy as np

from .api import PAIR_DTYPE, ROW_BLOCK


def owner_of_row(i, n_shards, row_block=ROW_BLOCK):
    """Rank that evaluates the pairs (i, j > i) (mirrors shard_of_group in csrc/prefilter.cuh)."""
    group = np.asarray(i, dtype=np.int64) // row_block
    rnd, pos = group // n_shards, group % n_shards
    return np.where(rnd % 2 == 1, n_shards - 1 - pos, pos)


def rows_of_shard(n, shard, n_shards, row_block=ROW_BLOCK):
    """Row indices owned by `shard` (ascending)."""
    rows = np.arange(n)
    return rows[owner_of_row(rows, n_shards, row_block) == shard]


def pairs_of_shard(n, shard, n_shards, row_block=ROW_BLOCK):
    """Number of (i, j) pairs with i < j that `shard` evaluates."""
    rows = rows_of_shard(n, shard, n_shards, row_block)
    return int(np.sum(n - 1 - rows))


def all_gather_table(local_table, local_counts, dist, device=None):
    """All-gather equally sized row slices into the full table (rank order = genome order).
    local_table: torch tensor (n_local, s) int64; local_counts: (n_local,) int32."""
    import torch
    world = dist.get_world_size()
    n_local, s = local_table.shape
    table = torch.empty((n_local * world, s), dtype=local_table.dtype, device=local_table.device)
    counts = torch.empty(n_local * world, dtype=local_counts.dtype, device=local_counts.device)
    dist.all_gather_into_tensor(table, local_table.contiguous())
    dist.all_gather_into_tensor(counts, local_counts.contiguous())
    return table, counts


def prefilter_sharded(local_table, local_counts, dist, shard_fn, k=21, min_ani=0.9):
    """Distributed finch prefilter.  `shard_fn(table, counts, k, min_ani, shard, n_shards)` evaluates
    one shard and returns PAIR_DTYPE records (on GPUs: galah_b200.prefilter_device on the gathered
    device table).  Returns the merged, (i, j)-sorted pair list on rank 0 and None elsewhere."""
    rank, world = dist.get_rank(), dist.get_world_size()
    table, counts = all_gather_table(local_table, local_counts, dist)
    mine = np.ascontiguousarray(shard_fn(table, counts, k, min_ani, rank, world), PAIR_DTYPE)
    gathered = [None] * world if rank == 0 else None
    dist.gather_object(mine, gathered, dst=0)
    if rank != 0:
        return None
    allp = np.concatenate(gathered) if gathered else np.zeros(0, PAIR_DTYPE)
    return allp[np.lexsort((allp["j"], allp["i"]))]


def gpu_shard_fn(gb):
    """shard_fn for prefilter_sharded that runs K2 on the gathered device table."""
    import torch

    def fn(table, counts, k, min_ani, shard, n_shards):
        n, s = table.shape
        return gb.prefilter_device(table.data_ptr(), counts.data_ptr(), n, s, k, min_ani, shard, n_shards,
                                   torch.cuda.current_stream().cuda_stream)
    return fn
