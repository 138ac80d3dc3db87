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


def ring_round_items(rank, world, blocks_per_rank):
    """Work lists of the ring exchange: for every round k the (rb, cb) block pairs, rb <= cb, that
    `rank` joins once the lists of peer (rank - k) % world are resident (round 0: its own slice).
    Block b is built by rank b // blocks_per_rank.  Pair {b1 <= b2} belongs to the builder of b1 if
    b1 + b2 is even, else to the builder of b2, so every pair has exactly one owner, every owner
    holds one of the two lists locally and all ranks get the same share.  Round 0 lists the
    diagonal pairs first, then by distance (they take longest).  Returns a list of (m, 2) int32 arrays."""
    r, nbp = rank, blocks_per_rank
    mine = np.arange(r * nbp, (r + 1) * nbp, dtype=np.int64)
    rounds = []
    for k in range(world):
        peer = (r - k) % world
        if k == 0:
            a, b = np.meshgrid(mine, mine, indexing="ij")
            keep = a <= b
            lo, hi = a[keep], b[keep]
            order = np.lexsort((hi, lo, hi - lo))
            lo, hi = lo[order], hi[order]
        else:
            theirs = np.arange(peer * nbp, (peer + 1) * nbp, dtype=np.int64)
            a, b = np.meshgrid(mine, theirs, indexing="ij")
            lo, hi = np.minimum(a, b).ravel(), np.maximum(a, b).ravel()
            owner = np.where((lo + hi) % 2 == 0, lo // nbp, hi // nbp)
            keep = owner == r
            lo, hi = lo[keep], hi[keep]
        rounds.append(np.ascontiguousarray(np.stack([lo, hi], axis=1).astype(np.int32)))
    return rounds


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


class ShardedPrefilter:
    """The N-GPU stage-1 prefilter behind one call (the public multi-GPU API; bench.py's `value` and
    `e2e` at --gpus > 1): every rank passes ITS slice of the sketch table; per call the rank uploads
    the slice (H2D), builds the block lists of its own rows, the lists are all-gathered over NVLink,
    every rank joins its row-block shard and finishes its candidates on the host.

    When the slice is a whole number of row blocks (n_local % ROW_BLOCK == 0) the build needs no
    gathered table: the ranks agree on the table-wide largest hash with one 8-byte all-reduce, and
    the all-gather of the sketch table (the join needs it for the survivors' exact `total`) runs
    on NCCL's stream WHILE the lists are built.  Otherwise the table is gathered first and every
    rank builds 1/G of the blocks from it.  Buffers are allocated once and re-used.

    Experimental (GALAH_B200_RING=1): with whole-block slices and G > 1 the exchange can be a RING
    instead of all-gathers: block pair
    {b1 <= b2} belongs to the rank that built b1 if b1 + b2 is even, else to the one that built b2
    (every rank gets the same share and every item has one local list).  In round k a rank sends
    its lists + table slice to rank r+k and receives those of rank r-k (NCCL send/recv over
    NVLink); the join of the items against peer r-k is enqueued as soon as that round has landed,
    so rounds k+1.. travel while round k is joined: the exchange hides behind the join."""

    def __init__(self, gb, dist, n_local, stride, device):
        import torch
        self.gb, self.dist, self.torch = gb, dist, torch
        self.world, self.rank = dist.get_world_size(), dist.get_rank()
        self.n_local, self.s, self.dev = n_local, stride, device
        n = self.n = n_local * self.world
        t = torch
        self.my_table = t.empty((n_local, stride), dtype=t.int64, device=device)
        self.my_counts = t.empty(n_local, dtype=t.int32, device=device)
        self.table = t.empty((n, stride), dtype=t.int64, device=device)
        self.counts = t.empty(n, dtype=t.int32, device=device)
        nb, epb, slack = gb.blocklist_layout(n, stride)
        self.local_build = n_local % ROW_BLOCK == 0
        self.nbp = nbp = n_local // ROW_BLOCK if self.local_build else (nb + self.world - 1) // self.world
        self.epb = epb
        self.my_hi = t.empty(nbp * epb, dtype=t.int32, device=device)
        self.my_lo = t.empty_like(self.my_hi)
        self.my_tags = t.empty(nbp * epb, dtype=t.uint8, device=device)
        self.my_len = t.empty(nbp, dtype=t.int32, device=device)
        self.all_hi = t.zeros(self.world * nbp * epb + slack, dtype=t.int32, device=device)
        self.all_lo = t.zeros_like(self.all_hi)
        self.all_tags = t.zeros(self.world * nbp * epb + slack, dtype=t.uint8, device=device)
        self.all_len = t.empty(self.world * nbp, dtype=t.int32, device=device)
        self.gmax = t.zeros(1, dtype=t.int64, device=device)  # uint64 bits of the largest valid hash
        self.cand_cap = max(1 << 20, 64 * n)
        self.d_cand = t.empty((self.cand_cap, 4), dtype=t.int32, device=device)
        self.d_ncand = t.zeros(1, dtype=t.int64, device=device)
        self.h_cand = t.empty((self.cand_cap, 4), dtype=t.int32).pin_memory()
        import os
        # experimental, off by default (GALAH_B200_RING=1): bit-exact (tools/check_sharded.py), but at
        # G = 2 it measured no faster than the all-gathers -- NCCL's send/recv kernels get no SM while
        # the persistent join CTAs run, so the rounds do not overlap the join yet
        self.ring = self.local_build and self.world > 1 and os.environ.get("GALAH_B200_RING") == "1"
        if self.ring:
            self._init_ring()

    def _grow_candidates(self, need):
        t = self.torch
        self.cand_cap = int(need) + (int(need) >> 2) + 1024
        self.d_cand = t.empty((self.cand_cap, 4), dtype=t.int32, device=self.dev)
        self.h_cand = t.empty((self.cand_cap, 4), dtype=t.int32).pin_memory()

    def _init_ring(self):
        """Per round k the explicit item list (rb, cb) this rank joins once peer (rank - k)'s lists are
        resident; round 0 = the items inside its own slice (diagonal items first: they take longest)."""
        t, G, r, nbp = self.torch, self.world, self.rank, self.nbp
        self.round_items, self.round_n = [], []
        for items in ring_round_items(r, G, nbp):
            self.round_n.append(len(items))
            self.round_items.append(t.from_numpy(items).to(self.dev) if len(items) else None)
        # views of this rank's slice inside the table-wide buffers: lists are built and the slice is
        # uploaded in place, peers' slices are received into theirs
        e, nl = nbp * self.epb, self.n_local
        self.v_table = lambda q: self.table[q * nl:(q + 1) * nl]
        self.v_counts = lambda q: self.counts[q * nl:(q + 1) * nl]
        self.v_hi = lambda q: self.all_hi[q * e:(q + 1) * e]
        self.v_lo = lambda q: self.all_lo[q * e:(q + 1) * e]
        self.v_tags = lambda q: self.all_tags[q * e:(q + 1) * e]
        self.v_len = lambda q: self.all_len[q * nbp:(q + 1) * nbp]

    def _step_ring(self, k, min_ani):
        t, gb, dist = self.torch, self.gb, self.dist
        st = t.cuda.current_stream().cuda_stream
        n, s, G, r = self.n, self.s, self.world, self.rank
        self.v_table(r).copy_(self.my_table, non_blocking=True)
        self.v_counts(r).copy_(self.my_counts, non_blocking=True)
        gb.table_max_device(self.my_table.data_ptr(), self.my_counts.data_ptr(), self.n_local, s,
                            self.gmax.data_ptr(), st)
        self.gmax.bitwise_xor_(-(1 << 63))  # the collective compares int64: map unsigned order onto signed
        dist.all_reduce(self.gmax, op=dist.ReduceOp.MAX)
        self.gmax.bitwise_xor_(-(1 << 63))
        gb.blocklist_build_local(self.my_table.data_ptr(), self.my_counts.data_ptr(), self.n_local, s,
                                 self.gmax.data_ptr(), self.nbp, self.v_hi(r).data_ptr(), self.v_lo(r).data_ptr(),
                                 self.v_tags(r).data_ptr(), self.v_len(r).data_ptr(), st)
        works = []
        for rnd in range(1, G):
            dst, src = (r + rnd) % G, (r - rnd) % G
            ops = []
            for view in (self.v_hi, self.v_lo, self.v_tags, self.v_len, self.v_table, self.v_counts):
                ops.append(dist.P2POp(dist.isend, view(r), dst))
                ops.append(dist.P2POp(dist.irecv, view(src), src))
            works.append(dist.batch_isend_irecv(ops))
        for rnd in range(G):
            if rnd:
                for w in works[rnd - 1]:
                    w.wait()
            if self.round_n[rnd]:
                gb.prefilter_join_items_enqueue(self.table.data_ptr(), self.counts.data_ptr(), n, s, k, min_ani,
                                                self.all_hi.data_ptr(), self.all_lo.data_ptr(),
                                                self.all_tags.data_ptr(), self.all_len.data_ptr(),
                                                self.round_items[rnd].data_ptr(), self.round_n[rnd], rnd == 0, st,
                                                self.d_cand.data_ptr(), self.cand_cap, self.d_ncand.data_ptr())
            elif rnd == 0:
                self.d_ncand.zero_()

    def step_device(self, k=21, min_ani=0.9, screen=None):
        """Everything after the upload, enqueued on the current stream: my_table / my_counts (device)
        -> candidates in d_cand / d_ncand.  No host synchronisation.  screen = None: the finch rule
        (Mash ANI >= min_ani); screen = faster_small (bool): the rows are marker sketches and the join
        applies the skani-style containment rule."""
        t, gb, dist = self.torch, self.gb, self.dist
        st = t.cuda.current_stream().cuda_stream
        n, s, w, r = self.n, self.s, self.world, self.rank
        m = w * self.nbp * self.epb
        if self.ring and screen is None:
            return self._step_ring(k, min_ani)
        if self.local_build:
            gb.table_max_device(self.my_table.data_ptr(), self.my_counts.data_ptr(), self.n_local, s,
                                self.gmax.data_ptr(), st)
            # the collective compares int64: flipping the top bit maps unsigned order onto signed order
            self.gmax.bitwise_xor_(-(1 << 63))
            dist.all_reduce(self.gmax, op=dist.ReduceOp.MAX)
            self.gmax.bitwise_xor_(-(1 << 63))
            work_t = dist.all_gather_into_tensor(self.table, self.my_table, async_op=True)
            work_c = dist.all_gather_into_tensor(self.counts, self.my_counts, async_op=True)
            gb.blocklist_build_local(self.my_table.data_ptr(), self.my_counts.data_ptr(), self.n_local, s,
                                     self.gmax.data_ptr(), self.nbp, self.my_hi.data_ptr(), self.my_lo.data_ptr(),
                                     self.my_tags.data_ptr(), self.my_len.data_ptr(), st)
            dist.all_gather_into_tensor(self.all_hi[:m], self.my_hi)
            dist.all_gather_into_tensor(self.all_lo[:m], self.my_lo)
            dist.all_gather_into_tensor(self.all_tags[:m], self.my_tags)
            dist.all_gather_into_tensor(self.all_len, self.my_len)
            work_t.wait()
            work_c.wait()
        else:
            dist.all_gather_into_tensor(self.table, self.my_table)
            dist.all_gather_into_tensor(self.counts, self.my_counts)
            gb.blocklist_build(self.table.data_ptr(), self.counts.data_ptr(), n, s, r * self.nbp, (r + 1) * self.nbp,
                               self.my_hi.data_ptr(), self.my_lo.data_ptr(), self.my_tags.data_ptr(),
                               self.my_len.data_ptr(), st)
            dist.all_gather_into_tensor(self.all_hi[:m], self.my_hi)
            dist.all_gather_into_tensor(self.all_lo[:m], self.my_lo)
            dist.all_gather_into_tensor(self.all_tags[:m], self.my_tags)
            dist.all_gather_into_tensor(self.all_len, self.my_len)
        if screen is not None:
            gb.prefilter_join_enqueue_screen(self.table.data_ptr(), self.counts.data_ptr(), n, s, bool(screen),
                                             self.all_hi.data_ptr(), self.all_lo.data_ptr(), self.all_tags.data_ptr(),
                                             self.all_len.data_ptr(), r, w, st, self.d_cand.data_ptr(), self.cand_cap,
                                             self.d_ncand.data_ptr())
            return
        gb.prefilter_join_enqueue(self.table.data_ptr(), self.counts.data_ptr(), n, s, k, min_ani,
                                  self.all_hi.data_ptr(), self.all_lo.data_ptr(), self.all_tags.data_ptr(),
                                  self.all_len.data_ptr(), r, w, st, self.d_cand.data_ptr(), self.cand_cap,
                                  self.d_ncand.data_ptr())

    def __call__(self, h_table, h_counts, k=21, min_ani=0.9):
        """h_table / h_counts: this rank's slice as (pinned) host torch tensors (int64 / int32 views of
        the uint64 / uint32 data).  Returns this rank's PAIR_DTYPE records, sorted by (i, j)."""
        self.my_table.copy_(h_table, non_blocking=True)
        self.my_counts.copy_(h_counts, non_blocking=True)
        self.step_device(k, min_ani)
        got = int(self.d_ncand.item())  # D2H + sync
        if got > self.cand_cap:  # grow and run again (as the single-GPU call does)
            self._grow_candidates(got)
            self.step_device(k, min_ani)
            got = int(self.d_ncand.item())
        self.h_cand[:got].copy_(self.d_cand[:got])
        return self.gb.finish_candidates(self.h_cand[:got].numpy().view(np.uint32), k, min_ani)


def route_hits(hits_i, n_local, world):
    """Rank that evaluates stage 2 of a hit (i, j), i < j: the owner of the QUERY genome i (genome
    slices are contiguous, n_local per rank).  Pure function (CPU-tested)."""
    return np.minimum(np.asarray(hits_i, dtype=np.int64) // n_local, world - 1)


def allgather_var(dist, dev, arr, dtype, width):
    """All-gather of per-rank (m_r, width) host arrays of `dtype` through one padded collective on `dev` (the sizes
    travel first).  Returns one array per rank."""
    import torch as t
    world = dist.get_world_size()
    m = t.tensor([len(arr)], dtype=t.int64, device=dev)
    ms = t.empty(world, dtype=t.int64, device=dev)
    dist.all_gather_into_tensor(ms, m)
    ms = ms.cpu().numpy()
    cap = int(ms.max())
    if cap == 0:
        return [np.zeros((0, width), dtype) for _ in range(world)]
    buf = np.zeros((cap, width), dtype)
    buf[: len(arr)] = arr
    mine = t.from_numpy(buf.view(np.uint8).reshape(-1)).to(dev)
    allb = t.empty(world * mine.numel(), dtype=t.uint8, device=dev)
    dist.all_gather_into_tensor(allb, mine)
    allb = allb.cpu().numpy().view(dtype).reshape(world, cap, width)
    return [allb[r, : int(ms[r])] for r in range(world)]


def cluster_in_waves_replicated(gb, dist, dev, hits, n, n_local, ani_threshold, evaluate_mine, stats=None):
    """Stage 2 + the greedy engine on every rank of a sharded run.  All ranks hold the same hit list, so their engines
    (galah_b200_cluster_from_distances_batched) ask for the same (representative, genome) pairs wave by wave -- exactly
    the pairs the reference's two passes evaluate (src/clusterer.rs:216-300, 350-449; the representative is the query).
    A request is evaluated by the rank that owns its QUERY genome (`evaluate_mine(q, r)`: global ids -> float32 ANIs),
    the values of a wave are all-gathered (8 B per request) and every engine moves on identically: one small
    collective per wave, no gather of all hits' values in both orientations.  Returns (clusters, info) on every rank."""
    import time
    rank, world = dist.get_rank(), dist.get_world_size()
    stats = stats if stats is not None else {}
    for k in ("pairs_ms", "gather_ms", "mine", "asked"):
        stats.setdefault(k, 0)

    def ani_batch(reps, genomes):
        ta = time.perf_counter()
        q, r = reps.astype(np.int64), genomes.astype(np.int64)
        my = np.nonzero(route_hits(q, n_local, world) == rank)[0]
        back = np.zeros((len(my), 2), np.uint32)
        back[:, 0] = my
        if len(my):
            back[:, 1] = np.ascontiguousarray(evaluate_mine(q[my], r[my]), np.float32).view(np.uint32)
        tb = time.perf_counter()
        out = np.zeros(len(reps), np.float32)
        for part in allgather_var(dist, dev, back, np.uint32, 2):
            out[part[:, 0]] = part[:, 1].view(np.float32)
        stats["pairs_ms"] += 1e3 * (tb - ta); stats["gather_ms"] += 1e3 * (time.perf_counter() - tb)
        stats["mine"] += len(my); stats["asked"] += len(reps)
        return out

    return gb.cluster_from_distances_batched(n, hits, ani_threshold, ani_batch)


class ShardedPipeline:
    """The whole two-stage path on G GPUs, one process per GPU (bench.py at --gpus > 1; BASELINE.json
    configs[3]): rank r owns genomes [r n_local, (r + 1) n_local).

      K1 + K3 index   every rank sketches and indexes ITS genomes (no communication)
      K2              ShardedPrefilter: 8-byte all-reduce, NCCL all-gather of the sketch table and of the
                      block lists over NVLink, boustrophedon row-block shard of the join per rank
      hits            every rank finishes its candidates in f64 on its host; the (small) hit lists are
                      all-gathered so that every rank knows all hits
      K3              both orientations of a hit (i, j) are evaluated, each by the owner of its query; when the
                      reference genome lives on another rank its hash table is read IN PLACE over NVLink
                      through a CUDA IPC mapping of the owner's table array (no gather of the 0.4 MB /
                      genome index)
      engine          ANI values are gathered to rank 0, which runs the greedy engine (host)
    Returns (clusters, info) on rank 0 and (None, info) elsewhere."""

    def __init__(self, gb, dist, n_local, stride, device):
        import torch
        self.gb, self.dist, self.torch = gb, dist, torch
        self.world, self.rank = dist.get_world_size(), dist.get_rank()
        self.n_local, self.s, self.dev = n_local, stride, device
        self.n = n_local * self.world
        self.sp = ShardedPrefilter(gb, dist, n_local, stride, device)
        self._idx = {}   # small_genomes -> AniIndex, cleared and re-used per step (no cudaMalloc / cudaFree in a step)

    def _allgather_var(self, arr, dtype, width):
        return allgather_var(self.dist, self.dev, arr, dtype, width)

    def _route_var(self, arr, dest, dtype, width):
        """Variable all-to-all of the rows of a host (m, width) array of `dtype`: row x goes to rank
        dest[x]; returns the rows this rank receives (grouped by source rank, source order kept).
        One NCCL all_to_all_single over byte buffers (the counts travel first)."""
        t, dist = self.torch, self.dist
        world = self.world
        order = np.argsort(dest, kind="stable")
        send = np.ascontiguousarray(arr[order])
        counts = np.bincount(dest, minlength=world).astype(np.int64)
        c_in = t.from_numpy(counts).to(self.dev)
        c_out = t.empty(world, dtype=t.int64, device=self.dev)
        dist.all_to_all_single(c_out, c_in)
        recv_counts = c_out.cpu().numpy()
        row = width * np.dtype(dtype).itemsize
        src = t.from_numpy(send.view(np.uint8).reshape(-1)).to(self.dev) if len(send) else t.empty(0, dtype=t.uint8, device=self.dev)
        dst = t.empty(int(recv_counts.sum()) * row, dtype=t.uint8, device=self.dev)
        dist.all_to_all_single(dst, src, output_split_sizes=[int(c) * row for c in recv_counts],
                               input_split_sizes=[int(c) * row for c in counts])
        return dst.cpu().numpy().view(dtype).reshape(-1, width)

    def _run(self, seq2, valid, d_base_off, base_off, lengths, device, min_ani, ani_pct, min_af, small_genomes=False):
        import time
        t, gb, dist, sp = self.torch, self.gb, self.dist, self.sp
        rank, world, n_local, n = self.rank, self.world, self.n_local, self.n
        info = {}
        t0 = time.perf_counter()
        if small_genomes not in self._idx:
            self._idx[small_genomes] = gb.AniIndex(small_genomes=small_genomes)
        idx = self._idx[small_genomes]
        idx.clear()
        if isinstance(valid, tuple):   # host buffers without the validity bitmap: (invalid begin, invalid end) arrays
            k1_ms, idx_ms = idx.ingest_packed_sparse(seq2, valid, base_off, lengths, sp.my_table.data_ptr(), sp.my_counts.data_ptr())
        else:
            k1_ms, idx_ms = idx.ingest_packed(seq2, valid, base_off, lengths, sp.my_table.data_ptr(), sp.my_counts.data_ptr(),
                                              device=device, d_base_off=d_base_off)
        t.cuda.synchronize()
        t1 = time.perf_counter()
        # ---- K2 (torch's current stream; the library's stream is idle: ingest_packed synchronised)
        sp.step_device(21, min_ani)
        got = int(sp.d_ncand.item())
        if got > sp.cand_cap:
            sp._grow_candidates(got)
            sp.step_device(21, min_ani)
            got = int(sp.d_ncand.item())
        sp.h_cand[:got].copy_(sp.d_cand[:got])
        mine = gb.finish_candidates(sp.h_cand[:got].numpy().view(np.uint32), 21, min_ani)
        t2 = time.perf_counter()
        tm = {}
        tq = time.perf_counter()
        # ---- every rank learns all hits (20 B each) and takes those whose query genome it owns
        packed = np.zeros((len(mine), 5), np.uint32)
        for c, f in enumerate(("i", "j", "common", "total")):
            packed[:, c] = mine[f]
        packed[:, 4] = mine["ani"].view(np.uint32)
        parts = self._allgather_var(packed, np.uint32, 5)
        allh = np.concatenate(parts) if parts else np.zeros((0, 5), np.uint32)
        order = np.argsort((allh[:, 0].astype(np.uint64) << np.uint64(32)) | allh[:, 1].astype(np.uint64), kind="stable")
        allh = allh[order]
        tm["hits_gather_sort_ms"] = 1e3 * (time.perf_counter() - tq); tq = time.perf_counter()
        # ---- stage 2 in waves, the greedy engine REPLICATED on every rank: all ranks hold the same hit list, so their
        # engines ask for the same (representative, genome) pairs wave by wave -- exactly the pairs the reference's two
        # passes evaluate (src/clusterer.rs:216-300, 350-449; calculate_ani(rep, genome) makes the representative the
        # query).  A request runs on the rank that owns its QUERY genome, reading the other genome's table in place on
        # its peer; the values of a wave are all-gathered (8 B per request) and every engine moves on identically.
        q_all = np.concatenate([allh[:, 0], allh[:, 1]]).astype(np.int64)
        r_all = np.concatenate([allh[:, 1], allh[:, 0]]).astype(np.int64)
        mine_possible = route_hits(q_all, n_local, world) == rank
        tm["jobs_ms"] = 1e3 * (time.perf_counter() - tq); tq = time.perf_counter()
        # ---- peer tables: IPC handles + per-genome offsets are exchanged (host metadata, 16 B / genome)
        handle, table_off, total_len = idx.export_tables()
        metas = [None] * world
        dist.all_gather_object(metas, (handle, table_off, total_len))
        tm["meta_exchange_ms"] = 1e3 * (time.perf_counter() - tq); tq = time.perf_counter()
        base = np.zeros(world, np.int64)      # id of a rank's first genome in this index's numbering
        for peer in sorted(set(int(x) for x in route_hits(r_all[mine_possible], n_local, world)) - {rank}):
            base[peer] = idx.attach_peer(*metas[peer])
        tm["attach_ms"] = 1e3 * (time.perf_counter() - tq); tq = time.perf_counter()
        hits = np.zeros(len(allh), PAIR_DTYPE)
        hits["i"], hits["j"], hits["common"], hits["total"] = allh[:, 0], allh[:, 1], allh[:, 2], allh[:, 3]
        hits["ani"] = allh[:, 4].view(np.float32)
        wave = {"chain_ms": 0.0, "remote": 0}

        def evaluate_mine(q, r):  # global genome ids of the pairs whose query this rank owns
            r_owner = route_hits(r, n_local, world)
            pairs = np.stack([q - rank * n_local, base[r_owner] + (r - r_owner * n_local)], axis=1).astype(np.uint32)
            res = idx.pairs(pairs, min_af)
            wave["chain_ms"] += idx.last_timing()[1]
            wave["remote"] += int(np.sum(r_owner != rank))
            return res["ani"]

        clusters, cinfo = cluster_in_waves_replicated(gb, dist, self.dev, hits, n, n_local, ani_pct, evaluate_mine, wave)
        t3 = time.perf_counter()
        dist.barrier()          # every peer has finished reading this rank's tables
        idx.clear()             # detaches the peers; the allocations stay for the next step
        chain_ms = wave["chain_ms"]
        tm["pairs_call_ms"] = wave["pairs_ms"]; tm["ani_gather_ms"] = wave["gather_ms"]
        tm["engine_call_ms"] = 1e3 * (t3 - tq) - wave["pairs_ms"] - wave["gather_ms"]
        info.update(cinfo)
        if rank != 0:
            clusters = None
        info["host_detail_ms"] = {k: round(v, 2) for k, v in tm.items()}
        t4 = time.perf_counter()
        # ani_ms: K3 launches + the per-wave exchanges; engine_ms: the replicated engine's own work + the closing barrier
        ani_total = (1e3 * (tq - t2)) + wave["pairs_ms"] + wave["gather_ms"]
        info.update(n_precluster_hits=len(allh), n_ani_pairs=wave["asked"], my_ani_pairs=wave["mine"],
                    remote_reference_pairs=wave["remote"], sketch_ms=k1_ms, index_ms=idx_ms,
                    ingest_ms=1e3 * (t1 - t0), prefilter_ms=1e3 * (t2 - t1), ani_ms=ani_total, ani_chain_ms=chain_ms,
                    engine_ms=1e3 * (t4 - t2) - ani_total, total_ms=1e3 * (t4 - t0))
        return clusters, info

    def run_skani(self, seq2, valid, d_base_off, base_off, lengths, device, threshold_pct=95.0, ani_pct=95.0, min_af=15.0,
                  small_genomes=True, individual_contigs=True):
        """cluster() with SkaniPreclusterer + SkaniClusterer / --cluster-contigs (skip_clusterer: the
        preclusterer's ANIs decide, src/clusterer.rs:32-44) on G GPUs -- BASELINE.json configs[4].  The
        pipeline's stride must be marker_row_capacity(longest unit, small_genomes).  Every rank makes the
        marker sketches + K3 index of ITS units in one pass; the marker table and block lists are
        all-gathered; the containment screen is row-block sharded; a screened pair (i, j) is evaluated
        by the owner of i (the query, src/skani.rs:109-225) reading j's table in place; rank 0 runs the engine."""
        import time
        t, gb, dist, sp = self.torch, self.gb, self.dist, self.sp
        rank, world, n_local, n = self.rank, self.world, self.n_local, self.n
        t0 = time.perf_counter()
        if small_genomes not in self._idx:
            self._idx[small_genomes] = gb.AniIndex(small_genomes=small_genomes)
        idx = self._idx[small_genomes]
        idx.clear()
        mk_ms, idx_ms = idx.ingest_packed_markers(seq2, valid, base_off, lengths, self.s, sp.my_table.data_ptr(),
                                                  sp.my_counts.data_ptr(), device=device, d_base_off=d_base_off)
        t.cuda.synchronize()
        if int((sp.my_counts == -1).sum().item()):
            raise RuntimeError("a unit holds more markers than the row stride")
        t1 = time.perf_counter()
        sp.step_device(21, 0.0, screen=small_genomes)
        got = int(sp.d_ncand.item())
        if got > sp.cand_cap:
            sp._grow_candidates(got)
            sp.step_device(21, 0.0, screen=small_genomes)
            got = int(sp.d_ncand.item())
        sp.h_cand[:got].copy_(sp.d_cand[:got])
        mine = sp.h_cand[:got].numpy().view(np.uint32).copy()
        t2 = time.perf_counter()
        tm = {}
        tq = time.perf_counter()
        # every screened pair travels ONCE, to the rank that owns its query genome (the lower index)
        myc = self._route_var(mine, route_hits(mine[:, 0], n_local, world), np.uint32, 4)
        tm["route_ms"] = 1e3 * (time.perf_counter() - tq); tq = time.perf_counter()
        myc = myc[np.argsort((myc[:, 0].astype(np.uint64) << np.uint64(32)) | myc[:, 1].astype(np.uint64), kind="stable")]
        tm["sort_ms"] = 1e3 * (time.perf_counter() - tq); tq = time.perf_counter()
        handle, table_off, total_len = idx.export_tables()
        metas = [None] * world
        dist.all_gather_object(metas, (handle, table_off, total_len))
        tm["meta_exchange_ms"] = 1e3 * (time.perf_counter() - tq); tq = time.perf_counter()
        r_owner = route_hits(myc[:, 1], n_local, world)
        base = np.zeros(world, np.int64)
        for peer in sorted(set(int(x) for x in r_owner) - {rank}):
            base[peer] = idx.attach_peer(*metas[peer])
        tm["attach_ms"] = 1e3 * (time.perf_counter() - tq); tq = time.perf_counter()
        q_local = myc[:, 0].astype(np.int64) - rank * n_local
        r_id = base[r_owner] + (myc[:, 1].astype(np.int64) - r_owner * n_local)
        res = idx.pairs(np.stack([q_local, r_id], axis=1).astype(np.uint32), min_af, individual_contigs=individual_contigs)
        chain_ms = idx.last_timing()[1]
        tm["pairs_call_ms"] = 1e3 * (time.perf_counter() - tq)
        t3 = time.perf_counter()
        keep = res["ani"] >= np.float32(threshold_pct)  # `if ani >= threshold`, src/skani.rs:205
        back = np.zeros((int(keep.sum()), 5), np.uint32)
        back[:, :4] = myc[keep]
        back[:, 4] = res["ani"][keep].view(np.uint32)
        n_screened = t.tensor([len(myc)], dtype=t.int64, device=self.dev)
        dist.all_reduce(n_screened)
        hits_all = self._route_var(back, np.zeros(len(back), np.int64), np.uint32, 5)  # the hits go to rank 0 only
        dist.barrier()
        idx.clear()
        tm["hits_route_ms"] = 1e3 * (time.perf_counter() - t3); tq = time.perf_counter()
        clusters, info = None, {}
        if rank == 0:
            hits_all = hits_all[np.argsort((hits_all[:, 0].astype(np.uint64) << np.uint64(32)) | hits_all[:, 1].astype(np.uint64),
                                           kind="stable")]
            hits = np.zeros(len(hits_all), PAIR_DTYPE)
            hits["i"], hits["j"], hits["common"], hits["total"] = hits_all[:, 0], hits_all[:, 1], hits_all[:, 2], hits_all[:, 3]
            hits["ani"] = hits_all[:, 4].view(np.float32)
            clusters, cinfo = gb.cluster_from_distances(n, hits, ani_pct, None, skip_clusterer=True)
            info.update(cinfo)
            info["n_hits"] = int(len(hits))
            tm["engine_call_ms"] = 1e3 * (time.perf_counter() - tq)
        info["host_detail_ms"] = {k: round(v, 2) for k, v in tm.items()}
        allc, my_rows = myc, np.arange(len(myc))
        t4 = time.perf_counter()
        info.update(n_screened=int(n_screened.item()), my_ani_pairs=int(len(my_rows)), remote_reference_pairs=int(np.sum(r_owner != rank)),
                    markers_ms=mk_ms, index_ms=idx_ms, ingest_ms=1e3 * (t1 - t0), screen_ms=1e3 * (t2 - t1),
                    ani_ms=1e3 * (t3 - t2), ani_chain_ms=chain_ms, engine_ms=1e3 * (t4 - t3), total_ms=1e3 * (t4 - t0))
        return clusters, info

    def step_device(self, d_seq2, d_valid, d_base_off, base_off, lengths, min_ani=0.9, ani_pct=95.0, min_af=15.0,
                    small_genomes=False):
        """This rank's genomes are resident in HBM (device pointers; base_off / lengths host arrays)."""
        return self._run(d_seq2, d_valid, d_base_off, base_off, lengths, True, min_ani, ani_pct, min_af, small_genomes)

    def step_host(self, h_seq2, h_valid, base_off, lengths, min_ani=0.9, ani_pct=95.0, min_af=15.0, small_genomes=False):
        """This rank's genomes are HOST buffers (addresses of pinned memory): uploaded inside the call.
        h_valid: address of the validity bitmap, or a (begin, end) tuple of invalid base ranges (no bitmap upload)."""
        return self._run(h_seq2, h_valid, 0, base_off, lengths, False, min_ani, ani_pct, min_af, small_genomes)
