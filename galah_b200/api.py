"""Python face of the C ABI: marshals numpy arrays / torch device pointers, nothing else.

Host-side logic (FASTA ingest, clustering) lives in C++ behind the same ABI; the CUDA kernels are
the only compute path.
"""
import ctypes
import os

import numpy as np

from . import _native
from ._native import GalahB200Error, Pair, check, lib

PAIR_DTYPE = np.dtype(
    [("i", "<u4"), ("j", "<u4"), ("common", "<u4"), ("total", "<u4"), ("ani", "<f4")]
)
ROW_BLOCK = 128
PAD = np.uint64(0xFFFFFFFFFFFFFFFF)

_bound_device = None


def init(device=0):
    """Bind this process to CUDA device `device` (galah_b200_init)."""
    global _bound_device
    check(lib().galah_b200_init(int(device)))
    _bound_device = int(device)
    return _bound_device


def device_count():
    return int(lib().galah_b200_device_count())


def launch_count():
    return int(lib().galah_b200_launch_count())


def prefilter_mode(mode=-1):
    """Select the K2 kernel path (0 = block-list join, 1 = pairwise warp merge); returns the
    previous mode.  Both are exact; mode 1 exists for cross-checks and as the large-s fallback."""
    return int(lib().galah_b200_prefilter_mode(int(mode)))


def prefilter_last_timing():
    """(build_ms, main_ms) of the most recent prefilter launch (library-recorded CUDA events)."""
    b, m = ctypes.c_float(0), ctypes.c_float(0)
    check(lib().galah_b200_prefilter_last_timing(ctypes.byref(b), ctypes.byref(m)))
    return float(b.value), float(m.value)


def prefilter_stream_chunks(chunks=-1):
    """Slices of the pipelined upload of host-buffer prefilter calls (<= 1 disables the pipeline;
    < 0 only queries).  Returns the previous setting."""
    return int(lib().galah_b200_prefilter_stream_chunks(int(chunks)))


def prefilter_last_host_timing():
    """Host wall-clock breakdown (ms) of the last host-buffer prefilter call:
    dict(enqueue, wait, d2h_extra, finish)."""
    ms = (ctypes.c_float * 4)()
    check(lib().galah_b200_prefilter_last_host_timing(ms))
    return dict(zip(("enqueue", "wait", "d2h_extra", "finish"), (float(x) for x in ms)))


def version():
    return lib().galah_b200_version().decode()


def _paths_array(paths):
    arr = (ctypes.c_char_p * len(paths))(*[os.fsencode(p) for p in paths])
    return arr


def _take_pairs(out, n_out):
    n = n_out.value
    res = np.zeros(n, PAIR_DTYPE)
    if n:
        ctypes.memmove(res.ctypes.data, out, n * ctypes.sizeof(Pair))
    lib().galah_b200_free(out)
    return res


def sketch_files(paths, k=21, s=1000, seed=0, threads=0):
    """GPU replacement for finch::sketch_files (reference src/finch.rs:55-69)."""
    n = len(paths)
    table = np.full((n, s), PAD, np.uint64)
    counts = np.zeros(n, np.uint32)
    check(lib().galah_b200_sketch_files(_paths_array(paths), n, k, s, seed, threads,
                                        table.ctypes.data_as(_native.u64p),
                                        counts.ctypes.data_as(_native.u32p)))
    return table, counts


def sketch_packed(seq2, valid, base_off, k=21, s=1000, seed=0):
    seq2 = np.ascontiguousarray(seq2, np.uint32)
    valid = np.ascontiguousarray(valid, np.uint32)
    base_off = np.ascontiguousarray(base_off, np.uint64)
    n = len(base_off) - 1
    table = np.full((n, s), PAD, np.uint64)
    counts = np.zeros(n, np.uint32)
    check(lib().galah_b200_sketch_packed(seq2.ctypes.data_as(_native.u32p),
                                         valid.ctypes.data_as(_native.u32p),
                                         base_off.ctypes.data_as(_native.u64p), n, k, s, seed,
                                         table.ctypes.data_as(_native.u64p),
                                         counts.ctypes.data_as(_native.u32p)))
    return table, counts


def sketch_packed_device(d_seq2, d_valid, d_base_off, n, k, s, seed, d_hashes, d_counts, stream=0):
    """All arguments are raw device pointers (ints); enqueues on `stream`."""
    check(lib().galah_b200_sketch_packed_device(d_seq2, d_valid, d_base_off, n, k, s, seed, d_hashes,
                                                d_counts, stream))


def prefilter(table, counts, k=21, min_ani=0.9, shard=0, n_shards=1):
    """GPU replacement for the pair loop at reference src/finch.rs:75-95 (host buffers).
    With n_shards > 1 only the row blocks shard, shard + n_shards, ... are evaluated."""
    table = np.ascontiguousarray(table, np.uint64)
    counts = np.ascontiguousarray(counts, np.uint32)
    n, stride = table.shape
    out = ctypes.POINTER(Pair)()
    n_out = ctypes.c_size_t(0)
    check(lib().galah_b200_prefilter_shard(table.ctypes.data_as(_native.u64p),
                                           counts.ctypes.data_as(_native.u32p), n, stride, k,
                                           ctypes.c_float(min_ani), shard, n_shards,
                                           ctypes.byref(out), ctypes.byref(n_out)))
    return _take_pairs(out, n_out)


def prefilter_device(d_hashes, d_counts, n, stride, k=21, min_ani=0.9, shard=0, n_shards=1, stream=0):
    out = ctypes.POINTER(Pair)()
    n_out = ctypes.c_size_t(0)
    check(lib().galah_b200_prefilter_device(d_hashes, d_counts, n, stride, k, ctypes.c_float(min_ani),
                                            shard, n_shards, stream, ctypes.byref(out),
                                            ctypes.byref(n_out)))
    return _take_pairs(out, n_out)


def prefilter_enqueue(d_hashes, d_counts, n, stride, k, min_ani, shard, n_shards, mode, stream, d_cand,
                      cand_cap, d_n_cand):
    check(lib().galah_b200_prefilter_enqueue(d_hashes, d_counts, n, stride, k, ctypes.c_float(min_ani),
                                             shard, n_shards, mode, stream, d_cand, cand_cap, d_n_cand))


def finish_candidates(cand, k=21, min_ani=0.9):
    """Device candidates (n x 4 uint32 {i, j, common, total}, already on the host) -> PAIR_DTYPE records
    after the reference's f64 formula / threshold / f32 store, sorted by (i, j)."""
    cand = np.ascontiguousarray(cand, np.uint32).reshape(-1, 4)
    out = ctypes.POINTER(Pair)()
    n_out = ctypes.c_size_t(0)
    check(lib().galah_b200_finish_candidates(cand.ctypes.data, len(cand), k, ctypes.c_float(min_ani),
                                             ctypes.byref(out), ctypes.byref(n_out)))
    return _take_pairs(out, n_out)


def blocklist_layout(n, stride):
    """(n_blocks, entries_per_block, slack) of the block lists of an n x stride sketch table."""
    a, b, c = ctypes.c_size_t(0), ctypes.c_size_t(0), ctypes.c_size_t(0)
    check(lib().galah_b200_blocklist_layout(n, stride, ctypes.byref(a), ctypes.byref(b), ctypes.byref(c)))
    return a.value, b.value, c.value


def blocklist_build(d_hashes, d_counts, n, stride, block_begin, block_end, d_hi, d_lo, d_tags, d_len, stream=0):
    """Build the block lists of blocks [block_begin, block_end) (slice-based device buffers)."""
    check(lib().galah_b200_blocklist_build(d_hashes, d_counts, n, stride, block_begin, block_end, d_hi, d_lo,
                                           d_tags, d_len, stream))


def prefilter_join_items_enqueue(d_hashes, d_counts, n, stride, k, min_ani, d_hi, d_lo, d_tags, d_len, d_items,
                                 n_items, reset_candidates, stream, d_cand, cand_cap, d_n_cand):
    """Join of an explicit device list of (rb, cb) block pairs; candidates are appended."""
    check(lib().galah_b200_prefilter_join_items_enqueue(d_hashes, d_counts, n, stride, k, ctypes.c_float(min_ani),
                                                        d_hi, d_lo, d_tags, d_len, d_items, n_items,
                                                        int(bool(reset_candidates)), stream, d_cand, cand_cap,
                                                        d_n_cand))


def table_max_device(d_hashes, d_counts, n, stride, d_max, stream=0):
    """Largest valid hash of a (slice of a) device sketch table -> device uint64 scalar d_max."""
    check(lib().galah_b200_table_max_device(d_hashes, d_counts, n, stride, d_max, stream))


def blocklist_build_local(d_rows, d_counts, n_rows, stride, d_table_max, n_blocks_out, d_hi, d_lo, d_tags, d_len,
                          stream=0):
    """Block lists of a local slice of whole row blocks, table-wide maximum supplied by the caller."""
    check(lib().galah_b200_blocklist_build_local(d_rows, d_counts, n_rows, stride, d_table_max, n_blocks_out,
                                                 d_hi, d_lo, d_tags, d_len, stream))


def marker_row_capacity(longest_unit, small_genomes=False):
    return int(lib().galah_b200_marker_row_capacity(int(longest_unit), int(bool(small_genomes))))


def prefilter_join_enqueue_screen(d_rows, d_counts, n, stride, faster_small, d_hi, d_lo, d_tags, d_len, shard, n_shards,
                                  stream, d_cand, cand_cap, d_n_cand):
    check(lib().galah_b200_prefilter_join_enqueue_screen(d_rows, d_counts, n, stride, int(bool(faster_small)), d_hi, d_lo,
                                                         d_tags, d_len, shard, n_shards, stream, d_cand, cand_cap, d_n_cand))


def prefilter_join_enqueue(d_hashes, d_counts, n, stride, k, min_ani, d_hi, d_lo, d_tags, d_len, shard, n_shards,
                           stream, d_cand, cand_cap, d_n_cand):
    check(lib().galah_b200_prefilter_join_enqueue(d_hashes, d_counts, n, stride, k, ctypes.c_float(min_ani), d_hi,
                                                  d_lo, d_tags, d_len, shard, n_shards, stream, d_cand, cand_cap,
                                                  d_n_cand))


def finch_distances(paths, min_ani=0.9, num_kmers=1000, kmer_length=21, threads=0):
    """GPU replacement for galah::finch::distances (reference src/finch.rs:48-97)."""
    out = ctypes.POINTER(Pair)()
    n_out = ctypes.c_size_t(0)
    check(lib().galah_b200_finch_distances(_paths_array(paths), len(paths), ctypes.c_float(min_ani),
                                           num_kmers, kmer_length, threads, ctypes.byref(out),
                                           ctypes.byref(n_out)))
    return _take_pairs(out, n_out)


ANI_RESULT_DTYPE = np.dtype([("ani", "<f4"), ("af_query", "<f4"), ("af_ref", "<f4"), ("estimator", "<u4"),
                             ("sum_fx", "<u8"), ("n_chunks", "<u4"), ("sum_m", "<u4"), ("span_m", "<u4"),
                             ("span_n", "<u4"), ("n_chains", "<u4"), ("cov_q", "<u4"), ("cov_r", "<u4"),
                             ("reserved", "<u4")])


class AniIndex:
    """Genomes indexed for stage-2 ANI (FracMinHash seeds + hash tables resident on the GPU).
    GPU replacement for SkaniClusterer::calculate_ani (reference src/skani.rs:689-788)."""

    def __init__(self, small_genomes=False):
        self._h = ctypes.c_void_p()
        check(lib().galah_b200_ani_index_create(int(bool(small_genomes)), ctypes.byref(self._h)))

    def close(self):
        if self._h:
            lib().galah_b200_ani_index_free(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __len__(self):
        return int(lib().galah_b200_ani_index_size(self._h))

    def clear(self):
        """Forget every genome (and detach peers) but keep the device allocations."""
        check(lib().galah_b200_ani_index_clear(self._h))

    def reserve(self, n_total_genomes):
        """Capacity hint after the first batch: the index will hold n_total_genomes like those added."""
        check(lib().galah_b200_ani_index_reserve(self._h, int(n_total_genomes)))

    def add_files(self, paths, threads=0):
        check(lib().galah_b200_ani_index_add_files(self._h, _paths_array(paths), len(paths), threads))

    def add_packed(self, seq2, valid, base_off, contig_off, contig_start, contig_len):
        seq2 = np.ascontiguousarray(seq2, np.uint32); valid = np.ascontiguousarray(valid, np.uint32)
        base_off = np.ascontiguousarray(base_off, np.uint64); contig_off = np.ascontiguousarray(contig_off, np.uint64)
        contig_start = np.ascontiguousarray(contig_start, np.uint32); contig_len = np.ascontiguousarray(contig_len, np.uint32)
        check(lib().galah_b200_ani_index_add_packed(
            self._h, seq2.ctypes.data_as(_native.u32p), valid.ctypes.data_as(_native.u32p),
            base_off.ctypes.data_as(_native.u64p), len(base_off) - 1, contig_off.ctypes.data_as(_native.u64p),
            contig_start.ctypes.data_as(_native.u32p), contig_len.ctypes.data_as(_native.u32p)))

    def add_packed_device(self, d_seq2, d_valid, d_base_off, base_off, lengths, stream=0):
        base_off = np.ascontiguousarray(base_off, np.uint64); lengths = np.ascontiguousarray(lengths, np.uint64)
        check(lib().galah_b200_ani_index_add_packed_device(
            self._h, d_seq2, d_valid, d_base_off, base_off.ctypes.data_as(_native.u64p),
            lengths.ctypes.data_as(_native.u64p), len(lengths), stream))

    def genome(self, g):
        ns, nc, tl = ctypes.c_uint64(0), ctypes.c_uint32(0), ctypes.c_uint64(0)
        check(lib().galah_b200_ani_index_genome(self._h, g, ctypes.byref(ns), ctypes.byref(nc), ctypes.byref(tl)))
        return {"n_seeds": ns.value, "n_chunks": nc.value, "total_len": tl.value}

    def seeds(self, g):
        n = self.genome(g)["n_seeds"]
        ks = np.zeros(max(n, 1), np.uint32); sp = np.zeros(max(n, 1), np.uint32); ch = np.zeros(max(n, 1), np.uint32)
        check(lib().galah_b200_ani_index_seeds(self._h, g, ks.ctypes.data_as(_native.u32p),
                                               sp.ctypes.data_as(_native.u32p), ch.ctypes.data_as(_native.u32p), max(n, 1)))
        return ks[:n], sp[:n], ch[:n]

    def pairs(self, pairs, min_af_pct=15.0, individual_contigs=False):
        """pairs: (n, 2) genome ids, (query, reference) -> ANI_RESULT_DTYPE records (ani as galah
        would parse it).  individual_contigs: the units are records (`skani triangle -i`)."""
        pairs = np.ascontiguousarray(pairs, np.uint32).reshape(-1, 2)
        out = np.zeros(len(pairs), ANI_RESULT_DTYPE)
        if len(pairs):
            check(lib().galah_b200_ani_pairs(self._h, pairs.ctypes.data_as(_native.u32p), len(pairs),
                                             ctypes.c_float(min_af_pct), int(bool(individual_contigs)),
                                             out.ctypes.data_as(ctypes.POINTER(_native.AniResult))))
        return out

    def ingest_packed(self, seq2, valid, base_off, lengths, d_hashes, d_counts, device=False, d_base_off=0):
        """Packed genomes -> K1 sketch rows in the device table (d_hashes / d_counts: device pointers,
        stride 1000) + this index.  device=False: seq2 / valid are host addresses (uploaded in batches
        behind the kernels); True: resident arrays (d_base_off = device copy of base_off).
        Returns (K1 device ms, index-build device ms)."""
        base_off = np.ascontiguousarray(base_off, np.uint64); lengths = np.ascontiguousarray(lengths, np.uint64)
        ms = (ctypes.c_float * 2)()
        check(lib().galah_b200_ingest_packed(int(seq2), int(valid), int(d_base_off), base_off.ctypes.data_as(_native.u64p),
                                             lengths.ctypes.data_as(_native.u64p), len(lengths), int(bool(device)),
                                             int(d_hashes), int(d_counts), self._h, ms))
        return float(ms[0]), float(ms[1])

    def ingest_packed_sparse(self, seq2, invalid_ranges, base_off, lengths, d_hashes, d_counts):
        """ingest_packed from HOST buffers without the validity bitmap: invalid_ranges = (begin, end) uint64 arrays
        of the invalid base ranges (absolute, sorted, disjoint; empty when every base is A/C/G/T)."""
        base_off = np.ascontiguousarray(base_off, np.uint64); lengths = np.ascontiguousarray(lengths, np.uint64)
        ib = np.ascontiguousarray(invalid_ranges[0], np.uint64); ie = np.ascontiguousarray(invalid_ranges[1], np.uint64)
        ms = (ctypes.c_float * 2)()
        check(lib().galah_b200_ingest_packed_sparse(int(seq2), ib.ctypes.data_as(_native.u64p), ie.ctypes.data_as(_native.u64p), len(ib),
                                                    base_off.ctypes.data_as(_native.u64p), lengths.ctypes.data_as(_native.u64p),
                                                    len(lengths), int(d_hashes), int(d_counts), self._h, ms))
        return float(ms[0]), float(ms[1])

    def ingest_packed_markers(self, seq2, valid, base_off, lengths, marker_stride, d_rows, d_counts, device=False,
                              d_base_off=0):
        """As ingest_packed, with FracMinHash marker rows (the skani-style screen's input) of stride marker_stride."""
        base_off = np.ascontiguousarray(base_off, np.uint64); lengths = np.ascontiguousarray(lengths, np.uint64)
        ms = (ctypes.c_float * 2)()
        check(lib().galah_b200_ingest_packed_markers(int(seq2), int(valid), int(d_base_off),
                                                     base_off.ctypes.data_as(_native.u64p), lengths.ctypes.data_as(_native.u64p),
                                                     len(lengths), int(bool(device)), int(marker_stride), int(d_rows),
                                                     int(d_counts), self._h, ms))
        return float(ms[0]), float(ms[1])

    def export_tables(self):
        """(64-byte CUDA IPC handle, table_off uint64[n+1], total_len uint64[n]) of this index's hash tables."""
        n = len(self)
        handle = (ctypes.c_uint8 * 64)()
        to = np.zeros(n + 1, np.uint64); tl = np.zeros(max(n, 1), np.uint64)
        check(lib().galah_b200_ani_index_export_tables(self._h, handle, to.ctypes.data_as(_native.u64p),
                                                       tl.ctypes.data_as(_native.u64p)))
        return bytes(handle), to, tl[:n]

    def attach_peer(self, handle, table_off, total_len):
        """Maps a peer process's hash tables (export_tables of ITS index); returns the id of its first
        genome in this index's numbering (usable as the reference of a pair only)."""
        table_off = np.ascontiguousarray(table_off, np.uint64); total_len = np.ascontiguousarray(total_len, np.uint64)
        h = (ctypes.c_uint8 * 64).from_buffer_copy(handle)
        first = ctypes.c_uint32(0)
        check(lib().galah_b200_ani_index_attach_peer(self._h, h, table_off.ctypes.data_as(_native.u64p),
                                                     total_len.ctypes.data_as(_native.u64p), len(total_len),
                                                     ctypes.byref(first)))
        return int(first.value)

    def last_timing(self):
        b, c = ctypes.c_float(0), ctypes.c_float(0)
        check(lib().galah_b200_ani_last_timing(self._h, ctypes.byref(b), ctypes.byref(c)))
        return float(b.value), float(c.value)


def cluster_from_distances(n_genomes, hits, ani_threshold, calculate_ani=None, skip_clusterer=False):
    """The clustering engine behind galah::clusterer::cluster() (reference src/clusterer.rs:56-151)
    on a precluster hit list.  `hits`: PAIR_DTYPE records (i, j, ani).  `calculate_ani(rep, genome)`
    returns a float or None (ClusterDistanceFinder::calculate_ani on indices).  Returns
    (clusters, info): clusters as lists of genome indices, representative first."""
    hits = np.ascontiguousarray(hits, PAIR_DTYPE)
    calls = {"exc": None}

    def _cb(_ctx, rep, genome, out):
        try:
            v = calculate_ani(int(rep), int(genome))
        except BaseException as e:  # never unwind through C
            calls["exc"] = e
            return 0
        if v is None:
            return 0
        out[0] = float(v)
        return 1

    cb = _native.ANI_FN(_cb) if calculate_ani is not None else _native.ANI_FN()
    res = _native.Clusters()
    rc = lib().galah_b200_cluster_from_distances(int(n_genomes), hits.ctypes.data, len(hits),
                                                 int(bool(skip_clusterer)), ctypes.c_float(ani_threshold),
                                                 cb, None, ctypes.byref(res))
    if calls["exc"] is not None or rc:
        lib().galah_b200_clusters_free(ctypes.byref(res))
        if calls["exc"] is not None:  # the callback's own failure comes first: the engine only saw a None
            raise calls["exc"]
        check(rc)
    return _take_clusters(res)


def cluster_from_distances_batched(n_genomes, hits, ani_threshold, calculate_ani_batch, max_waves=0):
    """The clustering engine with stage 2 asked for in waves: `calculate_ani_batch(reps, genomes)` gets two
    uint32 arrays (reps[x] is the query) and returns one float or None per pair (or a numpy float array when every
    pair has a value).  Same clusters, order and
    ani_calls as cluster_from_distances; info["ani_waves"] = batches asked."""
    hits = np.ascontiguousarray(hits, PAIR_DTYPE)
    calls = {"exc": None}

    def _cb(_ctx, reps, genomes, n, some, ani):
        try:
            r = np.ctypeslib.as_array(reps, shape=(n,)).copy()
            g = np.ctypeslib.as_array(genomes, shape=(n,)).copy()
            vals = calculate_ani_batch(r, g)
            if len(vals) != n:
                raise ValueError("calculate_ani_batch: one value per pair")
            if isinstance(vals, np.ndarray):  # a float array: every pair has a value (no per-element Python work)
                np.ctypeslib.as_array(ani, shape=(n,))[:] = vals.astype(np.float32, copy=False)
                np.ctypeslib.as_array(some, shape=(n,))[:] = 1
            else:
                for x, v in enumerate(vals):
                    some[x] = 0 if v is None else 1
                    ani[x] = 0.0 if v is None else float(v)
        except BaseException as e:  # never unwind through C
            calls["exc"] = e
            return 1
        return 0

    cb = _native.ANI_BATCH_FN(_cb)
    res = _native.Clusters()
    waves = ctypes.c_uint32(0)
    rc = lib().galah_b200_cluster_from_distances_batched(int(n_genomes), hits.ctypes.data, len(hits), ctypes.c_float(ani_threshold),
                                                         cb, None, int(max_waves), ctypes.byref(res), ctypes.byref(waves))
    if calls["exc"] is not None or rc:
        lib().galah_b200_clusters_free(ctypes.byref(res))
        if calls["exc"] is not None:
            raise calls["exc"]
        check(rc)
    clusters, info = _take_clusters(res)
    info["ani_waves"] = int(waves.value)
    return clusters, info


def cluster_lazy(mode):
    """Stage 2 of the one-call pipelines: 0 = every precluster hit up front, 1 / -1 = in waves (default)."""
    check(lib().galah_b200_cluster_lazy(int(mode)))


def cluster_from_ani_table(n_genomes, hits, ani, ani_threshold):
    """The clustering engine with calculate_ani served from a table: ani[x] (percent) belongs to
    hits[x]; hits sorted by (i, j) -- the batched form the Rust shim uses."""
    hits = np.ascontiguousarray(hits, PAIR_DTYPE)
    ani = np.ascontiguousarray(ani, np.float32)
    if len(ani) != len(hits):
        raise ValueError("one ANI value per hit")
    res = _native.Clusters()
    check(lib().galah_b200_cluster_from_ani_table(int(n_genomes), hits.ctypes.data, len(hits), ani.ctypes.data,
                                                  ctypes.c_float(ani_threshold), ctypes.byref(res)))
    return _take_clusters(res)


def cluster_from_ani_tables(n_genomes, hits, ani_fwd, ani_rev, ani_threshold):
    """The engine with both orientations of every hit (ani_fwd: query = hits.i, ani_rev: query = hits.j)."""
    hits = np.ascontiguousarray(hits, PAIR_DTYPE)
    ani_fwd = np.ascontiguousarray(ani_fwd, np.float32); ani_rev = np.ascontiguousarray(ani_rev, np.float32)
    if len(ani_fwd) != len(hits) or len(ani_rev) != len(hits):
        raise ValueError("one ANI value per hit and orientation")
    res = _native.Clusters()
    check(lib().galah_b200_cluster_from_ani_tables(int(n_genomes), hits.ctypes.data, len(hits), ani_fwd.ctypes.data,
                                                   ani_rev.ctypes.data, ctypes.c_float(ani_threshold), ctypes.byref(res)))
    return _take_clusters(res)


class ClusterList:
    """galah's Vec<Vec<usize>> as two arrays (members concatenated, representative first in each
    cluster; offsets): behaves like a read-only list of lists, without building a million Python
    lists unless asked to (`tolist()`)."""

    def __init__(self, members, offsets):
        self.members, self.offsets = members, offsets

    def __len__(self):
        return len(self.offsets) - 1

    def __getitem__(self, x):
        if isinstance(x, slice):
            return [self[y] for y in range(*x.indices(len(self)))]
        if x < 0:
            x += len(self)
        return self.members[self.offsets[x]:self.offsets[x + 1]].tolist()

    def __iter__(self):
        m, o = self.members.tolist(), self.offsets.tolist()
        return (m[o[x]:o[x + 1]] for x in range(len(o) - 1))

    def tolist(self):
        return list(self)

    def __eq__(self, other):
        if isinstance(other, ClusterList):
            return np.array_equal(self.members, other.members) and np.array_equal(self.offsets, other.offsets)
        return self.tolist() == other

    def __repr__(self):
        return repr(self.tolist()) if len(self) <= 64 else f"<ClusterList of {len(self)} clusters, {len(self.members)} members>"

    def sha1(self):
        """Digest of the cluster SET (clusters sorted by their member lists), for comparing runs."""
        import hashlib
        order = sorted(range(len(self)), key=lambda x: self.members[self.offsets[x]:self.offsets[x + 1]].tobytes()) \
            if len(self) < 4096 else np.lexsort((np.diff(self.offsets), self.members[self.offsets[:-1]]))
        h = hashlib.sha1()
        for x in order:
            h.update(self.members[self.offsets[x]:self.offsets[x + 1]].astype(np.uint32).tobytes()); h.update(b"|")
        return h.hexdigest()


def _take_clusters(res):
    try:
        nc = int(res.n_clusters)
        off = np.ctypeslib.as_array(res.offsets, shape=(nc + 1,)).astype(np.int64) if nc else np.zeros(1, np.int64)
        mem = np.ctypeslib.as_array(res.members, shape=(int(off[-1]),)).copy() if nc and off[-1] else np.zeros(0, np.uint32)
        clusters = ClusterList(mem, off)
        info = {"ani_calls": int(res.ani_calls), "n_preclusters": int(res.n_preclusters),
                "largest_precluster": int(res.largest_precluster)}
    finally:
        lib().galah_b200_clusters_free(ctypes.byref(res))
    return clusters, info


def cluster(genomes, precluster_ani=0.9, ani=95.0, min_aligned_fraction=15.0, small_genomes=False, threads=0):
    """galah::clusterer::cluster() with FinchPreclusterer + SkaniClusterer (reference
    src/clusterer.rs:14-152), entirely on the GPU path.  precluster_ani is a FRACTION (finch),
    ani and min_aligned_fraction are PERCENTAGES (skani), as in the reference's structs.
    Returns (clusters, info): clusters as lists of indices into `genomes`, representative first."""
    res = _native.Clusters()
    stats = _native.ClusterStats()
    check(lib().galah_b200_cluster_files(_paths_array(genomes), len(genomes), ctypes.c_float(precluster_ani),
                                         ctypes.c_float(ani), ctypes.c_float(min_aligned_fraction),
                                         int(bool(small_genomes)), threads, ctypes.byref(res), ctypes.byref(stats)))
    clusters, info = _take_clusters(res)
    info.update(_stats_dict(stats))
    return clusters, info


def cluster_multi(genomes, n_devices, precluster_ani=0.9, ani=95.0, min_aligned_fraction=15.0, small_genomes=False,
                  threads=0):
    """cluster() over n_devices GPUs of this process (init_devices(n_devices) first): the path list is cut
    into one slice per device, each read / decoded / sketched / indexed on its device."""
    res = _native.Clusters()
    stats = _native.ClusterStats()
    check(lib().galah_b200_cluster_files_multi(_paths_array(genomes), len(genomes), int(n_devices),
                                               ctypes.c_float(precluster_ani), ctypes.c_float(ani),
                                               ctypes.c_float(min_aligned_fraction), int(bool(small_genomes)), threads,
                                               ctypes.byref(res), ctypes.byref(stats)))
    clusters, info = _take_clusters(res)
    info.update(_stats_dict(stats))
    return clusters, info


def _stats_dict(stats):
    d = {"n_precluster_hits": int(stats.n_precluster_hits), "n_ani_pairs": int(stats.n_ani_pairs),
         "ani_waves": int(stats.ani_waves)}
    for f in ("ani_chain_ms", "ingest_ms", "sketch_ms", "index_ms", "prefilter_ms", "ani_ms", "engine_ms", "total_ms"):
        d[f] = float(getattr(stats, f))
    return d


def cluster_packed_sparse(seq2, invalid_ranges, base_off, lengths, precluster_ani=0.9, ani=95.0, min_aligned_fraction=15.0,
                          small_genomes=False):
    """cluster_packed on HOST buffers without the validity bitmap (a third of the bytes stays off PCIe):
    invalid_ranges = (begin, end) uint64 arrays of the invalid base ranges, absolute, sorted, disjoint."""
    base_off = np.ascontiguousarray(base_off, np.uint64); lengths = np.ascontiguousarray(lengths, np.uint64)
    ib = np.ascontiguousarray(invalid_ranges[0], np.uint64); ie = np.ascontiguousarray(invalid_ranges[1], np.uint64)
    res = _native.Clusters()
    stats = _native.ClusterStats()
    ptr = lambda a: a.ctypes.data if isinstance(a, np.ndarray) else int(a)
    check(lib().galah_b200_cluster_packed_sparse(ptr(seq2), ib.ctypes.data_as(_native.u64p), ie.ctypes.data_as(_native.u64p), len(ib),
                                                 base_off.ctypes.data_as(_native.u64p), lengths.ctypes.data_as(_native.u64p),
                                                 len(lengths), ctypes.c_float(precluster_ani), ctypes.c_float(ani),
                                                 ctypes.c_float(min_aligned_fraction), int(bool(small_genomes)),
                                                 ctypes.byref(res), ctypes.byref(stats)))
    clusters, info = _take_clusters(res)
    info.update(_stats_dict(stats))
    return clusters, info


def cluster_packed(seq2, valid, base_off, lengths, precluster_ani=0.9, ani=95.0, min_aligned_fraction=15.0,
                   small_genomes=False, device=False, d_base_off=None):
    """cluster() on genomes that are already packed (K1 / K3 layout, one contig per genome).
    device=False: seq2 / valid are host arrays (numpy uint32 or integer addresses of pinned buffers),
    uploaded in batches behind the kernels of the previous batch.  device=True: seq2 / valid /
    d_base_off are device pointers (ints).  base_off / lengths: host uint64 arrays."""
    base_off = np.ascontiguousarray(base_off, np.uint64); lengths = np.ascontiguousarray(lengths, np.uint64)
    res = _native.Clusters()
    stats = _native.ClusterStats()
    ptr = lambda a: a.ctypes.data if isinstance(a, np.ndarray) else int(a)
    common = (base_off.ctypes.data_as(_native.u64p), lengths.ctypes.data_as(_native.u64p), len(lengths),
              ctypes.c_float(precluster_ani), ctypes.c_float(ani), ctypes.c_float(min_aligned_fraction),
              int(bool(small_genomes)), ctypes.byref(res), ctypes.byref(stats))
    if device:
        check(lib().galah_b200_cluster_packed_device(ptr(seq2), ptr(valid), int(d_base_off), *common))
    else:
        check(lib().galah_b200_cluster_packed(ptr(seq2), ptr(valid), *common))
    clusters, info = _take_clusters(res)
    info.update(_stats_dict(stats))
    return clusters, info


def init_devices(n_devices):
    """Binds devices 0 .. n_devices-1 of THIS process for cluster_packed_multi (peer access opened)."""
    check(lib().galah_b200_init_devices(int(n_devices)))


def cluster_packed_multi(seq2, valid, base_off, lengths, n_devices, precluster_ani=0.9, ani=95.0,
                         min_aligned_fraction=15.0, small_genomes=False):
    """cluster() on packed genomes in HOST arrays over n_devices GPUs of this process (one host thread
    per device inside the library; init_devices(n_devices) first).  Same result as cluster_packed."""
    base_off = np.ascontiguousarray(base_off, np.uint64); lengths = np.ascontiguousarray(lengths, np.uint64)
    res = _native.Clusters()
    stats = _native.ClusterStats()
    ptr = lambda a: a.ctypes.data if isinstance(a, np.ndarray) else int(a)
    check(lib().galah_b200_cluster_packed_multi(ptr(seq2), ptr(valid), base_off.ctypes.data_as(_native.u64p),
                                                lengths.ctypes.data_as(_native.u64p), len(lengths), int(n_devices),
                                                ctypes.c_float(precluster_ani), ctypes.c_float(ani),
                                                ctypes.c_float(min_aligned_fraction), int(bool(small_genomes)),
                                                ctypes.byref(res), ctypes.byref(stats)))
    clusters, info = _take_clusters(res)
    info.update(_stats_dict(stats))
    return clusters, info


def skani_distances(paths, threshold=90.0, min_aligned_fraction=15.0, small_genomes=False, contigs=False, threads=0):
    """GPU replacement for SkaniPreclusterer::distances / distances_contigs (reference
    src/skani.rs:21-56).  Returns (PAIR_DTYPE hits with ANI in percent, number of units)."""
    out = ctypes.POINTER(Pair)()
    n_out, n_units = ctypes.c_size_t(0), ctypes.c_size_t(0)
    check(lib().galah_b200_skani_distances(_paths_array(paths), len(paths), ctypes.c_float(threshold),
                                           ctypes.c_float(min_aligned_fraction), int(bool(small_genomes)),
                                           int(bool(contigs)), threads, ctypes.byref(out), ctypes.byref(n_out),
                                           ctypes.byref(n_units)))
    return _take_pairs(out, n_out), int(n_units.value)


def cluster_skani(genomes, precluster_ani=90.0, ani=95.0, min_aligned_fraction=15.0, small_genomes=False,
                  cluster_contigs=False, threads=0):
    """galah::clusterer::cluster() with SkaniPreclusterer + SkaniClusterer (the CLI default), or contig
    clustering with cluster_contigs=True.  All thresholds are PERCENTAGES."""
    res = _native.Clusters()
    stats = _native.ClusterStats()
    check(lib().galah_b200_cluster_files_skani(_paths_array(genomes), len(genomes), ctypes.c_float(precluster_ani),
                                               ctypes.c_float(ani), ctypes.c_float(min_aligned_fraction),
                                               int(bool(small_genomes)), int(bool(cluster_contigs)), threads,
                                               ctypes.byref(res), ctypes.byref(stats)))
    clusters, info = _take_clusters(res)
    info.update(n_precluster_hits=int(stats.n_precluster_hits))
    return clusters, info


def skani_distances_multi(paths, n_devices, threshold=90.0, min_aligned_fraction=15.0, small_genomes=False, contigs=False,
                          threads=0):
    """skani_distances over n_devices GPUs of this process (init_devices first): device r ingests the r-th slice of
    the path list.  Same hit list as the single-GPU call."""
    out = ctypes.POINTER(Pair)()
    n_out, n_units = ctypes.c_size_t(0), ctypes.c_size_t(0)
    check(lib().galah_b200_skani_distances_multi(_paths_array(paths), len(paths), int(n_devices), ctypes.c_float(threshold),
                                                 ctypes.c_float(min_aligned_fraction), int(bool(small_genomes)),
                                                 int(bool(contigs)), threads, ctypes.byref(out), ctypes.byref(n_out),
                                                 ctypes.byref(n_units)))
    return _take_pairs(out, n_out), int(n_units.value)


def cluster_skani_multi(genomes, n_devices, precluster_ani=90.0, ani=95.0, min_aligned_fraction=15.0, small_genomes=False,
                        cluster_contigs=False, threads=0):
    """cluster_skani over n_devices GPUs of this process (init_devices first)."""
    res = _native.Clusters()
    stats = _native.ClusterStats()
    check(lib().galah_b200_cluster_files_skani_multi(_paths_array(genomes), len(genomes), int(n_devices),
                                                     ctypes.c_float(precluster_ani), ctypes.c_float(ani),
                                                     ctypes.c_float(min_aligned_fraction), int(bool(small_genomes)),
                                                     int(bool(cluster_contigs)), threads, ctypes.byref(res), ctypes.byref(stats)))
    clusters, info = _take_clusters(res)
    info.update(n_precluster_hits=int(stats.n_precluster_hits))
    return clusters, info


def pack_fasta_file(path):
    """Host ingest of one file -> (codes uint8 per base: 0..3 = ACGT, 4 = invalid, rec_start, rec_end) in
    packed coordinates (unpacked here for comparison with the oracle's load_codes)."""
    seq2, valid = _native.u32p(), _native.u32p()
    rs, re_ = _native.u64p(), _native.u64p()
    nb, nr = ctypes.c_uint64(0), ctypes.c_size_t(0)
    check(lib().galah_b200_pack_fasta_file(os.fsencode(path), ctypes.byref(seq2), ctypes.byref(valid), ctypes.byref(nb),
                                           ctypes.byref(rs), ctypes.byref(re_), ctypes.byref(nr)))
    try:
        n = nb.value
        w2 = np.ctypeslib.as_array(seq2, shape=((n + 15) // 16 + 1,)).astype(np.uint32)
        wv = np.ctypeslib.as_array(valid, shape=((n + 31) // 32 + 1,)).astype(np.uint32)
        idx = np.arange(n, dtype=np.uint64)
        codes = ((w2[idx >> np.uint64(4)] >> (np.uint32(2) * (idx & np.uint64(15)).astype(np.uint32))) & np.uint32(3)).astype(np.uint8)
        ok = ((wv[idx >> np.uint64(5)] >> (idx & np.uint64(31)).astype(np.uint32)) & np.uint32(1)).astype(bool)
        codes[~ok] = 4
        starts = np.ctypeslib.as_array(rs, shape=(max(nr.value, 1),))[: nr.value].copy()
        ends = np.ctypeslib.as_array(re_, shape=(max(nr.value, 1),))[: nr.value].copy()
    finally:
        for ptr in (seq2, valid, rs, re_):
            lib().galah_b200_free(ptr)
    return codes, starts, ends


def skani_distances_packed_device(d_seq2, d_valid, d_base_off, base_off, lengths, threshold=95.0,
                                  min_aligned_fraction=15.0, small_genomes=False, stream=0,
                                  individual_contigs=True):
    """SkaniPreclusterer on units already packed on the device (contig mode at scale).  Returns
    (PAIR_DTYPE hits with ani in PERCENT, info dict with n_screened and the stage times in ms)."""
    base_off = np.ascontiguousarray(base_off, np.uint64); lengths = np.ascontiguousarray(lengths, np.uint64)
    out = ctypes.POINTER(Pair)()
    n_out = ctypes.c_size_t(0)
    scr = ctypes.c_uint64(0)
    ms = (ctypes.c_float * 5)()
    check(lib().galah_b200_skani_distances_packed_device(
        d_seq2, d_valid, d_base_off, base_off.ctypes.data_as(_native.u64p), lengths.ctypes.data_as(_native.u64p),
        len(lengths), ctypes.c_float(threshold), ctypes.c_float(min_aligned_fraction), int(bool(small_genomes)),
        int(bool(individual_contigs)), stream, ctypes.byref(out), ctypes.byref(n_out), ctypes.byref(scr), ms))
    info = {"n_screened": int(scr.value),
            **dict(zip(("index_ms", "markers_ms", "screen_ms", "ani_ms", "total_ms"), (float(x) for x in ms)))}
    return _take_pairs(out, n_out), info


def skani_distances_packed_multi(seq2, valid, base_off, lengths, n_devices, threshold=95.0, min_aligned_fraction=15.0,
                                 small_genomes=False, individual_contigs=True):
    """SkaniPreclusterer on packed units in HOST arrays over n_devices GPUs of this process
    (init_devices first).  Returns (PAIR_DTYPE hits, {"n_screened": ...}), identical to the single-GPU call."""
    base_off = np.ascontiguousarray(base_off, np.uint64); lengths = np.ascontiguousarray(lengths, np.uint64)
    out = ctypes.POINTER(Pair)()
    n_out = ctypes.c_size_t(0)
    scr = ctypes.c_uint64(0)
    ptr = lambda a: a.ctypes.data if isinstance(a, np.ndarray) else int(a)
    check(lib().galah_b200_skani_distances_packed_multi(
        ptr(seq2), ptr(valid), base_off.ctypes.data_as(_native.u64p), lengths.ctypes.data_as(_native.u64p), len(lengths),
        int(n_devices), ctypes.c_float(threshold), ctypes.c_float(min_aligned_fraction), int(bool(small_genomes)),
        int(bool(individual_contigs)), ctypes.byref(out), ctypes.byref(n_out), ctypes.byref(scr)))
    return _take_pairs(out, n_out), {"n_screened": int(scr.value)}


def device_ingest(enable=-1):
    """K0 switch: decode FASTA bytes on the device (1, default) or pack on host threads (0);
    < 0 only queries.  Returns the previous setting."""
    return int(lib().galah_b200_device_ingest(int(enable)))


def decode_fasta_device(files, unpack=True):
    """K0 parity hook: a list of in-memory FASTA files (bytes) -> per file (codes uint8 per base:
    0..3 = ACGT, 4 = invalid; rec_start; rec_end; n_ambiguous; n_N), decoded on the device, plus the
    device time in ms.  unpack=False skips the per-base unpacking (timing runs)."""
    n = len(files)
    arr = (ctypes.c_char_p * n)(*files)
    lens = (ctypes.c_size_t * n)(*[len(f) for f in files])
    seq2, valid = _native.u32p(), _native.u32p()
    ptrs = [_native.u64p() for _ in range(7)]
    ms = ctypes.c_float(0)
    check(lib().galah_b200_decode_fasta_device(arr, lens, n, ctypes.byref(seq2), ctypes.byref(valid),
                                               *[ctypes.byref(p) for p in ptrs], ctypes.byref(ms)))
    base_off, n_bases, rec_off, rec_start, rec_end, n_amb, n_N = ptrs
    try:
        bo = np.ctypeslib.as_array(base_off, shape=(n + 1,)).copy()
        nb = np.ctypeslib.as_array(n_bases, shape=(max(n, 1),))[:n].copy()
        ro = np.ctypeslib.as_array(rec_off, shape=(n + 1,)).copy()
        nrec = int(ro[-1])
        rs = np.ctypeslib.as_array(rec_start, shape=(max(nrec, 1),))[:nrec].copy()
        re_ = np.ctypeslib.as_array(rec_end, shape=(max(nrec, 1),))[:nrec].copy()
        amb = np.ctypeslib.as_array(n_amb, shape=(max(n, 1),))[:n].copy()
        nn = np.ctypeslib.as_array(n_N, shape=(max(n, 1),))[:n].copy()
        total = int(bo[-1])
        w2 = np.ctypeslib.as_array(seq2, shape=(total // 16 + 4,)).astype(np.uint32)
        wv = np.ctypeslib.as_array(valid, shape=(total // 32 + 4,)).astype(np.uint32)
        out = []
        for f in range(n):
            if not unpack:
                out.append({"n_bases": int(nb[f]), "n_records": int(ro[f + 1] - ro[f]), "n_ambiguous": int(amb[f]),
                            "n_N": int(nn[f])})
                continue
            idx = np.arange(int(bo[f]), int(bo[f]) + int(nb[f]), dtype=np.uint64)
            codes = ((w2[idx >> np.uint64(4)] >> (np.uint32(2) * (idx & np.uint64(15)).astype(np.uint32))) & np.uint32(3)).astype(np.uint8)
            ok = ((wv[idx >> np.uint64(5)] >> (idx & np.uint64(31)).astype(np.uint32)) & np.uint32(1)).astype(bool)
            codes[~ok] = 4
            # nothing may be set in the padding up to the next multiple of 128
            pad = np.arange(int(bo[f]) + int(nb[f]), int(bo[f + 1]), dtype=np.uint64)
            pad_bits = (wv[pad >> np.uint64(5)] >> (pad & np.uint64(31)).astype(np.uint32)) & np.uint32(1)
            out.append({"codes": codes, "rec_start": rs[int(ro[f]):int(ro[f + 1])], "rec_end": re_[int(ro[f]):int(ro[f + 1])],
                        "n_ambiguous": int(amb[f]), "n_N": int(nn[f]), "padding_clean": not pad_bits.any()})
    finally:
        for ptr in [seq2, valid] + ptrs:
            lib().galah_b200_free(ptr)
    return out, float(ms.value)


GENOME_STATS_DTYPE = np.dtype([("num_contigs", "<u8"), ("num_ambiguous_bases", "<u8"), ("n50", "<u8")])


def genome_stats(paths, threads=0):
    """galah::genome_stats::calculate_genome_stats (reference src/genome_stats.rs:11-51) for each path
    (host side of the ingest pass; needs no device)."""
    out = np.zeros(len(paths), GENOME_STATS_DTYPE)
    check(lib().galah_b200_genome_stats(_paths_array(paths), len(paths), threads, out.ctypes.data))
    return out


def synth_layout(n, length):
    """Sizes (in uint32 / uint64 elements) of the packed buffers for n synthetic genomes."""
    padded = (length + 127) // 128 * 128
    return {"seq2_words": n * padded // 16 + 4, "valid_words": n * padded // 32 + 4,
            "base_off": n + 1, "padded": padded}


def stream():
    """The CUDA stream (integer handle) the library enqueues its work on."""
    return int(lib().galah_b200_stream() or 0)


def synth_packed_device_ex(seed, index_begin, n, length, family_size, rate_shift, d_seq2, d_valid, d_base_off, stream=0):
    check(lib().galah_b200_synth_packed_device_ex(seed, index_begin, n, length, family_size, rate_shift, d_seq2,
                                                  d_valid, d_base_off, stream))


def synth_packed_device(seed, index_begin, n, length, d_seq2, d_valid, d_base_off, stream=0):
    check(lib().galah_b200_synth_packed_device(seed, index_begin, n, length, d_seq2, d_valid,
                                               d_base_off, stream))


# ------------------------------------------------------------------------------------------------
# Host-side mirror of the reference's plugin interface (src/lib.rs:29-55, src/finch.rs:4-46,
# src/skani.rs:12-74, 689-716, src/clusterer.rs:14-21): same names, argument meaning and error
# behaviour, over the session entry points of the C ABI.  `cluster_with` below drives them in the
# order the reference's cluster() does, so the parity tests read like the reference's own tests.
# ------------------------------------------------------------------------------------------------
class Session:
    """State shared by the two trait objects of one cluster() call (include/galah_b200.h)."""

    def __init__(self):
        self._h = ctypes.c_void_p()
        check(lib().galah_b200_session_create(ctypes.byref(self._h)))

    def close(self):
        if self._h:
            lib().galah_b200_session_free(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def stats(self):
        a, b, c = ctypes.c_uint64(0), ctypes.c_uint64(0), ctypes.c_uint64(0)
        check(lib().galah_b200_session_stats(self._h, ctypes.byref(a), ctypes.byref(b), ctypes.byref(c)))
        return {"n_indexed": a.value, "n_pairs_computed": b.value, "n_launches": c.value}


def _session_pairs(fn, *args):
    out = ctypes.POINTER(Pair)()
    n_out = ctypes.c_size_t(0)
    check(fn(*args, ctypes.byref(out), ctypes.byref(n_out)))
    return _take_pairs(out, n_out)


class FinchPreclusterer:
    """galah::finch::FinchPreclusterer { min_ani (FRACTION), num_kmers, kmer_length, low_memory }."""

    def __init__(self, min_ani=0.9, num_kmers=1000, kmer_length=21, low_memory=False, session=None, threads=0):
        self.min_ani, self.num_kmers, self.kmer_length, self.low_memory = min_ani, num_kmers, kmer_length, low_memory
        self.session, self.threads = session or Session(), threads

    def method_name(self):
        return lib().galah_b200_finch_method_name().decode()

    def distances(self, genome_fasta_paths):
        return _session_pairs(lib().galah_b200_session_finch_distances, self.session._h, _paths_array(genome_fasta_paths),
                              len(genome_fasta_paths), ctypes.c_float(self.min_ani), self.num_kmers, self.kmer_length,
                              int(bool(self.low_memory)), self.threads)

    def distances_contigs(self, genome_fasta_paths, contig_names):
        return _session_pairs(lib().galah_b200_session_finch_distances_contigs, self.session._h,
                              _paths_array(genome_fasta_paths), len(genome_fasta_paths), _paths_array(contig_names),
                              len(contig_names))

    def distances_with_references(self, genome_fasta_paths, reference_genomes):
        return _session_pairs(lib().galah_b200_session_finch_distances_with_references, self.session._h,
                              _paths_array(genome_fasta_paths), len(genome_fasta_paths), _paths_array(reference_genomes),
                              len(reference_genomes))


class SkaniPreclusterer:
    """galah::skani::SkaniPreclusterer { threshold (PERCENT), min_aligned_threshold (FRACTION),
    small_genomes, threads, low_memory }."""

    def __init__(self, threshold=90.0, min_aligned_threshold=0.15, small_genomes=False, threads=0, low_memory=False,
                 session=None):
        self.threshold, self.min_aligned_threshold = threshold, min_aligned_threshold
        self.small_genomes, self.threads, self.low_memory = small_genomes, threads, low_memory
        self.session = session or Session()

    def method_name(self):
        return lib().galah_b200_skani_method_name().decode()

    def distances(self, genome_fasta_paths):
        return _session_pairs(lib().galah_b200_session_skani_distances, self.session._h, _paths_array(genome_fasta_paths),
                              len(genome_fasta_paths), ctypes.c_float(self.threshold),
                              ctypes.c_float(self.min_aligned_threshold), int(bool(self.small_genomes)),
                              int(bool(self.low_memory)), self.threads)

    def distances_contigs(self, genome_fasta_paths, contig_names):
        return _session_pairs(lib().galah_b200_session_skani_distances_contigs, self.session._h,
                              _paths_array(genome_fasta_paths), len(genome_fasta_paths), _paths_array(contig_names),
                              len(contig_names), ctypes.c_float(self.threshold), ctypes.c_float(self.min_aligned_threshold),
                              int(bool(self.small_genomes)), self.threads)

    def distances_with_references(self, genome_fasta_paths, reference_genomes):
        return _session_pairs(lib().galah_b200_session_skani_distances_with_references, self.session._h,
                              _paths_array(genome_fasta_paths), len(genome_fasta_paths), _paths_array(reference_genomes),
                              len(reference_genomes), ctypes.c_float(self.threshold),
                              ctypes.c_float(self.min_aligned_threshold), int(bool(self.small_genomes)), self.threads)


class SkaniClusterer:
    """galah::skani::SkaniClusterer { threshold (PERCENT), min_aligned_threshold (FRACTION), small_genomes }."""

    def __init__(self, threshold=95.0, min_aligned_threshold=0.15, small_genomes=False, session=None):
        self.threshold, self.min_aligned_threshold, self.small_genomes = threshold, min_aligned_threshold, small_genomes
        self.session = session or Session()

    def initialise(self):
        if not self.threshold > 1.0:  # assert!(self.threshold > 1.0), src/skani.rs:696-698
            raise GalahB200Error(-1, "assertion failed: self.threshold > 1.0")
        check(lib().galah_b200_session_set_clusterer(self.session._h, int(bool(self.small_genomes))))

    def method_name(self):
        return lib().galah_b200_skani_method_name().decode()

    def get_ani_threshold(self):
        return self.threshold

    def calculate_ani(self, fasta1, fasta2):
        ani, some = ctypes.c_float(0), ctypes.c_int(0)
        check(lib().galah_b200_session_calculate_ani(self.session._h, os.fsencode(fasta1), os.fsencode(fasta2),
                                                     ctypes.c_float(self.min_aligned_threshold),
                                                     int(bool(self.small_genomes)), ctypes.byref(ani), ctypes.byref(some)))
        return float(ani.value) if some.value else None


def contig_names(paths):
    """Record names as `galah cluster --cluster-contigs` collects them (duplicates raise the reference's panic text)."""
    out = ctypes.POINTER(ctypes.c_char_p)()
    n = ctypes.c_size_t(0)
    check(lib().galah_b200_contig_names(_paths_array(paths), len(paths), ctypes.byref(out), ctypes.byref(n)))
    try:
        return [out[x].decode() for x in range(n.value)]
    finally:
        lib().galah_b200_contig_names_free(out, n.value)


def cluster_with(genomes, preclusterer, clusterer, cluster_contigs=False, contig_names=None, reference_genomes=None):
    """galah::clusterer::cluster(genomes, &preclusterer, &clusterer, cluster_contigs, contig_names,
    reference_genomes) (src/clusterer.rs:14-152), driving the trait objects above exactly as the
    reference does: initialise, method names -> skip_clusterer, one distances* call, then the greedy
    engine calling clusterer.calculate_ani(representative path, genome path)."""
    clusterer.initialise()
    pre_name, cl_name = preclusterer.method_name(), clusterer.method_name()
    skip = cl_name == pre_name
    if cluster_contigs:
        if pre_name == "finch":
            raise GalahB200Error(-1, f"{pre_name} does not support contig comparisons.")  # src/clusterer.rs:39-41
        skip = True
    if reference_genomes is not None:
        cache = preclusterer.distances_with_references(genomes, reference_genomes)
    elif cluster_contigs:
        cache = preclusterer.distances_contigs(genomes, contig_names)
    else:
        cache = preclusterer.distances(genomes)
    names = contig_names if cluster_contigs else genomes
    fn = None if skip else (lambda rep, g: clusterer.calculate_ani(names[rep], names[g]))
    clusters, info = cluster_from_distances(len(names), cache, clusterer.get_ani_threshold(), fn, skip_clusterer=skip)
    return clusters, info
