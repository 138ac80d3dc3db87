"""ctypes binding to oracle/liboracle.so (test infrastructure; see oracle/__init__.py)."""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liboracle.so")

PAIR_DTYPE = np.dtype(
    [("i", "<u4"), ("j", "<u4"), ("common", "<u4"), ("total", "<u4"), ("ani", "<f4")]
)


def build(force=False):
    """Compile liboracle.so with the committed Makefile (gcc; a few seconds)."""
    srcs = [os.path.join(_HERE, f) for f in os.listdir(_HERE) if f.endswith(".c")]
    stale = force or not os.path.exists(_LIB_PATH) or any(
        os.path.getmtime(s) > os.path.getmtime(_LIB_PATH) for s in srcs
    )
    if stale:
        subprocess.check_call(["make", "-C", _HERE, "-B", "liboracle.so"], stdout=subprocess.DEVNULL)
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build()
        L = ctypes.CDLL(_LIB_PATH)
        u64p = ctypes.POINTER(ctypes.c_uint64)
        u32p = ctypes.POINTER(ctypes.c_uint32)
        u8p = ctypes.POINTER(ctypes.c_uint8)
        L.oracle_murmur3_x64_128_h1.restype = ctypes.c_uint64
        L.oracle_murmur3_x64_128_h1.argtypes = [ctypes.c_char_p, ctypes.c_int, ctypes.c_uint64]
        L.oracle_sketch_fasta.restype = ctypes.c_int
        L.oracle_sketch_fasta.argtypes = [ctypes.c_char_p, ctypes.c_int, ctypes.c_uint32,
                                          ctypes.c_uint64, u64p, u32p]
        L.oracle_sketch_records.restype = ctypes.c_int
        L.oracle_sketch_records.argtypes = [u8p, u64p, ctypes.c_uint32, ctypes.c_int,
                                            ctypes.c_uint32, ctypes.c_uint64, u64p, u32p]
        L.oracle_raw_distance.restype = None
        L.oracle_raw_distance.argtypes = [u64p, ctypes.c_uint32, u64p, ctypes.c_uint32, u64p, u64p]
        L.oracle_mash_ani.restype = ctypes.c_double
        L.oracle_mash_ani.argtypes = [ctypes.c_uint64, ctypes.c_uint64, ctypes.c_int]
        L.oracle_prefilter.restype = ctypes.c_size_t
        L.oracle_prefilter.argtypes = [u64p, u32p, ctypes.c_size_t, ctypes.c_size_t, ctypes.c_int,
                                       ctypes.c_float, ctypes.c_size_t, ctypes.c_size_t,
                                       ctypes.c_void_p, ctypes.c_size_t]
        L.oracle_prefilter_count_mt.restype = ctypes.c_size_t
        L.oracle_prefilter_count_mt.argtypes = [u64p, u32p, ctypes.c_size_t, ctypes.c_size_t,
                                                ctypes.c_int, ctypes.c_float, ctypes.c_size_t,
                                                ctypes.c_size_t]
        L.oracle_sketch_files_mt.restype = ctypes.c_int
        L.oracle_sketch_files_mt.argtypes = [ctypes.POINTER(ctypes.c_char_p), ctypes.c_size_t,
                                             ctypes.c_int, ctypes.c_uint32, ctypes.c_uint64, u64p, u32p]
        L.oracle_synth_block.restype = ctypes.c_uint64
        L.oracle_synth_block.argtypes = [ctypes.c_uint64, ctypes.c_uint64, ctypes.c_uint64]
        L.oracle_synth_genome.restype = None
        L.oracle_synth_genome.argtypes = [ctypes.c_uint64, ctypes.c_uint64, ctypes.c_uint64, u8p]
        L.oracle_sketch_synth_mt.restype = ctypes.c_int
        L.oracle_sketch_synth_mt.argtypes = [ctypes.c_uint64, ctypes.c_uint64, ctypes.c_uint64,
                                             ctypes.c_uint64, ctypes.c_int, ctypes.c_uint32,
                                             ctypes.c_uint64, u64p, u32p]
        _lib = L
    return _lib


def _p(a, ct):
    return a.ctypes.data_as(ctypes.POINTER(ct))


def murmur3_h1(data: bytes, seed: int = 0) -> int:
    return int(lib().oracle_murmur3_x64_128_h1(data, len(data), seed))


def sketch_fasta(path, k=21, s=1000, seed=0):
    """finch::sketch_files for one path (src/finch.rs:55-69). Returns ascending uint64 array."""
    out = np.zeros(s, np.uint64)
    cnt = ctypes.c_uint32(0)
    rc = lib().oracle_sketch_fasta(os.fsencode(path), k, s, seed, _p(out, ctypes.c_uint64),
                                   ctypes.byref(cnt))
    if rc:
        raise RuntimeError(f"oracle_sketch_fasta({path}) failed rc={rc}")
    return out[: cnt.value].copy()


def sketch_records(records, k=21, s=1000, seed=0):
    """Sketch a list of RAW record byte strings pooled into one sketch."""
    seq = np.frombuffer(b"".join(records) or b"\0", np.uint8).copy()
    off = np.zeros(len(records) + 1, np.uint64)
    off[1:] = np.cumsum([len(r) for r in records])
    out = np.zeros(s, np.uint64)
    cnt = ctypes.c_uint32(0)
    rc = lib().oracle_sketch_records(_p(seq, ctypes.c_uint8), _p(off, ctypes.c_uint64),
                                     len(records), k, s, seed, _p(out, ctypes.c_uint64),
                                     ctypes.byref(cnt))
    if rc:
        raise RuntimeError(f"oracle_sketch_records failed rc={rc}")
    return out[: cnt.value].copy()


def raw_distance(a, b):
    a = np.ascontiguousarray(a, np.uint64)
    b = np.ascontiguousarray(b, np.uint64)
    c = ctypes.c_uint64(0)
    t = ctypes.c_uint64(0)
    lib().oracle_raw_distance(_p(a, ctypes.c_uint64), len(a), _p(b, ctypes.c_uint64), len(b),
                              ctypes.byref(c), ctypes.byref(t))
    return c.value, t.value


def mash_ani(common, total, k=21):
    return float(lib().oracle_mash_ani(common, total, k))


def pack_table(sketches, s):
    """List of ascending uint64 arrays -> (n*s table padded with 2^64-1, counts)."""
    n = len(sketches)
    table = np.full((n, s), np.uint64(0xFFFFFFFFFFFFFFFF), np.uint64)
    counts = np.zeros(n, np.uint32)
    for g, sk in enumerate(sketches):
        table[g, : len(sk)] = sk
        counts[g] = len(sk)
    return table, counts


def prefilter(table, counts, k=21, min_ani=0.9, row_begin=0, row_end=None, cap=None):
    """src/finch.rs:75-95 on a sketch table. Returns structured array sorted by (i, j)."""
    table = np.ascontiguousarray(table, np.uint64)
    counts = np.ascontiguousarray(counts, np.uint32)
    n, stride = table.shape
    row_end = n if row_end is None else row_end
    cap = cap if cap is not None else max(1024, 64 * n)
    while True:
        out = np.zeros(cap, PAIR_DTYPE)
        got = lib().oracle_prefilter(_p(table, ctypes.c_uint64), _p(counts, ctypes.c_uint32), n,
                                     stride, k, ctypes.c_float(min_ani), row_begin, row_end,
                                     out.ctypes.data_as(ctypes.c_void_p), cap)
        if got <= cap:
            return out[:got]
        cap = got


def prefilter_count_mt(table, counts, k=21, min_ani=0.9, row_begin=0, row_end=None):
    table = np.ascontiguousarray(table, np.uint64)
    counts = np.ascontiguousarray(counts, np.uint32)
    n, stride = table.shape
    row_end = n if row_end is None else row_end
    return int(lib().oracle_prefilter_count_mt(_p(table, ctypes.c_uint64),
                                               _p(counts, ctypes.c_uint32), n, stride, k,
                                               ctypes.c_float(min_ani), row_begin, row_end))


def synth_genome(seed, index, length):
    out = np.zeros(length, np.uint8)
    lib().oracle_synth_genome(seed, index, length, _p(out, ctypes.c_uint8))
    return out.tobytes()


def synth_genome_ex(seed, index, length, family_size, rate_shift):
    L = lib()
    L.oracle_synth_genome_ex.restype = None
    L.oracle_synth_genome_ex.argtypes = [ctypes.c_uint64, ctypes.c_uint64, ctypes.c_uint64, ctypes.c_uint32,
                                         ctypes.c_uint32, ctypes.POINTER(ctypes.c_uint8)]
    out = np.zeros(length, np.uint8)
    L.oracle_synth_genome_ex(seed, index, length, family_size, rate_shift, _p(out, ctypes.c_uint8))
    return out.tobytes()


def synth_block(seed, index, block):
    return int(lib().oracle_synth_block(seed, index, block))


def sketch_synth(seed, index_begin, n, length, k=21, s=1000, hash_seed=0):
    """Sketches of synthetic genomes index_begin..index_begin+n (all host threads)."""
    table = np.full((n, s), np.uint64(0xFFFFFFFFFFFFFFFF), np.uint64)
    counts = np.zeros(n, np.uint32)
    lib().oracle_sketch_synth_mt(seed, index_begin, n, length, k, s, hash_seed,
                                 _p(table, ctypes.c_uint64), _p(counts, ctypes.c_uint32))
    return table, counts


# ---------------------------------------------------------------------------------------------
# stage 2 (skani_oracle.c)
# ---------------------------------------------------------------------------------------------
def _skani_sigs():
    L = lib()
    if getattr(L, "_skani_ready", False):
        return L
    u64p = ctypes.POINTER(ctypes.c_uint64)
    u32p = ctypes.POINTER(ctypes.c_uint32)
    u8p = ctypes.POINTER(ctypes.c_uint8)
    L.oracle_load_codes.restype = ctypes.c_int
    L.oracle_load_codes.argtypes = [ctypes.c_char_p, ctypes.POINTER(u8p), u64p, ctypes.POINTER(u64p),
                                    ctypes.POINTER(u64p), u32p]
    L.oracle_free.restype = None
    L.oracle_free.argtypes = [ctypes.c_void_p]
    L.skani_oracle_mm_hash64.restype = ctypes.c_uint64
    L.skani_oracle_mm_hash64.argtypes = [ctypes.c_uint64]
    L.skani_oracle_seeds.restype = ctypes.c_uint64
    L.skani_oracle_seeds.argtypes = [u8p, ctypes.c_uint64, u64p, u64p, ctypes.c_uint32, ctypes.c_uint32,
                                     u32p, u32p, u32p, ctypes.c_uint64, u32p, u64p]
    L.skani_oracle_chain.restype = ctypes.c_uint64
    L.skani_oracle_chain.argtypes = [u32p, u32p, u32p, ctypes.c_uint64, u32p, u32p, ctypes.c_uint64, u64p,
                                     u32p, ctypes.c_uint64]
    L.skani_oracle_chunk_identity_fx.restype = ctypes.c_uint64
    L.skani_oracle_chunk_identity_fx.argtypes = [ctypes.c_uint32, ctypes.c_uint32]
    L.skani_oracle_finish.restype = ctypes.c_float
    L.skani_oracle_finish.argtypes = [ctypes.c_uint64] * 9 + [ctypes.c_uint32, ctypes.c_int, ctypes.c_float,
                                      ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_double),
                                      ctypes.POINTER(ctypes.c_int)]
    L._skani_ready = True
    return L


def load_codes(path):
    """(codes uint8[n] with 0..3 = ACGT, 4 = other; rec_start; rec_end) in packed coordinates."""
    L = _skani_sigs()
    codes = ctypes.POINTER(ctypes.c_uint8)()
    rs = ctypes.POINTER(ctypes.c_uint64)()
    re_ = ctypes.POINTER(ctypes.c_uint64)()
    n = ctypes.c_uint64(0)
    nrec = ctypes.c_uint32(0)
    rc = L.oracle_load_codes(os.fsencode(path), ctypes.byref(codes), ctypes.byref(n), ctypes.byref(rs),
                             ctypes.byref(re_), ctypes.byref(nrec))
    if rc:
        raise RuntimeError(f"oracle_load_codes({path}) rc={rc}")
    try:
        c = np.ctypeslib.as_array(codes, shape=(max(n.value, 1),))[: n.value].copy()
        s = np.ctypeslib.as_array(rs, shape=(max(nrec.value, 1),))[: nrec.value].copy()
        e = np.ctypeslib.as_array(re_, shape=(max(nrec.value, 1),))[: nrec.value].copy()
    finally:
        L.oracle_free(codes); L.oracle_free(rs); L.oracle_free(re_)
    return c, s, e


def codes_from_ascii(seq: bytes):
    """One record of raw ACGT bytes -> (codes, rec_start, rec_end)."""
    lut = np.full(256, 4, np.uint8)
    for ch, v in zip(b"ACGTacgt", [0, 1, 2, 3, 0, 1, 2, 3]):
        lut[ch] = v
    c = lut[np.frombuffer(seq, np.uint8)]
    return c, np.array([0], np.uint64), np.array([len(c)], np.uint64)


class AniGenome:
    """Seeds of one genome under the skani_oracle.c specification."""

    def __init__(self, codes, rec_start, rec_end, c=125):
        L = _skani_sigs()
        codes = np.ascontiguousarray(codes, np.uint8)
        rs = np.ascontiguousarray(rec_start, np.uint64)
        re_ = np.ascontiguousarray(rec_end, np.uint64)
        cap = max(1024, int(len(codes) // c * 2 + 1024))
        while True:
            ks = np.zeros(cap, np.uint32); sp = np.zeros(cap, np.uint32); ch = np.zeros(cap, np.uint32)
            nch = ctypes.c_uint32(0); tl = ctypes.c_uint64(0)
            n = L.skani_oracle_seeds(_p(codes, ctypes.c_uint8), len(codes), _p(rs, ctypes.c_uint64),
                                     _p(re_, ctypes.c_uint64), len(rs), c, _p(ks, ctypes.c_uint32),
                                     _p(sp, ctypes.c_uint32), _p(ch, ctypes.c_uint32), cap,
                                     ctypes.byref(nch), ctypes.byref(tl))
            if n <= cap:
                break
            cap = int(n)
        self.kmer_strand, self.spread, self.chunk = ks[:n].copy(), sp[:n].copy(), ch[:n].copy()
        self.n_chunks, self.total_len = nch.value, tl.value

    @classmethod
    def from_file(cls, path, c=125):
        return cls(*load_codes(path), c=c)


def ani_pair_integers(a: "AniGenome", b: "AniGenome", chunks=False):
    """(sum_fx, n_chunks_counted, covq, covr, len_q, len_r, sum_m): the query is `a`, the FIRST genome
    (skani dist -q a -r b, src/skani.rs:733-744).  With chunks=True also the per-chunk (M, N) list."""
    L = _skani_sigs()
    q, r = a, b
    out = np.zeros(8, np.uint64)
    cap = q.n_chunks + 1
    mn = np.zeros(2 * cap, np.uint32)
    n = L.skani_oracle_chain(_p(q.kmer_strand, ctypes.c_uint32), _p(q.spread, ctypes.c_uint32),
                             _p(q.chunk, ctypes.c_uint32), len(q.kmer_strand),
                             _p(r.kmer_strand, ctypes.c_uint32), _p(r.spread, ctypes.c_uint32),
                             len(r.kmer_strand), _p(out, ctypes.c_uint64), _p(mn, ctypes.c_uint32), cap)
    res = (int(out[0]), int(out[1]), int(out[2]), int(out[3]), q.total_len, r.total_len, int(out[4]),
           int(out[5]), int(out[6]), int(out[7]))
    if chunks:
        return res, mn[: 2 * int(n)].reshape(-1, 2).copy()
    return res


def chunk_identity_fx(m, n):
    return int(_skani_sigs().skani_oracle_chunk_identity_fx(m, n))


def ani_finish(ints, min_af_pct, c=125, individual_contigs=False):
    """ints = ani_pair_integers(...) -> (ani f32 as galah parses it, af_q, af_r, unrounded ani, estimator)."""
    L = _skani_sigs()
    sum_fx, n_counted, covq, covr, len_q, len_r, _sum_m, span_m, span_n, n_chains = ints
    af = (ctypes.c_double * 2)()
    un = ctypes.c_double(0)
    est = ctypes.c_int(0)
    v = L.skani_oracle_finish(sum_fx, n_counted, covq, covr, len_q, len_r, span_m, span_n, n_chains, c,
                              int(bool(individual_contigs)), ctypes.c_float(min_af_pct), af, ctypes.byref(un),
                              ctypes.byref(est))
    return np.float32(v), af[0], af[1], un.value, est.value


def ani_pair(a, b, min_af_pct=15.0, c=125, individual_contigs=False):
    """ANI of query a vs reference b (both AniGenome built with the same c)."""
    return ani_finish(ani_pair_integers(a, b), min_af_pct, c, individual_contigs)


def markers(codes, rec_start, rec_end, c_marker=1000):
    """Ascending distinct marker hashes of one unit (skani_oracle.c)."""
    L = _skani_sigs()
    L.skani_oracle_markers.restype = ctypes.c_uint64
    L.skani_oracle_markers.argtypes = [ctypes.POINTER(ctypes.c_uint8), ctypes.POINTER(ctypes.c_uint64),
                                       ctypes.POINTER(ctypes.c_uint64), ctypes.c_uint32, ctypes.c_uint32,
                                       ctypes.POINTER(ctypes.c_uint64), ctypes.c_uint64]
    codes = np.ascontiguousarray(codes, np.uint8)
    rs = np.ascontiguousarray(rec_start, np.uint64); re_ = np.ascontiguousarray(rec_end, np.uint64)
    cap = max(1024, len(codes) // c_marker * 3 + 1024)
    while True:
        out = np.zeros(cap, np.uint64)
        n = L.skani_oracle_markers(_p(codes, ctypes.c_uint8), _p(rs, ctypes.c_uint64), _p(re_, ctypes.c_uint64),
                                   len(rs), c_marker, _p(out, ctypes.c_uint64), cap)
        if n <= cap:
            return out[:n].copy()
        cap = int(n)


def skani_distances(units, threshold, min_af_pct, small_genomes=False, individual_contigs=False,
                    variant="triangle", is_ref=None):
    """Oracle of SkaniPreclusterer::distances / ::distances_contigs / low-memory /
    ::distances_with_references on `units` = list of (codes, rec_start, rec_end): marker screen -> ANI
    -> keep ani >= threshold.  variant "triangle": query = lower index (src/skani.rs:109-225);
    "lowmem": sketch + search of everything against everything, the later record of a pair wins =
    query the HIGHER index (src/skani.rs:229-377); "references": only pairs of one reference and one
    non-reference (is_ref), query = the non-reference (src/skani.rs:502-687).
    Returns [(i, j, common, total, ani f32)], i < j."""
    import math
    L = _skani_sigs()
    L.skani_oracle_screen_fraction.restype = ctypes.c_double
    frac = L.skani_oracle_screen_fraction()
    c, cm = (30, 200) if small_genomes else (125, 1000)
    gen = [AniGenome(*u, c=c) for u in units]
    mk = [markers(*u, c_marker=cm) for u in units]
    out = []
    for i in range(len(units)):
        for j in range(i + 1, len(units)):
            if variant == "references" and bool(is_ref[i]) == bool(is_ref[j]):
                continue
            m = min(len(mk[i]), len(mk[j]))
            common, total = raw_distance(mk[i], mk[j]) if m else (0, 0)
            bypass = (not small_genomes) and m < 20  # skani without --faster-small
            if not bypass and (m == 0 or common < max(1, math.ceil(frac * m))):
                continue
            q, r = i, j
            if variant == "lowmem" or (variant == "references" and is_ref[i]):
                q, r = j, i
            ani = ani_pair(gen[q], gen[r], min_af_pct, c, individual_contigs)[0]
            if np.float32(ani) >= np.float32(threshold):
                out.append((i, j, common, total, np.float32(ani)))
    return out
