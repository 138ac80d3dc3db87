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


def synth_block(seed, index, block):
    return int(lib().oracle_synth_block(seed, index, block))


def sketch_synth(seed, index_begin, n, length, k=21, s=1000, hash_seed=0):
    """Sketches of synthetic genomes index_begin..index_begin+n (all host threads)."""
    table = np.full((n, s), np.uint64(0xFFFFFFFFFFFFFFFF), np.uint64)
    counts = np.zeros(n, np.uint32)
    lib().oracle_sketch_synth_mt(seed, index_begin, n, length, k, s, hash_seed,
                                 _p(table, ctypes.c_uint64), _p(counts, ctypes.c_uint32))
    return table, counts
