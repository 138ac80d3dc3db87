"""Shared helpers for the test-suite (synthetic sketch tables, FASTA writers)."""
import gzip
import os

import numpy as np

PAD = np.uint64(0xFFFFFFFFFFFFFFFF)


def random_family_table(n, s, rng, family=10, hi_bits=53, ragged=False):
    """Sorted-distinct uint64 sketch table where members of a family share a varying fraction."""
    table = np.full((n, s), PAD, np.uint64)
    counts = np.zeros(n, np.uint32)
    founder = None
    for g in range(n):
        if g % family == 0 or founder is None:
            founder = np.unique(rng.integers(0, 1 << hi_bits, size=2 * s, dtype=np.uint64))[: s]
            rng.shuffle(founder)
        keep = rng.uniform(0.0, 1.0)
        nk = int(round(keep * s))
        own = rng.integers(0, 1 << hi_bits, size=s, dtype=np.uint64)
        merged = np.unique(np.concatenate([founder[:nk], own[: s - nk]]))
        m = len(merged)
        if ragged:
            m = int(rng.integers(0, m + 1)) if rng.uniform() < 0.3 else m
        table[g, :m] = merged[:m]
        counts[g] = m
    return table, counts


def assert_pairs_equal(got, exp):
    assert len(got) == len(exp), f"{len(got)} pairs vs {len(exp)} expected"
    for f in ("i", "j", "common", "total"):
        assert np.array_equal(got[f], exp[f]), f
    assert np.array_equal(got["ani"].view(np.uint32), exp["ani"].view(np.uint32)), "ani bits"


def write_fasta(path, records, width=60, newline="\n", gz=False):
    """records: list of (name, bytes)."""
    out = bytearray()
    for name, seq in records:
        out += b">" + name.encode() + newline.encode()
        for p in range(0, len(seq), width):
            out += seq[p: p + width] + newline.encode()
    if gz:
        with gzip.open(path, "wb") as f:
            f.write(bytes(out))
    else:
        with open(path, "wb") as f:
            f.write(bytes(out))
    return path


def random_dna(n, rng):
    return bytes(np.frombuffer(b"ACGT", np.uint8)[rng.integers(0, 4, size=n)])
