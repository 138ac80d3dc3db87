"""CPU tests of the stage-2 oracle (oracle/skani_oracle.c): internal consistency.  The outcomes the
reference's own tests pin for skani on its fixture genomes are in tests/test_stage2_fixtures.py."""
import os

import numpy as np
import pytest

import oracle
from util import random_dna



def mutate(seq, rate, rng):
    a = np.frombuffer(seq, np.uint8).copy()
    hit = rng.uniform(size=len(a)) < rate
    lut = np.zeros(256, np.uint8)
    for x, y in zip(b"ACGT", b"CGTA"):
        lut[x] = y
    a[hit] = lut[a[hit]]
    return a.tobytes()


def revcomp(seq):
    return seq[::-1].translate(bytes.maketrans(b"ACGT", b"TGCA"))


def test_mm_hash64_is_invertible_mix():
    """minimap2's hash64 with a full 64-bit mask; spot values computed by hand from the definition."""
    L = oracle.binding._skani_sigs()
    def ref(key):
        m = (1 << 64) - 1
        key = (~key + (key << 21)) & m
        key ^= key >> 24
        key = (key + (key << 3) + (key << 8)) & m
        key ^= key >> 14
        key = (key + (key << 2) + (key << 4)) & m
        key ^= key >> 28
        key = (key + (key << 31)) & m
        return key
    for k in (0, 1, 2, 0x3FFFFFFF, 123456789, (1 << 30) - 5):
        assert L.skani_oracle_mm_hash64(k) == ref(k)


def test_identical_and_mutated_genomes():
    rng = np.random.default_rng(1)
    base = random_dna(400_000, rng)
    g0 = oracle.AniGenome(*oracle.codes_from_ascii(base))
    assert abs(len(g0.kmer_strand) - 400_000 / 125) < 400 and g0.n_chunks == 20 and g0.total_len == 400_000
    ani, afq, afr, _, _ = oracle.ani_pair(g0, g0, 15.0)
    assert ani == np.float32(100.0) and afq > 0.97 and afr > 0.97
    for rate, lo, hi in ((0.01, 98.8, 99.2), (0.05, 94.6, 95.4), (0.10, 89.3, 90.7)):
        g1 = oracle.AniGenome(*oracle.codes_from_ascii(mutate(base, rate, rng)))
        ani, afq, afr, un, _ = oracle.ani_pair(g0, g1, 15.0)
        assert lo < ani < hi, (rate, ani)
        assert afq > (0.85 if rate <= 0.05 else 0.6)
    # strand symmetry: the reverse complement is the same genome
    g2 = oracle.AniGenome(*oracle.codes_from_ascii(revcomp(base)))
    assert oracle.ani_pair(g0, g2, 15.0)[0] == np.float32(100.0)
    # unrelated genomes: no row
    g3 = oracle.AniGenome(*oracle.codes_from_ascii(random_dna(400_000, rng)))
    assert oracle.ani_pair(g0, g3, 15.0)[0] == np.float32(0.0)


def test_min_af_gate_and_two_decimal_print():
    rng = np.random.default_rng(2)
    base = random_dna(300_000, rng)
    half = base[:180_000] + random_dna(120_000, rng)
    a = oracle.AniGenome(*oracle.codes_from_ascii(base))
    b = oracle.AniGenome(*oracle.codes_from_ascii(half))
    ints = oracle.ani_pair_integers(a, b)
    ani20, afq, afr, un, est = oracle.ani_finish(ints, 20.0)
    ani60 = oracle.ani_finish(ints, 65.0)[0]
    assert 0.55 < afq < 0.62 and ani20 == np.float32(100.0) and ani60 == np.float32(0.0)
    assert est == 1  # c = 125, whole genomes, >= 150 kb aligned: the chain-span estimator
    # one chunk of 1000 matched seeds in a span of 2000: the raw estimator (contig mode)
    one = (oracle.chunk_identity_fx(1000, 2000), 1, 100, 100, 100, 100, 1000, 1000, 2000, 1)
    v = oracle.ani_finish(one, 0.0, 125, True)
    assert v[0] == np.float32(f"{100 * 0.5 ** (1 / 15):.2f}") and v[4] == 0
