"""CPU tests of the stage-2 oracle (oracle/skani_oracle.c): internal consistency everywhere, and --
only where /root/reference exists (the build container) -- the threshold-crossing behaviours the
reference's own tests pin for skani on its fixture genomes."""
import os

import numpy as np
import pytest

import oracle
from conftest import REFERENCE_DATA
from util import random_dna

needs_reference = pytest.mark.skipif(not os.path.isdir(REFERENCE_DATA), reason="reference fixtures not on this box")


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
    ani, afq, afr, _ = oracle.ani_pair(g0, g0, 15.0)
    assert ani == np.float32(100.0) and afq > 0.97 and afr > 0.97
    for rate, lo, hi in ((0.01, 98.8, 99.2), (0.05, 94.6, 95.4), (0.10, 89.3, 90.7)):
        g1 = oracle.AniGenome(*oracle.codes_from_ascii(mutate(base, rate, rng)))
        ani, afq, afr, un = oracle.ani_pair(g0, g1, 15.0)
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
    half = base[:150_000] + random_dna(150_000, rng)
    a = oracle.AniGenome(*oracle.codes_from_ascii(base))
    b = oracle.AniGenome(*oracle.codes_from_ascii(half))
    ints = oracle.ani_pair_integers(a, b)
    ani20, afq, afr, un = oracle.ani_finish(*ints[:6], 20.0)
    ani60 = oracle.ani_finish(*ints[:6], 60.0)[0]
    assert 0.4 < afq < 0.55 and ani20 == np.float32(100.0) and ani60 == np.float32(0.0)
    v = oracle.ani_finish(1000, 2000, 100, 100, 100, 100, 0.0)
    assert v[0] == np.float32(f"{100 * 0.5 ** (1 / 15):.2f}")


@needs_reference
def test_reference_pinned_behaviours_abisko4():
    """src/clusterer.rs:631-690: finch(0.9)+skani on the 4 abisko genomes is ONE cluster at 95 and
    {0,1,3},{2} at 99.  With the engine's greedy rule that needs: every pair >= 95; (0,1),(0,3) >= 99;
    (0,2) < 99."""
    D = os.path.join(REFERENCE_DATA, "abisko4")
    names = ["73.20120800_S1X.13.fna", "73.20120600_S2D.19.fna", "73.20120700_S3X.12.fna", "73.20110800_S2D.13.fna"]
    G = [oracle.AniGenome.from_file(os.path.join(D, n)) for n in names]
    ani = {(i, j): oracle.ani_pair(G[i], G[j], 20.0)[0] for i in range(4) for j in range(i + 1, 4)}
    assert all(v >= 95.0 for v in ani.values())
    assert ani[(0, 1)] >= 99.0 and ani[(0, 3)] >= 99.0 and ani[(0, 2)] < 99.0

    import galah_b200 as gb
    hits = np.zeros(6, gb.PAIR_DTYPE)
    for x, (i, j) in enumerate(ani):
        hits[x]["i"], hits[x]["j"], hits[x]["ani"] = i, j, 0.95  # all pass the 0.9 finch prefilter
    f = lambda r, g: float(ani[(min(r, g), max(r, g))])
    c95, _ = gb.cluster_from_distances(4, hits, 95.0, f)
    c99, _ = gb.cluster_from_distances(4, hits, 99.0, f)
    assert [sorted(c) for c in c95] == [[0, 1, 2, 3]]
    assert sorted(sorted(c) for c in c99) == [[0, 1, 3], [2]]


@needs_reference
def test_reference_pinned_behaviours_af():
    """tests/test_cmdline.rs:262-302: set2 1mbp vs 1mbp.half_aligned merge at min-AF 20 and split at
    60; :417-440: antonio MAG52/MAG189 merge at ANI 95, AF 60."""
    a = oracle.AniGenome.from_file(os.path.join(REFERENCE_DATA, "set2", "1mbp.fna"))
    b = oracle.AniGenome.from_file(os.path.join(REFERENCE_DATA, "set2", "1mbp.half_aligned.fna"))
    # f32 artefacts of parse_percentage: 20 -> 20, 60 -> 60.000004
    assert oracle.ani_pair(a, b, 20.0)[0] >= 95.0
    assert oracle.ani_pair(a, b, 60.000004)[0] == 0.0
    m52 = oracle.AniGenome.from_file(os.path.join(REFERENCE_DATA, "antonio_mags", "BE_RX_R2_MAG52.fna"))
    m189 = oracle.AniGenome.from_file(os.path.join(REFERENCE_DATA, "antonio_mags", "BE_RX_R3_MAG189.fna"))
    assert oracle.ani_pair(m52, m189, 60.000004)[0] >= 95.0
