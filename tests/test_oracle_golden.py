"""Pins the CPU oracle: reference known-answer test + committed golden vectors (CPU only)."""
import os

import numpy as np
import pytest

import oracle
from conftest import GOLDEN, REFERENCE_DATA


def test_murmur3_known_answers():
    # MurmurHash3_x64_128 h1, published vectors (seed 0)
    assert oracle.murmur3_h1(b"hello") == 0xCBD8A7B341BD9B02
    assert oracle.murmur3_h1(b"The quick brown fox jumps over the lazy dog") == 0xE34BBC7BBC071B6C
    assert oracle.murmur3_h1(b"") == 0


def test_reference_kat_from_committed_inputs():
    """/root/reference/src/finch.rs:107-129: distances([1mbp, 500kb], 0.9, 1000, 21) ==
    {(0,1): Some(0.9808188)}; with min_ani 0.99 the cache is empty."""
    a = oracle.sketch_fasta(os.path.join(GOLDEN, "set1_1mbp.fna.gz"))
    b = oracle.sketch_fasta(os.path.join(GOLDEN, "set1_500kb.fna.gz"))
    table, counts = oracle.pack_table([a, b], 1000)
    got = oracle.prefilter(table, counts, 21, 0.9)
    assert len(got) == 1
    assert (got[0]["i"], got[0]["j"], got[0]["common"], got[0]["total"]) == (0, 1, 502, 1000)
    assert got[0]["ani"] == np.float32(0.9808188)
    assert len(oracle.prefilter(table, counts, 21, 0.99)) == 0


@pytest.mark.skipif(not os.path.isdir(REFERENCE_DATA), reason="reference fixtures only exist in the build container")
def test_golden_matches_reference_fixtures(golden):
    names = [str(x) for x in golden["names"]]
    for idx in (0, 1, 7, 13, 29):
        sk = oracle.sketch_fasta(os.path.join(REFERENCE_DATA, names[idx]))
        assert np.array_equal(sk, golden["table"][idx][: golden["counts"][idx]])


def test_golden_pair_list_is_reproduced(golden):
    got = oracle.prefilter(golden["table"], golden["counts"], 21, 0.9)
    exp = golden["pairs_min_ani_0p9"]
    assert len(got) == len(exp)
    for f in ("i", "j", "common", "total"):
        assert np.array_equal(got[f], exp[f])
    assert np.array_equal(got["ani"].view(np.uint32), exp["ani"].view(np.uint32))
    # SURVEY.md Appendix B rows (derived with the restatement; row 1 is reference-pinned)
    names = [str(x) for x in golden["names"]]
    look = {(int(p["i"]), int(p["j"])): p for p in got}

    def pair(a, b):
        i, j = sorted((names.index(a), names.index(b)))
        return look[(i, j)]
    p = pair("set1/1mbp.fna", "set1/500kb.fna")
    assert (p["common"], p["total"], p["ani"]) == (502, 1000, np.float32(0.9808188))
    p = pair("set2/1mbp.fna", "set2/1mbp.half_aligned.fna")
    assert (p["common"], p["total"], p["ani"]) == (502, 1469, np.float32(0.96787864))
    p = pair("antonio_mags/BE_RX_R2_MAG52.fna", "antonio_mags/BE_RX_R3_MAG189.fna")
    assert (p["common"], p["total"], p["ani"]) == (460, 1155, np.float32(0.97320396))
    p = pair("abisko4/73.20110800_S2M.16.fna", "abisko4/73.20110800_S2M.16.fna.gz")
    assert (p["common"], p["total"], p["ani"]) == (1000, 1000, np.float32(1.0))


def test_raw_distance_edge_cases():
    a = np.array([1, 5, 9], np.uint64)
    assert oracle.raw_distance(a, a) == (3, 3)
    assert oracle.raw_distance(a, np.array([], np.uint64)) == (0, 0)
    assert oracle.raw_distance(a, np.array([10, 11], np.uint64)) == (0, 3)
    assert oracle.raw_distance(a, np.array([0, 1, 2, 9, 100], np.uint64)) == (2, 5)
    # total counts only what the merge consumed before one list ran out
    assert oracle.raw_distance(np.array([1, 2, 3], np.uint64), np.array([2, 50, 60, 70], np.uint64)) == (1, 3)
    assert oracle.mash_ani(0, 10) == 0.0
    assert oracle.mash_ani(10, 10) == 1.0
    assert oracle.mash_ani(0, 0) == 1.0  # 0/0 = NaN, f64::max/min ignore NaN (flagged in SURVEY App. A)


def test_sketch_normalisation_and_record_semantics():
    k, s = 5, 50
    # lower case == upper case; u == T; k-mers do not span records; N/- break k-mers; whitespace ignored
    assert np.array_equal(oracle.sketch_records([b"ACGTTGCAAC"], k, s), oracle.sketch_records([b"acgtugcaac"], k, s))
    assert np.array_equal(oracle.sketch_records([b"ACGTT\nGCA AC\r\n"], k, s), oracle.sketch_records([b"ACGTTGCAAC"], k, s))
    joined = oracle.sketch_records([b"ACGTTGCAACGGTA"], k, s)
    split = oracle.sketch_records([b"ACGTTGC", b"AACGGTA"], k, s)
    assert len(split) < len(joined)
    with_n = oracle.sketch_records([b"ACGTTGCNAACGGTA"], k, s)
    assert np.array_equal(with_n, split)
    assert np.array_equal(oracle.sketch_records([b"ACGTTGC-AACGGTA"], k, s), split)
    # canonical: a sequence and its reverse complement sketch identically
    rc = bytes(reversed(b"ACGTTGCAACGGTA".translate(bytes.maketrans(b"ACGT", b"TGCA"))))
    assert np.array_equal(oracle.sketch_records([rc], k, s), joined)
    # distinct: repeating the sequence changes nothing
    assert np.array_equal(oracle.sketch_records([b"ACGTTGCAACGGTA"] * 3, k, s), joined)
    # too short / empty
    assert len(oracle.sketch_records([b"ACG"], k, s)) == 0
    assert len(oracle.sketch_records([], k, s)) == 0


def test_synth_genomes_are_counter_based():
    g0 = oracle.synth_genome(1, 0, 1000)
    assert g0 == oracle.synth_genome(1, 0, 2000)[:1000]
    assert set(g0) <= set(b"ACGT")
    # member 0 has rate 0: equals the founder; other members differ by roughly their rate
    g5 = oracle.synth_genome(1, 5, 200000)
    f = oracle.synth_genome(1, 0, 200000)
    diff = np.mean(np.frombuffer(g5, np.uint8) != np.frombuffer(f, np.uint8))
    assert 0.035 < diff < 0.045  # member 5 -> 4 %
    other = oracle.synth_genome(1, 10, 200000)  # another family
    assert 0.70 < np.mean(np.frombuffer(other, np.uint8) != np.frombuffer(f, np.uint8)) < 0.80
