"""Host FASTA ingest (C++ packer) vs the oracle's reader: identical per-base codes and record
coordinates on the needletail edge cases the reference's fixtures exercise."""
import os

import numpy as np
import pytest

import galah_b200 as gb
import oracle
from conftest import GOLDEN
from util import random_dna, write_fasta


def same(path):
    codes, rs, re_ = gb.pack_fasta_file(path)
    ocodes, ors, ore = oracle.load_codes(path)
    assert np.array_equal(rs, ors) and np.array_equal(re_, ore)
    assert np.array_equal(codes, ocodes)
    return codes


def test_committed_real_genomes_plain_and_gzip():
    for n in sorted(os.listdir(os.path.join(GOLDEN, "abisko4"))):
        same(os.path.join(GOLDEN, "abisko4", n))
    same(os.path.join(GOLDEN, "set1_500kb.fna.gz"))


@pytest.mark.parametrize("width,newline,gz", [(60, "\n", False), (70, "\r\n", False), (10**9, "\n", True), (31, "\n", True)])
def test_synthetic_edge_cases(tmp_path, width, newline, gz):
    rng = np.random.default_rng(width % 97)
    recs = [("r0 with\ttab", random_dna(1000, rng) + b"NNNN" + random_dna(33, rng).lower()),
            ("r1", b""), ("r2", b"ACGURYKMacgu.-~*" * 9), ("r3", random_dna(31, rng)), ("r4", random_dna(32, rng)),
            ("r5", random_dna(4097, rng))]
    p = write_fasta(str(tmp_path / ("x.fna.gz" if gz else "x.fna")), recs, width=width, newline=newline, gz=gz)
    codes = same(p)
    assert len(codes) == sum(len(r[1]) for r in recs) + len(recs) - 1  # one separator between records


def test_truncated_gzip_fails_instead_of_yielding_a_shorter_genome(tmp_path):
    """needletail (the reference's reader, src/finch.rs:69) fails on a gzip stream cut mid-member."""
    import galah_b200 as gb
    rng = np.random.default_rng(5)
    p = write_fasta(str(tmp_path / "g.fna.gz"), [("r", random_dna(300_000, rng))], gz=True)
    data = open(p, "rb").read()
    cut = str(tmp_path / "cut.fna.gz")
    with open(cut, "wb") as f:
        f.write(data[: len(data) // 2])
    gb.pack_fasta_file(p)
    with pytest.raises(gb.GalahB200Error, match="truncated or corrupt gzip"):
        gb.pack_fasta_file(cut)
