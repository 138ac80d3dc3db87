"""Quality-ordering inputs (host side of the ingest pass) vs the reference's pinned values
(src/genome_stats.rs:63-95) and an independent restatement."""
import gzip
import os

import numpy as np
import pytest

import galah_b200 as gb
from conftest import GOLDEN, REFERENCE_DATA
from util import random_dna, write_fasta


def restated(path):
    """src/genome_stats.rs:11-51 in plain Python."""
    op = gzip.open if path.endswith(".gz") else open
    lens, n_amb, cur = [], 0, None
    with op(path, "rb") as f:
        for line in f:
            if line.startswith(b">"):
                if cur is not None:
                    lens.append(cur)
                cur = 0
            else:
                seq = line.strip()
                cur += len(seq)
                n_amb += seq.count(b"N") + seq.count(b"n")
    if cur is not None:
        lens.append(cur)
    total, s, n50 = sum(lens), 0, None
    for l in sorted(lens):
        s += l
        if s >= total // 2:
            n50 = l
            break
    return len(lens), n_amb, n50


@pytest.mark.skipif(not os.path.isdir(REFERENCE_DATA), reason="reference fixtures not on this box")
def test_reference_pinned_values():
    st = gb.genome_stats([os.path.join(REFERENCE_DATA, "abisko4", "73.20110600_S2D.10.fna"),
                          os.path.join(REFERENCE_DATA, "set1", "1mbp.fna")])
    assert tuple(st[0]) == (161, 6506, 8289)      # src/genome_stats.rs:68-76
    assert tuple(st[1]) == (1, 0, 1_000_000)      # src/genome_stats.rs:83-91


def test_committed_genomes_and_synthetic(tmp_path):
    rng = np.random.default_rng(3)
    paths = [os.path.join(GOLDEN, "abisko4", n) for n in sorted(os.listdir(os.path.join(GOLDEN, "abisko4")))]
    recs = [("a", random_dna(500, rng) + b"NNnn" + random_dna(50, rng)), ("b", b"ACGTRYK" * 30), ("c", random_dna(2000, rng)),
            ("d", b"")]
    paths.append(write_fasta(str(tmp_path / "x.fna"), recs))
    st = gb.genome_stats(paths, threads=2)
    for p, s in zip(paths, st):
        assert tuple(int(v) for v in s) == restated(p), p
    assert int(st[-1]["num_contigs"]) == 4 and int(st[-1]["num_ambiguous_bases"]) == 4
