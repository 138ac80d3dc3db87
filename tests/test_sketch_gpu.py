"""K1 parity: the CUDA sketch kernel vs the CPU oracle (identical uint64 hash lists)."""
import os

import numpy as np
import pytest

import oracle
from conftest import GOLDEN
from util import random_dna, write_fasta

pytestmark = pytest.mark.gpu


def _check_files(gb, paths, k=21, s=1000):
    table, counts = gb.sketch_files(paths, k=k, s=s)
    for g, p in enumerate(paths):
        exp = oracle.sketch_fasta(p, k=k, s=s)
        assert counts[g] == len(exp), (p, counts[g], len(exp))
        assert np.array_equal(table[g, : counts[g]], exp), p
        assert np.all(table[g, counts[g]:] == np.uint64(0xFFFFFFFFFFFFFFFF))


def test_reference_kat_end_to_end(gb):
    """/root/reference/src/finch.rs:107-129 through FASTA ingest + K1 + K2 on the GPU."""
    paths = [os.path.join(GOLDEN, "set1_1mbp.fna.gz"), os.path.join(GOLDEN, "set1_500kb.fna.gz")]
    got = gb.finch_distances(paths, 0.9, 1000, 21)
    assert len(got) == 1
    assert (got[0]["i"], got[0]["j"], got[0]["common"], got[0]["total"]) == (0, 1, 502, 1000)
    assert got[0]["ani"] == np.float32(0.9808188)
    assert len(gb.finch_distances(paths, 0.99, 1000, 21)) == 0
    _check_files(gb, paths)


def test_fasta_features(gb, tmp_path):
    rng = np.random.default_rng(5)
    base = random_dna(300_000, rng)
    mutated = bytearray(base)
    for p in rng.integers(0, len(base), size=3000):
        mutated[p] = b"ACGT"[rng.integers(0, 4)]
    with_n = bytearray(base)
    with_n[1000:1800] = b"N" * 800
    with_n[50_000:50_003] = b"RYK"
    with_n[70_000] = ord("-")
    paths = [
        write_fasta(str(tmp_path / "plain.fna"), [("c1", base)]),
        write_fasta(str(tmp_path / "multi.fna"), [(f"c{i} desc", base[i * 30_000:(i + 1) * 30_000]) for i in range(10)], width=70),
        write_fasta(str(tmp_path / "lower.fna"), [("c1", bytes(base).lower())]),
        write_fasta(str(tmp_path / "nruns.fna"), [("c1", bytes(with_n))]),
        write_fasta(str(tmp_path / "crlf.fna"), [("c1", base[:100_000]), ("c2", base[100_000:150_000])], newline="\r\n"),
        write_fasta(str(tmp_path / "gz.fna.gz"), [("c1", bytes(mutated))], gz=True),
        write_fasta(str(tmp_path / "short_records.fna"), [("a", base[:10]), ("b", base[10:30]), ("c", base[30:51]), ("d", base[51:5000])]),
        write_fasta(str(tmp_path / "oneline.fna"), [("c1", base)], width=10**9),
        write_fasta(str(tmp_path / "tiny.fna"), [("c1", base[:500])]),
        write_fasta(str(tmp_path / "uracil.fna"), [("c1", bytes(base[:50_000]).replace(b"T", b"U"))]),
    ]
    with open(tmp_path / "reads.fq", "wb") as f:
        for i in range(50):
            f.write(b"@r%d\n" % i + base[i * 150:(i + 1) * 150] + b"\n+\n" + b"I" * 150 + b"\n")
    paths.append(str(tmp_path / "reads.fq"))
    _check_files(gb, paths)
    # plain == lower == oneline
    table, counts = gb.sketch_files(paths[:3] + [paths[7]])
    assert np.array_equal(table[0], table[2]) and np.array_equal(table[0], table[3])


def test_fewer_than_s_distinct_and_repeats(gb, tmp_path):
    rng = np.random.default_rng(6)
    unit = random_dna(700, rng)
    paths = [
        write_fasta(str(tmp_path / "small.fna"), [("c1", unit)]),                      # < 1000 k-mers
        write_fasta(str(tmp_path / "tandem.fna"), [("c1", unit * 400)]),               # 280 kb, 700 distinct
        write_fasta(str(tmp_path / "polya.fna"), [("c1", b"A" * 100_000 + random_dna(3000, rng))]),
        write_fasta(str(tmp_path / "dup_heavy.fna"), [("c1", random_dna(40_000, rng) * 20)]),  # duplicates 20x
        write_fasta(str(tmp_path / "empty_seq.fna"), [("c1", b"")]),
        write_fasta(str(tmp_path / "all_n.fna"), [("c1", b"N" * 5000)]),
    ]
    _check_files(gb, paths)
    table, counts = gb.sketch_files(paths)
    assert counts[0] == 680 and counts[4] == 0 and counts[5] == 0


@pytest.mark.parametrize("k,s", [(21, 100), (15, 1000), (16, 200), (31, 1000), (32, 64), (8, 50), (17, 2), (25, 2000)])
def test_other_k_and_s(gb, tmp_path, k, s):
    rng = np.random.default_rng(k * 1000 + s)
    recs = [("a", random_dna(120_000, rng)), ("b", random_dna(999, rng))]
    _check_files(gb, [write_fasta(str(tmp_path / "x.fna"), recs)], k=k, s=s)


def test_many_files_batching(gb, tmp_path):
    rng = np.random.default_rng(8)
    paths = [write_fasta(str(tmp_path / f"g{i}.fna"), [("c", random_dna(int(rng.integers(1, 60_000)), rng))]) for i in range(40)]
    _check_files(gb, paths)


def test_synthetic_generator_and_sketch_on_device(gb):
    """On-device packed synthetic genomes == oracle's (seed, index) definition, and their
    sketches == oracle sketches of the same genomes."""
    import torch
    seed, n, L, begin = 1, 24, 150_003, 95
    lay = gb.synth_layout(n, L)
    d_seq = torch.zeros(lay["seq2_words"], dtype=torch.int32, device="cuda")
    d_val = torch.zeros(lay["valid_words"], dtype=torch.int32, device="cuda")
    d_off = torch.zeros(n + 1, dtype=torch.int64, device="cuda")
    d_h = torch.zeros((n, 1000), dtype=torch.int64, device="cuda")
    d_c = torch.zeros(n, dtype=torch.int32, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    gb.synth_packed_device(seed, begin, n, L, d_seq.data_ptr(), d_val.data_ptr(), d_off.data_ptr(), st)
    gb.sketch_packed_device(d_seq.data_ptr(), d_val.data_ptr(), d_off.data_ptr(), n, 21, 1000, 0,
                            d_h.data_ptr(), d_c.data_ptr(), st)
    torch.cuda.synchronize()
    seq = d_seq.cpu().numpy().view(np.uint32)
    off = d_off.cpu().numpy()
    assert np.array_equal(off, np.arange(n + 1) * lay["padded"])
    for g in (0, 7, 23):
        words = seq[off[g] // 16: off[g] // 16 + (L + 15) // 16]
        codes = ((words[:, None] >> (2 * np.arange(16, dtype=np.uint32))) & 3).reshape(-1)[:L]
        exp = np.frombuffer(oracle.synth_genome(seed, begin + g, L), np.uint8)
        assert np.array_equal(np.frombuffer(b"ACGT", np.uint8)[codes], exp)
    table, counts = oracle.sketch_synth(seed, begin, n, L)
    assert np.array_equal(d_c.cpu().numpy().view(np.uint32), counts)
    assert np.array_equal(d_h.cpu().numpy().view(np.uint64), table)
