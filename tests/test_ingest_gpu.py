"""K0 parity: FASTA bytes decoded on the device (csrc/ingest.cu) vs the host packer
(csrc/host/fasta.cpp, itself checked against the oracle's reader in test_ingest_host.py) and the
oracle's reader directly -- bit for bit: base codes, validity, record ranges, ambiguous / N counts."""
import gzip
import os

import numpy as np
import pytest

import oracle
from util import random_dna, write_fasta

pytestmark = pytest.mark.gpu


def host_pack(gb, tmp_path, data, name="x.fna"):
    p = str(tmp_path / name)
    with open(p, "wb") as f:
        f.write(data)
    return gb.pack_fasta_file(p)


def check_same(gb, tmp_path, files):
    got, ms = gb.decode_fasta_device(files)
    assert len(got) == len(files)
    for k, (data, g) in enumerate(zip(files, got)):
        codes, starts, ends = host_pack(gb, tmp_path, data, f"f{k}.fna")
        assert np.array_equal(g["codes"], codes), f"file {k}: codes"
        assert np.array_equal(g["rec_start"], starts) and np.array_equal(g["rec_end"], ends), f"file {k}: records"
        assert g["padding_clean"]
        # counts: ambiguous = invalid bases that are not record separators
        n_sep = max(len(starts) - 1, 0)
        assert g["n_ambiguous"] == int((codes == 4).sum()) - n_sep
    return got


def test_edge_cases_match_host_packer(gb, tmp_path):
    rng = np.random.default_rng(5)
    files = [
        b">a\nACGT\n",
        b">a\nACGT",                                   # no trailing newline
        b"\n\r\n>a desc\r\nACgtNnRYu\r\nUUAA\r\n>b\r\n\r\nGG\r\n",  # CRLF, blank lines, lower case, U, ambiguity codes
        b">only_header\n",
        b">h1\n>h2\nAC\n>h3\n",                        # empty records
        b">a\nAC>GT\nA>C\n>b\nTT\n",                   # '>' inside a sequence line is an ambiguous base
        b">a\n" + b"ACGT" * 5000 + b"\n",              # one long line (crosses chunks without a newline)
        b">" + b"h" * 20000 + b"\nACGTACGT\n>x\nAC\n", # a header longer than two chunks
        b"",                                           # empty file
        b"\n\n\r\n",                                   # only blank lines
        b">a\n \t A C\tG T \n",                         # blanks inside sequence lines
        b">a\n" + random_dna(8191, rng) + b"\n>b\n" + random_dna(8192, rng) + b"\n>c\n" + random_dna(3, rng),
    ]
    got = check_same(gb, tmp_path, files)
    assert got[2]["n_N"] == 2 and got[5]["n_ambiguous"] == 2


def test_random_multi_record_files(gb, tmp_path):
    rng = np.random.default_rng(11)
    files = []
    for k in range(24):
        recs = []
        for r in range(int(rng.integers(1, 40))):
            n = int(rng.integers(0, 30_000))
            seq = bytearray(random_dna(n, rng))
            for _ in range(int(rng.integers(0, 6))):   # N runs / IUPAC codes / lower case
                if n:
                    a = int(rng.integers(0, n)); b = min(n, a + int(rng.integers(1, 300)))
                    seq[a:b] = bytes(rng.choice(list(b"NnRYKMacgtu"), size=b - a).astype(np.uint8))
            recs.append((f"contig_{k}_{r} some description", bytes(seq)))
        width = int(rng.choice([60, 70, 80, 1 << 30]))
        nl = "\r\n" if k % 5 == 0 else "\n"
        p = write_fasta(str(tmp_path / f"r{k}.fna"), recs, width=min(width, 1 << 20), newline=nl)
        files.append(open(p, "rb").read())
    check_same(gb, tmp_path, files)


def test_matches_oracle_reader_and_committed_genomes(gb, tmp_path):
    """The committed gz copies of the reference's fixtures: device decode == oracle reader."""
    from conftest import GOLDEN
    paths = [os.path.join(GOLDEN, "set1_1mbp.fna.gz"), os.path.join(GOLDEN, "set1_500kb.fna.gz")]
    paths += sorted(os.path.join(GOLDEN, "abisko4", f) for f in os.listdir(os.path.join(GOLDEN, "abisko4")))[:2]
    files = [gzip.open(p, "rb").read() for p in paths]
    got, ms = gb.decode_fasta_device(files)
    for p, data, g in zip(paths, files, got):
        plain = str(tmp_path / (os.path.basename(p)[:-3]))
        with open(plain, "wb") as f:
            f.write(data)
        codes, rs, re_ = oracle.load_codes(plain)
        assert np.array_equal(g["codes"], codes)
        assert np.array_equal(g["rec_start"], rs) and np.array_equal(g["rec_end"], re_)


def test_file_entry_points_agree_with_host_ingest(gb, tmp_path):
    """sketch_files through K0 == sketch_files through the host packer (bit-exact tables)."""
    rng = np.random.default_rng(3)
    paths = []
    for k in range(6):
        recs = [(f"c{r}", random_dna(int(rng.integers(2_000, 60_000)), rng)) for r in range(int(rng.integers(1, 5)))]
        paths.append(write_fasta(str(tmp_path / f"g{k}.fna"), recs, gz=(k % 2 == 1)))
    prev = gb.device_ingest(1)
    try:
        t1, c1 = gb.sketch_files(paths, 21, 1000)
        gb.device_ingest(0)
        t0, c0 = gb.sketch_files(paths, 21, 1000)
    finally:
        gb.device_ingest(prev)
    assert np.array_equal(t1, t0) and np.array_equal(c1, c0)


def test_contig_mode_units_agree_with_host_packer(gb, tmp_path):
    """--cluster-contigs: every record is a unit.  The device path (K0 decode + record split) must
    give the same units as the host packer: identical skani-preclusterer output, including empty
    records, records shorter than one word, and records that are not multiples of 16 / 32 bases."""
    from test_skani_precluster_gpu import mutate
    rng = np.random.default_rng(21)
    base = random_dna(41_003, rng)
    recs1 = [("a", base), ("empty", b""), ("a95", mutate(base, 0.05, rng)), ("tiny", random_dna(7, rng)),
             ("b", random_dna(33_333, rng))]
    recs2 = [("b97", mutate(recs1[4][1], 0.03, rng)), ("c", random_dna(20_001, rng)), ("a99", mutate(base, 0.01, rng))]
    p1 = write_fasta(str(tmp_path / "m1.fna"), recs1, width=70)
    p2 = write_fasta(str(tmp_path / "m2.fna"), recs2, width=60, newline="\r\n")
    prev = gb.device_ingest(1)
    try:
        dev_hits, dev_units = gb.skani_distances([p1, p2], 90.0, 15.0, small_genomes=True, contigs=True)
        gb.device_ingest(0)
        host_hits, host_units = gb.skani_distances([p1, p2], 90.0, 15.0, small_genomes=True, contigs=True)
    finally:
        gb.device_ingest(prev)
    assert dev_units == host_units == len(recs1) + len(recs2)
    assert len(dev_hits) == len(host_hits) >= 3
    for f in ("i", "j", "common", "total"):
        assert np.array_equal(dev_hits[f], host_hits[f]), f
    assert np.array_equal(dev_hits["ani"].view(np.uint32), host_hits["ani"].view(np.uint32))


def test_packed_genomes_without_the_validity_bitmap(gb):
    """galah_b200_cluster_packed_sparse: the validity bitmap built on the device from the genomes' lengths and a list
    of invalid ranges gives the clusters (and hit counts) of the call that is handed the bitmap -- on genomes with
    N runs, a range that ends a genome, ranges sharing a bitmap word, and a genome whose length is not a multiple of 128."""
    rng = np.random.default_rng(11)
    n, L = 24, 150_037
    padded = (L + 127) // 128 * 128
    founders = [rng.integers(0, 4, L, dtype=np.uint8) for _ in range(6)]
    codes = np.zeros((n, padded), np.uint8)
    valid = np.zeros((n, padded), bool)
    begins, ends = [], []
    for g in range(n):
        seq = founders[g % 6].copy()
        pos = rng.choice(L, size=L // 50, replace=False)
        seq[pos] = (seq[pos] + 1 + (g // 6)) % 4
        codes[g, :L] = seq
        valid[g, :L] = True
        cuts = [(1000 + 37 * g, 1000 + 37 * g + 5), (1000 + 37 * g + 9, 1000 + 37 * g + 21), (70_000, 70_400)]
        if g % 5 == 0:
            cuts.append((L - 300, L))           # the genome's tail is N
        for b, e in cuts:
            valid[g, b:e] = False
            begins.append(g * padded + b); ends.append(g * padded + e)
    flat = codes.reshape(-1).astype(np.uint32)
    seq2 = np.zeros(n * padded // 16, np.uint32)
    for k in range(16):
        seq2 |= flat[k::16] << np.uint32(2 * k)
    vbits = np.packbits(valid.reshape(-1), bitorder="little").view(np.uint32)
    base_off = np.arange(n + 1, dtype=np.uint64) * np.uint64(padded)
    lengths = np.full(n, L, np.uint64)
    want, info_w = gb.cluster_packed(seq2, vbits, base_off, lengths)
    got, info_g = gb.cluster_packed_sparse(seq2, (np.array(begins, np.uint64), np.array(ends, np.uint64)), base_off, lengths)
    assert got == want
    assert info_g["n_precluster_hits"] == info_w["n_precluster_hits"] and info_g["n_ani_pairs"] == info_w["n_ani_pairs"]
    # bit for bit: the K1 sketch rows and the K3 seeds made from the device-built bitmap equal those made from the uploaded one
    import torch
    dev = torch.device("cuda", 0)
    rows, seeds = [], []
    for sparse in (False, True):
        idx = gb.AniIndex()
        table = torch.zeros((n, 1000), dtype=torch.int64, device=dev)
        counts = torch.zeros(n, dtype=torch.int32, device=dev)
        if sparse:
            idx.ingest_packed_sparse(seq2.ctypes.data, (np.array(begins, np.uint64), np.array(ends, np.uint64)), base_off, lengths,
                                     table.data_ptr(), counts.data_ptr())
        else:
            idx.ingest_packed(seq2.ctypes.data, vbits.ctypes.data, base_off, lengths, table.data_ptr(), counts.data_ptr())
        torch.cuda.synchronize()
        rows.append((table.cpu().numpy().copy(), counts.cpu().numpy().copy()))
        seeds.append([idx.seeds(g) for g in (0, 5, n - 1)])
        idx.close()
    assert np.array_equal(rows[0][0], rows[1][0]) and np.array_equal(rows[0][1], rows[1][1])
    for a, b in zip(seeds[0], seeds[1]):
        assert all(np.array_equal(x, y) for x, y in zip(a, b))
    with pytest.raises(gb.GalahB200Error):
        gb.cluster_packed_sparse(seq2, (np.array([10, 5], np.uint64), np.array([20, 8], np.uint64)), base_off, lengths)


def test_file_reads_straight_into_the_staging_buffer(gb, tmp_path):
    """Plain files are read by the host threads directly into the pinned K0 staging buffer (sizes from a probe),
    gzip twins are inflated first: an empty file, a file that starts with blank lines, a file shorter than the
    probe's 4 kB head, a 5 kB run of blank lines before the header (undecided head: the batch takes the host packer),
    plain / gz mixed -- same sketch rows as the host packer; a FASTQ file in the batch and a missing file behave as before."""
    rng = np.random.default_rng(11)
    recs = [("c0", random_dna(30_011, rng)), ("c1", random_dna(9_973, rng))]
    paths = [write_fasta(str(tmp_path / "plain.fna"), recs),
             write_fasta(str(tmp_path / "twin.fna.gz"), recs, gz=True),
             write_fasta(str(tmp_path / "short.fna"), [("s", random_dna(700, rng))])]
    empty = str(tmp_path / "empty.fna")
    open(empty, "wb").close()
    blank = str(tmp_path / "blank_start.fna")
    with open(blank, "wb") as f:
        f.write(b"\n\r\n\n" + open(paths[0], "rb").read())
    paths += [empty, blank]

    def both(plist):
        prev = gb.device_ingest(1)
        try:
            dev = gb.sketch_files(plist, 21, 1000)
            gb.device_ingest(0)
            host = gb.sketch_files(plist, 21, 1000)
        finally:
            gb.device_ingest(prev)
        assert np.array_equal(dev[0], host[0]) and np.array_equal(dev[1], host[1])
        return dev
    t, c = both(paths)
    assert c[0] == c[1] == c[4] == 1000 and np.array_equal(t[0], t[1]) and np.array_equal(t[0], t[4])
    assert c[3] == 0 and 0 < c[2] <= 680
    long_blank = str(tmp_path / "long_blank.fna")
    with open(long_blank, "wb") as f:
        f.write(b"\n" * 5000 + open(paths[0], "rb").read())
    t2, c2 = both(paths + [long_blank])
    assert np.array_equal(t2[5], t[0]) and np.array_equal(t2[:5], t)
    fq = str(tmp_path / "reads.fq")
    with open(fq, "wb") as f:
        f.write(b"@r1\n" + recs[0][1][:5000] + b"\n+\n" + b"I" * 5000 + b"\n")
    t3, c3 = both([paths[0], fq])
    assert np.array_equal(t3[0], t[0]) and c3[1] > 0
    with pytest.raises(gb.GalahB200Error) as e:
        gb.sketch_files([paths[0], str(tmp_path / "missing.fna")], 21, 1000)
    assert "Failed to open fasta file" in str(e.value)
