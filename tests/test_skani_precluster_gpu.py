"""a11: the skani-preclusterer replacement (marker screen + ANI) and contig mode vs the oracle."""
import os

import numpy as np
import pytest

import oracle
from conftest import GOLDEN
from test_ani_oracle import mutate
from util import random_dna, write_fasta

pytestmark = pytest.mark.gpu


def units_of(paths, per_record=False):
    units = []
    for p in paths:
        codes, rs, re_ = oracle.load_codes(p)
        if per_record:
            for a, b in zip(rs, re_):
                units.append((codes[int(a):int(b)].copy(), np.array([0], np.uint64), np.array([int(b - a)], np.uint64)))
        else:
            units.append((codes, rs, re_))
    return units


def check(got, exp):
    assert [(int(g["i"]), int(g["j"])) for g in got] == [(e[0], e[1]) for e in exp]
    for g, e in zip(got, exp):
        assert (int(g["common"]), int(g["total"])) == (e[2], e[3])
        assert np.float32(g["ani"]).view(np.uint32) == np.float32(e[4]).view(np.uint32)


@pytest.mark.parametrize("small", [False, True])
def test_genome_mode_matches_oracle(gb, tmp_path, small):
    rng = np.random.default_rng(11)
    founders = [random_dna(300_000, rng) for _ in range(3)]
    recs = []
    for f, base in enumerate(founders):
        for m, rate in enumerate((0.0, 0.02, 0.07, 0.13)):
            recs.append((f"f{f}m{m}", [mutate(base, rate, rng) if rate else base]))
    recs.append(("two_contigs", [founders[0][:140_000], founders[1][100_000:260_000]]))
    recs.append(("tiny", [random_dna(500, rng)]))
    paths = [write_fasta(str(tmp_path / f"{n}.fna"), [(f"{n}_{i}", r) for i, r in enumerate(rs)]) for n, rs in recs]
    for thr in (90.0, 97.0):
        got, n_units = gb.skani_distances(paths, thr, 15.0, small_genomes=small)
        assert n_units == len(paths)
        exp = oracle.skani_distances(units_of(paths), thr, 15.0, small_genomes=small)
        check(got, exp)
        assert len(exp) >= 6
    # the 13 % members are below every threshold the reference accepts; 2 % members pass at 97
    pairs = {(int(g["i"]), int(g["j"])) for g in gb.skani_distances(paths, 97.0, 15.0, small_genomes=small)[0]}
    assert (0, 1) in pairs and (0, 3) not in pairs and (0, 4) not in pairs


def test_threshold_below_85_mirrors_reference_panic(gb, tmp_path):
    p = write_fasta(str(tmp_path / "a.fna"), [("a", b"ACGT" * 100)])
    with pytest.raises(gb.GalahB200Error) as e:
        gb.skani_distances([p, p], 80.0, 15.0)
    # src/skani.rs:116-121, pinned by tests/test_cmdline.rs:409-414
    assert "Error: skani produces inaccurate results with ANI less than 85%. Provided: 80" in str(e.value)


def test_contig_mode_and_clusters(gb, tmp_path):
    """--cluster-contigs --small-contigs: every record is a unit; skip_clusterer greedy on the hits."""
    rng = np.random.default_rng(13)
    base = random_dna(60_000, rng)
    contigs = [("ref", base), ("100ANI", base), ("96ANI", mutate(base, 0.04, rng)), ("94ANI", mutate(base, 0.06, rng)),
               ("other", random_dna(60_000, rng)), ("other_98", None)]
    contigs[5] = ("other_98", mutate(contigs[4][1], 0.02, rng))
    p1 = write_fasta(str(tmp_path / "c1.fna"), contigs[:4])
    p2 = write_fasta(str(tmp_path / "c2.fna"), contigs[4:])
    got, n_units = gb.skani_distances([p1, p2], 95.0, 15.0, small_genomes=True, contigs=True)
    assert n_units == 6
    exp = oracle.skani_distances(units_of([p1, p2], per_record=True), 95.0, 15.0, small_genomes=True,
                                 individual_contigs=True)
    check(got, exp)
    clusters, info = gb.cluster_skani([p1, p2], precluster_ani=95.0, ani=95.0, min_aligned_fraction=15.0,
                                      small_genomes=True, cluster_contigs=True)
    # the shape the reference pins on contigs_specific.fna (tests/test_cmdline.rs:482-507): the 100 %
    # and 96 % contigs join the first, the 94 % one does not
    assert clusters == [[0, 1, 2], [4, 5], [3]]


def test_default_cli_path_on_reference_genomes(gb):
    """skani + skani (galah's CLI default) on the committed abisko4 genomes: one cluster at 95 %,
    {0,1,3},{2} at 99 % -- src/clusterer.rs:692-722 (test_skani_skani_two_clusters_same_ani: 90 / 99 / AF 0.2) pins the 99 % outcome."""
    names = ["73.20120800_S1X.13.fna.gz", "73.20120600_S2D.19.fna.gz", "73.20120700_S3X.12.fna.gz",
             "73.20110800_S2D.13.fna.gz"]
    paths = [os.path.join(GOLDEN, "abisko4", n) for n in names]
    c95, _ = gb.cluster_skani(paths, precluster_ani=90.0, ani=95.0, min_aligned_fraction=20.0)
    c99, _ = gb.cluster_skani(paths, precluster_ani=90.0, ani=99.0, min_aligned_fraction=20.0)
    assert c95 == [[0, 1, 2, 3]]
    assert c99 == [[0, 1, 3], [2]]


def test_device_resident_contigs_match_file_path_and_oracle(gb, tmp_path):
    """galah_b200_skani_distances_packed_device (units packed on the device, marker table never
    leaves it) == the file-based contig-mode call on the same synthetic contigs == the oracle."""
    import torch
    seed, n, L = 1, 60, 30_000
    lay = gb.synth_layout(n, L)
    dev = torch.device("cuda", 0)
    d_seq = torch.zeros(lay["seq2_words"], dtype=torch.int32, device=dev)
    d_val = torch.zeros(lay["valid_words"], dtype=torch.int32, device=dev)
    d_off = torch.zeros(n + 1, dtype=torch.int64, device=dev)
    st = torch.cuda.current_stream().cuda_stream
    gb.synth_packed_device(seed, 0, n, L, d_seq.data_ptr(), d_val.data_ptr(), d_off.data_ptr(), st)
    torch.cuda.synchronize()
    base_off = np.arange(n + 1, dtype=np.uint64) * np.uint64(lay["padded"])
    got, info = gb.skani_distances_packed_device(d_seq.data_ptr(), d_val.data_ptr(), d_off.data_ptr(), base_off,
                                                 np.full(n, L, np.uint64), 90.0, 15.0, small_genomes=True, stream=st)
    assert info["n_screened"] >= len(got) > 0
    path = write_fasta(str(tmp_path / "contigs.fna"), [(f"c{g}", oracle.synth_genome(seed, g, L)) for g in range(n)])
    via_files, n_units = gb.skani_distances([path], 90.0, 15.0, small_genomes=True, contigs=True)
    assert n_units == n
    assert len(got) == len(via_files)
    for f in ("i", "j", "common", "total"):
        assert np.array_equal(got[f], via_files[f]), f
    assert np.array_equal(got["ani"].view(np.uint32), via_files["ani"].view(np.uint32))
    exp = oracle.skani_distances(units_of([path], per_record=True), 90.0, 15.0, small_genomes=True,
                                 individual_contigs=True)
    check(got, exp)
    # members of a synthetic family of 10 only ever pair with each other
    assert np.all(got["i"] // 10 == got["j"] // 10)


def test_units_with_more_markers_than_one_sort_holds(gb, tmp_path):
    """The reference (skani) has no unit size limit (src/skani.rs:109-225).  A 3.6 Mbp genome at
    --small-genomes marker density (1/200) holds ~18,000 markers, more than one 16,384-entry
    shared-memory sort: its row is finished in two value-range partitions and the K2 join runs at
    a row stride of 32,768.  Marker intersections and ANI equal the oracle's."""
    rng = np.random.default_rng(23)
    base = random_dna(3_600_000, rng)
    recs = [("big0", [base]), ("big1", [mutate(base, 0.03, rng)]), ("big2_two_contigs", [base[:2_000_000], base[2_100_000:]]),
            ("other", [random_dna(300_000, rng)])]
    paths = [write_fasta(str(tmp_path / f"{n}.fna"), [(f"{n}_{i}", r) for i, r in enumerate(rs)]) for n, rs in recs]
    got, n_units = gb.skani_distances(paths, 90.0, 15.0, small_genomes=True)
    exp = oracle.skani_distances(units_of(paths), 90.0, 15.0, small_genomes=True)
    assert n_units == 4 and len(exp) == 3
    check(got, exp)
    assert min(int(g["common"]) for g in got) > 5000 and max(int(g["total"]) for g in got) > 16384
