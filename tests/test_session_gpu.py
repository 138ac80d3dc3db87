"""The trait-shaped boundary (include/galah_b200.h "session"; Python mirror in galah_b200/api.py):
an UNMODIFIED galah::clusterer::cluster() calls preclusterer.distances*(paths) once and then
clusterer.calculate_ani(fasta1, fasta2) per pair from many threads (src/clusterer.rs:14-152).
These tests drive the mirror classes in that order and read like src/clusterer.rs:537-824."""
import os
import shutil
import threading

import numpy as np
import pytest

import oracle
from conftest import GOLDEN
from stage2_cases import AB, MAG52

pytestmark = pytest.mark.gpu
ABP = [os.path.join(GOLDEN, f) for f in AB]


def test_finch_skani_through_the_traits_matches_the_reference_outcome(gb):
    """src/clusterer.rs:662-690 (finch 0.9 + skani 99 on abisko4): {0,1,3},{2}; the first calculate_ani
    evaluates every precluster hit in ONE launch, the rest are lookups."""
    session = gb.Session()
    pre = gb.FinchPreclusterer(min_ani=0.9, num_kmers=1000, kmer_length=21, session=session)
    cl = gb.SkaniClusterer(threshold=99.0, min_aligned_threshold=0.2, small_genomes=False, session=session)
    assert pre.method_name() == "finch" and cl.method_name() == "skani" and cl.get_ani_threshold() == 99.0
    clusters, info = gb.cluster_with(ABP, pre, cl)
    assert sorted(sorted(c) for c in clusters) == [[0, 1, 3], [2]]
    st = session.stats()
    assert st["n_indexed"] == 4 and st["n_launches"] <= 2 and st["n_pairs_computed"] >= 6
    # the same values as the one-call pipeline and as the oracle (query = the first path)
    gen = [oracle.AniGenome.from_file(p) for p in ABP]
    for a, b in ((0, 1), (2, 1), (3, 0)):
        exp = oracle.ani_pair(gen[a], gen[b], np.float32(0.2) * np.float32(100.0))[0]
        assert np.float32(cl.calculate_ani(ABP[a], ABP[b])) == np.float32(exp), (a, b)
    session.close()


def test_lazy_path_without_the_clusterer_hint_and_pairs_that_were_never_hits(gb):
    session = gb.Session()
    pre = gb.FinchPreclusterer(0.9, session=session)
    cl = gb.SkaniClusterer(95.0, 0.15, session=session)  # initialise() not called: no hint
    hits = pre.distances(ABP[:3])
    assert len(hits) == 3 and session.stats()["n_indexed"] == 0
    v01 = cl.calculate_ani(ABP[0], ABP[1])
    st = session.stats()
    assert st["n_indexed"] == 3 and st["n_launches"] == 1 and st["n_pairs_computed"] == 3
    # a genome the preclusterer never saw: indexed on demand
    v30 = cl.calculate_ani(ABP[3], ABP[0])
    assert session.stats()["n_indexed"] == 4 and v30 > 99.0 and v01 > 99.0
    # unrelated genomes: skani prints no row -> 0.0, still Some
    other = os.path.join(GOLDEN, MAG52)
    assert cl.calculate_ani(ABP[0], other) == 0.0
    session.close()


def test_calculate_ani_is_reentrant(gb):
    """src/clusterer.rs:267-293: calculate_ani is called from nested rayon workers."""
    session = gb.Session()
    pre = gb.FinchPreclusterer(0.9, session=session)
    cl = gb.SkaniClusterer(95.0, 0.15, session=session)
    cl.initialise()
    pre.distances(ABP)
    want = {(a, b): None for a in range(4) for b in range(4) if a != b}
    out, errs = {}, []

    def worker(keys):
        try:
            for a, b in keys:
                out[(a, b)] = cl.calculate_ani(ABP[a], ABP[b])
        except Exception as e:  # pragma: no cover
            errs.append(e)
    keys = list(want)
    th = [threading.Thread(target=worker, args=(keys[t::8] * 3,)) for t in range(8)]
    for t in th:
        t.start()
    for t in th:
        t.join()
    assert not errs and set(out) == set(want)
    serial = gb.Session()
    cl2 = gb.SkaniClusterer(95.0, 0.15, session=serial)
    for k in keys:
        assert np.float32(out[k]) == np.float32(cl2.calculate_ani(ABP[k[0]], ABP[k[1]])), k
    session.close(); serial.close()


def test_reference_panics_are_mirrored(gb, tmp_path):
    """src/finch.rs:15, 40; src/clusterer.rs:39-41; src/skani.rs:116-121, 243-245, 518-520, 696-698."""
    s = gb.Session()
    with pytest.raises(gb.GalahB200Error, match="Low-memory clustering currently only supported with skani preclusterer"):
        gb.FinchPreclusterer(0.9, low_memory=True, session=s).distances(ABP[:2])
    with pytest.raises(gb.GalahB200Error, match="Reference genome clustering currently only supported with skani preclusterer"):
        gb.FinchPreclusterer(0.9, session=s).distances_with_references(ABP[:2], ABP[:1])
    assert len(gb.FinchPreclusterer(0.9, session=s).distances_contigs(ABP[:1], ["a", "b"])) == 0
    with pytest.raises(gb.GalahB200Error, match="finch does not support contig comparisons"):
        gb.cluster_with(ABP[:1], gb.FinchPreclusterer(0.9, session=s), gb.SkaniClusterer(95.0, session=s),
                        cluster_contigs=True, contig_names=["a"])
    with pytest.raises(gb.GalahB200Error, match="less than 85%. Provided: 80"):
        gb.SkaniPreclusterer(80.0, session=s).distances(ABP[:2])
    with pytest.raises(gb.GalahB200Error, match="does not support small genomes with low-memory preclustering"):
        gb.SkaniPreclusterer(90.0, small_genomes=True, low_memory=True, session=s).distances(ABP[:2])
    with pytest.raises(gb.GalahB200Error, match="does not support small genomes with reference genome preclustering"):
        gb.SkaniPreclusterer(90.0, small_genomes=True, session=s).distances_with_references(ABP[:2], ABP[:1])
    with pytest.raises(gb.GalahB200Error, match="threshold > 1.0"):
        gb.SkaniClusterer(0.95, session=s).initialise()
    s.close()


def test_odd_num_kmers(gb):
    """finch::distances takes any usize num_kmers (src/finch.rs:48-53)."""
    s = gb.Session()
    got = gb.FinchPreclusterer(0.9, num_kmers=999, session=s).distances(ABP)
    sk = [oracle.sketch_fasta(p, 21, 999) for p in ABP]
    table, counts = oracle.pack_table(sk, 1000)
    exp = oracle.prefilter(table, counts, 21, 0.9)
    assert len(got) == len(exp) == 6
    for f in ("i", "j", "common", "total"):
        assert np.array_equal(got[f], exp[f]), f
    assert np.array_equal(got["ani"].view(np.uint32), exp["ani"].view(np.uint32))
    s.close()


def units_of(paths):
    return [oracle.load_codes(p) for p in paths]


def test_distances_with_references_matches_the_oracle(gb, tmp_path):
    """src/skani.rs:502-687 on the genomes of tests/test_cmdline.rs:734-748: the reference (set2/1mbp)
    first, then the two inputs; only reference x non-reference pairs, the non-reference is the query."""
    ref = str(tmp_path / "set2_1mbp.fna.gz")
    shutil.copy(os.path.join(GOLDEN, "set1_1mbp.fna.gz"), ref)  # set2/1mbp.fna is byte-identical to set1/1mbp.fna
    combined = [ref, os.path.join(GOLDEN, "set1_500kb.fna.gz"), os.path.join(GOLDEN, "set1_1mbp.fna.gz")]
    s = gb.Session()
    pre = gb.SkaniPreclusterer(threshold=95.0, min_aligned_threshold=0.15, session=s)
    got = pre.distances_with_references(combined, [ref])
    exp = oracle.skani_distances(units_of(combined), 95.0, np.float32(0.15) * np.float32(100.0), variant="references",
                                 is_ref=[1, 0, 0])
    assert [(int(g["i"]), int(g["j"])) for g in got] == [(e[0], e[1]) for e in exp] == [(0, 1), (0, 2)]
    for g, e in zip(got, exp):
        assert np.float32(g["ani"]).view(np.uint32) == np.float32(e[4]).view(np.uint32)
    clusters, _ = gb.cluster_with(combined, pre, gb.SkaniClusterer(95.0, 0.15, session=s), reference_genomes=[ref])
    assert clusters == [[0, 1, 2]]
    s.close()


def test_low_memory_takes_the_later_record(gb):
    """src/skani.rs:229-377: search of all against all, both records of a pair land on one key."""
    s = gb.Session()
    got = gb.SkaniPreclusterer(90.0, 0.2, low_memory=True, session=s).distances(ABP)
    exp = oracle.skani_distances(units_of(ABP), 90.0, 20.0, variant="lowmem")
    tri = oracle.skani_distances(units_of(ABP), 90.0, 20.0)
    assert len(got) == len(exp) == 6
    for g, e in zip(got, exp):
        assert (int(g["i"]), int(g["j"])) == (e[0], e[1])
        assert np.float32(g["ani"]).view(np.uint32) == np.float32(e[4]).view(np.uint32)
    assert any(a[4] != b[4] for a, b in zip(exp, tri))  # the orientation matters
    s.close()


def test_contig_names_and_their_order(gb, tmp_path):
    """src/cluster_argument_parsing.rs:596-629 + src/skani.rs:460-474: names are header lines up to the
    first TAB; the cache is indexed by position in contig_names, whatever the order."""
    p = os.path.join(GOLDEN, "contigs", "contigs_rep_bug.fna.gz")
    names = gb.contig_names([p])
    assert names == ["k141_313035 flag=1 multi=13.9893 len=27966", "k141_401621 flag=1 multi=12.7497 len=42088",
                     "NODE_1070_length_34582_cov_11.872969"]
    s = gb.Session()
    pre = gb.SkaniPreclusterer(95.0, 0.15000001, small_genomes=False, session=s)
    fwd = pre.distances_contigs([p], names)
    perm = [names[2], names[0], names[1]]
    rev = pre.distances_contigs([p], perm)
    pos = {n: x for x, n in enumerate(perm)}
    want = sorted((min(pos[names[int(h["i"])]], pos[names[int(h["j"])]]), max(pos[names[int(h["i"])]], pos[names[int(h["j"])]]),
                   float(h["ani"])) for h in fwd)
    assert [(int(h["i"]), int(h["j"]), float(h["ani"])) for h in rev] == want and len(fwd) == 3
    clusters, _ = gb.cluster_with([p], pre, gb.SkaniClusterer(95.0, 0.15000001, session=s), cluster_contigs=True,
                                  contig_names=names)
    assert clusters == [[0, 1, 2]]  # tests/test_cmdline.rs:569-588
    with pytest.raises(gb.GalahB200Error, match="Failed to find contig name in contig_names"):
        pre.distances_contigs([p], names[:2])
    dup = tmp_path / "dup.fna"
    dup.write_bytes(b">a\tx\nACGT\n>b\nACGT\n>a\ty\nACGT\n")
    with pytest.raises(gb.GalahB200Error, match="Duplicate contig name found in file"):
        gb.contig_names([str(dup)])
    s.close()
