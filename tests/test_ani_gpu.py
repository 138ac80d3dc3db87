"""K3 parity: the CUDA seed-and-chain ANI vs oracle/skani_oracle.c, through the C ABI.  Everything
the kernels produce is integer, so the bar is bit-exact: seed lists, the per-pair accumulators
(sum of fixed-point chunk identities, chunk / chain counts, span and coverage sums), and the f32
bits of the final ANI."""
import os

import numpy as np
import pytest

import oracle
from conftest import GOLDEN
from test_ani_oracle import mutate, revcomp
from util import random_dna, write_fasta

pytestmark = pytest.mark.gpu


def oracle_genome(records, c):
    """records: list of byte strings (may hold N / lowercase) -> AniGenome via a FASTA round trip."""
    import tempfile
    with tempfile.TemporaryDirectory() as d:
        p = write_fasta(os.path.join(d, "g.fna"), [(f"c{i}", r) for i, r in enumerate(records)])
        return oracle.AniGenome.from_file(p, c=c)


INT_FIELDS = ("sum_fx", "n_chunks", "cov_q", "cov_r", "sum_m", "span_m", "span_n", "n_chains")


def check_pairs(idx, genomes, pairs, min_af, c=125, contigs=False):
    got = idx.pairs(np.array(pairs, np.uint32), min_af, individual_contigs=contigs)
    for (a, b), g in zip(pairs, got):
        ints = oracle.ani_pair_integers(genomes[a], genomes[b])  # query = a, the first of the pair
        exp_ints = (ints[0], ints[1], ints[2], ints[3], ints[6], ints[7], ints[8], ints[9])
        assert tuple(int(g[f]) for f in INT_FIELDS) == exp_ints, (a, b)
        exp = oracle.ani_finish(ints, min_af, c, contigs)
        assert np.float32(g["ani"]).view(np.uint32) == np.float32(exp[0]).view(np.uint32), (a, b, g["ani"], exp[0])
        assert abs(g["af_query"] - exp[1]) < 1e-6 and abs(g["af_ref"] - exp[2]) < 1e-6
        assert int(g["estimator"]) == exp[4]
    return got


@pytest.mark.parametrize("small", [False, True])
def test_seeds_and_pairs_match_oracle_on_fasta_files(gb, tmp_path, small):
    rng = np.random.default_rng(3)
    c = 30 if small else 125
    base = random_dna(260_000, rng)
    variants = {
        "founder": [base],
        "mut2": [mutate(base, 0.02, rng)],
        "mut6_contigs": [mutate(base[:70_000], 0.06, rng), mutate(base[70_000:71_000], 0.06, rng),
                         mutate(base[90_000:260_000], 0.06, rng)],
        "revcomp_with_N": [revcomp(base[:120_000]) + b"N" * 37 + base[120_000:200_000].lower()],
        "rearranged": [base[130_000:] + base[:130_000]],
        "unrelated": [random_dna(180_000, rng)],
        "tiny": [random_dna(40, rng), b"ACGT", b""],
    }
    paths, genomes = [], []
    for name, recs in variants.items():
        paths.append(write_fasta(str(tmp_path / f"{name}.fna"), [(f"{name}_{i}", r) for i, r in enumerate(recs)]))
        genomes.append(oracle.AniGenome.from_file(paths[-1], c=c))
    idx = gb.AniIndex(small_genomes=small)
    idx.add_files(paths[:3], threads=2)
    idx.add_files(paths[3:], threads=2)  # second batch appends
    assert len(idx) == len(paths)
    for g, og in enumerate(genomes):
        info = idx.genome(g)
        assert (info["n_seeds"], info["n_chunks"], info["total_len"]) == (len(og.kmer_strand), og.n_chunks, og.total_len)
        ks, sp, ch = idx.seeds(g)
        assert np.array_equal(ks, og.kmer_strand) and np.array_equal(sp, og.spread) and np.array_equal(ch, og.chunk)
    n = len(paths)
    pairs = [(a, b) for a in range(n) for b in range(n) if a != b]
    got = check_pairs(idx, genomes, pairs, 15.0, c)
    res = {p: g for p, g in zip(pairs, got)}
    assert res[(0, 1)]["ani"] > 97.5 and res[(0, 3)]["ani"] == np.float32(100.0) and res[(0, 5)]["ani"] == 0.0
    assert res[(0, 4)]["ani"] == np.float32(100.0)
    # the query is the pair's first genome: swapping the arguments swaps the aligned fractions
    assert abs(res[(0, 2)]["af_query"] - res[(2, 0)]["af_ref"]) < 0.02
    assert abs(float(res[(0, 2)]["ani"]) - float(res[(2, 0)]["ani"])) < 0.3
    check_pairs(idx, genomes, pairs[:8], 99.5, c)  # AF gate
    check_pairs(idx, genomes, pairs[:12], 15.0, c, contigs=True)  # `-i` units: never the span estimator
    idx.close()


def test_repeats_and_max_occurrence_rule(gb, tmp_path):
    """A reference holding 3 and 12 copies of segments: <= 8 occurrences chain, > 8 are masked."""
    rng = np.random.default_rng(5)
    seg_a, seg_b = random_dna(30_000, rng), random_dna(30_000, rng)
    spacer = lambda: random_dna(5_000, rng)
    ref = b"".join(seg_a + spacer() for _ in range(3)) + b"".join(seg_b + spacer() for _ in range(12))
    qry = seg_a + spacer() + seg_b + random_dna(50_000, rng)
    paths = [write_fasta(str(tmp_path / "ref.fna"), [("r", ref)]), write_fasta(str(tmp_path / "q.fna"), [("q", qry)])]
    genomes = [oracle.AniGenome.from_file(p) for p in paths]
    idx = gb.AniIndex()
    idx.add_files(paths)
    got = check_pairs(idx, genomes, [(0, 1), (1, 0)], 0.0)
    assert got[0]["span_n"] > 0 and got[1]["n_chunks"] > 0
    idx.close()


def test_synthetic_families_on_device(gb):
    """Device-resident synthetic genomes (bench path): one family, all 45 pairs, against the oracle
    regenerating the same genomes from (seed, index)."""
    import torch
    seed, n, L = 1, 10, 300_000
    lay = gb.synth_layout(n, L)
    dev = torch.device("cuda", 0)
    d_seq = torch.zeros(lay["seq2_words"], dtype=torch.int32, device=dev)
    d_val = torch.zeros(lay["valid_words"], dtype=torch.int32, device=dev)
    d_off = torch.zeros(n + 1, dtype=torch.int64, device=dev)
    st = torch.cuda.current_stream().cuda_stream
    gb.synth_packed_device(seed, 0, n, L, d_seq.data_ptr(), d_val.data_ptr(), d_off.data_ptr(), st)
    torch.cuda.synchronize()
    base_off = d_off.cpu().numpy().astype(np.uint64)
    idx = gb.AniIndex()
    idx.add_packed_device(d_seq.data_ptr(), d_val.data_ptr(), d_off.data_ptr(), base_off, np.full(n, L, np.uint64), st)
    genomes = [oracle.AniGenome(*oracle.codes_from_ascii(oracle.synth_genome(seed, g, L))) for g in range(n)]
    pairs = [(a, b) for a in range(n) for b in range(a + 1, n)]
    got = check_pairs(idx, genomes, pairs, 15.0)
    # member m has substitution rate {0,.5,1,2,3,4,5,6,8,10} %: ANI to the founder tracks 100 - rate
    rates = [0, .5, 1, 2, 3, 4, 5, 6, 8, 10]
    for (a, b), g in zip(pairs, got):
        if a == 0:
            assert abs(float(g["ani"]) - (100 - rates[b])) < 0.6, (b, g["ani"])
    build_ms, chain_ms = idx.last_timing()
    assert build_ms > 0 and chain_ms > 0
    idx.close()


def test_committed_real_genomes_reproduce_reference_clusters(gb):
    """End to end on the reference's own abisko4 genomes (committed gz copies): finch prefilter on the
    GPU -> stage-2 ANI on the GPU -> greedy engine gives the cluster outputs the reference pins at
    src/clusterer.rs:631-690 (one cluster at 95 %, {0,1,3},{2} at 99 %)."""
    names = ["73.20120800_S1X.13.fna.gz", "73.20120600_S2D.19.fna.gz", "73.20120700_S3X.12.fna.gz",
             "73.20110800_S2D.13.fna.gz"]
    paths = [os.path.join(GOLDEN, "abisko4", n) for n in names]
    hits = gb.finch_distances(paths, 0.9, 1000, 21)
    assert len(hits) == 6  # every pair passes the 0.9 MinHash prefilter
    idx = gb.AniIndex()
    idx.add_files(paths)
    pairs = np.stack([hits["i"], hits["j"]], axis=1)
    res = idx.pairs(pairs, 20.0)
    genomes = [oracle.AniGenome.from_file(p) for p in paths]
    check_pairs(idx, genomes, [tuple(map(int, p)) for p in pairs], 20.0)
    table = {(int(i), int(j)): float(a) for (i, j), a in zip(pairs, res["ani"])}
    f = lambda r, g: table[(min(r, g), max(r, g))]
    c95, _ = gb.cluster_from_distances(4, hits, 95.0, f)
    c99, _ = gb.cluster_from_distances(4, hits, 99.0, f)
    assert [sorted(c) for c in c95] == [[0, 1, 2, 3]]
    assert sorted(sorted(c) for c in c99) == [[0, 1, 3], [2]]
    idx.close()
    # the same through the one-call drop-in for galah::clusterer::cluster()
    c95, info = gb.cluster(paths, precluster_ani=0.9, ani=95.0, min_aligned_fraction=20.0)
    assert c95 == [[0, 1, 2, 3]] and info["n_precluster_hits"] == 6 and info["n_preclusters"] == 1
    c99, _ = gb.cluster(paths, precluster_ani=0.9, ani=99.0, min_aligned_fraction=20.0)
    assert c99 == [[0, 1, 3], [2]]  # representative first, clusters in representative order
    with pytest.raises(gb.GalahB200Error) as e:
        gb.cluster(paths, ani=0.95)  # SkaniClusterer::initialise asserts a percentage (src/skani.rs:696-698)
    assert "threshold > 1.0" in str(e.value)
