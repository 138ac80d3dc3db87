"""Host clustering engine (C++ behind the C ABI) vs the pure-Python restatement of
galah::clusterer::cluster() in oracle/cluster_oracle.py.  No GPU needed: the engine is host logic."""
import numpy as np
import pytest

import galah_b200 as gb
from oracle import cluster_oracle as co


def make_hits(pairs):
    h = np.zeros(len(pairs), gb.PAIR_DTYPE)
    for x, (i, j, a) in enumerate(pairs):
        h[x]["i"], h[x]["j"], h[x]["ani"] = i, j, a
    return h


def test_transform_ids_reference_vectors():
    """src/sorted_pair_genome_distance_cache.rs:69-114 (the reference's own unit tests)."""
    c = co.SortedPairGenomeDistanceCache()
    c.insert((1, 2), np.float32(0.5))
    c.insert((3, 2), np.float32(0.6))
    assert c.get((2, 1)) == (True, np.float32(0.5)) and c.get((2, 3)) == (True, np.float32(0.6))
    assert not c.contains_key((1, 3))
    t = c.transform_ids([1, 2])
    assert t.internal == {(0, 1): np.float32(0.5)}
    t = c.transform_ids([2, 3, 5])
    assert t.internal == {(0, 1): np.float32(0.6)}


def test_hand_worked_two_stage():
    """5 genomes, one precluster {0,1,2,3} + singleton {4}.  Stage-2 ANI: 1~0 96, 2~0 94, 2~1 97,
    3~0 95.5, 3~2 99.  Representatives at 95: 0 and 2 (1 joins 0's candidates only; 2 sees only
    rep 0 at 94; 3 tries rep 0 first -- lowest precluster ANI -- and stops at 95.5).  Membership is
    by BEST ANI over all representatives with a precluster hit: 1 -> 2 (97 > 96), 3 -> 2 (99)."""
    hits = make_hits([(0, 1, 0.97), (0, 2, 0.91), (1, 2, 0.95), (0, 3, 0.93), (2, 3, 0.99)])
    ani = {(0, 1): 96.0, (0, 2): 94.0, (1, 2): 97.0, (0, 3): 95.5, (2, 3): 99.0}
    calls = []

    def f(rep, g):
        calls.append((rep, g))
        return ani[(min(rep, g), max(rep, g))]

    clusters, info = gb.cluster_from_distances(5, hits, 95.0, f)
    assert clusters == [[0], [2, 1, 3], [4]]
    assert info["n_preclusters"] == 2 and info["largest_precluster"] == 4
    # rep stage: (0,1); (0,2); (0,3) passes -> stop, (2,3) not computed there; membership: (2,1), (2,3)
    assert calls == [(0, 1), (0, 2), (0, 3), (2, 1), (2, 3)]
    exp, einfo = co.cluster(5, [(h["i"], h["j"], h["ani"]) for h in hits], 95.0, f)
    assert exp == clusters and einfo["ani_calls"] == info["ani_calls"]


def test_skip_clusterer_uses_precluster_ani():
    hits = make_hits([(0, 1, 96.0), (1, 2, 97.0), (0, 2, 90.0), (3, 4, 99.5)])
    clusters, info = gb.cluster_from_distances(6, hits, 95.0, None, skip_clusterer=True)
    exp, _ = co.cluster(6, [(h["i"], h["j"], h["ani"]) for h in hits], 95.0, None, skip_clusterer=True)
    # precluster {0,1,2}: 1 is not a rep (96 vs rep 0); 2 vs rep 0 is 90 -> rep; 1 then joins its
    # BEST rep, 2 (97 > 96); then {3,4}; then {5}
    assert clusters == exp == [[0], [2, 1], [3, 4], [5]]
    assert info["ani_calls"] == 0


def test_none_ani_and_unwrap_panic():
    """calculate_ani -> None everywhere: every genome becomes a representative (None never reaches
    the threshold), so nobody needs a membership and nothing panics."""
    hits = make_hits([(0, 1, 0.95), (1, 2, 0.95)])
    clusters, _ = gb.cluster_from_distances(3, hits, 95.0, lambda r, g: None)
    assert clusters == [[0], [1], [2]]
    # a non-rep whose only ANIs are None at membership time cannot happen in the reference either:
    # it became a non-rep because some ANI >= threshold was cached.


def test_zero_genomes_mirrors_reference_panic():
    with pytest.raises(gb.GalahB200Error) as e:
        gb.cluster_from_distances(0, make_hits([]), 95.0, lambda r, g: 99.0)
    assert "index out of bounds" in str(e.value)


@pytest.mark.parametrize("seed", range(12))
def test_random_graphs_match_oracle(seed):
    rng = np.random.default_rng(seed)
    n = int(rng.integers(1, 60))
    fam = rng.integers(0, max(1, n // 4), size=n)
    pairs, ani2 = [], {}
    for i in range(n):
        for j in range(i + 1, n):
            same = fam[i] == fam[j]
            if rng.uniform() < (0.7 if same else 0.03):
                pairs.append((i, j, float(np.float32(rng.uniform(0.9, 1.0)))))
            # stage-2 ANI table with ties and Nones
            v = rng.choice([None, 94.0, 95.0, 96.5, 99.0, float(np.float32(rng.uniform(90, 100)))],
                           p=[0.05, 0.15, 0.2, 0.2, 0.1, 0.3])
            ani2[(i, j)] = v
    hits = make_hits(pairs)

    def f(rep, g):
        return ani2[(min(rep, g), max(rep, g))]

    for skip in (False, True):
        thr = 95.0 if not skip else 0.95
        try:
            exp, einfo = co.cluster(n, pairs, thr, f, skip_clusterer=skip)
        except RuntimeError as ex:
            with pytest.raises(gb.GalahB200Error) as e:
                gb.cluster_from_distances(n, hits, thr, f, skip_clusterer=skip)
            assert "Option::unwrap()" in str(e.value) and "unwrap" in str(ex)
            continue
        got, info = gb.cluster_from_distances(n, hits, thr, f, skip_clusterer=skip)
        assert got == exp
        assert info == einfo
        assert sorted(x for c in got for x in c) == list(range(n))


def test_large_sparse_is_linear_in_hits():
    """100k genomes in families of 10 (450k hits): the engine must not be quadratic in N."""
    import time
    n = 100_000
    pairs = [(f * 10 + a, f * 10 + b, 0.97) for f in range(n // 10) for a in range(10) for b in range(a + 1, 10)]
    hits = make_hits(pairs)
    t0 = time.time()
    clusters, info = gb.cluster_from_distances(n, hits, 0.95, None, skip_clusterer=True)
    dt = time.time() - t0
    assert len(clusters) == n // 10 and all(len(c) == 10 for c in clusters)
    assert info["n_preclusters"] == n // 10 and info["largest_precluster"] == 10
    assert dt < 20.0


def test_ani_table_entry_matches_callback_entry():
    """galah_b200_cluster_from_ani_table == the callback engine fed from the same table == the
    Python restatement, on random family-structured hit lists; unsorted hits are rejected."""
    rng = np.random.default_rng(17)
    n = 400
    pairs = []
    for base in range(0, n, 8):
        fam = list(range(base, min(n, base + 8)))
        for a in fam:
            for b in fam:
                if a < b and rng.uniform() < 0.7:
                    pairs.append((a, b, float(np.float32(rng.uniform(0.9, 1.0)))))
    pairs.sort()
    hits = make_hits(pairs)
    ani = np.round(rng.uniform(90.0, 100.0, len(hits)), 2).astype(np.float32)
    table = {(int(h["i"]), int(h["j"])): float(a) for h, a in zip(hits, ani)}
    f = lambda rep, g: table[(min(rep, g), max(rep, g))]
    got, ginfo = gb.cluster_from_ani_table(n, hits, ani, 95.0)
    via_cb, cinfo = gb.cluster_from_distances(n, hits, 95.0, f)
    exp, einfo = co.cluster(n, [(h["i"], h["j"], h["ani"]) for h in hits], 95.0, f)
    assert got == via_cb == exp
    assert ginfo["ani_calls"] == cinfo["ani_calls"] == einfo["ani_calls"]
    with pytest.raises(gb.GalahB200Error):
        gb.cluster_from_ani_table(n, hits[::-1].copy(), ani, 95.0)
    empty, _ = gb.cluster_from_ani_table(3, hits[:0], ani[:0], 95.0)
    assert empty == [[0], [1], [2]]


def test_threaded_sweeps_match_the_serial_engine():
    """With ANI values served from tables the two sweeps run on the host threads (>= 4096 preclusters);
    a Python callback keeps the serial path: both give the same clusters, order and call counts, also for
    the two-orientation tables and for skip_clusterer."""
    rng = np.random.default_rng(23)
    n = 48_000
    pairs = []
    for base in range(0, n, 8):
        for a in range(base, base + 8):
            for b in range(a + 1, base + 8):
                if rng.uniform() < 0.6:
                    pairs.append((a, b, float(np.float32(rng.uniform(0.9, 1.0)))))
    hits = make_hits(pairs)
    fwd = np.round(rng.uniform(92.0, 100.0, len(hits)), 2).astype(np.float32)
    rev = np.round(rng.uniform(92.0, 100.0, len(hits)), 2).astype(np.float32)
    key = {(int(h["i"]), int(h["j"])): x for x, h in enumerate(hits)}

    def f2(rep, g):
        x = key[(min(rep, g), max(rep, g))]
        return float(fwd[x] if rep < g else rev[x])
    got, ginfo = gb.cluster_from_ani_tables(n, hits, fwd, rev, 95.0)
    want, winfo = gb.cluster_from_distances(n, hits, 95.0, f2)
    assert got == want and ginfo["ani_calls"] == winfo["ani_calls"]
    assert ginfo["n_preclusters"] >= 4096
    one, oinfo = gb.cluster_from_ani_table(n, hits, fwd, 95.0)
    want1, w1info = gb.cluster_from_distances(n, hits, 95.0, lambda rep, g: float(fwd[key[(min(rep, g), max(rep, g))]]))
    assert one == want1 and oinfo["ani_calls"] == w1info["ani_calls"]
    pct = hits.copy()
    pct["ani"] = fwd
    skip, sinfo = gb.cluster_from_distances(n, pct, 95.0, None, skip_clusterer=True)
    # blocks of 8 genomes are closed under the hits: the oracle (quadratic) checks the first 1,600 genomes
    m = 1600
    sub = pct[pct["j"] < m]
    exp, _ = co.cluster(m, [(int(h["i"]), int(h["j"]), h["ani"]) for h in sub], 95.0, None, skip_clusterer=True)
    assert [c for c in skip if c[0] < m] == exp and sinfo["ani_calls"] == 0
    assert sorted(g for c in skip for g in c) == list(range(n))


def _batch_of(f, log=None):
    def batch(reps, genomes):
        if log is not None:
            log.append(list(zip(reps.tolist(), genomes.tolist())))
        return [f(int(r), int(g)) for r, g in zip(reps, genomes)]
    return batch


@pytest.mark.parametrize("seed", range(16))
def test_wave_engine_matches_serial_engine_and_oracle(seed):
    """galah_b200_cluster_from_distances_batched (stage 2 asked for in waves) == the callback engine == the
    Python restatement: clusters, order, ani_calls -- with orientation-dependent values, ties at the threshold
    and Nones; every pair is asked for once, with the representative as the query; also when the wave budget
    runs out (max_waves 1, 2, 3: the one-batch finish)."""
    rng = np.random.default_rng(1000 + seed)
    n = int(rng.integers(1, 90))
    fam = rng.integers(0, max(1, n // 6), size=n)
    pairs, ani2 = [], {}
    p_none = 0.0 if seed % 2 else 0.06
    for i in range(n):
        for j in range(i + 1, n):
            if rng.uniform() < (0.8 if fam[i] == fam[j] else 0.02):
                pairs.append((i, j, float(np.float32(rng.uniform(0.9, 1.0)))))
            for key in ((i, j), (j, i)):
                ani2[key] = rng.choice([None, 94.0, 95.0, 96.5, 99.0, float(np.float32(rng.uniform(90, 100)))],
                                       p=[p_none, 0.2 - p_none, 0.2, 0.2, 0.1, 0.3])
    hits = make_hits(pairs)
    f = lambda rep, g: ani2[(rep, g)]
    try:
        exp, einfo = co.cluster(n, pairs, 95.0, f)
    except RuntimeError:
        with pytest.raises(gb.GalahB200Error) as e:
            gb.cluster_from_distances_batched(n, hits, 95.0, _batch_of(f))
        assert "Option::unwrap()" in str(e.value)
        return
    serial, sinfo = gb.cluster_from_distances(n, hits, 95.0, f)
    log = []
    got, info = gb.cluster_from_distances_batched(n, hits, 95.0, _batch_of(f, log))
    assert got == serial == exp
    assert info["ani_calls"] == sinfo["ani_calls"] == einfo["ani_calls"]
    assert info["ani_waves"] == len(log)
    asked = [p for wave in log for p in wave]
    hit_set = {(i, j) for i, j, _ in pairs}
    assert len(asked) == len(set(asked)) and all((min(p), max(p)) in hit_set for p in asked)
    reps = {c[0] for c in got}
    assert all(r in reps for r, _ in asked)  # the query is always a representative
    for w in (1, 2, 3):
        short, winfo = gb.cluster_from_distances_batched(n, hits, 95.0, _batch_of(f), max_waves=w)
        assert short == exp and winfo["ani_waves"] <= w + 2


def test_wave_engine_on_one_clade_asks_for_representatives_only():
    """One clade of 600 genomes, every pair a precluster hit (179,700 hits), everything within the threshold of
    genome 0: the reference evaluates 599 pairs (genome 0 against everybody), and so do the waves -- in one batch."""
    n = 600
    pairs = [(a, b, 0.99) for a in range(n) for b in range(a + 1, n)]
    hits = make_hits(pairs)
    log = []
    got, info = gb.cluster_from_distances_batched(n, hits, 95.0, _batch_of(lambda r, g: 98.0, log))
    assert got == [list(range(n))]
    assert info["ani_calls"] == n - 1 and info["ani_waves"] == 1 and log[0] == [(0, g) for g in range(1, n)]
    # two sub-clades that are far from each other: genome 0 and the first genome of the other sub-clade
    far = lambda r, g: 98.0 if (r % 2) == (g % 2) else 80.0
    got2, info2 = gb.cluster_from_distances_batched(n, hits, 95.0, _batch_of(far))
    ser2, sinfo2 = gb.cluster_from_distances(n, hits, 95.0, far)
    assert got2 == ser2 == [list(range(0, n, 2)), list(range(1, n, 2))]
    assert info2["ani_calls"] == sinfo2["ani_calls"] == (n - 1) + (n - 2) and info2["ani_waves"] == 2


def test_wave_engine_callback_failure_aborts():
    hits = make_hits([(0, 1, 0.95), (1, 2, 0.95)])

    def boom(reps, genomes):
        raise KeyError("backend down")
    with pytest.raises(KeyError):
        gb.cluster_from_distances_batched(3, hits, 95.0, boom)


def test_large_dense_list_threaded_adjacency():
    """> 2^19 hits: the adjacency rows are filled by several host threads (row ranges).  One clade of 1,100 genomes
    (604,450 hits) in two far-apart halves; table engine, wave engine and skip_clusterer agree with the closed form."""
    n = 1100
    iu = np.triu_indices(n, 1)
    hits = np.zeros(len(iu[0]), gb.PAIR_DTYPE)
    hits["i"], hits["j"] = iu[0], iu[1]
    same = (hits["i"] % 2) == (hits["j"] % 2)
    hits["ani"] = np.where(same, 98.0, 80.0).astype(np.float32)
    assert len(hits) > (1 << 19)
    want = [list(range(0, n, 2)), list(range(1, n, 2))]
    ani = hits["ani"].copy()
    got, info = gb.cluster_from_ani_tables(n, hits, ani, ani, 95.0)
    assert got == want and info["ani_calls"] == (n - 1) + (n - 2)
    waves, winfo = gb.cluster_from_distances_batched(
        n, hits, 95.0, lambda reps, genomes: np.where((reps % 2) == (genomes % 2), 98.0, 80.0).tolist())
    assert waves == want and winfo["ani_calls"] == info["ani_calls"] and winfo["ani_waves"] == 2
    skip, _ = gb.cluster_from_distances(n, hits, 95.0, None, skip_clusterer=True)
    assert skip == want
    # unsorted input with a duplicate key: the later record wins (BTreeMap::insert)
    shuffled = hits[np.random.default_rng(5).permutation(len(hits))]
    dup = shuffled[:1].copy()
    extra = np.concatenate([dup, shuffled])
    extra[0]["ani"] = 1.0  # overwritten by the later record of the same key
    again, _ = gb.cluster_from_distances(n, extra, 95.0, None, skip_clusterer=True)
    assert again == want


def test_wave_engine_long_chains_take_the_one_batch_finish():
    """A path graph of 60 genomes whose neighbours are all BELOW the threshold: every genome is a representative and
    genome i is only settled once i - 1 has been applied -- 60 dependent waves.  The default budget (16 waves) ends in
    the one-batch finish; a budget of 100 walks the whole chain.  Same clusters and -- with the budget large enough --
    the same calculate_ani count as the serial engine.  Then the same path with every third link ABOVE the threshold."""
    n = 60
    pairs = [(i, i + 1, 0.95) for i in range(n - 1)]
    hits = make_hits(pairs)
    for rule in (lambda r, g: 90.0, lambda r, g: 97.0 if min(r, g) % 3 == 0 else 90.0):
        serial, sinfo = gb.cluster_from_distances(n, hits, 95.0, rule)
        exp, einfo = co.cluster(n, pairs, 95.0, rule)
        assert serial == exp and sinfo["ani_calls"] == einfo["ani_calls"]
        short, winfo = gb.cluster_from_distances_batched(n, hits, 95.0, _batch_of(rule))
        assert short == serial and winfo["ani_waves"] <= 18
        full, finfo = gb.cluster_from_distances_batched(n, hits, 95.0, _batch_of(rule), max_waves=100)
        assert full == serial and finfo["ani_calls"] == sinfo["ani_calls"]
    # all below the threshold: one wave per genome of the chain but the last
    alone, ainfo = gb.cluster_from_distances_batched(n, hits, 95.0, _batch_of(lambda r, g: 90.0), max_waves=100)
    assert alone == [[g] for g in range(n)] and ainfo["ani_waves"] == n - 1
