"""Every stage-2 outcome the reference's own tests pin (tests/stage2_cases.py), through the CUDA
path behind the C ABI (galah_b200_cluster_files_skani / _cluster_files / _skani_distances /
_ani_pairs) on the committed copies of the reference's fixture genomes, and the committed golden
integers of tests/golden/stage2_golden.json."""
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN
from stage2_cases import CASES, paths_of

pytestmark = pytest.mark.gpu


def gpu_clusters(gb, case):
    paths = paths_of(GOLDEN, case)
    if case.get("low_memory"):
        session = gb.Session()
        pre = gb.SkaniPreclusterer(case["pre_thr"], case["min_af"] / 100.0, case["small"], low_memory=True, session=session)
        cl_ = gb.SkaniClusterer(case["ani"], case["min_af"] / 100.0, case["small"], session=session)
        cl, info = gb.cluster_with(paths, pre, cl_)
    elif case["pre"] == "skani":
        cl, info = gb.cluster_skani(paths, precluster_ani=case["pre_thr"], ani=case["ani"],
                                    min_aligned_fraction=case["min_af"], small_genomes=case["small"],
                                    cluster_contigs=case["contigs"])
    else:
        cl, info = gb.cluster(paths, precluster_ani=case["pre_thr"], ani=case["ani"],
                              min_aligned_fraction=case["min_af"], small_genomes=case["small"])
    return cl


@pytest.mark.parametrize("name", sorted(CASES))
def test_reference_pinned_outcome_on_gpu(gb, name):
    case = CASES[name]
    cl = gpu_clusters(gb, case)
    if case["clusters"] is None:
        assert any(c[0] == 0 for c in cl)
        return
    assert sorted(sorted(c) for c in cl) == case["clusters"], name
    if "order" in case:
        assert sorted(cl) == case["order"], name


@pytest.mark.parametrize("name", ["skani_skani_two_preclusters", "cli_antonio_af60", "cli_min_aligned_fraction_0.2",
                                  "cli_small_genomes_pair"])
def test_genome_pairs_match_golden_integers(gb, name):
    """Whole-genome cases: the kernel's accumulators and the finished ANI == the committed oracle
    output, pair by pair (bit-exact; f32 bits for the ANI)."""
    case = CASES[name]
    gold = json.load(open(os.path.join(GOLDEN, "stage2_golden.json")))[name]
    idx = gb.AniIndex(small_genomes=case["small"])
    idx.add_files(paths_of(GOLDEN, case))
    pairs = np.array([(r["i"], r["j"]) for r in gold], np.uint32)
    got = idx.pairs(pairs, case["min_af"])
    for g, r in zip(got, gold):
        for f in ("sum_fx", "n_chunks", "cov_q", "cov_r", "sum_m", "span_m", "span_n", "n_chains", "estimator"):
            assert int(g[f]) == r[f], (name, r["i"], r["j"], f)
        assert np.float32(g["ani"]).view(np.uint32) == np.float32(r["ani"]).view(np.uint32)
    idx.close()


@pytest.mark.parametrize("name", ["cli_contig_rep_bug_large", "cli_contig_rep_bug_small",
                                  "cli_contig_cluster_specific_small", "cli_contig_cluster_multiple_files_small"])
def test_contig_hits_match_golden(gb, name):
    """Contig cases: the preclusterer's hit list (pairs with ANI >= threshold) == the golden rows at
    or above the threshold that also pass the marker screen; ANI bit-exact."""
    case = CASES[name]
    gold = json.load(open(os.path.join(GOLDEN, "stage2_golden.json")))[name]
    hits, n_units = gb.skani_distances(paths_of(GOLDEN, case), case["pre_thr"], case["min_af"],
                                       small_genomes=case["small"], contigs=True)
    want = {(r["i"], r["j"]): r["ani"] for r in gold if np.float32(r["ani"]) >= np.float32(case["pre_thr"])}
    got = {(int(h["i"]), int(h["j"])): float(h["ani"]) for h in hits}
    # a golden pair can be missing only if the marker screen dropped it; none of the pinned cases has one
    assert got == {k: float(np.float32(v)) for k, v in want.items()}, name


def _both_modes(gb, run):
    """run() with stage 2 on every precluster hit, then in waves (galah_b200_cluster_lazy)."""
    try:
        gb.cluster_lazy(0)
        eager = run()
        gb.cluster_lazy(1)
        waves = run()
    finally:
        gb.cluster_lazy(-1)
    return eager, waves


@pytest.mark.parametrize("name", ["finch_skani_abisko4_95", "finch_skani_abisko4_99", "cli_min_aligned_fraction_0.2",
                                  "cli_min_aligned_fraction_0.6"])
def test_stage2_in_waves_gives_the_pinned_outcome(gb, name):
    """The one-call pipeline with stage 2 asked for in waves (only the pairs the reference's two passes
    evaluate): the reference-pinned clusters, in the same order as with every hit evaluated up front."""
    case = CASES[name]
    run = lambda: gb.cluster(paths_of(GOLDEN, case), precluster_ani=case["pre_thr"], ani=case["ani"],
                             min_aligned_fraction=case["min_af"], small_genomes=case["small"])
    (eager, ei), (waves, wi) = _both_modes(gb, run)
    assert waves == eager and sorted(sorted(c) for c in waves) == case["clusters"]
    assert ei["ani_waves"] == 0 and (wi["ani_waves"] >= 1 or wi["n_precluster_hits"] == 0)
    assert wi["n_ani_pairs"] <= ei["n_ani_pairs"]


def test_stage2_in_waves_on_synthetic_families(gb):
    """Four synthetic families of 10 genomes (substitution rates 0 .. 10 %: several representatives per family at
    95 %, so the waves go several rounds deep) resident on the device: clusters, order and precluster statistics
    equal the eager pipeline's; fewer pairs are evaluated; a wave budget of one batch is covered by the CPU tests."""
    import torch
    seed, n, L = 3, 40, 300_000
    lay = gb.synth_layout(n, L)
    dev = torch.device("cuda", 0)
    d_seq = torch.zeros(lay["seq2_words"], dtype=torch.int32, device=dev)
    d_val = torch.zeros(lay["valid_words"], dtype=torch.int32, device=dev)
    d_off = torch.zeros(n + 1, dtype=torch.int64, device=dev)
    st = torch.cuda.current_stream().cuda_stream
    gb.synth_packed_device(seed, 0, n, L, d_seq.data_ptr(), d_val.data_ptr(), d_off.data_ptr(), st)
    torch.cuda.synchronize()
    base_off = d_off.cpu().numpy().astype(np.uint64)
    lengths = np.full(n, L, np.uint64)
    run = lambda: gb.cluster_packed(d_seq.data_ptr(), d_val.data_ptr(), base_off, lengths, precluster_ani=0.9, ani=95.0,
                                    min_aligned_fraction=15.0, device=True, d_base_off=d_off.data_ptr())
    (eager, ei), (waves, wi) = _both_modes(gb, run)
    assert waves == eager
    assert sorted(g for c in waves for g in c) == list(range(n)) and all(c[0] // 10 == g // 10 for c in waves for g in c)
    assert len(waves) > 4  # more than one representative per family
    for key in ("n_precluster_hits", "n_preclusters", "largest_precluster"):
        assert wi[key] == ei[key]
    assert wi["ani_waves"] >= 2 and 0 < wi["n_ani_pairs"] < ei["n_ani_pairs"]
    # the engine's own count of calculate_ani invocations is the reference's in both modes
    assert wi["ani_calls"] == ei["ani_calls"] == wi["n_ani_pairs"]
