"""K2 parity: the CUDA all-pairs prefilter vs the CPU oracle (bit-exact), through the C ABI."""
import numpy as np
import pytest

import oracle
from util import PAD, assert_pairs_equal, random_family_table

pytestmark = pytest.mark.gpu


@pytest.fixture(params=[0, 1], ids=["join", "pairwise"], autouse=True)
def kernel_mode(gb, request):
    """Every parity test runs through both exact kernel paths of the C ABI."""
    prev = gb.prefilter_mode(request.param)
    yield request.param
    gb.prefilter_mode(prev)


def test_golden_fixture_pair_list(gb, golden):
    """Sketches of the reference's own test genomes -> identical pair list (i, j, common, total,
    ani bits), including the reference-pinned (502, 1000, 0.9808188) row."""
    got = gb.prefilter(golden["table"], golden["counts"], 21, 0.9)
    assert_pairs_equal(got, golden["pairs_min_ani_0p9"])
    assert (got[0]["i"], got[0]["j"], got[0]["common"], got[0]["total"]) == (0, 1, 502, 1000)
    assert got[0]["ani"] == np.float32(0.9808188)


@pytest.mark.parametrize("n,s,seed", [(2, 1000, 0), (9, 1000, 1), (64, 1000, 2), (65, 1000, 6), (301, 1000, 3),
                                      (130, 10, 4), (77, 500, 5), (200, 2, 8), (129, 4608, 9)])
def test_random_families_match_oracle(gb, n, s, seed):
    rng = np.random.default_rng(seed)
    table, counts = random_family_table(n, s, rng)
    for min_ani in (0.9, 0.95):
        assert_pairs_equal(gb.prefilter(table, counts, 21, min_ani), oracle.prefilter(table, counts, 21, min_ani))


def test_ragged_and_empty_sketches(gb):
    rng = np.random.default_rng(7)
    table, counts = random_family_table(120, 1000, rng, ragged=True)
    counts[3] = 0; table[3] = PAD
    counts[17] = 0; table[17] = PAD
    counts[40] = 1; table[40, 1:] = PAD
    exp = oracle.prefilter(table, counts, 21, 0.9)
    got = gb.prefilter(table, counts, 21, 0.9)
    assert_pairs_equal(got, exp)
    # two empty sketches: 0/0 -> NaN -> clamped to distance 0 -> ANI 1.0 (flagged reference quirk)
    hit = [(int(p["i"]), int(p["j"])) for p in got if p["total"] == 0]
    assert (3, 17) in hit


@pytest.mark.parametrize("min_ani", [0.0, 0.5, 0.99, 1.0])
def test_threshold_extremes(gb, min_ani):
    rng = np.random.default_rng(11)
    table, counts = random_family_table(60, 1000, rng)
    table[5] = table[4]; counts[5] = counts[4]  # an identical pair (ANI exactly 1.0)
    exp = oracle.prefilter(table, counts, 21, min_ani)
    got = gb.prefilter(table, counts, 21, min_ani)
    assert_pairs_equal(got, exp)
    if min_ani == 0.0:
        assert len(got) == 60 * 59 // 2
    if min_ani == 1.0:
        assert (4, 5) in [(int(p["i"]), int(p["j"])) for p in got]


def test_genuine_max_hash_value(gb):
    """2^64-1 is also the padding value; a real hash of that value must still be counted."""
    rng = np.random.default_rng(13)
    table, counts = random_family_table(20, 1000, rng)
    for g in (0, 1, 2):
        table[g, counts[g] - 1] = PAD  # genuine element (still sorted, still distinct)
    assert_pairs_equal(gb.prefilter(table, counts, 21, 0.5), oracle.prefilter(table, counts, 21, 0.5))


def test_whole_clade_identical(gb):
    """Worst case for the join's match path: 150 identical sketches (every value matches in every
    block pair) next to 50 unrelated ones."""
    rng = np.random.default_rng(31)
    table, counts = random_family_table(200, 1000, rng)
    table[:150] = table[0]; counts[:150] = counts[0]
    exp = oracle.prefilter(table, counts, 21, 0.9)
    assert_pairs_equal(gb.prefilter(table, counts, 21, 0.9), exp)
    assert len(exp) >= 150 * 149 // 2


def test_duplicate_heavy_small_universe(gb):
    """Values drawn from a tiny universe: long equal runs inside and across block lists."""
    rng = np.random.default_rng(37)
    n, s = 260, 64
    table = np.full((n, s), PAD, np.uint64)
    counts = np.zeros(n, np.uint32)
    for g in range(n):
        m = int(rng.integers(0, s + 1))
        v = np.sort(rng.choice(96, size=m, replace=False)).astype(np.uint64)
        table[g, :m] = v; counts[g] = m
    for min_ani in (0.9, 0.97):
        assert_pairs_equal(gb.prefilter(table, counts, 21, min_ani), oracle.prefilter(table, counts, 21, min_ani))


def test_modes_agree_on_candidates(gb):
    """The two kernel paths emit the same integer candidate set {i, j, common, total}."""
    import torch
    rng = np.random.default_rng(41)
    n, s = 700, 1000
    table, counts = random_family_table(n, s, rng, ragged=True)
    d_t = torch.from_numpy(table.view(np.int64)).cuda()
    d_c = torch.from_numpy(counts.view(np.int32)).cuda()
    cap = 1 << 20
    out = []
    for mode in (0, 1):
        d_cand = torch.zeros((cap, 4), dtype=torch.int32, device="cuda")
        d_n = torch.zeros(1, dtype=torch.int64, device="cuda")
        torch.cuda.synchronize()
        gb.prefilter_enqueue(d_t.data_ptr(), d_c.data_ptr(), n, s, 21, 0.9, 0, 1, mode,
                             torch.cuda.current_stream().cuda_stream, d_cand.data_ptr(), cap, d_n.data_ptr())
        torch.cuda.synchronize()
        m = int(d_n.item())
        c = d_cand[:m].cpu().numpy().view(np.uint32)
        out.append(c[np.lexsort((c[:, 1], c[:, 0]))])
    assert out[0].shape == out[1].shape and np.array_equal(out[0], out[1])


def test_dense_clade_small_launch_takes_the_pairwise_fallback_and_is_exact(gb):
    """One clade (every sketch shares most hashes with every other): a launch of few join items hands
    the table to the pairwise kernel through a device flag (join_build_and_launch); whichever
    kernel runs, the pair list is the oracle's, bit for bit."""
    rng = np.random.default_rng(47)
    n, s = 420, 1000
    pool = np.unique(rng.integers(1, 1 << 62, size=1400, dtype=np.uint64))
    table = np.full((n, s), np.uint64(0xFFFFFFFFFFFFFFFF), np.uint64)
    counts = np.zeros(n, np.uint32)
    for g in range(n):
        keep = np.sort(rng.choice(pool, size=int(rng.integers(900, 1001)), replace=False))
        table[g, : len(keep)] = keep
        counts[g] = len(keep)
    got = gb.prefilter(table, counts, 21, 0.9)
    exp = oracle.prefilter(table, counts, 21, 0.9)
    assert len(exp) > n * (n - 1) // 4
    assert_pairs_equal(got, exp)
    gb.prefilter_mode(1)
    try:
        assert_pairs_equal(gb.prefilter(table, counts, 21, 0.9), exp)
    finally:
        gb.prefilter_mode(0)


def test_sliced_blocklist_build_then_join_matches(gb, kernel_mode):
    """The multi-GPU form: lists built in 3 equal slices (last one padded past the table), gathered,
    then joined per shard -- the union of the shards' candidates equals the one-call result."""
    if kernel_mode != 0:
        pytest.skip("block lists belong to the join path")
    import torch
    rng = np.random.default_rng(43)
    n, s, G = 500, 1000, 3
    table, counts = random_family_table(n, s, rng, ragged=True)
    d_t = torch.from_numpy(table.view(np.int64)).cuda()
    d_c = torch.from_numpy(counts.view(np.int32)).cuda()
    st = torch.cuda.current_stream().cuda_stream
    nb, epb, slack = gb.blocklist_layout(n, s)
    nbp = (nb + G - 1) // G
    hi = torch.zeros(G * nbp * epb + slack, dtype=torch.int32, device="cuda")
    lo = torch.zeros_like(hi)
    tags = torch.zeros(G * nbp * epb + slack, dtype=torch.uint8, device="cuda")
    ln = torch.zeros(G * nbp, dtype=torch.int32, device="cuda")
    for r in range(G):  # what each rank would build, written straight into its slot of the gathered arrays
        off = r * nbp * epb
        gb.blocklist_build(d_t.data_ptr(), d_c.data_ptr(), n, s, r * nbp, (r + 1) * nbp,
                           hi[off:].data_ptr(), lo[off:].data_ptr(), tags[off:].data_ptr(), ln[r * nbp:].data_ptr(), st)
    cap = 1 << 20
    got = []
    for shard in range(G):
        d_cand = torch.zeros((cap, 4), dtype=torch.int32, device="cuda")
        d_n = torch.zeros(1, dtype=torch.int64, device="cuda")
        gb.prefilter_join_enqueue(d_t.data_ptr(), d_c.data_ptr(), n, s, 21, 0.9, hi.data_ptr(), lo.data_ptr(),
                                  tags.data_ptr(), ln.data_ptr(), shard, G, st, d_cand.data_ptr(), cap, d_n.data_ptr())
        torch.cuda.synchronize()
        got.append(d_cand[: int(d_n.item())].cpu().numpy().view(np.uint32))
    got = np.concatenate(got)
    got = got[np.lexsort((got[:, 1], got[:, 0]))]
    d_cand = torch.zeros((cap, 4), dtype=torch.int32, device="cuda")
    d_n = torch.zeros(1, dtype=torch.int64, device="cuda")
    gb.prefilter_enqueue(d_t.data_ptr(), d_c.data_ptr(), n, s, 21, 0.9, 0, 1, 0, st, d_cand.data_ptr(), cap, d_n.data_ptr())
    torch.cuda.synchronize()
    exp = d_cand[: int(d_n.item())].cpu().numpy().view(np.uint32)
    exp = exp[np.lexsort((exp[:, 1], exp[:, 0]))]
    assert got.shape == exp.shape and np.array_equal(got, exp) and len(exp) > 0


@pytest.mark.parametrize("n,top_bit", [(400, False), (512, False), (640, True)])
def test_sharded_prefilter_driver_single_rank(gb, kernel_mode, n, top_bit):
    """galah_b200.distributed.ShardedPrefilter (the multi-GPU public API) on a 1-rank NCCL group:
    n = 400 takes the gather-first path, multiples of the row block build from the local slice with
    the all-reduced largest hash (a hash with the top bit set checks the unsigned MAX)."""
    if kernel_mode != 0:
        pytest.skip("driver uses the join path")
    import torch
    import torch.distributed as dist
    from galah_b200.distributed import ShardedPrefilter
    if not dist.is_initialized():
        dist.init_process_group("nccl", init_method="tcp://127.0.0.1:29533", rank=0, world_size=1,
                                device_id=torch.device("cuda", 0))
    rng = np.random.default_rng(47 + n)
    s = 1000
    table, counts = random_family_table(n, s, rng, hi_bits=64 if top_bit else 53)
    if top_bit:
        assert int(table[counts > 0].max()) >> 63 == 1 or int(table[table != PAD].max()) >> 63 == 1
    sp = ShardedPrefilter(gb, dist, n, s, torch.device("cuda", 0))
    assert sp.local_build == (n % gb.ROW_BLOCK == 0)
    h_t = torch.from_numpy(table.view(np.int64)).pin_memory()
    h_c = torch.from_numpy(counts.view(np.int32)).pin_memory()
    for _ in range(2):
        assert_pairs_equal(sp(h_t, h_c, 21, 0.9), oracle.prefilter(table, counts, 21, 0.9))


def test_large_sketches_use_generic_kernel(gb):
    rng = np.random.default_rng(17)
    table, counts = random_family_table(24, 2000, rng)
    assert_pairs_equal(gb.prefilter(table, counts, 21, 0.9), oracle.prefilter(table, counts, 21, 0.9))


def test_other_kmer_lengths(gb):
    rng = np.random.default_rng(19)
    table, counts = random_family_table(50, 1000, rng)
    for k in (15, 31):
        assert_pairs_equal(gb.prefilter(table, counts, k, 0.9), oracle.prefilter(table, counts, k, 0.9))


def test_degenerate_sizes(gb):
    t = np.full((1, 1000), PAD, np.uint64)
    assert len(gb.prefilter(t, np.zeros(1, np.uint32), 21, 0.9)) == 0
    assert len(gb.prefilter(np.zeros((0, 1000), np.uint64), np.zeros(0, np.uint32), 21, 0.9)) == 0


def test_row_shards_partition_the_pair_list(gb):
    import torch
    rng = np.random.default_rng(23)
    n, s = 333, 1000
    table, counts = random_family_table(n, s, rng)
    exp = oracle.prefilter(table, counts, 21, 0.9)
    d_t = torch.from_numpy(table.view(np.int64)).cuda()
    d_c = torch.from_numpy(counts.view(np.int32)).cuda()
    torch.cuda.synchronize()
    stream = torch.cuda.current_stream().cuda_stream
    for n_shards in (1, 2, 4, 8):
        parts = [gb.prefilter_device(d_t.data_ptr(), d_c.data_ptr(), n, s, 21, 0.9, r, n_shards, stream)
                 for r in range(n_shards)]
        allp = np.concatenate(parts)
        allp = allp[np.lexsort((allp["j"], allp["i"]))]
        assert_pairs_equal(allp, exp)
        # shard r owns the row blocks the boustrophedon rule gives it
        from galah_b200 import distributed as gd
        for r, part in enumerate(parts):
            assert np.all(gd.owner_of_row(part["i"], n_shards) == r)


def test_medium_table_sampled_rows(gb):
    """N=3000 (4.5 M pairs) on the GPU; the oracle checks a row sample and the all-core count."""
    rng = np.random.default_rng(29)
    n, s = 3000, 1000
    table, counts = random_family_table(n, s, rng)
    got = gb.prefilter(table, counts, 21, 0.9)
    assert len(got) == oracle.prefilter_count_mt(table, counts, 21, 0.9)
    for r0 in (0, 1492, 2990):
        exp = oracle.prefilter(table, counts, 21, 0.9, row_begin=r0, row_end=r0 + 8)
        sub = got[(got["i"] >= r0) & (got["i"] < r0 + 8)]
        assert_pairs_equal(sub, exp)
    # size-independent properties: keys strictly increasing, i < j, common <= total
    key = got["i"].astype(np.int64) * n + got["j"]
    assert np.all(np.diff(key) > 0) and np.all(got["i"] < got["j"]) and np.all(got["common"] <= got["total"])


@pytest.mark.parametrize("n,chunks", [(512, 8), (700, 3), (1300, 8), (3000, 16), (2049, 64)])
def test_streamed_upload_pipeline_matches_single_upload(gb, kernel_mode, n, chunks):
    """Host-buffer calls upload the table in slices and build + join each slice as it lands
    (one wave of items per slice).  The pair list must not depend on the slicing."""
    if kernel_mode != 0:
        pytest.skip("the upload pipeline belongs to the join path")
    rng = np.random.default_rng(1000 + n)
    table, counts = random_family_table(n, 1000, rng, ragged=True)
    prev = gb.prefilter_stream_chunks(1)
    try:
        base = gb.prefilter(table, counts, 21, 0.9)
        gb.prefilter_stream_chunks(chunks)
        for _ in range(2):  # second call re-uses the workspace of the first
            assert_pairs_equal(gb.prefilter(table, counts, 21, 0.9), base)
        t = gb.prefilter_last_host_timing()
        assert set(t) == {"enqueue", "wait", "d2h_extra", "finish"}
    finally:
        gb.prefilter_stream_chunks(prev)
    rows = min(n, 40)
    exp = oracle.prefilter(table, counts, 21, 0.9, row_begin=0, row_end=rows)
    assert_pairs_equal(base[base["i"] < rows], exp)


def test_streamed_upload_candidate_overflow_retries(gb, kernel_mode):
    """More survivors than the first candidate buffer holds: the retry runs on the resident table."""
    if kernel_mode != 0:
        pytest.skip("the upload pipeline belongs to the join path")
    n, s = 640, 1000
    row = np.arange(1, s + 1, dtype=np.uint64) * np.uint64(1 << 40)
    table = np.tile(row, (n, 1))
    counts = np.full(n, s, np.uint32)
    got = gb.prefilter(table, counts, 21, 0.9)  # all n(n-1)/2 = 204,480 pairs pass (> 65,536)
    assert len(got) == n * (n - 1) // 2
    assert np.all(got["common"] == s) and np.all(got["total"] == s) and np.all(got["ani"] == 1.0)
    key = got["i"].astype(np.int64) * n + got["j"]
    assert np.all(np.diff(key) > 0)


def test_full_size_config_properties(gb, kernel_mode):
    """BASELINE.json configs[1] at full size (10,000 synthetic 2 Mbp genomes sketched on the
    device, 49,995,000 pairs): size-independent properties.  The two exact kernel paths agree on
    the whole candidate set; the pair list is strictly (i, j)-ordered; shards partition it; every
    pair is within a family of 10 (cross-family genomes share ~0 hashes); a row sample is
    bit-exact against the oracle."""
    if kernel_mode != 0:
        pytest.skip("runs both paths itself")
    import torch
    n, L, s = 10_000, 2_000_000, 1000
    dev = torch.device("cuda", 0)
    st = torch.cuda.current_stream().cuda_stream
    d_h = torch.empty((n, s), dtype=torch.int64, device=dev)
    d_c = torch.empty(n, dtype=torch.int32, device=dev)
    batch = 500
    lay = gb.synth_layout(batch, L)
    d_seq = torch.zeros(lay["seq2_words"], dtype=torch.int32, device=dev)
    d_val = torch.zeros(lay["valid_words"], dtype=torch.int32, device=dev)
    d_off = torch.zeros(batch + 1, dtype=torch.int64, device=dev)
    for b0 in range(0, n, batch):
        gb.synth_packed_device(1, b0, batch, L, d_seq.data_ptr(), d_val.data_ptr(), d_off.data_ptr(), st)
        gb.sketch_packed_device(d_seq.data_ptr(), d_val.data_ptr(), d_off.data_ptr(), batch, 21, s, 0,
                                d_h[b0:].data_ptr(), d_c[b0:].data_ptr(), st)
    torch.cuda.synchronize()
    del d_seq, d_val
    cap = 1 << 20
    cands = []
    for mode in (0, 1):
        d_cand = torch.zeros((cap, 4), dtype=torch.int32, device=dev)
        d_n = torch.zeros(1, dtype=torch.int64, device=dev)
        gb.prefilter_enqueue(d_h.data_ptr(), d_c.data_ptr(), n, s, 21, 0.9, 0, 1, mode, st, d_cand.data_ptr(), cap,
                             d_n.data_ptr())
        torch.cuda.synchronize()
        c = d_cand[: int(d_n.item())].cpu().numpy().view(np.uint32)
        cands.append(c[np.lexsort((c[:, 1], c[:, 0]))])
    assert cands[0].shape == cands[1].shape and np.array_equal(cands[0], cands[1])
    full = gb.prefilter_device(d_h.data_ptr(), d_c.data_ptr(), n, s, 21, 0.9, 0, 1, st)
    assert 20_000 < len(full) <= len(cands[0])
    key = full["i"].astype(np.int64) * n + full["j"]
    assert np.all(np.diff(key) > 0) and np.all(full["i"] < full["j"])
    assert np.all(full["i"] // 10 == full["j"] // 10), "a pair across synthetic families passed"
    assert np.all(full["common"] <= full["total"]) and np.all(full["total"] <= 2 * s) and np.all(full["total"] >= s)
    parts = [gb.prefilter_device(d_h.data_ptr(), d_c.data_ptr(), n, s, 21, 0.9, r, 3, st) for r in range(3)]
    merged = np.concatenate(parts)
    merged = merged[np.lexsort((merged["j"], merged["i"]))]
    assert_pairs_equal(merged, full)
    table = d_h.cpu().numpy().view(np.uint64)
    counts = d_c.cpu().numpy().view(np.uint32)
    # host-buffer call (pipelined upload) == device-resident call
    assert_pairs_equal(gb.prefilter(table, counts, 21, 0.9), full)
    for r0 in (0, 4990, 9980):
        exp = oracle.prefilter(table, counts, 21, 0.9, row_begin=r0, row_end=r0 + 10)
        assert_pairs_equal(full[(full["i"] >= r0) & (full["i"] < r0 + 10)], exp)
