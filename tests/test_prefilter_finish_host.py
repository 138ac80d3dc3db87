"""Host finish of stage 1 (no GPU): galah_b200_finish_candidates -- the reference's f64 formula, `>= min_ani as f64`
and the f32 store (src/finch.rs:78-93) over the device's integer candidates, then the (i, j) order of the
SortedPairGenomeDistanceCache -- against the oracle's own evaluation, on short lists (one thread) and on lists long
enough for the threaded formula pass (>= 32,768 candidates) and the threaded per-row sort (>= 131,072 survivors)."""
import numpy as np
import pytest

import galah_b200 as gb
import oracle


def reference_finish(cand, k, min_ani):
    ani = np.array([oracle.mash_ani(int(c), int(t), k) for c, t in cand[:, 2:4]], np.float64)
    keep = ani >= np.float64(np.float32(min_ani))
    c = cand[keep]
    order = np.lexsort((c[:, 1], c[:, 0]))
    return c[order], ani[keep][order].astype(np.float32)


@pytest.mark.parametrize("n_rows,per_row,min_ani", [(40, 30, 0.9), (300, 130, 0.9), (600, 400, 0.0), (500, 300, 0.99)])
def test_finish_matches_the_oracle_formula(n_rows, per_row, min_ani):
    rng = np.random.default_rng(n_rows)
    rows = []
    for i in range(n_rows):
        js = i + 1 + rng.choice(5 * per_row, size=per_row, replace=False)
        common = rng.integers(0, 1001, per_row)
        total = np.maximum(common, rng.integers(1000, 2001, per_row))
        rows.append(np.stack([np.full(per_row, i), js, common, total], axis=1))
    cand = np.concatenate(rows).astype(np.uint32)
    cand = cand[rng.permutation(len(cand))]  # the kernel appends in no particular order
    got = gb.finish_candidates(cand, 21, min_ani)
    exp_c, exp_ani = reference_finish(cand, 21, min_ani)
    assert len(got) == len(exp_c)
    for x, f in enumerate(("i", "j", "common", "total")):
        assert np.array_equal(got[f], exp_c[:, x]), f
    assert np.array_equal(got["ani"].view(np.uint32), exp_ani.view(np.uint32))


def test_zero_over_zero_and_empty():
    """common = total = 0 (two empty sketches): 0/0 is NaN in f64, NaN.max(0) is 0 in Rust, so the distance is 0 and the
    pair passes with ANI 1.0 -- the reference's quirk (src/finch.rs:78-92), kept."""
    got = gb.finish_candidates(np.array([[0, 1, 0, 0], [0, 2, 0, 1000]], np.uint32), 21, 0.9)
    assert len(got) == 1 and int(got["j"][0]) == 1 and float(got["ani"][0]) == 1.0
    assert len(gb.finish_candidates(np.zeros((0, 4), np.uint32), 21, 0.9)) == 0
