"""Host finish of stage 2 (no GPU): the exact integer route to strtof(sprintf("%.2f", v)) vs the
text route, and galah_b200_ani_finish vs the oracle's skani_oracle_finish (reference
src/skani.rs:760-779: two-decimal TSV column parsed as f32, 0.0 when skani prints no row)."""
import ctypes

import numpy as np

import galah_b200 as gb
import oracle
from galah_b200 import _native


def text_route(v):
    return np.float32(float("%.2f" % v))


def fast(v):
    return np.float32(_native.lib().galah_b200_print2_parse_f32(ctypes.c_double(v)))


def test_every_two_decimal_number_parses_like_strtof():
    """(float)(n / 100.0) == the float nearest to the decimal n/100 for every printable value."""
    n = np.arange(0, 200_001, dtype=np.int64)
    via_double = (n.astype(np.float64) / 100.0).astype(np.float32)
    via_text = np.array([np.float32(f"{x // 100}.{x % 100:02d}") for x in n], np.float32)
    assert np.array_equal(via_double.view(np.uint32), via_text.view(np.uint32))


def test_print2_parse_matches_text_route():
    rng = np.random.default_rng(3)
    vals = list(rng.uniform(0.0, 100.0, 200_000)) + list(rng.uniform(94.0, 100.0, 100_000))
    # exact ties (k/8 + 0.005 is never exact, but x.125, x.375, x.625, x.875 are: ties to even)
    vals += [x + f for x in range(0, 100) for f in (0.125, 0.375, 0.625, 0.875, 0.5, 0.25, 0.75)]
    # neighbours of every two-decimal boundary n/100 + 0.005 in [90, 100]
    for n in range(9000, 10000):
        b = (n + 0.5) / 100.0
        vals += [b, np.nextafter(b, 0.0), np.nextafter(b, 1e9)]
    vals += [100.0, 99.995, 99.99499999999999, 1e-9, 5e-3, 4.999999e-3, 1e-300, 0.0, 123456.785, 2.5e7]
    for v in vals:
        a, b = fast(float(v)), text_route(float(v))
        assert a.view(np.uint32) == b.view(np.uint32), (v, a, b)


def test_ani_finish_matches_oracle_finish():
    rng = np.random.default_rng(9)
    ints, out = _native.AniResult(), _native.AniResult()
    for _ in range(20_000):
        n_chunks = int(rng.integers(0, 150))
        fx = [oracle.chunk_identity_fx(int(m), int(n)) for n in rng.integers(1, 700, n_chunks)
              for m in [rng.integers(0, n + 1)]]
        n_chains = int(rng.integers(0, 300))
        span_n = int(rng.integers(0, 20_000)) + 3 * n_chains
        span_m = int(rng.integers(3 * n_chains, span_n + 1)) if span_n else 0
        len_q, len_r = int(rng.integers(1, 3_000_000)), int(rng.integers(1, 3_000_000))
        cov_q, cov_r = int(rng.integers(0, len_q * 1.1 + 1)), int(rng.integers(0, len_r * 1.1 + 1))
        min_af = float(rng.choice([0.0, 15.0, 50.0, 60.0]))
        small, contigs = bool(rng.integers(0, 2)), bool(rng.integers(0, 2))
        ints.sum_fx, ints.n_chunks, ints.sum_m = sum(fx), n_chunks, 0
        ints.span_m, ints.span_n, ints.n_chains, ints.cov_q, ints.cov_r = span_m, span_n, n_chains, cov_q, cov_r
        _native.check(_native.lib().galah_b200_ani_finish(ctypes.byref(ints), len_q, len_r, int(small), int(contigs),
                                                          ctypes.c_float(min_af), ctypes.byref(out)))
        exp = oracle.ani_finish((sum(fx), n_chunks, cov_q, cov_r, len_q, len_r, 0, span_m, span_n, n_chains),
                                min_af, 30 if small else 125, contigs)
        assert np.float32(out.ani).view(np.uint32) == np.float32(exp[0]).view(np.uint32), (span_m, span_n, n_chunks)
        assert out.estimator == exp[4]


def test_chunk_identity_fixed_point_matches_oracle():
    """round(2^40 (m/n)^(1/15)): the product's table entries == the oracle's per-chunk terms."""
    rng = np.random.default_rng(4)
    f = _native.lib().galah_b200_chunk_identity_fx
    for n in list(range(1, 40)) + [int(x) for x in rng.integers(40, 30_000, 400)]:
        for m in {0, 1, n // 3, n // 2, n - 1, n, int(rng.integers(0, n + 1))}:
            assert int(f(m, n)) == oracle.chunk_identity_fx(m, n), (m, n)
    assert int(f(5, 5)) == 1 << 40 and int(f(0, 7)) == 0
