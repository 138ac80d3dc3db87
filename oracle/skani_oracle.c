/*
 * oracle/skani_oracle.c -- CPU statement of the stage-2 ANI this repository computes in place of
 * the `skani dist` subprocess that /root/reference/src/skani.rs:718-788 spawns per genome pair.
 *
 * THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE (see oracle/__init__.py).
 *
 * PARITY UNPINNED.  The reference obtains ANI from an external executable (skani 0.2.2,
 * pixi.lock:124) whose source is not under /root/reference and which is not installed in this
 * image; no reference test asserts an ANI value.  What follows is therefore a
 * restatement of skani's PUBLISHED method (Shaw & Yu, Nat. Methods 2023: FracMinHash seeds of
 * k = 15 at density 1/c, c = 125 [30 with --small-genomes]; exact seed matches as anchors;
 * banded colinear chaining with anchor score 20 and gap-difference penalty; the query cut into
 * ~20 kb chunks, each chunk's identity (matched seeds / seeds)^(1/k), ANI = their mean; aligned
 * fraction from chained spans; for c >= 70, >= 150 kb aligned and whole genomes a learned
 * regression replaces the raw mean), with every free choice fixed here so that the CUDA kernels
 * can be checked bit-exactly against this file.  skani's regression weights live inside the
 * binary; the chain-span ratio below stands in for it.  What the reference's own tests DO pin for
 * this stage are cluster outcomes on its fixture genomes; the free choices of this file were
 * fixed so that ALL of them hold (tests/test_stage2_fixtures.py runs them on committed copies):
 *   src/clusterer.rs:631-690, 692-723, 725-757, 793-823; tests/test_cmdline.rs:262-302, 417-440,
 *   460-480, 482-507, 546-567, 569-588 (rep_bug --large-contigs), 590-609 (--small-contigs).
 *
 * Specification (all integer until the last line)
 *   seeds      every window of K = 15 valid bases at packed position p: fwd = MSB-first 2-bit
 *              integer, rev = reverse complement, canon = min; strand = (rev < fwd); selected iff
 *              mm_hash64(canon) < (2^64-1)/c.  Seeds are kept in position order.
 *              spread position = p + contig_index * (BAND+1)  (no chain can cross a contig break)
 *              chunk = chunk_base[contig] + (p - contig_start) / CHUNK,  CHUNK = 20000.
 *   query      the FIRST genome of the pair, as given: `skani dist -q fasta1 -r fasta2`
 *              (src/skani.rs:733-744; galah passes the representative, i.e. the lower index, first,
 *              src/clusterer.rs:262-296); for the triangle of the preclusterer the lower index.
 *   anchors    per query chunk, for its seeds x = 0,1,.. in order: all reference seeds with the
 *              same canonical k-mer, unless that k-mer occurs more than MAXOCC = 8 times in the
 *              reference (then the seed counts as unmatched); matches in ascending reference
 *              spread position.  rel = strand_q xor strand_r.
 *   chaining   f(a) = max(ALPHA, max_b f(b) + ALPHA - |dq - dr|) over the previous H = 16 anchors b
 *              (most recent first, scan stops at the first b with dq > BAND) that have the same
 *              rel, 0 < dq <= BAND, 0 < dr <= BAND (dr measured in chain direction) and
 *              |dq - dr| <= MAXGAP = 300; strict improvement only, so ties keep the more recent b.
 *              cnt(a) = cnt(best b) + 1 (1 without a predecessor).  An anchor with cnt >= 3 is
 *              CHAINED; a chain QUALIFIES at its third anchor.
 *   per chunk  counted iff it holds a chained anchor.  first = the smallest first-seed index over
 *              its qualified chains, last = the seed of its last chained anchor;
 *              N = last - first + 1 seeds, M = matched seeds (1..MAXOCC occurrences) among them;
 *              identity = (M/N)^(1/K), carried as round(2^40 * identity) so that sums are integer.
 *              coverage: a qualifying chain adds its span so far + K to covq / covr, every later
 *              chained anchor its link's dq / dr; span_m / span_n likewise count chained anchors
 *              (3 at qualification, then 1 per anchor) and the seeds inside the chain spans.
 *   per pair   raw ANI = 100 * mean over counted chunks of identity (skani's default estimator:
 *              mean, not --median / --robust).  If c >= 70, covq >= 150,000 and the units are
 *              whole genomes (not `-i` records) -- skani's condition for its learned ANI -- the
 *              chain-span ratio 100 * ((span_m - 2 chains) / (span_n - 2 chains))^(1/K) is used
 *              instead (end anchors of a chain match by construction).  AF = cov / total length
 *              (capped at 1).
 *   output     as galah sees skani's TSV: no row (ANI 0.0, src/skani.rs:760) when no chunk is
 *              counted or max(AFq, AFr) * 100 < min_af; otherwise ANI printed with two decimals and
 *              parsed back as f32 (src/skani.rs:773-779).
 *   screen     (preclusterer) marker sketches: k = 21, 1/1000 (1/200 with --small-genomes); a pair
 *              is compared iff |A n B| >= max(1, ceil(0.8^21 min(|A|,|B|))), or -- without
 *              --small-genomes, whose alias includes --faster-small -- min(|A|,|B|) < 20.
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define SK_K 15
#define SK_CHUNK 20000u
#define SK_BAND 2500
#define SK_MAXGAP 300
#define SK_ALPHA 20
#define SK_H 16
#define SK_MAXOCC 8
#define SK_MIN_ANCHORS 3
#define SK_FX_BITS 40
#define SK_LEARNED_MIN_C 70u
#define SK_LEARNED_MIN_BASES 150000u

static inline uint64_t mm_hash64(uint64_t key) {
    key = ~key + (key << 21);
    key = key ^ (key >> 24);
    key = (key + (key << 3)) + (key << 8);
    key = key ^ (key >> 14);
    key = (key + (key << 2)) + (key << 4);
    key = key ^ (key >> 28);
    key = key + (key << 31);
    return key;
}
uint64_t skani_oracle_mm_hash64(uint64_t key) { return mm_hash64(key); }

/* Seeds of one genome.  Outputs (caller-allocated, capacity cap): kmer_strand = canon << 1 | strand,
 * spread, chunk.  Returns the number of seeds (may exceed cap: nothing is written past cap).
 * *n_chunks and *total_len are always set. */
uint64_t skani_oracle_seeds(const uint8_t *codes, uint64_t n, const uint64_t *rec_start,
                            const uint64_t *rec_end, uint32_t nrec, uint32_t c, uint32_t *kmer_strand,
                            uint32_t *spread, uint32_t *chunk, uint64_t cap, uint32_t *n_chunks,
                            uint64_t *total_len) {
    (void)n;
    const uint64_t thr = UINT64_MAX / c;
    const uint64_t mask = (1ULL << (2 * SK_K)) - 1;
    uint64_t out = 0, tl = 0;
    uint32_t chunk_base = 0;
    for (uint32_t r = 0; r < nrec; r++) {
        const uint64_t s = rec_start[r], e = rec_end[r];
        tl += e - s;
        uint64_t fwd = 0, rev = 0;
        uint32_t run = 0;
        for (uint64_t p = s; p < e; p++) {
            const uint8_t cd = codes[p];
            if (cd > 3) { run = 0; fwd = rev = 0; continue; }
            fwd = ((fwd << 2) | cd) & mask;
            rev = (rev >> 2) | ((uint64_t)(3 - cd) << (2 * (SK_K - 1)));
            if (++run < SK_K) continue;
            const uint64_t start = p + 1 - SK_K;
            const uint64_t canon = fwd < rev ? fwd : rev;
            const uint32_t strand = rev < fwd ? 1u : 0u;
            if (mm_hash64(canon) < thr) {
                if (out < cap) {
                    kmer_strand[out] = (uint32_t)(canon << 1) | strand;
                    spread[out] = (uint32_t)(start + (uint64_t)r * (SK_BAND + 1));
                    chunk[out] = chunk_base + (uint32_t)((start - s) / SK_CHUNK);
                }
                out++;
            }
        }
        chunk_base += (uint32_t)((e - s + SK_CHUNK - 1) / SK_CHUNK);
    }
    *n_chunks = chunk_base;
    *total_len = tl;
    return out;
}

typedef struct { uint32_t kmer, spread, strand; } ref_entry;
static int cmp_ref(const void *a, const void *b) {
    const ref_entry *x = (const ref_entry *)a, *y = (const ref_entry *)b;
    if (x->kmer != y->kmer) return x->kmer < y->kmer ? -1 : 1;
    if (x->spread != y->spread) return x->spread < y->spread ? -1 : 1;
    return 0;
}

/* 2^40 * (m/n)^(1/15), rounded to nearest: the per-chunk identity as a fixed-point integer, so
 * that the per-pair sum is an integer sum (order-independent on the device).  glibc pow. */
uint64_t skani_oracle_chunk_identity_fx(uint32_t m, uint32_t n) {
    if (m == 0 || n == 0) return 0;
    if (m >= n) return 1ULL << SK_FX_BITS;
    return (uint64_t)llround(pow((double)m / (double)n, 1.0 / SK_K) * (double)(1ULL << SK_FX_BITS));
}

/* Integer core for one (query, reference) pair; the query is the FIRST genome, as given.
 * out[0..4] = sum of per-chunk fixed-point identities, number of chunks counted, covq, covr,
 *             total matched seeds M over counted chunks (diagnostic; not used by the finish).
 * chunk_mn (optional, capacity chunk_cap pairs): per counted chunk (M, N) in chunk order. */
uint64_t skani_oracle_chain(const uint32_t *q_ks, const uint32_t *q_spread, const uint32_t *q_chunk,
                            uint64_t nq, const uint32_t *r_ks, const uint32_t *r_spread, uint64_t nr,
                            uint64_t *out, uint32_t *chunk_mn, uint64_t chunk_cap) {
    ref_entry *ref = (ref_entry *)malloc((nr ? nr : 1) * sizeof(ref_entry));
    for (uint64_t x = 0; x < nr; x++) {
        ref[x].kmer = r_ks[x] >> 1; ref[x].strand = r_ks[x] & 1; ref[x].spread = r_spread[x];
    }
    qsort(ref, nr, sizeof(ref_entry), cmp_ref);
    uint64_t sum_fx = 0, n_counted = 0, covq = 0, covr = 0, sum_m = 0, span_m = 0, span_n = 0, n_chains = 0;
    struct { int32_t q, r, f; uint32_t rel, cnt, first_x, first_mb, x; int32_t first_r; } ring[SK_H];
    uint64_t c0 = 0;
    while (c0 < nq) {
        uint64_t c1 = c0;
        while (c1 < nq && q_chunk[c1] == q_chunk[c0]) c1++;
        uint32_t n_anchor = 0, m_run = 0;
        int qualified = 0;
        uint32_t first_q = 0, mb_first = 0, last_q = 0, m_at_last = 0;
        for (uint64_t x = c0; x < c1; x++) {
            const uint32_t km = q_ks[x] >> 1, qs = q_ks[x] & 1;
            uint64_t lo = 0, hi = nr;
            while (lo < hi) { uint64_t mid = (lo + hi) >> 1; if (ref[mid].kmer < km) lo = mid + 1; else hi = mid; }
            uint64_t e = lo;
            while (e < nr && ref[e].kmer == km) e++;
            const uint64_t occ = e - lo;
            if (occ == 0 || occ > SK_MAXOCC) continue;
            const uint32_t m_before = m_run++;
            const uint32_t xi = (uint32_t)(x - c0);
            for (uint64_t m = lo; m < e; m++) {
                const int32_t q = (int32_t)q_spread[x], r = (int32_t)ref[m].spread;
                const uint32_t rel = qs ^ ref[m].strand;
                int32_t f = SK_ALPHA; uint32_t cnt = 1, first_x = xi, first_mb = m_before; int32_t first_r = r;
                int32_t link_dq = 0, link_dr = 0; uint32_t link_dx = 0;
                const uint32_t look = n_anchor < SK_H ? n_anchor : SK_H;
                for (uint32_t b = 1; b <= look; b++) {
                    const uint32_t slot = (n_anchor - b) % SK_H;
                    const int32_t dq = q - ring[slot].q;
                    if (dq > SK_BAND) break;
                    if (dq <= 0 || ring[slot].rel != rel) continue;
                    const int32_t dr = rel ? ring[slot].r - r : r - ring[slot].r;
                    if (dr <= 0 || dr > SK_BAND) continue;
                    const int32_t gap = dq > dr ? dq - dr : dr - dq;
                    if (gap > SK_MAXGAP) continue;
                    const int32_t cand = ring[slot].f + SK_ALPHA - gap;
                    if (cand > f) {
                        f = cand; cnt = ring[slot].cnt + 1; first_x = ring[slot].first_x;
                        first_mb = ring[slot].first_mb; first_r = ring[slot].first_r;
                        link_dq = dq; link_dr = dr; link_dx = xi - ring[slot].x;
                    }
                }
                const uint32_t slot = n_anchor % SK_H;
                ring[slot].q = q; ring[slot].r = r; ring[slot].f = f; ring[slot].rel = rel;
                ring[slot].cnt = cnt; ring[slot].first_x = first_x; ring[slot].first_mb = first_mb;
                ring[slot].first_r = first_r; ring[slot].x = xi;
                n_anchor++;
                if (cnt >= SK_MIN_ANCHORS) {
                    if (!qualified || first_x < first_q) { first_q = first_x; mb_first = first_mb; }
                    qualified = 1;
                    last_q = xi; m_at_last = m_before + 1;
                    if (cnt == SK_MIN_ANCHORS) {
                        covq += (uint64_t)(q - (int32_t)q_spread[c0 + first_x]) + SK_K;
                        covr += (uint64_t)(r > first_r ? r - first_r : first_r - r) + SK_K;
                        span_m += SK_MIN_ANCHORS; span_n += xi - first_x + 1; n_chains++;
                    } else {
                        covq += (uint64_t)link_dq; covr += (uint64_t)link_dr;
                        span_m += 1; span_n += link_dx;
                    }
                }
            }
        }
        if (qualified) {
            const uint32_t M = m_at_last - mb_first, N = last_q - first_q + 1;
            sum_fx += skani_oracle_chunk_identity_fx(M, N);
            sum_m += M;
            if (chunk_mn && n_counted < chunk_cap) { chunk_mn[2 * n_counted] = M; chunk_mn[2 * n_counted + 1] = N; }
            n_counted++;
        }
        c0 = c1;
    }
    free(ref);
    out[0] = sum_fx; out[1] = n_counted; out[2] = covq; out[3] = covr; out[4] = sum_m;
    out[5] = span_m; out[6] = span_n; out[7] = n_chains;
    return n_counted;
}

/* Host finish shared in spirit with the product (galah_b200/csrc/ani.cu: ani_finish): integers ->
 * the f32 galah would parse from skani's TSV.  af_out[0..1] = AFq, AFr as fractions;
 * *estimator = 0 (mean of chunk identities) or 1 (chain-span ratio). */
float skani_oracle_finish(uint64_t sum_fx, uint64_t n_counted, uint64_t covq, uint64_t covr, uint64_t len_q,
                          uint64_t len_r, uint64_t span_m, uint64_t span_n, uint64_t n_chains, uint32_t c,
                          int individual_contigs, float min_af_pct, double *af_out, double *ani_unrounded,
                          int *estimator) {
    double afq = len_q ? (double)covq / (double)len_q : 0.0, afr = len_r ? (double)covr / (double)len_r : 0.0;
    if (afq > 1.0) afq = 1.0;
    if (afr > 1.0) afr = 1.0;
    if (af_out) { af_out[0] = afq; af_out[1] = afr; }
    if (ani_unrounded) *ani_unrounded = 0.0;
    if (estimator) *estimator = 0;
    if (n_counted == 0 || sum_fx == 0) return 0.0f;
    const int span = c >= SK_LEARNED_MIN_C && !individual_contigs && covq >= SK_LEARNED_MIN_BASES &&
                     span_n > 2 * n_chains && span_m > 2 * n_chains;
    if (estimator) *estimator = span;
    const double ani = span ? 100.0 * pow((double)(span_m - 2 * n_chains) / (double)(span_n - 2 * n_chains), 1.0 / SK_K)
                            : 100.0 * ((double)sum_fx / ((double)n_counted * (double)(1ULL << SK_FX_BITS)));
    if (ani_unrounded) *ani_unrounded = ani;
    const double best_af = afq > afr ? afq : afr;
    if (best_af * 100.0 < (double)min_af_pct) return 0.0f;
    char buf[64];
    snprintf(buf, sizeof(buf), "%.2f", ani);
    return strtof(buf, NULL);
}

/* ------------------------------------------------------------------------------------------ */
/* Marker sketches and the screen used in place of `skani triangle`'s own (src/skani.rs:109-225) */
/*   markers  every window of 21 valid bases: canon = min(fwd, revcomp) as MSB-first 2-bit        */
/*            integers; kept iff mm_hash64(canon) < (2^64-1)/c_marker (c_marker = 1000, or 200   */
/*            with --small-genomes); the sketch is the ascending list of DISTINCT hashes.        */
/*   screen   a pair is compared iff min(|A|,|B|) > 0 and                                        */
/*            |A n B| >= max(1, ceil(0.8^21 * min(|A|,|B|)))  (marker containment ~ 80 % ANI).   */
/* PARITY UNPINNED like the rest of this file: skani's own screen is not reproduced bit for bit. */
/* ------------------------------------------------------------------------------------------ */
static int cmp_u64_sk(const void *a, const void *b) {
    uint64_t x = *(const uint64_t *)a, y = *(const uint64_t *)b;
    return x < y ? -1 : (x > y);
}
uint64_t skani_oracle_markers(const uint8_t *codes, const uint64_t *rec_start, const uint64_t *rec_end,
                              uint32_t nrec, uint32_t c_marker, uint64_t *out, uint64_t cap) {
    const int K = 21;
    const uint64_t thr = UINT64_MAX / c_marker;
    const uint64_t mask = (1ULL << (2 * K)) - 1;
    uint64_t n = 0;
    for (uint32_t r = 0; r < nrec; r++) {
        uint64_t fwd = 0, rev = 0; uint32_t run = 0;
        for (uint64_t p = rec_start[r]; p < rec_end[r]; p++) {
            const uint8_t cd = codes[p];
            if (cd > 3) { run = 0; fwd = rev = 0; continue; }
            fwd = ((fwd << 2) | cd) & mask;
            rev = (rev >> 2) | ((uint64_t)(3 - cd) << (2 * (K - 1)));
            if (++run < (uint32_t)K) continue;
            const uint64_t h = mm_hash64(fwd < rev ? fwd : rev);
            if (h < thr) { if (n < cap) out[n] = h; n++; }
        }
    }
    if (n > cap) return n;
    qsort(out, n, sizeof(uint64_t), cmp_u64_sk);
    uint64_t m = 0;
    for (uint64_t i = 0; i < n; i++) if (m == 0 || out[m - 1] != out[i]) out[m++] = out[i];
    return m;
}
double skani_oracle_screen_fraction(void) { return pow(0.80, 21.0); }
