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
 * banded colinear chaining with anchor score 20 and gap-difference penalty; per-~20 kb-chunk
 * identity (matched seeds / seeds)^(1/k); aligned fraction from chained spans), WITHOUT skani's
 * learned regression correction (its weights live inside the binary), with every free choice
 * fixed here so that the CUDA kernels can be checked bit-exactly against this file.  What the
 * reference's own tests pin for this stage are threshold-crossing behaviours on its fixture
 * genomes (src/clusterer.rs:631-690, tests/test_cmdline.rs:262-302,417-440); tests/
 * test_ani_oracle_fixtures.py checks those against this restatement.
 *
 * Specification (all integer until the last line)
 *   seeds      every window of K = 15 valid bases at packed position p: fwd = MSB-first 2-bit
 *              integer, rev = reverse complement, canon = min; strand = (rev < fwd); selected iff
 *              mm_hash64(canon) < (2^64-1)/c.  Seeds are kept in position order.
 *              spread position = p + contig_index * (BAND+1)  (no chain can cross a contig break)
 *              chunk = chunk_base[contig] + (p - contig_start) / CHUNK,  CHUNK = 20000.
 *   query      the genome with the smaller total length (ties: the first argument).
 *   anchors    per query chunk, for its seeds x = 0,1,.. in order: all reference seeds with the
 *              same canonical k-mer, unless that k-mer occurs more than MAXOCC = 8 times in the
 *              reference; matches in ascending reference spread position.
 *              rel = strand_q xor strand_r.
 *   chaining   f(a) = max(ALPHA, max_b f(b) + ALPHA - |dq - dr|) over the previous H = 16 anchors b
 *              (most recent first, scan stops at the first b with dq > BAND) that have the same
 *              rel, 0 < dq <= BAND, 0 < dr <= BAND (dr measured in chain direction) and
 *              |dq - dr| <= MAXGAP = 300; strict improvement only, so ties keep the more recent b.
 *              Each anchor carries (count, first seed index, first reference position) of its
 *              best chain.  The chunk's chain is the first anchor attaining the maximum f.
 *   per chunk  accepted iff the chain has M >= 3 anchors; N = seeds from its first to its last
 *              anchor inclusive; covq = query span + K; covr = reference span + K.
 *   per pair   sumM = sum (M - 2), sumN = sum (N - 2) (end anchors are matches by construction);
 *              ANI = 100 * (sumM / sumN)^(1/K) in f64; AF = cov / total length (capped at 1).
 *   output     as galah sees skani's TSV: no row (ANI 0.0, src/skani.rs:760) when sumN == 0 or
 *              max(AFq, AFr) * 100 < min_af; otherwise ANI printed with two decimals and parsed
 *              back as f32 (src/skani.rs:773-779).
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

/* Integer core for one (query, reference) orientation ALREADY chosen by the caller.
 * out[0..3] = sumM, sumN, covq, covr. */
void skani_oracle_chain(const uint32_t *q_ks, const uint32_t *q_spread, const uint32_t *q_chunk,
                        uint64_t nq, const uint32_t *r_ks, const uint32_t *r_spread, uint64_t nr,
                        uint64_t *out) {
    ref_entry *ref = (ref_entry *)malloc((nr ? nr : 1) * sizeof(ref_entry));
    for (uint64_t x = 0; x < nr; x++) {
        ref[x].kmer = r_ks[x] >> 1; ref[x].strand = r_ks[x] & 1; ref[x].spread = r_spread[x];
    }
    qsort(ref, nr, sizeof(ref_entry), cmp_ref);
    uint64_t sumM = 0, sumN = 0, covq = 0, covr = 0;
    struct { int32_t q, r, f; uint32_t rel, cnt, first_x; int32_t first_r; } ring[SK_H];
    uint64_t c0 = 0;
    while (c0 < nq) {
        uint64_t c1 = c0;
        while (c1 < nq && q_chunk[c1] == q_chunk[c0]) c1++;
        uint32_t n_anchor = 0;
        int32_t best_f = 0; uint32_t best_cnt = 0, best_first_x = 0, best_last_x = 0;
        int32_t best_first_r = 0, best_last_r = 0;
        for (uint64_t x = c0; x < c1; x++) {
            const uint32_t km = q_ks[x] >> 1, qs = q_ks[x] & 1;
            /* equal range in ref */
            uint64_t lo = 0, hi = nr;
            while (lo < hi) { uint64_t mid = (lo + hi) >> 1; if (ref[mid].kmer < km) lo = mid + 1; else hi = mid; }
            uint64_t e = lo;
            while (e < nr && ref[e].kmer == km) e++;
            const uint64_t occ = e - lo;
            if (occ == 0 || occ > SK_MAXOCC) continue;
            for (uint64_t m = lo; m < e; m++) {
                const int32_t q = (int32_t)q_spread[x], r = (int32_t)ref[m].spread;
                const uint32_t rel = qs ^ ref[m].strand;
                int32_t f = SK_ALPHA; uint32_t cnt = 1, first_x = (uint32_t)(x - c0); int32_t first_r = r;
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
                    if (cand > f) { f = cand; cnt = ring[slot].cnt + 1; first_x = ring[slot].first_x; first_r = ring[slot].first_r; }
                }
                const uint32_t slot = n_anchor % SK_H;
                ring[slot].q = q; ring[slot].r = r; ring[slot].f = f; ring[slot].rel = rel;
                ring[slot].cnt = cnt; ring[slot].first_x = first_x; ring[slot].first_r = first_r;
                n_anchor++;
                if (f > best_f) {
                    best_f = f; best_cnt = cnt; best_first_x = first_x; best_last_x = (uint32_t)(x - c0);
                    best_first_r = first_r; best_last_r = r;
                }
            }
        }
        if (best_cnt >= SK_MIN_ANCHORS) {
            const uint32_t N = best_last_x - best_first_x + 1;
            sumM += best_cnt - 2; sumN += N - 2;
            covq += (uint64_t)(q_spread[c0 + best_last_x] - q_spread[c0 + best_first_x]) + SK_K;
            const int32_t span = best_last_r > best_first_r ? best_last_r - best_first_r : best_first_r - best_last_r;
            covr += (uint64_t)span + SK_K;
        }
        c0 = c1;
    }
    free(ref);
    out[0] = sumM; out[1] = sumN; out[2] = covq; out[3] = covr;
}

/* Host finish shared in spirit with the product (galah_b200/csrc/ani.cu: ani_finish): integers ->
 * the f32 galah would parse from skani's TSV.  af_out[0..1] = AFq, AFr as fractions. */
float skani_oracle_finish(uint64_t sumM, uint64_t sumN, uint64_t covq, uint64_t covr, uint64_t len_q,
                          uint64_t len_r, float min_af_pct, double *af_out, double *ani_unrounded) {
    double afq = len_q ? (double)covq / (double)len_q : 0.0, afr = len_r ? (double)covr / (double)len_r : 0.0;
    if (afq > 1.0) afq = 1.0;
    if (afr > 1.0) afr = 1.0;
    if (af_out) { af_out[0] = afq; af_out[1] = afr; }
    if (ani_unrounded) *ani_unrounded = 0.0;
    if (sumN == 0 || sumM == 0) return 0.0f;
    const double ani = 100.0 * pow((double)sumM / (double)sumN, 1.0 / SK_K);
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
