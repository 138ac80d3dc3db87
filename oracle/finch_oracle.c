/*
 * oracle/finch_oracle.c -- CPU restatement of Galah's stage-1 (finch MinHash prefilter) path.
 *
 * THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may load it.  The product (galah_b200/) never
 * links, imports or calls anything in oracle/.
 *
 * What it restates (all citations into /root/reference):
 *   - src/finch.rs:48-97   `distances()`: sketch parameters (Mash scheme, kmers_to_sketch = final_size
 *                          = num_kmers, no_strict, hash_seed 0, filters off), the serial i<j pair loop,
 *                          `1.0 - mash_distance >= min_ani as f64`, `Some(distance as f32)`.
 *   - The arithmetic itself lives in third-party crates that are NOT vendored in the reference
 *     (Cargo.toml:31,33: finch = "0.6.*", needletail = "0.5.*"; murmurhash3 transitively).  Their
 *     published algorithms are restated here:
 *       needletail 0.5 `normalize(false)` + `canonical_kmers`   -> norm_table / oracle_sketch_*()
 *       murmurhash3 `murmurhash3_x64_128(bytes, seed).0`        -> murmur3_x64_128_h1()
 *       finch 0.6 MashSketcher (bottom-s over DISTINCT hashes)  -> sketcher_*()
 *       finch 0.6 distance::raw_distance / distance()           -> oracle_raw_distance / oracle_mash_ani
 *
 * Parity pin: the reference's only numeric known-answer test on this path, src/finch.rs:107-129
 * ((set1/1mbp.fna, set1/500kb.fna) -> Some(0.9808188); empty at min_ani 0.99), is reproduced by
 * tests/test_oracle_golden.py, which also pins the derived vectors in tests/golden/.
 *
 * Build: see oracle/Makefile (gcc -O2 -shared -fPIC ... -lz -lm -fopenmp).
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

/* zlib: the image ships libz.so but not zlib.h; these are the three stable entry points we use.
 * gzread() transparently passes through files that are not gzip-compressed, which matches
 * needletail's magic-byte sniffing for the two formats the reference fixtures use (plain, gzip). */
typedef struct gzFile_s *gzFile;
extern gzFile gzopen(const char *path, const char *mode);
extern int gzread(gzFile file, void *buf, unsigned len);
extern int gzclose(gzFile file);

/* ------------------------------------------------------------------------------------------ */
/* MurmurHash3_x64_128, first 64-bit word (finch `hash_f`: murmurhash3_x64_128(kmer, seed).0)  */
/* ------------------------------------------------------------------------------------------ */
static inline uint64_t rotl64(uint64_t x, int r) { return (x << r) | (x >> (64 - r)); }
static inline uint64_t fmix64(uint64_t k) {
    k ^= k >> 33; k *= 0xff51afd7ed558ccdULL;
    k ^= k >> 33; k *= 0xc4ceb9fe1a85ec53ULL;
    k ^= k >> 33; return k;
}

uint64_t oracle_murmur3_x64_128_h1(const uint8_t *data, int len, uint64_t seed) {
    const uint64_t c1 = 0x87c37b91114253d5ULL, c2 = 0x4cf5ad432745937fULL;
    uint64_t h1 = seed, h2 = seed;
    int nblocks = len / 16;
    for (int i = 0; i < nblocks; i++) {
        uint64_t k1, k2;
        memcpy(&k1, data + 16 * i, 8);
        memcpy(&k2, data + 16 * i + 8, 8);
        k1 *= c1; k1 = rotl64(k1, 31); k1 *= c2; h1 ^= k1;
        h1 = rotl64(h1, 27); h1 += h2; h1 = h1 * 5 + 0x52dce729ULL;
        k2 *= c2; k2 = rotl64(k2, 33); k2 *= c1; h2 ^= k2;
        h2 = rotl64(h2, 31); h2 += h1; h2 = h2 * 5 + 0x38495ab5ULL;
    }
    const uint8_t *tail = data + nblocks * 16;
    uint64_t k1 = 0, k2 = 0;
    int rem = len & 15;
    for (int i = rem - 1; i >= 8; i--) k2 ^= (uint64_t)tail[i] << (8 * (i - 8));
    if (rem > 8) { k2 *= c2; k2 = rotl64(k2, 33); k2 *= c1; h2 ^= k2; }
    for (int i = (rem > 8 ? 7 : rem - 1); i >= 0; i--) k1 ^= (uint64_t)tail[i] << (8 * i);
    if (rem > 0) { k1 *= c1; k1 = rotl64(k1, 31); k1 *= c2; h1 ^= k1; }
    h1 ^= (uint64_t)len; h2 ^= (uint64_t)len;
    h1 += h2; h2 += h1;
    h1 = fmix64(h1); h2 = fmix64(h2);
    h1 += h2;
    return h1;
}

/* ------------------------------------------------------------------------------------------ */
/* needletail 0.5 normalize(seq, allow_iupac = false)                                          */
/*   ACGTN- kept; acg -> upper; t,u,U -> T; '.' '~' -> '-'; space/tab/CR/LF removed (0 here);  */
/*   every other byte -> N.                                                                    */
/* ------------------------------------------------------------------------------------------ */
static uint8_t norm_table[256];
static int norm_ready = 0;
static void norm_init(void) {
    if (norm_ready) return;
    for (int i = 0; i < 256; i++) norm_table[i] = 'N';
    norm_table['A'] = 'A'; norm_table['C'] = 'C'; norm_table['G'] = 'G'; norm_table['T'] = 'T';
    norm_table['N'] = 'N'; norm_table['-'] = '-';
    norm_table['a'] = 'A'; norm_table['c'] = 'C'; norm_table['g'] = 'G';
    norm_table['t'] = 'T'; norm_table['u'] = 'T'; norm_table['U'] = 'T';
    norm_table['.'] = '-'; norm_table['~'] = '-';
    norm_table[' '] = 0; norm_table['\t'] = 0; norm_table['\r'] = 0; norm_table['\n'] = 0;
    norm_ready = 1;
}
static inline int base_code(uint8_t c) { /* A0 C1 G2 T3, else -1 */
    switch (c) { case 'A': return 0; case 'C': return 1; case 'G': return 2; case 'T': return 3; }
    return -1;
}

/* ------------------------------------------------------------------------------------------ */
/* finch 0.6 MashSketcher: keep the `size` smallest DISTINCT hashes seen (BinaryHeap + count   */
/* map; a repeated hash only bumps its count).  Restated as threshold + periodic compaction,   */
/* which yields the same final set.                                                            */
/* ------------------------------------------------------------------------------------------ */
typedef struct {
    uint64_t *buf; size_t n, cap; uint32_t size; uint64_t thr; int thr_valid;
} sketcher_t;

static int cmp_u64(const void *a, const void *b) {
    uint64_t x = *(const uint64_t *)a, y = *(const uint64_t *)b;
    return x < y ? -1 : (x > y);
}
static void sketcher_compact(sketcher_t *s) {
    qsort(s->buf, s->n, sizeof(uint64_t), cmp_u64);
    size_t m = 0;
    for (size_t i = 0; i < s->n; i++)
        if (m == 0 || s->buf[m - 1] != s->buf[i]) s->buf[m++] = s->buf[i];
    if (m > s->size) m = s->size;
    s->n = m;
    if (m == s->size && m > 0) { s->thr = s->buf[m - 1]; s->thr_valid = 1; }
}
static void sketcher_init(sketcher_t *s, uint32_t size) {
    s->size = size; s->cap = (size_t)size * 8 + 64; s->n = 0; s->thr = 0; s->thr_valid = 0;
    s->buf = (uint64_t *)malloc(s->cap * sizeof(uint64_t));
}
static inline void sketcher_push(sketcher_t *s, uint64_t h) {
    if (s->thr_valid && h > s->thr) return; /* finch: add iff new_hash <= heap max, or heap not full */
    s->buf[s->n++] = h;
    if (s->n == s->cap) sketcher_compact(s);
}

/* Feed one RAW (un-normalised) record.  k-mers never span records (finch calls
 * sketcher.process() once per SequenceRecord).  Canonical k-mer = lexicographic min of the
 * forward and reverse-complement ASCII strings; with A<C<G<T this equals comparing the
 * MSB-first 2-bit integers.  Windows containing any non-ACGT byte are skipped. */
static void sketcher_feed(sketcher_t *s, const uint8_t *raw, size_t len, int k, uint64_t seed) {
    norm_init();
    const uint64_t mask = (k == 32) ? ~0ULL : ((1ULL << (2 * k)) - 1);
    uint64_t fwd = 0, rev = 0; int valid = 0;
    uint8_t kmer[32];
    static const char ACGT[4] = {'A', 'C', 'G', 'T'};
    for (size_t p = 0; p < len; p++) {
        uint8_t c = norm_table[raw[p]];
        if (c == 0) continue; /* whitespace is removed before k-mers are formed */
        int code = base_code(c);
        if (code < 0) { valid = 0; fwd = rev = 0; continue; }
        fwd = ((fwd << 2) | (uint64_t)code) & mask;
        rev = (rev >> 2) | ((uint64_t)(3 - code) << (2 * (k - 1)));
        if (++valid < k) continue;
        uint64_t canon = fwd < rev ? fwd : rev;
        for (int t = 0; t < k; t++) kmer[t] = (uint8_t)ACGT[(canon >> (2 * (k - 1 - t))) & 3];
        sketcher_push(s, oracle_murmur3_x64_128_h1(kmer, k, seed));
    }
}
static uint32_t sketcher_finish(sketcher_t *s, uint64_t *out) {
    sketcher_compact(s);
    memcpy(out, s->buf, s->n * sizeof(uint64_t));
    uint32_t n = (uint32_t)s->n;
    free(s->buf); s->buf = NULL;
    return n;
}

/* Sketch a set of raw records held in memory.  rec_off has nrec+1 entries into `seq`. */
int oracle_sketch_records(const uint8_t *seq, const uint64_t *rec_off, uint32_t nrec, int k,
                          uint32_t s, uint64_t seed, uint64_t *out, uint32_t *count) {
    if (k < 1 || k > 32) return 1;
    sketcher_t sk; sketcher_init(&sk, s);
    for (uint32_t r = 0; r < nrec; r++)
        sketcher_feed(&sk, seq + rec_off[r], (size_t)(rec_off[r + 1] - rec_off[r]), k, seed);
    *count = sketcher_finish(&sk, out);
    return 0;
}

/* ------------------------------------------------------------------------------------------ */
/* FASTA / FASTQ reader with needletail-like record semantics (plain or gzip).                 */
/* ------------------------------------------------------------------------------------------ */
typedef struct { uint8_t *d; size_t n, cap; } bytes_t;
static void bytes_push(bytes_t *b, const uint8_t *p, size_t n) {
    if (b->n + n > b->cap) {
        while (b->n + n > b->cap) b->cap = b->cap ? b->cap * 2 : (1 << 20);
        b->d = (uint8_t *)realloc(b->d, b->cap);
    }
    memcpy(b->d + b->n, p, n); b->n += n;
}
static int slurp(const char *path, bytes_t *out) {
    gzFile f = gzopen(path, "rb");
    if (!f) return 1;
    static const unsigned CH = 1 << 20;
    uint8_t *tmp = (uint8_t *)malloc(CH);
    int got;
    while ((got = gzread(f, tmp, CH)) > 0) bytes_push(out, tmp, (size_t)got);
    free(tmp); gzclose(f);
    return got < 0 ? 2 : 0;
}

/* Calls cb(seq_ptr, seq_len) per record; sequence bytes may still contain line breaks
 * (normalisation removes them, exactly as normalize() does for the reference). */
typedef void (*record_cb)(void *ctx, const uint8_t *seq, size_t len);
static int for_each_record(const uint8_t *d, size_t n, record_cb cb, void *ctx) {
    size_t p = 0;
    while (p < n && (d[p] == '\n' || d[p] == '\r')) p++;
    if (p >= n) return 0;
    if (d[p] == '>') {
        while (p < n) {
            /* header line */
            while (p < n && d[p] != '\n') p++;
            if (p < n) p++;
            size_t start = p;
            /* sequence runs to the next line that starts with '>' */
            while (p < n) {
                if (d[p] == '>' && (p == start || d[p - 1] == '\n')) break;
                p++;
            }
            cb(ctx, d + start, p - start);
        }
        return 0;
    } else if (d[p] == '@') {
        while (p < n) {
            while (p < n && d[p] != '\n') p++; /* @header */
            if (p < n) p++;
            size_t start = p;
            while (p < n && d[p] != '\n') p++; /* sequence (single line) */
            size_t end = p;
            if (p < n) p++;
            while (p < n && d[p] != '\n') p++; /* + */
            if (p < n) p++;
            while (p < n && d[p] != '\n') p++; /* quality */
            if (p < n) p++;
            cb(ctx, d + start, end - start);
            while (p < n && (d[p] == '\n' || d[p] == '\r')) p++;
        }
        return 0;
    }
    return 3; /* not FASTA/FASTQ */
}

typedef struct { sketcher_t sk; int k; uint64_t seed; uint64_t bases; uint32_t nrec; } feed_ctx;
static void feed_cb(void *c, const uint8_t *seq, size_t len) {
    feed_ctx *f = (feed_ctx *)c;
    sketcher_feed(&f->sk, seq, len, f->k, f->seed);
    f->nrec++;
}

/* finch::sketch_files for ONE path with SketchParams::Mash{kmers_to_sketch=s, final_size=s,
 * no_strict=true, kmer_length=k, hash_seed=seed} and filters off (src/finch.rs:55-69).
 * out must hold s values; *count receives how many (<= s: no_strict allows fewer). */
int oracle_sketch_fasta(const char *path, int k, uint32_t s, uint64_t seed, uint64_t *out,
                        uint32_t *count) {
    if (k < 1 || k > 32) return 1;
    bytes_t b = {0, 0, 0};
    int rc = slurp(path, &b);
    if (rc) { free(b.d); return 10 + rc; }
    feed_ctx f; sketcher_init(&f.sk, s); f.k = k; f.seed = seed; f.bases = 0; f.nrec = 0;
    rc = for_each_record(b.d, b.n, feed_cb, &f);
    *count = sketcher_finish(&f.sk, out);
    free(b.d);
    return rc ? 20 + rc : 0;
}

/* ------------------------------------------------------------------------------------------ */
/* finch 0.6 distance::raw_distance (Mash scheme => scale 0 => no tail advance)                */
/*   two-pointer merge while BOTH lists have elements; total = i + j - common.                 */
/* ------------------------------------------------------------------------------------------ */
void oracle_raw_distance(const uint64_t *a, uint32_t na, const uint64_t *b, uint32_t nb,
                         uint64_t *common_out, uint64_t *total_out) {
    uint32_t i = 0, j = 0; uint64_t common = 0;
    while (i < na && j < nb) {
        if (a[i] < b[j]) i++;
        else if (a[i] > b[j]) j++;
        else { common++; i++; j++; }
    }
    *common_out = common;
    *total_out = (uint64_t)i + j - common;
}

/* finch distance(): jaccard = common/total; mash_distance = -ln(2j/(1+j))/k clamped with
 * f64::min(1, f64::max(0, d)); galah: ani = 1.0 - mash_distance (src/finch.rs:78-86).
 * Rust's f64::max/min return the non-NaN operand, so 0/0 (two empty sketches) gives
 * mash_distance 0 and ANI 1.0 -- restated with fmax/fmin, which have the same NaN rule. */
double oracle_mash_ani(uint64_t common, uint64_t total, int k) {
    double jaccard = (double)common / (double)total;
    double md = -1.0 * log((2.0 * jaccard) / (1.0 + jaccard)) / (double)k;
    md = fmin(1.0, fmax(0.0, md));
    return 1.0 - md;
}

typedef struct { uint32_t i, j, common, total; float ani; } oracle_pair_t;

/* src/finch.rs:75-95: serial nested loop over i<j, rows restricted to [row_begin,row_end) so a
 * bounded sample can be timed.  Returns the number of passing pairs (written up to cap). */
size_t oracle_prefilter(const uint64_t *hashes, const uint32_t *counts, size_t n, size_t stride,
                        int k, float min_ani, size_t row_begin, size_t row_end, oracle_pair_t *out,
                        size_t cap) {
    size_t n_out = 0;
    if (row_end > n) row_end = n;
    for (size_t i = row_begin; i < row_end; i++) {
        for (size_t j = i + 1; j < n; j++) {
            uint64_t common, total;
            oracle_raw_distance(hashes + i * stride, counts[i], hashes + j * stride, counts[j],
                                &common, &total);
            double ani = oracle_mash_ani(common, total, k);
            if (ani >= (double)min_ani) {
                if (n_out < cap) {
                    out[n_out].i = (uint32_t)i; out[n_out].j = (uint32_t)j;
                    out[n_out].common = (uint32_t)common; out[n_out].total = (uint32_t)total;
                    out[n_out].ani = (float)ani;
                }
                n_out++;
            }
        }
    }
    return n_out;
}

/* Same loop, rows spread over all host threads (NOT how the reference runs -- its pair loop is a
 * plain nested `for` -- reported separately as the "all cores" fairness variant).  Counts only. */
size_t oracle_prefilter_count_mt(const uint64_t *hashes, const uint32_t *counts, size_t n,
                                 size_t stride, int k, float min_ani, size_t row_begin,
                                 size_t row_end) {
    size_t n_out = 0;
    if (row_end > n) row_end = n;
#pragma omp parallel for schedule(dynamic, 1) reduction(+ : n_out)
    for (size_t i = row_begin; i < row_end; i++) {
        for (size_t j = i + 1; j < n; j++) {
            uint64_t common, total;
            oracle_raw_distance(hashes + i * stride, counts[i], hashes + j * stride, counts[j],
                                &common, &total);
            if (oracle_mash_ani(common, total, k) >= (double)min_ani) n_out++;
        }
    }
    return n_out;
}

/* finch::sketch_files over many paths, one thread per file (finch uses rayon par_iter). */
int oracle_sketch_files_mt(const char *const *paths, size_t n, int k, uint32_t s, uint64_t seed,
                           uint64_t *hashes, uint32_t *counts) {
    int err = 0;
#pragma omp parallel for schedule(dynamic, 1)
    for (size_t g = 0; g < n; g++) {
        int rc = oracle_sketch_fasta(paths[g], k, s, seed, hashes + g * (size_t)s, &counts[g]);
        if (rc) {
#pragma omp critical
            err = rc;
        }
    }
    return err;
}

/* ------------------------------------------------------------------------------------------ */
/* Synthetic genomes: counter-based, regenerable from (seed, index).            */
/*   family f = index / 10, member m = index % 10, substitution rate RATE[m].                  */
/*   founder base block b (32 bases, 2 bits each, LSB first) = mix(key(seed, 2f, b)).          */
/*   per-base mutation draws for genome g come from 9 words mix(key(seed, 2g+1, 16b + w)):     */
/*     words 0..7 give a 16-bit uniform per base (mutate iff u16 < thr), word 8 gives 2 bits   */
/*     per base choosing which of the 3 other bases (new = (old + 1 + r % 3) & 3).             */
/* The CUDA generator in galah_b200/csrc/synth.cu implements the same definition.              */
/* ------------------------------------------------------------------------------------------ */
static inline uint64_t splitmix64(uint64_t x) {
    x += 0x9e3779b97f4a7c15ULL;
    x = (x ^ (x >> 30)) * 0xbf58476d1ce4e5b9ULL;
    x = (x ^ (x >> 27)) * 0x94d049bb133111ebULL;
    return x ^ (x >> 31);
}
static inline uint64_t synth_word(uint64_t seed, uint64_t stream, uint64_t ctr) {
    return splitmix64(splitmix64(seed ^ (stream * 0xd1342543de82ef95ULL)) + ctr * 0x2545f4914f6cdd1dULL);
}
static const uint32_t SYNTH_RATE_U16[10] = {
    /* round(rate * 65536) for {0, .5, 1, 2, 3, 4, 5, 6, 8, 10} % */
    0, 328, 655, 1311, 1966, 2621, 3277, 3932, 5243, 6554};

uint64_t oracle_synth_block_ex(uint64_t seed, uint64_t index, uint64_t block, uint32_t family_size,
                               uint32_t rate_shift);
uint64_t oracle_synth_block(uint64_t seed, uint64_t index, uint64_t block) {
    return oracle_synth_block_ex(seed, index, block, 10, 0);
}
uint64_t oracle_synth_block_ex(uint64_t seed, uint64_t index, uint64_t block, uint32_t family_size,
                               uint32_t rate_shift) {
    uint64_t fam = index / family_size, mem = (index % family_size) % 10;
    uint64_t w = synth_word(seed, 2 * fam, block);
    uint32_t thr = SYNTH_RATE_U16[mem] >> rate_shift;
    if (thr == 0) return w;
    uint64_t sel = synth_word(seed, 2 * index + 1, 16 * block + 8);
    for (int q = 0; q < 8; q++) {
        uint64_t u = synth_word(seed, 2 * index + 1, 16 * block + q);
        for (int t = 0; t < 4; t++) {
            uint32_t u16 = (uint32_t)(u >> (16 * t)) & 0xffff;
            if (u16 < thr) {
                int pos = q * 4 + t;
                uint64_t old = (w >> (2 * pos)) & 3;
                uint64_t r = (sel >> (2 * pos)) & 3;
                uint64_t nw = (old + 1 + (r % 3)) & 3;
                w = (w & ~(3ULL << (2 * pos))) | (nw << (2 * pos));
            }
        }
    }
    return w;
}

/* Writes L ASCII bases of genome `index`. */
void oracle_synth_genome(uint64_t seed, uint64_t index, uint64_t L, uint8_t *out) {
    static const char ACGT[4] = {'A', 'C', 'G', 'T'};
    for (uint64_t b = 0; b * 32 < L; b++) {
        uint64_t w = oracle_synth_block(seed, index, b);
        for (int t = 0; t < 32 && b * 32 + t < L; t++) out[b * 32 + t] = (uint8_t)ACGT[(w >> (2 * t)) & 3];
    }
}

void oracle_synth_genome_ex(uint64_t seed, uint64_t index, uint64_t L, uint32_t family_size, uint32_t rate_shift,
                            uint8_t *out) {
    static const char ACGT[4] = {'A', 'C', 'G', 'T'};
    for (uint64_t b = 0; b * 32 < L; b++) {
        uint64_t w = oracle_synth_block_ex(seed, index, b, family_size, rate_shift);
        for (int t = 0; t < 32 && b * 32 + t < L; t++) out[b * 32 + t] = (uint8_t)ACGT[(w >> (2 * t)) & 3];
    }
}

/* Sketch synthetic genome `index` directly (one record, pure ACGT). */
int oracle_sketch_synth(uint64_t seed, uint64_t index, uint64_t L, int k, uint32_t s,
                        uint64_t hash_seed, uint64_t *out, uint32_t *count) {
    uint8_t *seq = (uint8_t *)malloc(L);
    oracle_synth_genome(seed, index, L, seq);
    uint64_t off[2] = {0, L};
    int rc = oracle_sketch_records(seq, off, 1, k, s, hash_seed, out, count);
    free(seq);
    return rc;
}

int oracle_sketch_synth_mt(uint64_t seed, uint64_t index_begin, uint64_t n, uint64_t L, int k,
                           uint32_t s, uint64_t hash_seed, uint64_t *hashes, uint32_t *counts) {
#pragma omp parallel for schedule(dynamic, 1)
    for (uint64_t g = 0; g < n; g++)
        oracle_sketch_synth(seed, index_begin + g, L, k, s, hash_seed, hashes + g * (uint64_t)s, &counts[g]);
    return 0;
}

/* ------------------------------------------------------------------------------------------ */
/* Per-base codes of a FASTA/FASTQ file, in the packed coordinate system the product uses      */
/* (galah_b200/csrc/host/fasta.cpp): records concatenated, ONE invalid base between records,   */
/* whitespace dropped, 0..3 = ACGT (after normalize(false)), 4 = any other base.               */
/* Used by skani_oracle.c.  Caller frees *codes, *rec_start, *rec_end with oracle_free().       */
/* ------------------------------------------------------------------------------------------ */
typedef struct { bytes_t codes; uint64_t *rs, *re; uint32_t nrec, cap; } codes_ctx;
static void codes_cb(void *c, const uint8_t *seq, size_t len) {
    codes_ctx *x = (codes_ctx *)c;
    norm_init();
    if (x->nrec == x->cap) {
        x->cap = x->cap ? x->cap * 2 : 64;
        x->rs = (uint64_t *)realloc(x->rs, x->cap * sizeof(uint64_t));
        x->re = (uint64_t *)realloc(x->re, x->cap * sizeof(uint64_t));
    }
    if (x->nrec > 0) { uint8_t sep = 4; bytes_push(&x->codes, &sep, 1); }
    x->rs[x->nrec] = x->codes.n;
    uint8_t buf[4096]; size_t m = 0;
    for (size_t i = 0; i < len; i++) {
        uint8_t ch = norm_table[seq[i]];
        if (ch == 0) continue;
        int code = base_code(ch);
        buf[m++] = code < 0 ? 4 : (uint8_t)code;
        if (m == sizeof(buf)) { bytes_push(&x->codes, buf, m); m = 0; }
    }
    if (m) bytes_push(&x->codes, buf, m);
    x->re[x->nrec] = x->codes.n;
    x->nrec++;
}
int oracle_load_codes(const char *path, uint8_t **codes, uint64_t *n, uint64_t **rec_start,
                      uint64_t **rec_end, uint32_t *nrec) {
    bytes_t b = {0, 0, 0};
    int rc = slurp(path, &b);
    if (rc) { free(b.d); return 10 + rc; }
    codes_ctx x; memset(&x, 0, sizeof(x));
    rc = for_each_record(b.d, b.n, codes_cb, &x);
    free(b.d);
    if (rc) { free(x.codes.d); free(x.rs); free(x.re); return 20 + rc; }
    *codes = x.codes.d; *n = x.codes.n; *rec_start = x.rs; *rec_end = x.re; *nrec = x.nrec;
    return 0;
}
void oracle_free(void *p) { free(p); }
