// Stage 1a (K1): per-genome bottom-s MinHash sketch on sm_100a.
//
// Replaces `finch::sketch_files(paths, SketchParams::Mash{kmers_to_sketch: s, final_size: s,
// no_strict: true, kmer_length: k, hash_seed: seed}, filters off)` called at
// /root/reference/src/finch.rs:55-69.  Third-party semantics restated (finch 0.6, needletail
// 0.5, murmurhash3; none vendored in the reference):
//   * every window of k bases that are all ACGT (after needletail normalize(false)); k-mers
//     never span records;
//   * canonical k-mer = lexicographic min of the forward and reverse-complement ASCII strings;
//   * hash = MurmurHash3_x64_128(canonical ASCII bytes, seed), first 64-bit word;
//   * sketch = the s smallest DISTINCT hashes of the file, ascending (fewer allowed).
//
// Input layout (HBM): 2 bits/base LSB-first in uint32 words (A0 C1 G2 T3) + a validity bitmap
// (1 bit/base).  Record breaks and non-ACGT bases are invalid bases.  Genome g covers bases
// [base_off[g], base_off[g+1]); base_off[g] % 128 == 0.  Buffers carry >= 16 bytes of padding.
//
// Kernels (one launch sequence per batch of genomes, no host round trip in between):
//   sketch_plan_kernel    per genome: number of CHUNK-position chunks, hash threshold T_g set so that
//                         about 1.3 s + 64 of the genome's k-mers hash below it, counters zeroed;
//                         exclusive scan -> chunk_off.  sketch_items_kernel: chunk -> genome table.
//   sketch_scan_kernel    one CTA per chunk (so the grid is as wide as the batch is long, not as
//                         wide as it has genomes).  Every thread extracts a k-mer straight from
//                         the packed stream (no rolling state, so no warm-up and no divergence):
//                             V = 2k bits at bit offset 2p           (LSB-first forward k-mer)
//                             F = reverse-2-bit-groups(V)            (MSB-first forward integer)
//                             R = ~V & mask                          (MSB-first reverse-complement)
//                             canonical LSB-first word W = F < R ? V : ~F & mask
//                         W is expanded to ASCII 8 bases at a time (bit spread + PRMT against the
//                         constant "ACGT"), hashed, and hashes <= T_g are appended to the genome's
//                         candidate buffer (one global atomic per candidate, ~1.4 k per genome).
//   sketch_select_kernel  one CTA per genome: candidates -> shared memory, bitonic sort, adjacent
//                         de-duplication, first s distinct written out.  A genome whose
//                         candidates overflowed or held fewer than s distinct values (repeats, N
//                         runs, tiny T) is put on a redo list instead.
//   sketch_kernel         the exact single-CTA-per-genome kernel (threshold bisection + shared
//                         hash set, re-scans until it holds >= s distinct values) runs over the
//                         redo list only, so the result is exact for any input.
#include "sketch.cuh"

#include <stdlib.h>

#include <algorithm>

#include "common.cuh"

namespace gb200 {

struct SketchKernelParams {
    const uint32_t *seq2;
    const uint32_t *valid;
    const uint64_t *base_off;
    uint32_t n;
    int k;
    uint32_t s;
    uint64_t seed;
    uint64_t *hashes;
    uint32_t *counts;
    uint32_t out_stride;
    uint32_t cap;  // hash-set slots (power of two)
    unsigned long long *work_counter;
    const uint32_t *redo_list;  // if non-null: genomes to process (count in *redo_n)
    const uint32_t *redo_n;
};

__device__ __forceinline__ uint64_t rotl64(uint64_t x, int r) { return (x << r) | (x >> (64 - r)); }
__device__ __forceinline__ uint64_t fmix64(uint64_t k) {
    k ^= k >> 33; k *= 0xff51afd7ed558ccdull;
    k ^= k >> 33; k *= 0xc4ceb9fe1a85ec53ull;
    k ^= k >> 33; return k;
}
// 8 bases (16 bits, LSB-first 2-bit codes) -> 8 ASCII bytes, base 0 in the lowest byte.
__device__ __forceinline__ uint64_t ascii8(uint32_t x) {
    x &= 0xFFFFu;
    x = (x | (x << 8)) & 0x00FF00FFu;
    x = (x | (x << 4)) & 0x0F0F0F0Fu;
    x = (x | (x << 2)) & 0x33333333u;
    const uint32_t lo = __byte_perm(0x54474341u, 0u, x);
    const uint32_t hi = __byte_perm(0x54474341u, 0u, x >> 16);
    return (uint64_t)lo | ((uint64_t)hi << 32);
}
__device__ __forceinline__ uint64_t byte_mask(int nbytes) {  // 1..8
    return nbytes >= 8 ? ~0ull : ((1ull << (8 * nbytes)) - 1);
}

// MurmurHash3_x64_128(h1 only) of the k ASCII bases encoded LSB-first in W.
template <int KT>
__device__ __forceinline__ uint64_t murmur_kmer(uint64_t W, int k_rt, uint64_t seed) {
    const int k = KT ? KT : k_rt;
    const uint64_t c1 = 0x87c37b91114253d5ull, c2 = 0x4cf5ad432745937full;
    uint64_t h1 = seed, h2 = seed;
    const int nblocks = k >> 4, rem = k & 15;
    uint32_t w16 = 0;  // index of the next 8-base group
    for (int b = 0; b < nblocks; b++) {
        uint64_t k1 = ascii8((uint32_t)(W >> (16 * w16)));
        uint64_t k2 = ascii8((uint32_t)(W >> (16 * (w16 + 1))));
        w16 += 2;
        k1 *= c1; k1 = rotl64(k1, 31); k1 *= c2; h1 ^= k1;
        h1 = rotl64(h1, 27); h1 += h2; h1 = h1 * 5 + 0x52dce729ull;
        k2 *= c2; k2 = rotl64(k2, 33); k2 *= c1; h2 ^= k2;
        h2 = rotl64(h2, 31); h2 += h1; h2 = h2 * 5 + 0x38495ab5ull;
    }
    if (rem > 8) {
        uint64_t k2 = ascii8((uint32_t)(W >> (16 * (w16 + 1)))) & byte_mask(rem - 8);
        k2 *= c2; k2 = rotl64(k2, 33); k2 *= c1; h2 ^= k2;
    }
    if (rem > 0) {
        uint64_t k1 = ascii8((uint32_t)(w16 < 4 ? (W >> (16 * w16)) : 0)) & byte_mask(rem > 8 ? 8 : rem);
        k1 *= c1; k1 = rotl64(k1, 31); k1 *= c2; h1 ^= k1;
    }
    h1 ^= (uint64_t)k; h2 ^= (uint64_t)k;
    h1 += h2; h2 += h1;
    h1 = fmix64(h1); h2 = fmix64(h2);
    h1 += h2;
    return h1;
}

// Canonical k-mer hash at base position `pos` (absolute, in the concatenated arrays).
// Returns false when the window holds an invalid base.
template <int KT>
__device__ __forceinline__ bool kmer_hash(const uint32_t *__restrict__ seq2,
                                          const uint32_t *__restrict__ valid, uint64_t pos, int k_rt,
                                          uint64_t seed, uint64_t &h) {
    const int k = KT ? KT : k_rt;
    const uint64_t vw = pos >> 5;
    const uint32_t vsh = (uint32_t)pos & 31u;
    const uint32_t m0 = __ldg(valid + vw), m1 = __ldg(valid + vw + 1);
    const uint32_t vm = __funnelshift_r(m0, m1, vsh);
    const uint32_t need = k >= 32 ? 0xFFFFFFFFu : ((1u << k) - 1u);
    if ((vm & need) != need) return false;
    const uint64_t w = pos >> 4;
    const uint32_t sh = ((uint32_t)pos & 15u) * 2u;
    const uint32_t a0 = __ldg(seq2 + w), a1 = __ldg(seq2 + w + 1), a2 = __ldg(seq2 + w + 2);
    const uint32_t v0 = __funnelshift_r(a0, a1, sh), v1 = __funnelshift_r(a1, a2, sh);
    const uint64_t mask = k >= 32 ? ~0ull : ((1ull << (2 * k)) - 1ull);
    const uint64_t V = ((uint64_t)v0 | ((uint64_t)v1 << 32)) & mask;
    uint64_t F = __brevll(V);
    F = ((F >> 1) & 0x5555555555555555ull) | ((F & 0x5555555555555555ull) << 1);
    F >>= (64 - 2 * k);
    const uint64_t R = ~V & mask;
    const uint64_t W = F < R ? V : (~F & mask);
    h = murmur_kmer<KT>(W, k_rt, seed);
    return true;
}

// skani-style marker hash: mm_hash64 (minimap2's invertible mix) of the canonical k-mer taken as an
// MSB-first 2-bit integer (oracle/skani_oracle.c).  Same window extraction as kmer_hash.
// The shift-and-add steps are written as multiplications by the equal constants (2^21 - 1, 265, 21,
// 2^31 + 1): the scan kernels are bound by the ALU pipe (shifts, logic), and integer
// multiply-adds issue on the FMA pipe, which has room.
__device__ __forceinline__ uint64_t mm_hash64_dev(uint64_t key) {
    key = key * 0x1FFFFFull - 1ull;        // ~key + (key << 21)
    key = key ^ (key >> 24);
    key = key * 265ull;                     // (key + (key << 3)) + (key << 8)
    key = key ^ (key >> 14);
    key = key * 21ull;                      // (key + (key << 2)) + (key << 4)
    key = key ^ (key >> 28);
    key = key * 0x80000001ull;              // key + (key << 31)
    return key;
}
template <int KT>
__device__ __forceinline__ bool kmer_mmhash(const uint32_t *__restrict__ seq2, const uint32_t *__restrict__ valid,
                                            uint64_t pos, int k_rt, uint64_t &h) {
    const int k = KT ? KT : k_rt;
    const uint64_t vw = pos >> 5;
    const uint32_t vsh = (uint32_t)pos & 31u;
    const uint32_t vm = __funnelshift_r(__ldg(valid + vw), __ldg(valid + vw + 1), vsh);
    const uint32_t need = k >= 32 ? 0xFFFFFFFFu : ((1u << k) - 1u);
    if ((vm & need) != need) return false;
    const uint64_t w = pos >> 4;
    const uint32_t sh = ((uint32_t)pos & 15u) * 2u;
    const uint32_t a0 = __ldg(seq2 + w), a1 = __ldg(seq2 + w + 1), a2 = __ldg(seq2 + w + 2);
    const uint32_t v0 = __funnelshift_r(a0, a1, sh), v1 = __funnelshift_r(a1, a2, sh);
    const uint64_t mask = k >= 32 ? ~0ull : ((1ull << (2 * k)) - 1ull);
    const uint64_t V = ((uint64_t)v0 | ((uint64_t)v1 << 32)) & mask;
    uint64_t F = __brevll(V);
    F = ((F >> 1) & 0x5555555555555555ull) | ((F & 0x5555555555555555ull) << 1);
    F >>= (64 - 2 * k);
    const uint64_t R = ~V & mask;
    h = mm_hash64_dev(F < R ? F : R);
    return true;
}

constexpr int kSketchThreads = 256;

template <int KT>
__global__ void __launch_bounds__(kSketchThreads) sketch_kernel(const SketchKernelParams p) {
    extern __shared__ __align__(16) uint64_t set[];  // p.cap slots
    __shared__ unsigned long long s_genome;
    __shared__ uint32_t s_distinct, s_overflow, s_has_max;
    const uint32_t tid = threadIdx.x;
    const uint32_t cap = p.cap, cap_mask = cap - 1, limit = cap / 2;
    const int k = KT ? KT : p.k;

    for (;;) {
        if (tid == 0) s_genome = atomicAdd(p.work_counter, 1ull);
        __syncthreads();
        unsigned long long g = s_genome;
        if (p.redo_list) {
            if (g >= *p.redo_n) break;
            g = p.redo_list[g];
        } else if (g >= p.n) break;
        const uint64_t b0 = p.base_off[g], b1 = p.base_off[g + 1];
        const uint64_t len = b1 - b0;
        const uint64_t npos = len >= (uint64_t)k ? len - k + 1 : 0;

        // initial threshold: about (1.3 s + 64) candidates expected among npos hashes
        uint64_t T = kPad;
        {
            const double want = 1.3 * (double)p.s + 64.0;
            if ((double)npos > want) T = (uint64_t)(18446744073709551616.0 * (want / (double)npos));
        }
        uint64_t T_lo = 0, T_hi = 0;  // T_lo: known too small; T_hi: known to overflow
        bool have_lo = false, have_hi = false;
        uint32_t distinct = 0;
        bool has_max = false;
        for (int round = 0; round < 80; round++) {
            for (uint32_t x = tid; x < cap; x += kSketchThreads) set[x] = kPad;
            if (tid == 0) { s_distinct = 0; s_overflow = 0; s_has_max = 0; }
            __syncthreads();
            for (uint64_t q = tid; q < npos; q += kSketchThreads) {
                uint64_t h;
                if (!kmer_hash<KT>(p.seq2, p.valid, b0 + q, p.k, p.seed, h)) continue;
                if (h > T) continue;
                if (h == kPad) { s_has_max = 1; continue; }
                if (*(volatile uint32_t *)&s_overflow) continue;
                uint32_t slot = (uint32_t)((h * 0x9E3779B97F4A7C15ull) >> 40) & cap_mask;
                for (;;) {
                    const unsigned long long prev =
                        atomicCAS(reinterpret_cast<unsigned long long *>(&set[slot]), kPad, h);
                    if (prev == h) break;  // duplicate
                    if (prev == kPad) {    // inserted
                        if (atomicAdd(&s_distinct, 1u) + 1 > limit) s_overflow = 1;
                        break;
                    }
                    slot = (slot + 1) & cap_mask;
                }
            }
            __syncthreads();
            distinct = s_distinct;
            const bool overflow = s_overflow != 0;
            has_max = s_has_max != 0;
            __syncthreads();
            if (overflow) {
                T_hi = T; have_hi = true;
                T = have_lo ? T_lo + (T_hi - T_lo) / 2 : T / 4;
                continue;
            }
            if (distinct + (has_max ? 1u : 0u) < p.s && T != kPad) {
                T_lo = T; have_lo = true;
                if (have_hi) T = T_lo + (T_hi - T_lo) / 2;
                else T = T > (kPad >> 3) ? kPad : T << 3;
                continue;
            }
            break;
        }
        // bitonic sort of the whole table; empties (2^64-1) sink to the end
        for (uint32_t size = 2; size <= cap; size <<= 1) {
            for (uint32_t str = size >> 1; str > 0; str >>= 1) {
                for (uint32_t x = tid; x < (cap >> 1); x += kSketchThreads) {
                    const uint32_t lo = 2 * x - (x & (str - 1));
                    const uint32_t hi = lo + str;
                    const bool up = (lo & size) == 0;
                    const uint64_t a = set[lo], b = set[hi];
                    if ((a > b) == up) { set[lo] = b; set[hi] = a; }
                }
                __syncthreads();
            }
        }
        uint32_t total = distinct;
        if (has_max && total < cap) {  // a genuine 2^64-1 hash is the largest possible value
            total += 1;                  // (slot `distinct` already holds 2^64-1)
        }
        const uint32_t out_n = min(total, p.s);
        uint64_t *out = p.hashes + (size_t)g * p.out_stride;
        for (uint32_t x = tid; x < p.out_stride; x += kSketchThreads) out[x] = x < out_n ? set[x] : kPad;
        if (tid == 0) p.counts[g] = out_n;
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------
// chunked pipeline: plan -> scan -> select (-> exact kernel over the redo list)
// ------------------------------------------------------------------------------------------
constexpr uint32_t kChunk = 16384;  // k-mer start positions per scan CTA

struct ChunkParams {
    const uint32_t *seq2;
    const uint32_t *valid;
    const uint64_t *base_off;
    uint32_t n;
    int k;
    uint32_t s;
    uint64_t seed;
    uint32_t cap;            // candidate slots per genome (power of two)
    uint64_t *chunk_off;     // [n + 1]
    uint64_t *thr;           // [n]
    uint32_t *cand_n;        // [n]
    uint32_t *has_max;       // [n]
    uint32_t *item_genome;   // [max_items]
    uint64_t max_items;
    uint64_t *cand;          // [n * cap]
    uint32_t *redo_list;     // [n]
    uint32_t *redo_n;        // [1]
    uint64_t *hashes;
    uint32_t *counts;
    uint32_t out_stride;
    uint64_t fixed_thr;      // != 0: FracMinHash mode -- keep every hash <= fixed_thr (no bottom-s cap)
    // FracMinHash mode only: the hash range [0, fixed_thr] is cut into n_parts equal value ranges,
    // each with its own candidate buffer of `cap` slots (cand[(g * n_parts + part) * cap ..],
    // cand_n[g * n_parts + part]), so a row may hold n_parts * cap markers although one sort in
    // shared memory holds cap.  part(h) = mulhi(h * frac_c, n_parts), exact and monotone in h.
    uint32_t n_parts;        // 0 / 1: a single buffer per genome
    uint32_t part;           // marker_select_kernel: the partition this launch finishes
    uint64_t frac_c;         // the FracMinHash compression factor c (h * c < 2^64 for every kept h)
    // fused K3 seed marking (scan21v2_kernel): one bit per base position, relative to sel_first_base
    int plan_k;              // k-mer length the chunk plan covers (0: k); 15 when seeds are marked too
    uint32_t *sel;
    uint64_t sel_first_base;
    uint64_t seed_thr;       // a 15-mer is a seed iff mm_hash64(canonical) < seed_thr
    uint32_t *seed_count;    // [n] or null: seeds per genome, added up by the scan
};

__global__ void __launch_bounds__(1024) sketch_plan_kernel(const ChunkParams p) {
    __shared__ uint64_t s_part[1024];
    const uint32_t tid = threadIdx.x;
    const uint32_t per = (p.n + 1023) / 1024;
    const uint32_t g0 = min(tid * per, p.n), g1 = min(g0 + per, p.n);
    uint64_t sum = 0;
    for (uint32_t g = g0; g < g1; g++) {
        const uint64_t len = p.base_off[g + 1] - p.base_off[g];
        const uint64_t npos = len >= (uint64_t)p.k ? len - p.k + 1 : 0;
        const uint64_t pk = p.plan_k ? (uint64_t)p.plan_k : (uint64_t)p.k;
        const uint64_t nplan = len >= pk ? len - pk + 1 : 0;
        sum += (nplan + kChunk - 1) / kChunk;
        uint64_t T = kPad;
        const double want = 1.3 * (double)p.s + 64.0;
        if ((double)npos > want) T = (uint64_t)(18446744073709551616.0 * (want / (double)npos));
        if (p.fixed_thr) T = p.fixed_thr;
        p.thr[g] = T;
        p.cand_n[g] = 0;
        p.has_max[g] = 0;
    }
    s_part[tid] = sum;
    __syncthreads();
    for (uint32_t off = 1; off < 1024; off <<= 1) {  // inclusive Hillis-Steele scan
        const uint64_t v = tid >= off ? s_part[tid - off] : 0;
        __syncthreads();
        s_part[tid] += v;
        __syncthreads();
    }
    uint64_t run = s_part[tid] - sum;  // exclusive
    for (uint32_t g = g0; g < g1; g++) {
        p.chunk_off[g] = run;
        const uint64_t len = p.base_off[g + 1] - p.base_off[g];
        const uint64_t pk = p.plan_k ? (uint64_t)p.plan_k : (uint64_t)p.k;
        const uint64_t nplan = len >= pk ? len - pk + 1 : 0;
        run += (nplan + kChunk - 1) / kChunk;
    }
    if (tid == 1023) p.chunk_off[p.n] = s_part[1023];
    if (tid == 0) *p.redo_n = 0;
}

__global__ void __launch_bounds__(256) sketch_items_kernel(const ChunkParams p) {
    // one warp per genome writes that genome's index into its item slots
    const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= p.n) return;
    const uint64_t c0 = p.chunk_off[warp], c1 = min(p.chunk_off[warp + 1], p.max_items);
    for (uint64_t c = c0 + lane; c < c1; c += 32) p.item_genome[c] = warp;
}

template <int KT, int HK>  // HK: 0 = MurmurHash3 of the ASCII k-mer (finch), 1 = mm_hash64 of the 2-bit k-mer
__global__ void __launch_bounds__(256) sketch_scan_kernel(const ChunkParams p) {
    const uint64_t item = blockIdx.x;
    if (item >= p.chunk_off[p.n]) return;
    const uint32_t g = p.item_genome[item];
    const uint64_t b0 = p.base_off[g], len = p.base_off[g + 1] - b0;
    const int k = KT ? KT : p.k;
    const uint64_t npos = len - k + 1;  // >= 1: genomes without k-mers have no items
    const uint64_t q0 = (item - p.chunk_off[g]) * kChunk;
    const uint64_t q1 = min(q0 + (uint64_t)kChunk, npos);
    const uint64_t T = p.thr[g];
    uint64_t *cand = p.cand + (size_t)g * p.cap;
    for (uint64_t q = q0 + threadIdx.x; q < q1; q += 256) {
        uint64_t h;
        if (HK == 0) { if (!kmer_hash<KT>(p.seq2, p.valid, b0 + q, p.k, p.seed, h)) continue; }
        else { if (!kmer_mmhash<KT>(p.seq2, p.valid, b0 + q, p.k, h)) continue; }
        if (h > T) continue;
        if (h == kPad) { p.has_max[g] = 1; continue; }
        if (HK == 1 && p.n_parts > 1) {
            const uint32_t part = (uint32_t)__umul64hi(h * p.frac_c, (uint64_t)p.n_parts);
            const size_t buf = (size_t)g * p.n_parts + part;
            const uint32_t slot = atomicAdd(&p.cand_n[buf], 1u);
            if (slot < p.cap) p.cand[buf * p.cap + slot] = h;
            continue;
        }
        const uint32_t slot = atomicAdd(&p.cand_n[g], 1u);
        if (slot < p.cap) cand[slot] = h;
    }
}

// FracMinHash rows wider than one shared-memory sort: launch `part` = 0 .. n_parts - 1 in order;
// every launch sorts one value range of every genome, drops duplicates and appends the distinct
// values behind what the earlier ranges wrote (counts[g] is the running length; 0xFFFFFFFF flags
// a buffer or row overflow).  The last launch pads the row.  One CTA per genome, p.cap uint64 of
// dynamic shared memory.
__global__ void __launch_bounds__(256) marker_select_kernel(const ChunkParams p) {
    extern __shared__ __align__(16) uint64_t sel[];
    __shared__ uint32_t s_warp_sum[8];
    const uint32_t g = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const size_t buf = (size_t)g * p.n_parts + p.part;
    const uint32_t got = p.cand_n[buf];
    const uint32_t base = p.part == 0 ? 0u : p.counts[g];
    const bool bad = got > p.cap || base == 0xFFFFFFFFu;
    const uint32_t m = min(got, p.cap);
    uint32_t cap2 = 256;
    while (cap2 < m) cap2 <<= 1;
    const uint64_t *cand = p.cand + buf * p.cap;
    for (uint32_t x = tid; x < cap2; x += 256) sel[x] = x < m ? cand[x] : kPad;
    __syncthreads();
    for (uint32_t size = 2; size <= cap2; size <<= 1) {
        for (uint32_t str = size >> 1; str > 0; str >>= 1) {
            for (uint32_t x = tid; x < (cap2 >> 1); x += 256) {
                const uint32_t lo = 2 * x - (x & (str - 1));
                const uint32_t hi = lo + str;
                const bool up = (lo & size) == 0;
                const uint64_t a = sel[lo], b = sel[hi];
                if ((a > b) == up) { sel[lo] = b; sel[hi] = a; }
            }
            __syncthreads();
        }
    }
    const uint32_t per = cap2 / 256;
    const uint32_t x0 = tid * per;
    uint32_t mine = 0;
    for (uint32_t x = x0; x < x0 + per; x++) {
        const uint64_t v = sel[x];
        mine += (v != kPad && (x == 0 || sel[x - 1] != v)) ? 1u : 0u;
    }
    uint32_t incl = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
        if ((int)lane >= o) incl += t;
    }
    if (lane == 31) s_warp_sum[warp] = incl;
    __syncthreads();
    uint32_t before = 0, distinct = 0;
    for (uint32_t w = 0; w < 8; w++) {
        if (w < warp) before += s_warp_sum[w];
        distinct += s_warp_sum[w];
    }
    if (bad || base + distinct > p.out_stride) {
        if (tid == 0) p.counts[g] = 0xFFFFFFFFu;
        return;
    }
    uint64_t *out = p.hashes + (size_t)g * p.out_stride + base;
    uint32_t pos = before + incl - mine;
    for (uint32_t x = x0; x < x0 + per; x++) {
        const uint64_t v = sel[x];
        if (v != kPad && (x == 0 || sel[x - 1] != v)) out[pos++] = v;
    }
    if (p.part + 1 == p.n_parts)
        for (uint32_t x = base + distinct + tid; x < p.out_stride; x += 256) p.hashes[(size_t)g * p.out_stride + x] = kPad;
    if (tid == 0) p.counts[g] = base + distinct;
}

// ------------------------------------------------------------------------------------------
// k = 21 scan with ROLLING state (the hot kernel of the whole path): a thread owns 64 consecutive
// k-mer start positions of its chunk and keeps, per position, in registers
//   * the 21 ASCII bytes of the forward k-mer (a byte-wise shift register over 6 words) and of its
//     reverse complement (shifted the other way): what MurmurHash3 reads -- no per-position bit
//     reversal / ASCII expansion;
//   * the MSB-first 2-bit integers of both strands, LEFT-aligned in 64 bits (canonical choice; the
//     15-mer of the K3 seed is the top 30 bits of the forward integer and bits 22..51 of the reverse
//     one).  Bases fall off the top by themselves (no mask), and since k is odd a k-mer never equals
//     its reverse complement, so the stale bits under the reverse integer cannot change
//     `forward < reverse`.
// One pass feeds K1 (MODE 0: MurmurHash3 candidates under the genome's threshold) or the marker
// sketch (MODE 1: mm_hash64 of the canonical 21-mer under the fixed threshold) AND, with SEEDS, the
// K3 seed selection bits (mm_hash64 of the canonical 15-mer < seed_thr).
// Loads are three aligned vector loads per thread (64 + 32 bases of sequence and validity); bases
// past the genome's (128-padded) span are never read.
// Instruction diet (ncu: r2_ncu_full_k1_scan21.txt -> r2_ncu_full_k1_scan21v2.txt, 204 -> 152
// thread-instructions per k-mer; the kernel is bound by the ALU pipe, 0.5 warp-instructions per
// clock per sub-partition for LOP3 / SHF / PRMT / SEL as measured by tools/pipe_bench.cu):
//   * no warm-up: the state after the first 20 bases is built directly from the loaded words
//     (bit-pair reversal for the forward integer, PRMT look-ups for the ASCII bytes);
//   * validity of all 21-base / 15-base windows of the thread as two 64-bit masks from five
//     shift-and-AND steps on the 96 validity bits, one bit test per position in the loop;
//   * 64-bit multiplies by constants as one wide multiply + two multiply-adds, MurmurHash3
//     specialised for seed 0;
//   * the last xor-shift of both fmix64 and the low half of the final sum are only evaluated for
//     hashes whose HIGH words can still be under the threshold (7 of 10,000).
// Tried and measured without gain: byte shifts as IMAD.HI + IMAD (IMAD.HI / IMAD.WIDE issue at
// 0.25 per clock, half of IMAD), strand choice by predicated multiplies (ptxas turns them back into
// SEL), 6 CTAs per SM at 40 registers.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint64_t pack64(uint32_t lo, uint32_t hi) { return (uint64_t)hi << 32 | lo; }
template <uint64_t C>
__device__ __forceinline__ uint64_t mul64c(uint64_t a) {
    const uint32_t alo = (uint32_t)a, ahi = (uint32_t)(a >> 32);
    uint64_t w;
    asm("mul.wide.u32 %0, %1, %2;" : "=l"(w) : "r"(alo), "r"((uint32_t)C));
    uint32_t lo = (uint32_t)w, hi = (uint32_t)(w >> 32);
    asm("mad.lo.u32 %0, %1, %2, %0;" : "+r"(hi) : "r"(alo), "r"((uint32_t)(C >> 32)));
    asm("mad.lo.u32 %0, %1, %2, %0;" : "+r"(hi) : "r"(ahi), "r"((uint32_t)C));
    return pack64(lo, hi);
}
// a * M + add, M < 2^32 (ptxas turns a constant `add` into IMAD.WIDE + IADD3 + IMAD.X whatever
// form it is handed in: immediate, constant bank or register pair -- measured, not fought further)
template <uint32_t M>
__device__ __forceinline__ uint64_t mad64s(uint64_t a, uint64_t add) {
    const uint32_t alo = (uint32_t)a, ahi = (uint32_t)(a >> 32);
    uint64_t w;
    asm("mad.wide.u32 %0, %1, %2, %3;" : "=l"(w) : "r"(alo), "r"(M), "l"(add));
    uint32_t hi = (uint32_t)(w >> 32);
    asm("mad.lo.u32 %0, %1, %2, %0;" : "+r"(hi) : "r"(ahi), "r"(M));
    return pack64((uint32_t)w, hi);
}
__device__ __forceinline__ uint64_t rotl64f(uint64_t x, int r) {  // r in 1..63, r != 32
    const uint32_t lo = (uint32_t)x, hi = (uint32_t)(x >> 32);
    if (r < 32) return pack64(__funnelshift_l(hi, lo, r), __funnelshift_l(lo, hi, r));
    return pack64(__funnelshift_l(lo, hi, r - 32), __funnelshift_l(hi, lo, r - 32));
}
// x ^= x >> 33 only touches the low word
__device__ __forceinline__ uint64_t xsr33(uint64_t x) {
    const uint32_t hi = (uint32_t)(x >> 32);
    return pack64((uint32_t)x ^ (hi >> 1), hi);
}
__device__ __forceinline__ uint64_t mm_hash64_v2(uint64_t key, uint64_t minus1) {
    key = mad64s<0x1FFFFFu>(key, minus1);  // ~key + (key << 21)
    key ^= key >> 24;
    key = mad64s<265u>(key, 0ull);
    key ^= key >> 14;
    key = mad64s<21u>(key, 0ull);
    key ^= key >> 28;
    key = mul64c<0x80000001ull>(key);
    return key;
}
// The same for a 30-bit key (the canonical 15-mer of a K3 seed), used ONLY for `hash < seed_thr`:
// key * (2^21 - 1) - 1 is below 2^51 for key >= 1, so the first xor-shift (by 24) leaves the high
// word alone -- two ALU-pipe instructions less.  key = 0 (poly-A) takes a different value this way
// (0xa992.. instead of 0x77cf..), both far above any threshold 2^64 / c with c >= 3, so the selection
// is the same (the host refuses a seed density above 1/3; checked against the full hash in
// tests/test_sketch_gpu.py through the seed lists of the K3 index).
__device__ __forceinline__ uint64_t mm_hash64_key30(uint32_t key30, uint64_t minus1) {
    uint64_t key = mad64s<0x1FFFFFu>((uint64_t)key30, minus1);
    const uint32_t lo = (uint32_t)key, hi = (uint32_t)(key >> 32);
    key = pack64(lo ^ __funnelshift_r(lo, hi, 24), hi);
    key = mad64s<265u>(key, 0ull);
    key ^= key >> 14;
    key = mad64s<21u>(key, 0ull);
    key ^= key >> 28;
    key = mul64c<0x80000001ull>(key);
    return key;
}
// the 96 validity bits ANDed with themselves shifted right by s
__device__ __forceinline__ void and_shr96(uint32_t &x0, uint32_t &x1, uint32_t &x2, int s) {
    const uint32_t y0 = __funnelshift_r(x0, x1, s), y1 = __funnelshift_r(x1, x2, s), y2 = x2 >> s;
    x0 &= y0; x1 &= y1; x2 &= y2;
}
// 2-bit groups of a word in reverse order
__device__ __forceinline__ uint32_t rev2_32(uint32_t v) {
    v = __brev(v);
    return ((v >> 1) & 0x55555555u) | ((v & 0x55555555u) << 1);
}
// 8 bases (low 16 bits of x) -> nibble selectors of two PRMTs (base 0 in nibble 0)
__device__ __forceinline__ uint32_t nibbles8(uint32_t x) {
    x = __byte_perm(x, 0u, 0x4140);               // bytes: b0, 0, b1, 0
    x = (x | (x << 4)) & 0x0F0F0F0Fu;
    return (x | (x << 2)) & 0x33333333u;
}

// a < b on 64-bit operands as the borrow of a - b: 0xFFFFFFFF if a < b, else 0.  Three add-with-carry
// instructions (IADD3 issues at twice the rate of the ALU pipe and beside it, profiles/r2_pipe_rates_b200.txt)
// in place of two ISETP; the result is used arithmetically, see sel32.
__device__ __forceinline__ uint32_t neg_borrow64(uint32_t alo, uint32_t ahi, uint32_t blo, uint32_t bhi) {
    uint32_t nb;
    asm("{\n\t.reg .u32 t0, t1;\n\tsub.cc.u32 t0, %1, %3;\n\tsubc.cc.u32 t1, %2, %4;\n\tsubc.u32 %0, 0, 0;\n\t}"
        : "=r"(nb) : "r"(alo), "r"(ahi), "r"(blo), "r"(bhi));
    return nb;
}
// nb == 0xFFFFFFFF ? f : r  as  r + nb * (r - f): a subtract and a multiply-add (FMA pipe) instead of
// a SEL (ALU pipe, the one the scan saturates)
__device__ __forceinline__ uint32_t sel32(uint32_t nb, uint32_t f, uint32_t r) { return r + nb * (r - f); }

template <int MODE>
__device__ __forceinline__ void scan_emit(const ChunkParams &p, uint32_t g, uint64_t T, uint64_t h) {
    if (h > T) return;
    if (h == kPad) { p.has_max[g] = 1; return; }
    if (MODE == 1 && p.n_parts > 1) {
        const uint32_t part = (uint32_t)__umul64hi(h * p.frac_c, (uint64_t)p.n_parts);
        const size_t buf = (size_t)g * p.n_parts + part;
        const uint32_t slot = atomicAdd(&p.cand_n[buf], 1u);
        if (slot < p.cap) p.cand[buf * p.cap + slot] = h;
    } else {
        const uint32_t slot = atomicAdd(&p.cand_n[g], 1u);
        if (slot < p.cap) p.cand[(size_t)g * p.cap + slot] = h;
    }
}

template <int MODE, bool SEEDS, bool SEED0>
__global__ void __launch_bounds__(256) scan21v2_kernel(const ChunkParams p) {
    const uint64_t item = blockIdx.x;
    if (item >= p.chunk_off[p.n]) return;
    const uint32_t g = p.item_genome[item];
    const uint64_t b0 = p.base_off[g], len = p.base_off[g + 1] - b0;  // the 128-padded span
    const uint64_t q0 = (item - p.chunk_off[g]) * kChunk + (uint64_t)threadIdx.x * 64;
    if (q0 >= len) return;
    const uint64_t P0 = b0 + q0;  // multiple of 64: the vector loads below are aligned
    const uint4 s03 = __ldg(reinterpret_cast<const uint4 *>(p.seq2 + (P0 >> 4)));
    const uint2 v01 = __ldg(reinterpret_cast<const uint2 *>(p.valid + (P0 >> 5)));
    uint2 s45 = make_uint2(0u, 0u);
    uint32_t v2 = 0u;
    if (q0 + 64 < len) {  // then q0 + 128 <= len: the next 32 bases belong to this genome
        s45 = __ldg(reinterpret_cast<const uint2 *>(p.seq2 + (P0 >> 4) + 4));
        v2 = __ldg(p.valid + (P0 >> 5) + 2);
    }
    const uint64_t T = MODE == 0 ? p.thr[g] : p.fixed_thr;
    uint32_t Th1 = min((uint32_t)(T >> 32), 0xFFFFFFFEu) + 1u;
    asm volatile("" : "+r"(Th1));  // keep it in a register: re-deriving it per position costs two ALU slots
    const uint64_t addA = 0x52dce729ull, addB = 0x38495ab5ull, minus1 = ~0ull;

    // ---- window validity: KV bit t = bases t .. t+20 valid, SV bit t = bases t .. t+14 valid
    uint32_t kv0, kv1, sv0, sv1;
    {
        uint32_t a0 = v01.x, a1 = v01.y, a2 = v2;
        and_shr96(a0, a1, a2, 1);   // windows of 2
        and_shr96(a0, a1, a2, 2);   // 4
        and_shr96(a0, a1, a2, 4);   // 8
        sv0 = a0 & __funnelshift_r(a0, a1, 7);   // [t, t+8) and [t+7, t+15)
        sv1 = a1 & __funnelshift_r(a1, a2, 7);
        and_shr96(a0, a1, a2, 8);   // 16
        kv0 = a0 & __funnelshift_r(a0, a1, 5);   // [t, t+16) and [t+5, t+21)
        kv1 = a1 & __funnelshift_r(a1, a2, 5);
    }
    if ((kv0 | kv1 | (SEEDS ? sv0 | sv1 : 0u)) == 0u) {  // no window of this thread is valid
        if (SEEDS) { uint32_t *out = p.sel + ((P0 - p.sel_first_base) >> 5); out[0] = 0u; out[1] = 0u; }
        return;
    }

    // ---- state after bases 0 .. 19 (what 20 rolling steps from zero would leave)
    uint32_t f0, f1, f2, f3, f4, f5;   // forward ASCII bytes: byte i + 1 = base i
    uint32_t r0, r1, r2, r3, r4, r5;   // reverse-complement ASCII bytes: byte j = comp(base 19 - j)
    uint32_t Flo, Fhi, Rlo, Rhi;        // left-aligned 2-bit integers (42 significant bits on top)
    {
        const uint32_t n0 = nibbles8(s03.x), n1 = nibbles8(s03.x >> 16), n2 = nibbles8(s03.y);
        const uint32_t A0 = __byte_perm(0x54474341u, 0u, n0), A1 = __byte_perm(0x54474341u, 0u, n0 >> 16);
        const uint32_t A2 = __byte_perm(0x54474341u, 0u, n1), A3 = __byte_perm(0x54474341u, 0u, n1 >> 16);
        const uint32_t A4 = __byte_perm(0x54474341u, 0u, n2);
        f0 = A0 << 8; f1 = __funnelshift_l(A0, A1, 8); f2 = __funnelshift_l(A1, A2, 8);
        f3 = __funnelshift_l(A2, A3, 8); f4 = __funnelshift_l(A3, A4, 8); f5 = A4 >> 24;
        const uint32_t B0 = __byte_perm(0x41434754u, 0u, n0), B1 = __byte_perm(0x41434754u, 0u, n0 >> 16);
        const uint32_t B2 = __byte_perm(0x41434754u, 0u, n1), B3 = __byte_perm(0x41434754u, 0u, n1 >> 16);
        const uint32_t B4 = __byte_perm(0x41434754u, 0u, n2);
        r0 = __byte_perm(B4, 0u, 0x0123); r1 = __byte_perm(B3, 0u, 0x0123); r2 = __byte_perm(B2, 0u, 0x0123);
        r3 = __byte_perm(B1, 0u, 0x0123); r4 = __byte_perm(B0, 0u, 0x0123); r5 = 0u;
        // forward: base i at bits 60 - 2 i (base 19 at 22..23)
        const uint64_t Fr = pack64(rev2_32(s03.y & 0xFFu), rev2_32(s03.x)) >> 2;
        Flo = (uint32_t)Fr; Fhi = (uint32_t)(Fr >> 32);
        // reverse: complement of base i at bits 24 + 2 i (base 19 on top)
        const uint64_t Rr = pack64(~s03.x, ~s03.y & 0xFFu) << 24;
        Rlo = (uint32_t)Rr; Rhi = (uint32_t)(Rr >> 32);
    }
    uint32_t sel0 = 0, sel1 = 0;
#pragma unroll
    for (int wi = 0; wi < 4; wi++) {
        // the 16 bases that enter during this block: bases 16 wi + 20 .. 16 wi + 35
        const uint32_t wa = wi == 0 ? s03.y : wi == 1 ? s03.z : wi == 2 ? s03.w : s45.x;
        const uint32_t wb = wi == 0 ? s03.z : wi == 1 ? s03.w : wi == 2 ? s45.x : s45.y;
        uint32_t inc = __funnelshift_r(wa, wb, 8);
        // validity bits of the block, at the bit positions `bit` walks over (no shifting in the loop)
        const uint32_t kv16 = wi < 2 ? kv0 : kv1, sv16 = wi < 2 ? sv0 : sv1;
        // `bit` walks over this block's 16 positions of the seed word and ends the loop (no counter)
        uint32_t bit = (wi & 1) ? 0x10000u : 1u;
        const uint32_t bit_end = (wi & 1) ? 0u : 0x10000u;
#pragma unroll 1
        do {
            const uint32_t code = inc & 3u;
            const uint32_t fa = __byte_perm(0x54474341u, 0u, code | 0x4440u);  // "ACGT"[code]
            f0 = __funnelshift_r(f0, f1, 8); f1 = __funnelshift_r(f1, f2, 8); f2 = __funnelshift_r(f2, f3, 8);
            f3 = __funnelshift_r(f3, f4, 8); f4 = __funnelshift_r(f4, f5, 8); f5 = fa;
            r5 = r4 >> 24; r4 = __funnelshift_l(r3, r4, 8); r3 = __funnelshift_l(r2, r3, 8);
            r2 = __funnelshift_l(r1, r2, 8); r1 = __funnelshift_l(r0, r1, 8);
            r0 = __byte_perm(r0, 0x41434754u, code | 0x2104u);  // bytes: "TGCA"[code], r0.b0, r0.b1, r0.b2
            // the integer updates as multiply-adds (FMA pipe; the ALU pipe is the one that is full):
            // forward  lo' = lo * 4 + code * 2^22
            // reverse  hi' = ((hi >> 2) | 3 << 30) - code * 2^30  (the funnel shift brings the two ones in)
            Fhi = __funnelshift_l(Flo, Fhi, 2);
            Flo = Flo * 4u + code * 0x00400000u;
            Rlo = __funnelshift_r(Rlo, Rhi, 2);
            Rhi = __funnelshift_r(Rhi, 0xFFFFFFFFu, 2) + code * 0xC0000000u;
            inc >>= 2;
            const uint32_t cur = bit;
            bit <<= 1;
            if (SEEDS && (sv16 & cur)) {
                const uint32_t F15 = Fhi >> 2, R15 = __funnelshift_r(Rlo, Rhi, 22) & 0x3FFFFFFFu;
                const uint64_t hs = mm_hash64_key30(min(F15, R15), minus1);
                const uint32_t pick = cur & neg_borrow64((uint32_t)hs, (uint32_t)(hs >> 32), (uint32_t)p.seed_thr,
                                                         (uint32_t)(p.seed_thr >> 32));
                if (wi < 2) sel0 |= pick; else sel1 |= pick;
            }
            if (kv16 & cur) {
                const uint32_t nb = neg_borrow64(Flo, Fhi, Rlo, Rhi);  // all ones: the forward strand is canonical
                const bool fwd = nb != 0u;
                if (MODE == 0) {
                    const uint64_t c1 = 0x87c37b91114253d5ull, c2 = 0x4cf5ad432745937full;
                    // three of the six words by multiply-add, three by SEL: measured best (both pipes level)
                    uint64_t k1 = mul64c<c1>(pack64(sel32(nb, f0, r0), sel32(nb, f1, r1)));
                    uint64_t k2 = mul64c<c2>(pack64(sel32(nb, f2, r2), fwd ? f3 : r3));
                    uint64_t kt = mul64c<c1>(pack64(fwd ? f4 : r4, fwd ? f5 : r5));
                    k1 = rotl64f(k1, 31); k1 = mul64c<c2>(k1);
                    k2 = rotl64f(k2, 33); k2 = mul64c<c1>(k2);
                    kt = rotl64f(kt, 31); kt = mul64c<c2>(kt);
                    uint64_t h1, h2;
                    if (SEED0) {
                        h1 = rotl64f(k1, 27); h1 = mad64s<5u>(h1, addA);
                        h2 = rotl64f(k2, 31); h2 += h1; h2 = mad64s<5u>(h2, addB);
                    } else {
                        h1 = p.seed ^ k1; h1 = rotl64f(h1, 27); h1 += p.seed; h1 = mad64s<5u>(h1, addA);
                        h2 = p.seed ^ k2; h2 = rotl64f(h2, 31); h2 += h1; h2 = mad64s<5u>(h2, addB);
                    }
                    h1 ^= kt;
                    h1 ^= 21ull; h2 ^= 21ull;
                    h1 += h2; h2 += h1;
                    h1 = mul64c<0xc4ceb9fe1a85ec53ull>(xsr33(mul64c<0xff51afd7ed558ccdull>(xsr33(h1))));
                    h2 = mul64c<0xc4ceb9fe1a85ec53ull>(xsr33(mul64c<0xff51afd7ed558ccdull>(xsr33(h2))));
                    // the last xor-shift leaves the high words as they are: hash.hi is s or s + 1
                    const uint32_t s1 = (uint32_t)(h1 >> 32) + (uint32_t)(h2 >> 32) + 1u;
                    if (s1 <= Th1) scan_emit<MODE>(p, g, T, xsr33(h1) + xsr33(h2));
                } else {
                    scan_emit<MODE>(p, g, T, mm_hash64_v2((fwd ? pack64(Flo, Fhi) : pack64(Rlo, Rhi)) >> 22, minus1));
                }
            }
        } while (bit != bit_end);
    }
    if (SEEDS) {
        uint32_t *out = p.sel + ((P0 - p.sel_first_base) >> 5);
        out[0] = sel0; out[1] = sel1;
        if (p.seed_count) {  // the lanes that are still here add up their seeds: one atomic per warp
            const unsigned lanes = __activemask();
            const uint32_t tot = __reduce_add_sync(lanes, (uint32_t)(__popc(sel0) + __popc(sel1)));
            if ((threadIdx.x & 31u) == (uint32_t)(__ffs((int)lanes) - 1) && tot) atomicAdd(p.seed_count + g, tot);
        }
    }
}

// One CTA per genome.  Dynamic shared memory: cap uint64.
__global__ void __launch_bounds__(256) sketch_select_kernel(const ChunkParams p) {
    extern __shared__ __align__(16) uint64_t sel[];
    __shared__ uint32_t s_warp_sum[8];
    const uint32_t g = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t got = p.cand_n[g];
    const bool overflow = got > p.cap;
    const uint32_t m = min(got, p.cap);
    // sort only the power of two that covers m
    uint32_t cap2 = 256;
    while (cap2 < m) cap2 <<= 1;
    const uint64_t *cand = p.cand + (size_t)g * p.cap;
    for (uint32_t x = tid; x < cap2; x += 256) sel[x] = x < m ? cand[x] : kPad;
    __syncthreads();
    for (uint32_t size = 2; size <= cap2; size <<= 1) {
        for (uint32_t str = size >> 1; str > 0; str >>= 1) {
            for (uint32_t x = tid; x < (cap2 >> 1); x += 256) {
                const uint32_t lo = 2 * x - (x & (str - 1));
                const uint32_t hi = lo + str;
                const bool up = (lo & size) == 0;
                const uint64_t a = sel[lo], b = sel[hi];
                if ((a > b) == up) { sel[lo] = b; sel[hi] = a; }
            }
            __syncthreads();
        }
    }
    // adjacent de-duplication: thread t owns the contiguous slice [t*per, (t+1)*per)
    const uint32_t per = cap2 / 256;
    const uint32_t x0 = tid * per;
    uint32_t mine = 0;
    for (uint32_t x = x0; x < x0 + per; x++) {
        const uint64_t v = sel[x];
        mine += (v != kPad && (x == 0 || sel[x - 1] != v)) ? 1u : 0u;
    }
    uint32_t incl = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
        if ((int)lane >= o) incl += t;
    }
    if (lane == 31) s_warp_sum[warp] = incl;
    __syncthreads();
    uint32_t base = 0, distinct = 0;
    for (uint32_t w = 0; w < 8; w++) {
        if (w < warp) base += s_warp_sum[w];
        distinct += s_warp_sum[w];
    }
    const bool has_max = p.has_max[g] != 0;
    const uint32_t total = distinct + (has_max ? 1u : 0u);
    const bool complete = !overflow && (p.fixed_thr != 0 || total >= p.s || p.thr[g] == kPad);
    if (!complete) {
        // adaptive mode: the exact kernel re-does the genome; FracMinHash mode: the candidate
        // buffer was too small -- flagged with an impossible count, the host reports it
        if (tid == 0) {
            if (p.fixed_thr) p.counts[g] = 0xFFFFFFFFu;
            else p.redo_list[atomicAdd(p.redo_n, 1u)] = g;
        }
        return;
    }
    const uint32_t out_n = min(total, p.s);
    uint64_t *out = p.hashes + (size_t)g * p.out_stride;
    uint32_t pos = base + incl - mine;
    // distinct values first (each thread writes the ones it owns) ...
    for (uint32_t x = x0; x < x0 + per; x++) {
        const uint64_t v = sel[x];
        if (v != kPad && (x == 0 || sel[x - 1] != v)) {
            if (pos < out_n) out[pos] = v;
            pos++;
        }
    }
    // ... then padding (a genuine 2^64-1 hash, if any, is the largest value and equals the pad)
    for (uint32_t x = min(distinct, out_n) + tid; x < p.out_stride; x += 256) out[x] = kPad;
    if (tid == 0) p.counts[g] = out_n;
}

uint32_t sketch_set_capacity(uint32_t s) {
    // expected candidates 1.3 s + 64; inserts stop at cap / 2 (+ one in flight per thread)
    uint32_t need = (uint32_t)(2.5 * (1.3 * s + 64.0));
    uint32_t cap = 1024;
    while (cap < need) cap <<= 1;
    return cap;
}

template <typename T>
static int sk_ensure(T *&ptr, size_t &cap, size_t need) {
    if (need <= cap) return 0;
    if (ptr) GB_CUDA(cudaFree(ptr));
    ptr = nullptr; cap = 0;
    GB_CUDA(cudaMalloc(&ptr, need * sizeof(T)));
    cap = need;
    return 0;
}

int sketch_enqueue(SketchWorkspace &ws, const uint32_t *d_seq2, const uint32_t *d_valid,
                   const uint64_t *d_base_off, size_t n, int k, uint32_t s, uint64_t seed,
                   uint64_t *d_hashes, uint32_t *d_counts, size_t out_stride, cudaStream_t stream,
                   const SeedSink *seeds, const uint64_t *host_span) {
    if (k < 1 || k > 32) { set_error("sketch: k must be in 1..32"); return 3; }
    if (seeds && k != 21) { set_error("sketch: fused seed marking needs k = 21"); return 3; }
    if (seeds && seeds->thr > ~0ull / 3) { set_error("sketch: seed density above 1/3 is not supported by the fused marking"); return 3; }
    if (s == 0) { set_error("sketch: s must be > 0"); return 3; }
    if (out_stride < s) { set_error("sketch: out_stride < s"); return 3; }
    if (n >= 0x7FFFFFFFull) { set_error("sketch: too many genomes in one batch"); return 3; }
    if (n == 0) return 0;
    const uint32_t cap = sketch_set_capacity(s);
    const size_t smem = (size_t)cap * 8;
    int dev = 0, sms = kNumSMsFallback, max_smem = 0;
    GB_CUDA(cudaGetDevice(&dev));
    GB_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    GB_CUDA(cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
    if (smem + 1024 > (size_t)max_smem) {
        set_error("sketch: num_kmers too large for the shared-memory candidate set");
        return 5;
    }
    // the scan grid is sized from the batch length: one 8-byte read of base_off[n]
    uint64_t total_bases = 0, first_base = 0;
    if (host_span) {
        first_base = host_span[0]; total_bases = host_span[1];
    } else {
        GB_CUDA(cudaMemcpyAsync(&total_bases, d_base_off + n, 8, cudaMemcpyDeviceToHost, stream));
        GB_CUDA(cudaMemcpyAsync(&first_base, d_base_off, 8, cudaMemcpyDeviceToHost, stream));
        GB_CUDA(cudaStreamSynchronize(stream));
    }
    const uint64_t max_items = (total_bases - first_base) / kChunk + n;
    if (max_items > 0x7FFFFFFFull) { set_error("sketch: batch too long for one launch"); return 3; }

    if (!ws.d_work_counter) GB_CUDA(cudaMalloc(&ws.d_work_counter, sizeof(unsigned long long)));
    if (!ws.d_redo_n) GB_CUDA(cudaMalloc(&ws.d_redo_n, sizeof(uint32_t)));
    if (sk_ensure(ws.d_chunk_off, ws.cap_chunk_off, n + 1) || sk_ensure(ws.d_thr, ws.cap_thr, n) ||
        sk_ensure(ws.d_cand_n, ws.cap_cand_n, n) || sk_ensure(ws.d_has_max, ws.cap_has_max, n) ||
        sk_ensure(ws.d_redo_list, ws.cap_redo, n) ||
        sk_ensure(ws.d_item_genome, ws.cap_items, (size_t)max_items + 1) ||
        sk_ensure(ws.d_cand, ws.cap_cand, n * (size_t)cap))
        return 2;
    GB_CUDA(cudaMemsetAsync(ws.d_work_counter, 0, sizeof(unsigned long long), stream));

    ChunkParams c;
    c.seq2 = d_seq2; c.valid = d_valid; c.base_off = d_base_off; c.n = (uint32_t)n; c.k = k; c.s = s;
    c.seed = seed; c.cap = cap; c.chunk_off = ws.d_chunk_off; c.thr = ws.d_thr; c.cand_n = ws.d_cand_n;
    c.has_max = ws.d_has_max; c.item_genome = ws.d_item_genome; c.max_items = max_items;
    c.cand = ws.d_cand; c.redo_list = ws.d_redo_list; c.redo_n = ws.d_redo_n;
    c.hashes = d_hashes; c.counts = d_counts; c.out_stride = (uint32_t)out_stride;
    c.fixed_thr = 0; c.n_parts = 0; c.part = 0; c.frac_c = 0;
    c.plan_k = seeds ? 15 : 0; c.sel = seeds ? seeds->d_sel : nullptr;
    c.sel_first_base = seeds ? seeds->first_base : 0; c.seed_thr = seeds ? seeds->thr : 0;
    c.seed_count = seeds ? seeds->d_seed_count : nullptr;
    if (c.seed_count) GB_CUDA(cudaMemsetAsync(c.seed_count, 0, n * sizeof(uint32_t), stream));

    sketch_plan_kernel<<<1, 1024, 0, stream>>>(c);
    GB_LAUNCH_CHECK();
    sketch_items_kernel<<<(uint32_t)((n * 32 + 255) / 256), 256, 0, stream>>>(c);
    GB_LAUNCH_CHECK();
    if (max_items > 0) {
        const bool generic = getenv("GALAH_B200_OLD_SCAN") != nullptr;  // the position-per-thread kernel, for A/B timing
        if (k == 21 && seeds && seed == 0) scan21v2_kernel<0, true, true><<<(uint32_t)max_items, 256, 0, stream>>>(c);
        else if (k == 21 && seeds) scan21v2_kernel<0, true, false><<<(uint32_t)max_items, 256, 0, stream>>>(c);
        else if (k == 21 && !generic && seed == 0) scan21v2_kernel<0, false, true><<<(uint32_t)max_items, 256, 0, stream>>>(c);
        else if (k == 21 && !generic) scan21v2_kernel<0, false, false><<<(uint32_t)max_items, 256, 0, stream>>>(c);
        else if (k == 21) sketch_scan_kernel<21, 0><<<(uint32_t)max_items, 256, 0, stream>>>(c);
        else sketch_scan_kernel<0, 0><<<(uint32_t)max_items, 256, 0, stream>>>(c);
        GB_LAUNCH_CHECK();
    }
    GB_CUDA(cudaFuncSetAttribute(sketch_select_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    sketch_select_kernel<<<(uint32_t)n, 256, smem, stream>>>(c);
    GB_LAUNCH_CHECK();

    // exact kernel over the redo list (normally empty: the CTAs read *redo_n == 0 and leave)
    SketchKernelParams p;
    p.seq2 = d_seq2; p.valid = d_valid; p.base_off = d_base_off; p.n = (uint32_t)n; p.k = k;
    p.s = s; p.seed = seed; p.hashes = d_hashes; p.counts = d_counts;
    p.out_stride = (uint32_t)out_stride; p.cap = cap; p.work_counter = ws.d_work_counter;
    p.redo_list = ws.d_redo_list; p.redo_n = ws.d_redo_n;
    const int ctas_per_sm = std::max(1, std::min(4, (int)((size_t)max_smem / (smem + 2048))));
    const uint32_t grid = (uint32_t)std::min<size_t>(n, (size_t)sms * ctas_per_sm);
    if (k == 21) {
        GB_CUDA(cudaFuncSetAttribute(sketch_kernel<21>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int)smem));
        sketch_kernel<21><<<grid, kSketchThreads, smem, stream>>>(p);
    } else {
        GB_CUDA(cudaFuncSetAttribute(sketch_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int)smem));
        sketch_kernel<0><<<grid, kSketchThreads, smem, stream>>>(p);
    }
    GB_LAUNCH_CHECK();
    return 0;
}

// FracMinHash marker sketches (skani-style, k = 21): every distinct mm_hash64(canonical k-mer) below
// (2^64-1)/c, ascending, one row of `cap` uint64 per genome (cap a power of two, <= 16384 so that the
// per-genome sort fits shared memory).  counts[g] == 0xFFFFFFFF flags a genome with more than cap
// markers.  Shares the plan / items / scan / select pipeline of K1.
int marker_sketch_enqueue(SketchWorkspace &ws, const uint32_t *d_seq2, const uint32_t *d_valid,
                          const uint64_t *d_base_off, size_t n, int k, uint32_t c_marker, uint32_t cap,
                          uint64_t *d_rows, uint32_t *d_counts, cudaStream_t stream, const SeedSink *seeds,
                          const uint64_t *host_span) {
    if (k < 1 || k > 32) { set_error("marker sketch: k must be in 1..32"); return 3; }
    if (seeds && k != 21) { set_error("marker sketch: fused seed marking needs k = 21"); return 3; }
    if (seeds && seeds->thr > ~0ull / 3) { set_error("marker sketch: seed density above 1/3 is not supported by the fused marking"); return 3; }
    if (cap < 256 || (cap & (cap - 1)) || cap > kMarkerMaxCap) { set_error("marker sketch: bad capacity"); return 3; }
    const uint32_t n_parts = cap > kMarkerPartCap ? cap / kMarkerPartCap : 1u;  // value-range partitions of a wide row
    const uint32_t part_cap = cap > kMarkerPartCap ? kMarkerPartCap : cap;
    if (n >= 0x7FFFFFFFull) { set_error("marker sketch: too many genomes in one batch"); return 3; }
    if (n == 0) return 0;
    uint64_t total_bases = 0, first_base = 0;
    if (host_span) {
        first_base = host_span[0]; total_bases = host_span[1];
    } else {
        GB_CUDA(cudaMemcpyAsync(&total_bases, d_base_off + n, 8, cudaMemcpyDeviceToHost, stream));
        GB_CUDA(cudaMemcpyAsync(&first_base, d_base_off, 8, cudaMemcpyDeviceToHost, stream));
        GB_CUDA(cudaStreamSynchronize(stream));
    }
    const uint64_t max_items = (total_bases - first_base) / kChunk + n;
    if (max_items > 0x7FFFFFFFull) { set_error("marker sketch: batch too long for one launch"); return 3; }
    if (!ws.d_redo_n) GB_CUDA(cudaMalloc(&ws.d_redo_n, sizeof(uint32_t)));
    if (sk_ensure(ws.d_chunk_off, ws.cap_chunk_off, n + 1) || sk_ensure(ws.d_thr, ws.cap_thr, n) ||
        sk_ensure(ws.d_cand_n, ws.cap_cand_n, n * (size_t)n_parts) || sk_ensure(ws.d_has_max, ws.cap_has_max, n) ||
        sk_ensure(ws.d_redo_list, ws.cap_redo, n) ||
        sk_ensure(ws.d_item_genome, ws.cap_items, (size_t)max_items + 1) ||
        sk_ensure(ws.d_cand, ws.cap_cand, n * (size_t)cap))
        return 2;
    ChunkParams c;
    c.seq2 = d_seq2; c.valid = d_valid; c.base_off = d_base_off; c.n = (uint32_t)n; c.k = k; c.s = cap;
    c.n_parts = n_parts; c.part = 0; c.frac_c = c_marker;
    c.seed = 0; c.cap = part_cap; c.chunk_off = ws.d_chunk_off; c.thr = ws.d_thr; c.cand_n = ws.d_cand_n;
    c.has_max = ws.d_has_max; c.item_genome = ws.d_item_genome; c.max_items = max_items;
    c.cand = ws.d_cand; c.redo_list = ws.d_redo_list; c.redo_n = ws.d_redo_n;
    c.hashes = d_rows; c.counts = d_counts; c.out_stride = cap;
    c.fixed_thr = ~0ull / c_marker - 1;  // keep h < (2^64-1)/c  <=>  h <= that - 1
    c.plan_k = seeds ? 15 : 0; c.sel = seeds ? seeds->d_sel : nullptr;
    c.sel_first_base = seeds ? seeds->first_base : 0; c.seed_thr = seeds ? seeds->thr : 0;
    c.seed_count = seeds ? seeds->d_seed_count : nullptr;
    if (c.seed_count) GB_CUDA(cudaMemsetAsync(c.seed_count, 0, n * sizeof(uint32_t), stream));
    sketch_plan_kernel<<<1, 1024, 0, stream>>>(c);
    GB_LAUNCH_CHECK();
    if (n_parts > 1) GB_CUDA(cudaMemsetAsync(ws.d_cand_n, 0, n * (size_t)n_parts * sizeof(uint32_t), stream));
    sketch_items_kernel<<<(uint32_t)((n * 32 + 255) / 256), 256, 0, stream>>>(c);
    GB_LAUNCH_CHECK();
    if (max_items > 0) {
        if (k == 21 && seeds) scan21v2_kernel<1, true, true><<<(uint32_t)max_items, 256, 0, stream>>>(c);
        else if (k == 21 && !getenv("GALAH_B200_OLD_SCAN")) scan21v2_kernel<1, false, true><<<(uint32_t)max_items, 256, 0, stream>>>(c);
        else if (k == 21) sketch_scan_kernel<21, 1><<<(uint32_t)max_items, 256, 0, stream>>>(c);
        else sketch_scan_kernel<0, 1><<<(uint32_t)max_items, 256, 0, stream>>>(c);
        GB_LAUNCH_CHECK();
    }
    const size_t smem = (size_t)part_cap * 8;
    if (n_parts == 1) {
        GB_CUDA(cudaFuncSetAttribute(sketch_select_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        sketch_select_kernel<<<(uint32_t)n, 256, smem, stream>>>(c);
        GB_LAUNCH_CHECK();
        return 0;
    }
    // the plan kernel zeroed cand_n[0..n) only: the partitioned buffers count in n * n_parts slots
    GB_CUDA(cudaFuncSetAttribute(marker_select_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    for (uint32_t part = 0; part < n_parts; part++) {
        c.part = part;
        marker_select_kernel<<<(uint32_t)n, 256, smem, stream>>>(c);
        GB_LAUNCH_CHECK();
    }
    return 0;
}

int SketchWorkspace::release() {
    cudaFree(d_work_counter); cudaFree(d_redo_n); cudaFree(d_chunk_off); cudaFree(d_thr);
    cudaFree(d_cand_n); cudaFree(d_has_max); cudaFree(d_redo_list); cudaFree(d_item_genome);
    cudaFree(d_cand);
    *this = SketchWorkspace();
    return 0;
}

}  // namespace gb200
