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
// Kernel: persistent CTAs take genomes from an atomic counter.  Every thread extracts a k-mer
// straight from the packed stream (no rolling state, so no warm-up and no divergence):
//     V = 2k bits at bit offset 2p           (LSB-first forward k-mer)
//     F = reverse-2-bit-groups(V)            (MSB-first forward integer)
//     R = ~V & mask                          (MSB-first reverse-complement integer)
//     canonical LSB-first word W = F < R ? V : ~F & mask
// W is expanded to ASCII 8 bases at a time (bit spread + PRMT against the constant "ACGT"),
// hashed, and hashes <= a per-genome threshold T go into a shared-memory open-addressing set
// (atomicCAS, duplicates collapse).  T starts at ~1.3 s/n_kmers of the hash range; if the set
// ends with fewer than s distinct values (or overflows) T is bisected and the genome is
// re-scanned, so the result is exact for any input.  The set is then bitonic-sorted in place
// (empty slots hold 2^64-1 and sink to the end) and the first s values are written out.
#include "sketch.cuh"

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
        const unsigned long long g = s_genome;
        if (g >= p.n) break;
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

uint32_t sketch_set_capacity(uint32_t s) {
    // expected candidates 1.3 s + 64; inserts stop at cap / 2 (+ one in flight per thread)
    uint32_t need = (uint32_t)(2.5 * (1.3 * s + 64.0));
    uint32_t cap = 1024;
    while (cap < need) cap <<= 1;
    return cap;
}

int sketch_enqueue(SketchWorkspace &ws, const uint32_t *d_seq2, const uint32_t *d_valid,
                   const uint64_t *d_base_off, size_t n, int k, uint32_t s, uint64_t seed,
                   uint64_t *d_hashes, uint32_t *d_counts, size_t out_stride, cudaStream_t stream) {
    if (k < 1 || k > 32) { set_error("sketch: k must be in 1..32"); return 3; }
    if (s == 0) { set_error("sketch: s must be > 0"); return 3; }
    if (out_stride < s) { set_error("sketch: out_stride < s"); return 3; }
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
    if (!ws.d_work_counter) GB_CUDA(cudaMalloc(&ws.d_work_counter, sizeof(unsigned long long)));
    GB_CUDA(cudaMemsetAsync(ws.d_work_counter, 0, sizeof(unsigned long long), stream));
    SketchKernelParams p;
    p.seq2 = d_seq2; p.valid = d_valid; p.base_off = d_base_off; p.n = (uint32_t)n; p.k = k;
    p.s = s; p.seed = seed; p.hashes = d_hashes; p.counts = d_counts;
    p.out_stride = (uint32_t)out_stride; p.cap = cap; p.work_counter = ws.d_work_counter;
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

int SketchWorkspace::release() {
    cudaFree(d_work_counter);
    d_work_counter = nullptr;
    return 0;
}

}  // namespace gb200
