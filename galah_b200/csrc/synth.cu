// Synthetic genome generator: counter-based, so any genome can be regenerated
// from (seed, index) -- the CPU oracle holds the same definition for parity checks.
//   family f = index / F, member m = (index % F) % 10, substitution rate {0,.5,1,2,3,4,5,6,8,10} % >> rate_shift
//   (F = 10, rate_shift = 0: the sparse benchmark families; F = thousands, rate_shift = 2: one dense clade)
//   founder block b (32 bases, 2 bits each, LSB first) = mix(key(seed, 2f, b))
//   mutation draws for genome g: words mix(key(seed, 2g+1, 16b + w)), w = 0..7 give one 16-bit
//   uniform per base (mutate iff u16 < round(rate * 65536)), w = 8 picks the replacement base
//   (new = (old + 1 + r % 3) & 3).
// Output is written directly in the packed layout the sketch kernel reads.
#include "common.cuh"
#include "sketch.cuh"

namespace gb200 {

__device__ __forceinline__ uint64_t splitmix64(uint64_t x) {
    x += 0x9e3779b97f4a7c15ull;
    x = (x ^ (x >> 30)) * 0xbf58476d1ce4e5b9ull;
    x = (x ^ (x >> 27)) * 0x94d049bb133111ebull;
    return x ^ (x >> 31);
}
__device__ __forceinline__ uint64_t synth_word(uint64_t seed, uint64_t stream, uint64_t ctr) {
    return splitmix64(splitmix64(seed ^ (stream * 0xd1342543de82ef95ull)) + ctr * 0x2545f4914f6cdd1dull);
}
__constant__ uint32_t kSynthRateU16[10] = {0, 328, 655, 1311, 1966, 2621, 3277, 3932, 5243, 6554};

__global__ void __launch_bounds__(256) synth_kernel(uint64_t seed, uint64_t index_begin, uint32_t n,
                                                    uint64_t length, uint64_t blocks_per_genome,
                                                    uint32_t family_size, uint32_t rate_shift,
                                                    uint2 *__restrict__ seq2, uint32_t *__restrict__ valid,
                                                    uint64_t *__restrict__ base_off) {
    const uint64_t gid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint64_t total = (uint64_t)n * blocks_per_genome;
    if (gid <= n) base_off[gid] = gid * blocks_per_genome * 32;
    for (uint64_t x = gid; x < total; x += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t g = x / blocks_per_genome, b = x % blocks_per_genome;
        const uint64_t index = index_begin + g;
        const uint64_t fam = index / family_size, mem = (index % family_size) % 10;
        uint64_t w = synth_word(seed, 2 * fam, b);
        const uint32_t thr = kSynthRateU16[mem] >> rate_shift;
        if (thr != 0) {
            const uint64_t sel = synth_word(seed, 2 * index + 1, 16 * b + 8);
#pragma unroll
            for (int q = 0; q < 8; q++) {
                const uint64_t u = synth_word(seed, 2 * index + 1, 16 * b + q);
#pragma unroll
                for (int t = 0; t < 4; t++) {
                    const uint32_t u16 = (uint32_t)(u >> (16 * t)) & 0xffffu;
                    if (u16 < thr) {
                        const int pos = q * 4 + t;
                        const uint64_t old = (w >> (2 * pos)) & 3;
                        const uint64_t r = (sel >> (2 * pos)) & 3;
                        const uint64_t nw = (old + 1 + (r % 3)) & 3;
                        w = (w & ~(3ull << (2 * pos))) | (nw << (2 * pos));
                    }
                }
            }
        }
        // bases at or beyond `length` are padding: invalid
        const uint64_t first = b * 32;
        uint32_t vm = 0xFFFFFFFFu;
        if (first >= length) vm = 0;
        else if (first + 32 > length) vm = (1u << (uint32_t)(length - first)) - 1u;
        seq2[x] = make_uint2((uint32_t)w, (uint32_t)(w >> 32));
        valid[x] = vm;
    }
}

int synth_enqueue(uint64_t seed, uint64_t index_begin, size_t n, uint64_t length, uint32_t *d_seq2,
                  uint32_t *d_valid, uint64_t *d_base_off, cudaStream_t stream, uint32_t family_size,
                  uint32_t rate_shift) {
    if (n == 0) return 0;
    if (family_size == 0 || rate_shift > 15) { set_error("synth: bad family_size / rate_shift"); return 3; }
    if (n >= 0xFFFFFFFFull) { set_error("synth: n too large"); return 3; }
    const uint64_t padded = (length + 127) / 128 * 128;
    const uint64_t bpg = padded / 32;
    const uint64_t total = (uint64_t)n * bpg;
    uint64_t grid = (total + 255) / 256;
    if (grid < (n + 256) / 256) grid = (n + 256) / 256;
    if (grid > 148ull * 64) grid = 148ull * 64;
    // base_off needs gid <= n covered by the first pass of the grid
    if (grid * 256 <= n) { set_error("synth: n too large for base_off pass"); return 3; }
    synth_kernel<<<(uint32_t)grid, 256, 0, stream>>>(seed, index_begin, (uint32_t)n, length, bpg, family_size, rate_shift,
                                                     reinterpret_cast<uint2 *>(d_seq2), d_valid,
                                                     d_base_off);
    GB_LAUNCH_CHECK();
    return 0;
}

}  // namespace gb200
