// Stage 2 (K3): FracMinHash seed-and-chain ANI on sm_100a.
//
// Replaces the `skani dist` subprocess spawned per genome pair at
// /root/reference/src/skani.rs:718-788 (ClusterDistanceFinder::calculate_ani, src/lib.rs:54).
// The algorithm is specified, constant by constant, in oracle/skani_oracle.c (a restatement of
// skani's published method without its learned correction; numeric parity with the skani binary
// is unpinned).  Everything on the device is integer arithmetic, so the kernels
// are bit-exact against that restatement.
//
// Index build (once per genome, genomes stay resident in HBM):
//   ani_mark_kernel   every k-mer start position of the batch: canonical 15-mer straight from the
//                     packed 2-bit stream (funnel shift + bit reversal, as K1), mm_hash64,
//                     selected iff hash < (2^64-1)/c; one ballot word per 32 positions.
//   ani_count_kernel  seeds per genome (popc reduction) -> host sizes the arrays.
//   ani_emit_kernel   one CTA (32 warps) per genome: per-warp counts + warp scans give every seed its rank,
//                     so seeds land in POSITION ORDER without a sort; per seed the contig (binary
//                     search), spread position and chunk id; then the chunk -> seed-range table
//                     and the genome's open-addressing hash table
//                     (entry = kmer << 33 | strand << 32 | spread, atomicCAS linear probing).
// Pair evaluation:
//   ani_chain_kernel  one THREAD per (pair, query chunk): walks the chunk's seeds in order,
//                     probes the reference genome's hash table (<= 8 occurrences, emitted in
//                     ascending reference position), and runs the banded chaining DP over a ring
//                     of the last 16 anchors held in shared memory ([slot][field][thread], so a
//                     warp's accesses never conflict).  The walk is a state machine (advance to
//                     the next anchor, then one chaining step) so that the lanes of a warp --
//                     different chunks -- run the chaining step converged.  Each anchor carries (count, first seed,
//                     first reference position) of its best chain, so no backtracking pass.
//                     Accepted chunks add (M-2, N-2, covq, covr) to the pair's accumulators.
//                     Bound: latency of random 8-byte reads of the reference table (L2/HBM);
//                     algorithmic bytes per pair = 2 * (L/c) * 12 B.
#include "ani.cuh"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <chrono>
#include <mutex>
#include <thread>

#include "common.cuh"

namespace gb200 {

template <typename T>
int DevVec<T>::reserve(size_t need, cudaStream_t st) {
    if (need <= cap) return 0;
    size_t ncap = std::max(need, cap + cap / 2 + 1024);
    T *np = nullptr;
    GB_CUDA(cudaMalloc(&np, ncap * sizeof(T)));
    if (p && n) GB_CUDA(cudaMemcpyAsync(np, p, n * sizeof(T), cudaMemcpyDeviceToDevice, st));
    GB_CUDA(cudaStreamSynchronize(st));
    if (p) GB_CUDA(cudaFree(p));
    p = np; cap = ncap;
    return 0;
}

constexpr unsigned long long kEmpty = ~0ull;
// bit 63 of a bucket's LAST slot: some key whose home is this bucket lives beyond it (k-mers are 30
// bits, so bits 33..62 hold them and bit 63 is free; an empty slot is all ones)
constexpr unsigned long long kOverflowFlag = 1ull << 63;

__device__ __forceinline__ uint64_t mm_hash64(uint64_t key) {
    key = ~key + (key << 21);
    key = key ^ (key >> 24);
    key = (key + (key << 3)) + (key << 8);
    key = key ^ (key >> 14);
    key = (key + (key << 2)) + (key << 4);
    key = key ^ (key >> 28);
    key = key + (key << 31);
    return key;
}

// Canonical 15-mer at absolute packed position P; false if the window holds an invalid base.
__device__ __forceinline__ bool canon15(const uint32_t *__restrict__ seq2, const uint32_t *__restrict__ valid,
                                        uint64_t P, uint32_t &canon, uint32_t &strand) {
    const uint64_t vw = P >> 5;
    const uint32_t vm = __funnelshift_r(__ldg(valid + vw), __ldg(valid + vw + 1), (uint32_t)P & 31u);
    if ((vm & 0x7FFFu) != 0x7FFFu) return false;
    const uint64_t w = P >> 4;
    const uint32_t V = __funnelshift_r(__ldg(seq2 + w), __ldg(seq2 + w + 1), ((uint32_t)P & 15u) * 2u) & 0x3FFFFFFFu;
    uint32_t F = __brev(V);
    F = ((F >> 1) & 0x55555555u) | ((F & 0x55555555u) << 1);
    F >>= 2;
    const uint32_t R = ~V & 0x3FFFFFFFu;
    canon = min(F, R);
    strand = R < F ? 1u : 0u;
    return true;
}

__global__ void __launch_bounds__(256) ani_mark_kernel(const uint32_t *__restrict__ seq2,
                                                       const uint32_t *__restrict__ valid, uint64_t first_base,
                                                       uint64_t end_base, uint64_t thr,
                                                       uint32_t *__restrict__ sel) {
    const uint64_t step = (uint64_t)gridDim.x * blockDim.x;
    // first_base and end_base are multiples of 128, so every warp covers whole mask words
    for (uint64_t P = first_base + (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; P < end_base; P += step) {
        uint32_t canon, strand;
        bool pick = canon15(seq2, valid, P, canon, strand);
        if (pick) pick = mm_hash64(canon) < thr;
        const uint32_t word = __ballot_sync(0xffffffffu, pick);
        if ((threadIdx.x & 31) == 0) sel[(P - first_base) >> 5] = word;
    }
}

// Mask word w (absolute word index relative to first_base) of a genome whose k-mer start
// positions end at `limit` (relative to first_base): a window must not run into the next genome.
__device__ __forceinline__ uint32_t sel_word(const uint32_t *__restrict__ sel, uint64_t w, uint64_t limit) {
    const uint64_t wp = w << 5;
    if (wp >= limit) return 0u;
    const uint32_t word = sel[w];
    return limit - wp >= 32 ? word : word & ((1u << (uint32_t)(limit - wp)) - 1u);
}

__global__ void __launch_bounds__(256) ani_count_kernel(const uint32_t *__restrict__ sel,
                                                        const uint64_t *__restrict__ base_off, uint64_t first_base,
                                                        uint32_t *__restrict__ count) {
    __shared__ uint32_t s_sum[8];
    const uint32_t g = blockIdx.x;
    const uint64_t w0 = (base_off[g] - first_base) >> 5, w1 = (base_off[g + 1] - first_base) >> 5;
    const uint64_t len = base_off[g + 1] - base_off[g];
    const uint64_t limit = (base_off[g] - first_base) + (len >= (uint64_t)kAniK ? len - kAniK + 1 : 0);
    uint32_t sum = 0;
    for (uint64_t w = w0 + threadIdx.x; w < w1; w += 256) sum += __popc(sel_word(sel, w, limit));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    if ((threadIdx.x & 31) == 0) s_sum[threadIdx.x >> 5] = sum;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t t = 0;
        for (int w = 0; w < 8; w++) t += s_sum[w];
        count[g] = t;
    }
}

struct EmitParams {
    const uint32_t *seq2, *valid, *sel;
    const uint64_t *base_off;      // batch-local [n + 1]
    uint64_t first_base;
    const uint64_t *contig_off;    // batch-local [n + 1]
    const uint32_t *contig_start;  // per contig, relative to the genome's first base
    const uint32_t *contig_chunk_base;
    const uint64_t *seed_off;      // batch-local [n + 1], absolute offsets into ks/spread
    const uint64_t *cso_off;       // batch-local [n + 1], absolute offsets into cso
    const uint32_t *n_chunks;      // batch-local [n]
    const uint64_t *table_off;     // batch-local [n + 1], absolute offsets into table
    uint2 *kq;  // per seed: x = kmer << 1 | strand, y = spread position (one 8-byte read per seed)
    uint32_t *cso, *chunk_tmp;
    unsigned long long *table;
    uint32_t *mismatch;  // set to genome + 1 if the selection bits hold another number of seeds than the host was told
};

// One bucket (four 8-byte slots, 32-byte aligned) in ONE 256-bit load (SASS LDG.E.256, sm_100):
// per lane a gather costs the LSU one transaction per instruction, so two 128-bit loads of the
// same sector would cost two.
__device__ __forceinline__ void load_bucket(const unsigned long long *p, unsigned long long (&v)[4]) {
    asm volatile("ld.global.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(v[0]), "=l"(v[1]), "=l"(v[2]), "=l"(v[3]) : "l"(p));
}

// Home slot of a k-mer: the start of a bucket of four slots (one aligned 32-byte sector).  Keys
// probe linearly from there, so a reader sees a whole bucket per memory transaction, and at a
// load factor <= 1/4 (about one key per bucket) a run almost never leaves its first sector.
__device__ __forceinline__ uint32_t table_slot(uint32_t km, uint32_t mask) {
    return (uint32_t)(((uint64_t)km * 0x9E3779B97F4A7C15ull) >> 32) & mask & ~3u;
}

// One CTA of 32 warps per genome.  A warp owns a contiguous 1/32 of the genome's mask words: it
// counts its seeds first (one pass of popcounts), the 32 counts are scanned through shared memory,
// and every warp then ranks and writes its own seeds with warp shuffles only -- the 244 block-wide
// barrier rounds of the 256-thread version (1.09 ms per 512-genome batch, all of it latency) are
// gone and four times as many dependent load chains are in flight per genome.
constexpr int kEmitThreads = 1024;
constexpr int kEmitWords = 4;     // mask words per lane and round
constexpr int kEmitStage = 256;   // staged seeds per warp and round (8x the expectation at 1 seed per 30 positions)
__global__ void __launch_bounds__(kEmitThreads, 2) ani_emit_kernel(const EmitParams p) {
    __shared__ uint32_t s_warp[32];
    __shared__ uint32_t s_stage[32][kEmitStage];
    const uint32_t g = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint64_t b0 = p.base_off[g];
    const uint64_t w0 = (b0 - p.first_base) >> 5, w1 = (p.base_off[g + 1] - p.first_base) >> 5;
    const uint64_t so = p.seed_off[g];
    const uint32_t n_seeds = (uint32_t)(p.seed_off[g + 1] - so);
    const uint64_t c0 = p.contig_off[g];
    const uint32_t n_contigs = (uint32_t)(p.contig_off[g + 1] - c0);
    const uint64_t glen = p.base_off[g + 1] - b0;
    const uint64_t limit = (b0 - p.first_base) + (glen >= (uint64_t)kAniK ? glen - kAniK + 1 : 0);
    const uint64_t per = ((w1 - w0 + 31) / 32 + 31) & ~31ull;  // mask words per warp, a multiple of 32
    const uint64_t wa = min(w0 + warp * per, w1), wz = min(wa + per, w1);
    uint32_t mine = 0;
    for (uint64_t w = wa + lane; w < wz; w += 32) mine += __popc(sel_word(p.sel, w, limit));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mine += __shfl_xor_sync(0xffffffffu, mine, o);
    if (lane == 0) s_warp[warp] = mine;
    __syncthreads();
    uint32_t running = 0;
    for (uint32_t x = 0; x < warp; x++) running += s_warp[x];
    if (warp == 31) {  // the arrays were sized from a count made elsewhere (the scan's own, or ani_count_kernel)
        if (running + mine != n_seeds) { if (lane == 0) atomicExch(p.mismatch, g + 1u); return; }
    } else {
        uint32_t total = running;
        for (uint32_t x = warp; x < 32; x++) total += s_warp[x];
        if (total != n_seeds) return;  // whole CTA leaves: nothing is written past the space the host reserved
    }
    // One seed: canonical 15-mer, contig, spread position, chunk -> slot `rank` of the genome's arrays.
    auto emit_seed = [&](const uint32_t rel, const uint32_t rank) {
        uint32_t canon = 0, strand = 0;
        canon15(p.seq2, p.valid, b0 + rel, canon, strand);
        uint32_t lo = 0, hi = n_contigs;  // last contig with start <= rel
        while (hi - lo > 1) {
            const uint32_t mid = (lo + hi) >> 1;
            if (p.contig_start[c0 + mid] <= rel) lo = mid; else hi = mid;
        }
        p.kq[so + rank] = make_uint2((canon << 1) | strand, rel + lo * (uint32_t)(kAniBand + 1));
        p.chunk_tmp[so + rank] = p.contig_chunk_base[c0 + lo] + (rel - p.contig_start[c0 + lo]) / kAniChunk;
    };
    // A warp takes 128 mask words per round (four per lane, loaded together), ranks their set bits with
    // four warp scans and STAGES the seed positions in shared memory in rank order; the lanes then take
    // one staged seed each, so the dependent loads of a round (the k-mer's bases, the contig table) are
    // in flight for 32 seeds at once instead of behind each other in a per-lane loop over a word's bits.
    uint32_t *stage = s_stage[warp];
    const uint64_t rel0 = (b0 - p.first_base);  // bit position of the genome's first base in the mask
    for (uint64_t wb = wa; wb < wz; wb += 32 * kEmitWords) {
        uint32_t word[kEmitWords], at[kEmitWords];
        uint32_t tot = 0;
#pragma unroll
        for (int u = 0; u < kEmitWords; u++) {
            const uint64_t w = wb + 32 * u + lane;
            word[u] = w < wz ? sel_word(p.sel, w, limit) : 0u;
        }
#pragma unroll
        for (int u = 0; u < kEmitWords; u++) {
            const uint32_t pc = __popc(word[u]);
            uint32_t incl = pc;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
                if ((int)lane >= o) incl += t;
            }
            at[u] = tot + incl - pc;
            tot += __shfl_sync(0xffffffffu, incl, 31);
        }
        if (tot <= (uint32_t)kEmitStage) {
#pragma unroll
            for (int u = 0; u < kEmitWords; u++) {
                uint32_t bits = word[u], x = at[u];
                const uint32_t wrel = (uint32_t)(((wb + 32 * u + lane) << 5) - rel0);
                while (bits) {
                    stage[x++] = wrel + (uint32_t)(__ffs(bits) - 1);
                    bits &= bits - 1;
                }
            }
            __syncwarp();
            for (uint32_t x = lane; x < tot; x += 32) emit_seed(stage[x], running + x);
            __syncwarp();
        } else {  // a round denser than the staging area (never at 1 seed per 30 .. 125 positions of real sequence)
#pragma unroll
            for (int u = 0; u < kEmitWords; u++) {
                uint32_t bits = word[u], x = running + at[u];
                const uint32_t wrel = (uint32_t)(((wb + 32 * u + lane) << 5) - rel0);
                while (bits) {
                    emit_seed(wrel + (uint32_t)(__ffs(bits) - 1), x++);
                    bits &= bits - 1;
                }
            }
        }
        running += tot;
    }
    __syncthreads();  // every seed of the genome is written before the tables below read them
    // chunk -> first seed table (cso[t] = index of the first seed with chunk >= t; cso[n_chunks] = n_seeds)
    const uint32_t nch = p.n_chunks[g];
    uint32_t *cso = p.cso + p.cso_off[g];
    if (n_seeds == 0) {
        for (uint32_t t = tid; t <= nch; t += kEmitThreads) cso[t] = 0;
    } else {
        for (uint32_t x = tid; x < n_seeds; x += kEmitThreads) {
            const uint32_t ch = p.chunk_tmp[so + x];
            const uint32_t first = x == 0 ? 0 : p.chunk_tmp[so + x - 1] + 1;
            for (uint32_t t = first; t <= ch; t++) cso[t] = x;
            if (x == n_seeds - 1) for (uint32_t t = ch + 1; t <= nch; t++) cso[t] = n_seeds;
        }
    }
    // hash table of this genome's seeds
    const uint64_t to = p.table_off[g];
    const uint32_t mask = (uint32_t)(p.table_off[g + 1] - to) - 1;
    unsigned long long *table = p.table + to;
    for (uint32_t x = tid; x < n_seeds; x += kEmitThreads) {
        const uint2 kq = p.kq[so + x];
        const uint32_t ks = kq.x;
        const unsigned long long e = ((unsigned long long)(ks >> 1) << 33) | ((unsigned long long)(ks & 1) << 32) | kq.y;
        const uint32_t home = table_slot(ks >> 1, mask);
        uint32_t slot = home;
        while (atomicCAS(&table[slot], kEmpty, e) != kEmpty) slot = (slot + 1) & mask;
        // a key that left its home bucket flags the bucket (its last slot is occupied by now: this
        // insert walked over it), so that a reader of a full bucket WITHOUT the flag can stop there
        if (((slot - home) & mask) >= 4u) atomicOr(&table[home + 3], kOverflowFlag);
    }
}

struct ChainParams {
    const uint32_t *pairs;  // (query, reference) genome ids
    const uint2 *kq;  // per seed: x = kmer << 1 | strand, y = spread position
    const uint32_t *cso;
    const unsigned long long *const *pair_table;  // per pair: the reference genome's hash table (local or peer-mapped)
    const uint32_t *pair_mask;                    // per pair: its slot count - 1
    const uint64_t *seed_off, *cso_off;
    const uint32_t *n_chunks;
    const uint32_t *unit_prefix;  // [n_pairs + 1] running sum of the query genomes' chunk counts
    uint32_t n_pairs, n_units;
    const unsigned long long *idtab;  // [kIdTabN * (kIdTabN + 1) / 2]: 2^40 (M/N)^(1/15) at N (N + 1) / 2 + M
    unsigned long long *acc_fx;       // per pair: sum of the per-chunk fixed-point identities
    uint32_t *acc;                    // kAccWords per pair: chunks, covq, covr, sum M, span M, span N, chains
    uint32_t *overflow;               // (pair, M, N) of chunks with N >= kIdTabN; overflow[0] = count
    uint32_t overflow_cap;
};

constexpr int kChainThreads = 128;
constexpr int kChainCtasPerSm = 4;  // 40 kB anchor ring + 16 kB bucket slots per CTA
constexpr int kRingFields = 5;  // q, r, f, rel<<31 | cnt<<29 | acc_dr<<15 | x, first_x<<16 | first_mb

// Four seeds (one aligned 32-byte sector of the interleaved (k-mer, position) array) in one load.
__device__ __forceinline__ void load_seeds4(const uint2 *p, unsigned long long (&k)[4]) {
    asm volatile("ld.global.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(k[0]), "=l"(k[1]), "=l"(k[2]), "=l"(k[3]) : "l"(p));
}

__global__ void __launch_bounds__(kChainThreads, kChainCtasPerSm) ani_chain_kernel(const ChainParams p) {
    extern __shared__ int ring[];  // [kAniH][kRingFields][kChainThreads]
    const uint32_t tid = threadIdx.x;
    // work units = (pair, query chunk), flattened over the batch so every thread has one
    const uint32_t u = blockIdx.x * kChainThreads + tid;
    // no early exits: every lane of a warp stays in the lock-step walk (idle lanes walk 0 seeds)
    uint32_t pair = 0, x0 = 0, x1 = 0;
    const uint2 *qkq = p.kq;
    uint64_t so = 0;
    const unsigned long long *table = nullptr;
    uint32_t mask = 0;
    if (u < p.n_units) {
        uint32_t lo = 0, hi = p.n_pairs;  // last pair with unit_prefix[pair] <= u
        while (hi - lo > 1) {
            const uint32_t mid = (lo + hi) >> 1;
            if (p.unit_prefix[mid] <= u) lo = mid; else hi = mid;
        }
        pair = lo;
        const uint32_t t = u - p.unit_prefix[pair];
        const uint32_t q = p.pairs[2 * pair];
        const uint32_t *cso = p.cso + p.cso_off[q];
        x0 = cso[t]; x1 = cso[t + 1];
        if (x1 - x0 < (uint32_t)kAniMinAnchors) x1 = x0;
        so = p.seed_off[q];
        qkq = p.kq + so;
        table = p.pair_table[pair];
        mask = p.pair_mask[pair];
    }
#define RING(slot, field) ring[((slot) * kRingFields + (field)) * kChainThreads + tid]
    uint32_t n_anchor = 0;
    // per-chunk result state (oracle/skani_oracle.c, "per chunk")
    uint32_t m_run = 0, first_q = 0xFFFFFFFFu, mb_first = 0, last_q = 0, m_at_last = 0;
    uint32_t covq = 0, covr = 0, span_m = 0, span_n = 0, n_chains = 0;
    int f_seen = 0;  // largest score of any anchor so far: bounds every score still in the ring
    // The lanes of a warp are different chunks and walk their seeds in lock step: one seed per
    // lane per step.  The probe of the reference table (the random read this kernel is bound
    // by) then runs converged across the warp; only lanes whose seed occurs 1..8 times in the
    // reference go on to the chaining step(s), which the f_seen bound keeps to one or two
    // look-backs for colinear anchors.
    // The probe sequence is read four slots (one 32-byte sector) at a time: linear probing keeps a
    // cluster contiguous, so one read usually holds the whole run up to its empty slot and the
    // chain of dependent reads -- max over the warp's lanes -- is one or two long instead of
    // the cluster length.
    // One seed: count its occurrences in the reference table (first bucket already in v), then chain.
    auto process = [&](const unsigned long long seed, unsigned long long (&v)[4], const uint32_t x) {
        const uint32_t ks = (uint32_t)seed;
        const int qpos = (int)(uint32_t)(seed >> 32);
        const uint32_t km = ks >> 1, qs = ks & 1;
        const uint32_t home = table_slot(km, mask);
        uint32_t c = 0;
        unsigned long long only = kEmpty;  // the match of a seed that occurs exactly once (the common case)
        {
            uint32_t g = home;
            bool open = true;
            for (;;) {
#pragma unroll
                for (uint32_t i = 0; i < 4; i++) {
                    if (open) {
                        if (v[i] == kEmpty) open = false;
                        else if (((uint32_t)(v[i] >> 33) & 0x3FFFFFFFu) == km) { only = v[i]; c++; }
                    }
                }
                if (!open) break;
                // a full home bucket: its keys go on beyond it only if it carries the overflow flag
                if (g == home && !(v[3] & kOverflowFlag)) break;
                g = (g + 4) & mask;
                load_bucket(table + g, v);
            }
        }
        const uint32_t occ = c > (uint32_t)kAniMaxOcc ? 0u : c;
        const uint32_t m_before = m_run;  // matched seeds of this chunk before x
        m_run += occ ? 1u : 0u;
        long long prev = -1;
        for (uint32_t m = 0; m < occ; m++) {
            // occurrence m: the smallest reference position above the previous one
            unsigned long long pick = only;
            if (occ > 1) {
                pick = kEmpty;
                for (uint32_t slot = home;; slot = (slot + 1) & mask) {
                    const unsigned long long e2 = table[slot];
                    if (e2 == kEmpty) break;
                    if (((uint32_t)(e2 >> 33) & 0x3FFFFFFFu) != km) continue;
                    const long long sp = (long long)(uint32_t)e2;
                    if (sp > prev && (pick == kEmpty || sp < (long long)(uint32_t)pick)) pick = e2;
                }
            }
            prev = (long long)(uint32_t)pick;
            const int rpos = (int)(uint32_t)pick;
            const uint32_t rel = qs ^ (uint32_t)((pick >> 32) & 1);
            const uint32_t xi = x - x0;
            int f = kAniAlpha;
            uint32_t pmeta = 0, pfirst = 0;  // best predecessor's words 3 and 4 (pmeta == 0: none)
            int link_dq = 0, link_dr = 0;
            const uint32_t look = min(n_anchor, (uint32_t)kAniH);
            for (uint32_t b = 1; b <= look; b++) {
                // a later predecessor can reach at most f_seen + alpha and must be strictly better
                if (f >= f_seen + kAniAlpha) break;
                const uint32_t slot = (n_anchor - b) % kAniH;
                const int dq = qpos - RING(slot, 0);
                if (dq > kAniBand) break;
                const uint32_t meta = (uint32_t)RING(slot, 3);
                if (dq <= 0 || (meta >> 31) != rel) continue;
                const int rb = RING(slot, 1);
                const int dr = rel ? rb - rpos : rpos - rb;
                if (dr <= 0 || dr > kAniBand) continue;
                const int gap = abs(dq - dr);
                if (gap > kAniMaxGap) continue;
                const int cand = RING(slot, 2) + kAniAlpha - gap;
                if (cand > f) {
                    f = cand; pmeta = meta; pfirst = (uint32_t)RING(slot, 4);
                    link_dq = dq; link_dr = dr;
                }
            }
            // cnt code: 1, 2 or 3 (= three or more); words 3 / 4 as in kRingFields' comment
            const uint32_t pcnt = (pmeta >> 29) & 3u;  // 0 when there is no predecessor
            const uint32_t cnt = min(pcnt + 1u, 3u);
            uint32_t first_x = xi, first_mb = m_before, acc_dr = 0;
            if (pcnt) { first_x = pfirst >> 16; first_mb = pfirst & 0xFFFFu; }
            if (pcnt == 1) acc_dr = (uint32_t)link_dr;
            const uint32_t slot = n_anchor % kAniH;
            RING(slot, 0) = qpos; RING(slot, 1) = rpos; RING(slot, 2) = f;
            RING(slot, 3) = (int)((rel << 31) | (cnt << 29) | (acc_dr << 15) | xi);
            RING(slot, 4) = (int)((first_x << 16) | first_mb);
            n_anchor++;
            f_seen = max(f_seen, f);
            if (cnt == 3) {
                last_q = xi; m_at_last = m_before + 1;
                if (pcnt == 2) {  // the chain's third anchor: it qualifies now
                    if (first_q == 0xFFFFFFFFu || first_x < first_q) { first_q = first_x; mb_first = first_mb; }
                    covq += (uint32_t)(qpos - (int)qkq[x0 + first_x].y) + kAniK;
                    covr += ((pmeta >> 15) & 0x3FFFu) + (uint32_t)link_dr + kAniK;
                    span_m += 3; span_n += xi - first_x + 1; n_chains++;
                } else {
                    covq += (uint32_t)link_dq; covr += (uint32_t)link_dr;
                    span_m += 1; span_n += xi - (pmeta & 0x7FFFu);
                }
            }
        }
    };
    // Software pipeline, four steps deep: while seed X is processed, the home buckets of seeds
    // X+1 .. X+4 are in flight as cp.async copies into a per-thread slot ring in shared memory, and the
    // seed groups (aligned groups of four seeds = one 32-byte sector, absolute seed indices) up to
    // twelve seeds ahead as register loads.  The kernel is bound by the latency of the bucket reads;
    // cp.async groups complete in order and are waited for by COUNT (wait_group 3 = "the copy issued
    // four steps ago has landed"), which register loads cannot express: with four LDG sites sharing
    // the warp's six scoreboards ptxas made every step wait for the copy issued one step earlier
    // (ncu: three of the four unrolled consumers stalled, the first one did not).
    uint4 *bk = reinterpret_cast<uint4 *>(ring + kAniH * kRingFields * kChainThreads);  // [4 slots][2 halves][threads]
    auto bucket_issue = [&](const int j, const unsigned long long seed) {
        const unsigned long long *src = table + table_slot((uint32_t)seed >> 1, mask);
        const uint32_t dst = (uint32_t)__cvta_generic_to_shared(bk + (j * 2) * kChainThreads + tid);
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src));
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + (uint32_t)(kChainThreads * 16)), "l"(src + 2));
    };
    const uint64_t X0 = so + x0, X1 = so + x1;
    const uint64_t XA = X0 & ~3ull;
    const uint32_t n_groups = x1 > x0 ? (uint32_t)((X1 - XA + 3) >> 2) : 0u;
    const uint32_t max_groups = __reduce_max_sync(0xffffffffu, n_groups);
    unsigned long long kA[4] = {0, 0, 0, 0}, kB[4] = {0, 0, 0, 0}, kC[4] = {0, 0, 0, 0};
    if (n_groups > 0) load_seeds4(p.kq + XA, kA);
    if (n_groups > 1) load_seeds4(p.kq + XA + 4, kB);
    if (n_groups > 2) load_seeds4(p.kq + XA + 8, kC);
#pragma unroll
    for (int j = 0; j < 4; j++) {
        const uint64_t X = XA + j;
        if (X >= X0 && X < X1) bucket_issue(j, kA[j]);
        asm volatile("cp.async.commit_group;");
    }
    // The trip count is the warp's longest chunk and every step starts with a warp barrier, so the
    // lanes re-join after the divergent chaining step whatever the compiler's own reconvergence
    // points are.
    for (uint32_t G = 0; G < max_groups; G++) {
#pragma unroll
        for (int j = 0; j < 4; j++) {
            __syncwarp();
            asm volatile("cp.async.wait_group 3;" ::: "memory");
            const uint64_t X = XA + 4ull * G + j;
            if (X >= X0 && X < X1) {
                const uint4 h0 = bk[(j * 2) * kChainThreads + tid], h1 = bk[(j * 2 + 1) * kChainThreads + tid];
                unsigned long long v[4] = {(unsigned long long)h0.y << 32 | h0.x, (unsigned long long)h0.w << 32 | h0.z,
                                           (unsigned long long)h1.y << 32 | h1.x, (unsigned long long)h1.w << 32 | h1.z};
                process(kA[j], v, (uint32_t)(X - so));
            }
            kA[j] = kB[j];
            if (X + 4 >= X0 && X + 4 < X1) bucket_issue(j, kA[j]);
            asm volatile("cp.async.commit_group;");
        }
#pragma unroll
        for (int j = 0; j < 4; j++) kB[j] = kC[j];
        if (G + 3 < n_groups) load_seeds4(p.kq + XA + 4ull * (G + 3), kC);
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
#undef RING
    if (first_q != 0xFFFFFFFFu) {
        const uint32_t M = m_at_last - mb_first, N = last_q - first_q + 1;
        if (N < (uint32_t)kIdTabN) {
            atomicAdd(&p.acc_fx[pair], p.idtab[(size_t)N * (N + 1) / 2 + M]);
        } else {
            const uint32_t at = atomicAdd(&p.overflow[0], 1u);
            if (at < p.overflow_cap) {
                p.overflow[1 + 3 * at] = pair; p.overflow[2 + 3 * at] = M; p.overflow[3 + 3 * at] = N;
            }
        }
        uint32_t *acc = p.acc + (size_t)kAccWords * pair;
        atomicAdd(&acc[0], 1u);
        atomicAdd(&acc[1], covq);
        atomicAdd(&acc[2], covr);
        atomicAdd(&acc[3], M);
        atomicAdd(&acc[4], span_m);
        atomicAdd(&acc[5], span_n);
        atomicAdd(&acc[6], n_chains);
    }
}

// ------------------------------------------------------------------------------------------
// host
// ------------------------------------------------------------------------------------------
// strtof(sprintf("%.2f", v)) without the text: v = M * 2^-k exactly (M < 2^53), so 100 v is the
// integer P = 100 M over 2^k and the two-decimal rounding printf does (nearest, ties to even, on
// the exact binary value) is integer arithmetic on P.  The printed number is then n / 100 with
// n an integer; (float)((double)n / 100.0) is the float nearest to it (the double quotient is
// within 2^-53 relative of n / 100, which is either a float itself or > 2^-27 relative away from
// every float midpoint, so rounding twice cannot differ from rounding once).  Values outside
// the range the ANI formula can produce take the text route.
float print2_parse_f32(double v) {
    if (v > 0.0 && v < 1.0e6) {
        int exp2 = 0;
        const double m = frexp(v, &exp2);  // v = m * 2^exp2, m in [0.5, 1)
        const uint64_t M = (uint64_t)ldexp(m, 53);
        const int k = 53 - exp2;  // v = M * 2^-k
        if (k >= 1 && k <= 62) {
            const unsigned __int128 P = (unsigned __int128)M * 100u;
            uint64_t n = (uint64_t)(P >> k);
            const unsigned __int128 r = P & (((unsigned __int128)1 << k) - 1), half = (unsigned __int128)1 << (k - 1);
            if (r > half || (r == half && (n & 1))) n++;
            return (float)((double)n / 100.0);
        }
    }
    char buf[64];
    snprintf(buf, sizeof(buf), "%.2f", v);
    return strtof(buf, nullptr);
}

uint64_t chunk_identity_fx(uint32_t m, uint32_t n) {
    if (m == 0 || n == 0) return 0;
    if (m >= n) return 1ull << kAniFxBits;
    return (uint64_t)llround(pow((double)m / (double)n, 1.0 / kAniK) * (double)(1ull << kAniFxBits));
}

AniPairResult ani_finish(const AniPairInts &v, uint64_t len_q, uint64_t len_r, uint32_t c, bool individual_contigs,
                         float min_af_pct) {
    AniPairResult res;
    memset(&res, 0, sizeof(res));
    double afq = len_q ? (double)v.cov_q / (double)len_q : 0.0, afr = len_r ? (double)v.cov_r / (double)len_r : 0.0;
    afq = std::min(afq, 1.0); afr = std::min(afr, 1.0);
    res.af_query = (float)afq; res.af_ref = (float)afr;
    res.sum_fx = v.sum_fx; res.n_chunks = v.n_chunks; res.sum_m = v.sum_m; res.span_m = v.span_m;
    res.span_n = v.span_n; res.n_chains = v.n_chains; res.cov_q = v.cov_q; res.cov_r = v.cov_r;
    if (v.n_chunks == 0 || v.sum_fx == 0) return res;
    // skani replaces its raw estimate by a learned regression "for c >= 70 and >= 150,000 bases
    // aligned and not on individual contigs"; its weights are not available, the chain-span
    // ratio (insensitive to unaligned stretches inside a chunk) stands in for it.
    const bool span = c >= kAniLearnedMinC && !individual_contigs && v.cov_q >= kAniLearnedMinBases &&
                      v.span_n > 2 * v.n_chains && v.span_m > 2 * v.n_chains;
    res.estimator = span ? 1u : 0u;
    const double ani = span ? 100.0 * pow((double)(v.span_m - 2 * v.n_chains) / (double)(v.span_n - 2 * v.n_chains), 1.0 / kAniK)
                            : 100.0 * ((double)v.sum_fx / ((double)v.n_chunks * (double)(1ull << kAniFxBits)));
    if (std::max(afq, afr) * 100.0 < (double)min_af_pct) return res;  // skani prints no row
    res.ani = print2_parse_f32(ani);  // skani prints {:.2}; galah parses the text as f32
    return res;
}


// Grow-only scratch buffer (cudaMalloc / cudaFree are device-wide synchronisations, so the
// per-batch temporaries are kept across calls).
template <typename T>
struct TmpBuf {
    T *p = nullptr;
    size_t cap = 0;
    ~TmpBuf() { if (p) cudaFree(p); }
    int alloc(size_t n) {
        n = std::max<size_t>(n, 1);
        if (n <= cap) return 0;
        if (p) GB_CUDA(cudaFree(p));
        p = nullptr; cap = 0;
        const size_t want = n + n / 4;
        GB_CUDA(cudaMalloc(&p, want * sizeof(T)));
        cap = want;
        return 0;
    }
    int upload(const std::vector<T> &v, cudaStream_t st) {
        if (alloc(v.size())) return 2;
        if (!v.empty()) GB_CUDA(cudaMemcpyAsync(p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice, st));
        return 0;
    }
};

// Grow-only pinned host buffer: device -> host copies into pageable memory are staged by the driver at
// ~2 GB/s (2.4 ms for the 4 MB of accumulators of 148k pairs); pinned, the same copy takes 0.2 ms.
template <typename T>
struct PinnedBuf {
    T *p = nullptr;
    size_t cap = 0;
    ~PinnedBuf() { if (p) cudaFreeHost(p); }
    int alloc(size_t n) {
        if (n <= cap) return 0;
        if (p) GB_CUDA(cudaFreeHost(p));
        p = nullptr; cap = 0;
        const size_t want = n + n / 4 + 16;
        GB_CUDA(cudaMallocHost(&p, want * sizeof(T)));
        cap = want;
        return 0;
    }
};

struct AniScratch {
    PinnedBuf<uint32_t> h_acc;
    PinnedBuf<unsigned long long> h_fx;
    TmpBuf<uint32_t> sel, count, contig_start, chunk_base, chunk_tmp, nch, pairs, acc, unit_prefix, overflow, mismatch;
    TmpBuf<uint64_t> contig_off, seed_off_b, cso_off_b, table_off_b;
    TmpBuf<unsigned long long> acc_fx, pair_table;
    TmpBuf<uint32_t> pair_mask;
};

// 2^40 (M/N)^(1/15) for every M <= N < kIdTabN, evaluated once per device on the host (glibc pow,
// the same expression as the oracle's) so that the chain kernel only adds integers.  4.2 MB,
// resident for the life of the process.
static int identity_table(const unsigned long long **out, cudaStream_t st) {
    static std::mutex mu;
    static unsigned long long *tabs[64] = {nullptr};
    int dev = 0;
    GB_CUDA(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64) { set_error("ani: device ordinal out of range"); return 3; }
    std::lock_guard<std::mutex> lock(mu);
    if (!tabs[dev]) {
        std::vector<unsigned long long> tab((size_t)kIdTabN * (kIdTabN + 1) / 2 + kIdTabN, 0);
        for (uint32_t N = 1; N < (uint32_t)kIdTabN; N++)
            for (uint32_t M = 0; M <= N; M++) tab[(size_t)N * (N + 1) / 2 + M] = chunk_identity_fx(M, N);
        unsigned long long *d = nullptr;
        GB_CUDA(cudaMalloc(&d, tab.size() * 8));
        GB_CUDA(cudaMemcpyAsync(d, tab.data(), tab.size() * 8, cudaMemcpyHostToDevice, st));
        GB_CUDA(cudaStreamSynchronize(st));
        tabs[dev] = d;
    }
    *out = tabs[dev];
    return 0;
}

int AniIndex::export_tables(cudaIpcMemHandle_t *handle, std::vector<uint64_t> &table_off,
                            std::vector<uint64_t> &total_len) const {
    if (!d_table_.p) { set_error("ani index: nothing to export"); return 3; }
    GB_CUDA(cudaIpcGetMemHandle(handle, d_table_.p));
    table_off = table_off_;
    total_len = total_len_;
    return 0;
}

int AniIndex::attach_peer(const cudaIpcMemHandle_t &handle, const uint64_t *table_off, const uint64_t *total_len,
                          size_t n, uint32_t *first_id) {
    // Opening a mapping costs hundreds of milliseconds for a multi-GB array (measured: 750 ms per step
    // on BASELINE configs[4]) and a peer re-uses its allocation from step to step: mappings are kept
    // open, keyed by the handle, until this index is destroyed.
    void *base = nullptr;
    for (const auto &c : ipc_cache_)
        if (memcmp(&c.first, &handle, sizeof(handle)) == 0) base = c.second;
    if (!base) {
        GB_CUDA(cudaIpcOpenMemHandle(&base, handle, cudaIpcMemLazyEnablePeerAccess));
        ipc_cache_.emplace_back(handle, base);
    }
    PeerGroup pg;
    pg.base = (const unsigned long long *)base;
    pg.table_off.assign(table_off, table_off + n + 1);
    peers_.push_back(std::move(pg));
    *first_id = (uint32_t)size() + peer_first_.back();
    peer_first_.push_back(peer_first_.back() + (uint32_t)n);
    peer_total_len_.insert(peer_total_len_.end(), total_len, total_len + n);
    return 0;
}

int AniIndex::attach_peer_direct(const unsigned long long *base, const uint64_t *table_off, const uint64_t *total_len,
                                 size_t n, uint32_t *first_id) {
    PeerGroup pg;
    pg.base = base;
    pg.table_off.assign(table_off, table_off + n + 1);
    pg.ipc = false;
    peers_.push_back(std::move(pg));
    *first_id = (uint32_t)size() + peer_first_.back();
    peer_first_.push_back(peer_first_.back() + (uint32_t)n);
    peer_total_len_.insert(peer_total_len_.end(), total_len, total_len + n);
    return 0;
}

void AniIndex::clear() {
    peers_.clear(); peer_first_.assign(1, 0); peer_total_len_.clear();
    seed_off_.assign(1, 0); cso_off_.assign(1, 0); table_off_.assign(1, 0);
    total_len_.clear(); n_chunks_.clear();
    d_kq_.n = d_cso_.n = d_table_.n = 0;
    d_seed_off_.n = d_cso_off_.n = d_table_off_.n = d_n_chunks_.n = 0;
}

AniIndex::~AniIndex() {
    for (auto &c : ipc_cache_) cudaIpcCloseMemHandle(c.second);
    d_kq_.release(); d_cso_.release(); d_table_.release();
    d_seed_off_.release(); d_cso_off_.release(); d_table_off_.release(); d_n_chunks_.release();
    for (int x = 0; x < 2; x++) if (ev_[x]) cudaEventDestroy(ev_[x]);
    delete scratch_;
}

int AniIndex::add_packed_device(const uint32_t *d_seq2, const uint32_t *d_valid, const uint64_t *d_base_off,
                                size_t n, const std::vector<uint64_t> &base_off, const std::vector<uint64_t> &contig_off,
                                const std::vector<uint32_t> &contig_start, const std::vector<uint32_t> &contig_len,
                                cudaStream_t st, const uint32_t *d_sel_in, const uint32_t *d_count_in) {
    if (n == 0) return 0;
    if (base_off.size() != n + 1 || contig_off.size() != n + 1) { set_error("ani index: bad offset arrays"); return 3; }
    for (size_t g = 0; g <= n; g++)
        if (base_off[g] % 128) { set_error("ani index: base_off must be multiples of 128"); return 3; }
    if (!ev_[0]) { GB_CUDA(cudaEventCreate(&ev_[0])); GB_CUDA(cudaEventCreate(&ev_[1])); }
    int dev = 0, sms = kNumSMsFallback;
    GB_CUDA(cudaGetDevice(&dev));
    GB_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    const uint64_t first = base_off[0], end = base_off[n];
    const uint64_t n_words = (end - first) / 32;

    // per-contig chunk bases, per-genome chunk counts and lengths (host)
    std::vector<uint32_t> chunk_base(contig_start.size(), 0), n_chunks(n, 0);
    std::vector<uint64_t> total_len(n, 0);
    for (size_t g = 0; g < n; g++) {
        uint32_t base = 0;
        if (base_off[g + 1] - base_off[g] >= (1ull << 31)) { set_error("ani index: genome longer than 2^31 bases"); return 3; }
        for (uint64_t c = contig_off[g]; c < contig_off[g + 1]; c++) {
            chunk_base[c] = base;
            base += (contig_len[c] + kAniChunk - 1) / kAniChunk;
            total_len[g] += contig_len[c];
        }
        n_chunks[g] = base;
    }

    GB_CUDA(cudaEventRecord(ev_[0], st));
    if (!scratch_) scratch_ = new AniScratch();
    TmpBuf<uint32_t> &d_sel = scratch_->sel, &d_count = scratch_->count, &d_contig_start = scratch_->contig_start,
                     &d_chunk_base = scratch_->chunk_base, &d_chunk_tmp = scratch_->chunk_tmp, &d_nch = scratch_->nch;
    if (scratch_->mismatch.alloc(1)) return 2;
    GB_CUDA(cudaMemsetAsync(scratch_->mismatch.p, 0, 4, st));
    TmpBuf<uint64_t> &d_contig_off = scratch_->contig_off, &d_seed_off_b = scratch_->seed_off_b,
                     &d_cso_off_b = scratch_->cso_off_b, &d_table_off_b = scratch_->table_off_b;
    if ((!d_sel_in && d_sel.alloc(n_words + 1)) || d_count.alloc(n)) return 2;
    const uint64_t thr = ~0ull / c_;
    const uint32_t *sel_p = d_sel_in ? d_sel_in : d_sel.p;
    {
        if (!d_sel_in) {
            const uint64_t blocks = std::min<uint64_t>((end - first + 255) / 256, (uint64_t)sms * 32);
            ani_mark_kernel<<<(uint32_t)std::max<uint64_t>(blocks, 1), 256, 0, st>>>(d_seq2, d_valid, first, end, thr, d_sel.p);
            GB_LAUNCH_CHECK();
        }
        if (!(d_sel_in && d_count_in)) {
            ani_count_kernel<<<(uint32_t)n, 256, 0, st>>>(sel_p, d_base_off, first, d_count.p);
            GB_LAUNCH_CHECK();
        }
    }
    std::vector<uint32_t> count(n);
    GB_CUDA(cudaMemcpyAsync(count.data(), d_sel_in && d_count_in ? d_count_in : d_count.p, n * 4, cudaMemcpyDeviceToHost, st));
    GB_CUDA(cudaStreamSynchronize(st));

    // absolute offsets of the new genomes in the index arrays
    const size_t g0 = size();
    std::vector<uint64_t> seed_off(n + 1), cso_off(n + 1), table_off(n + 1);
    seed_off[0] = seed_off_.back(); cso_off[0] = cso_off_.back(); table_off[0] = table_off_.back();
    for (size_t g = 0; g < n; g++) {
        seed_off[g + 1] = seed_off[g] + count[g];
        cso_off[g + 1] = cso_off[g] + n_chunks[g] + 1;
        uint64_t slots = 16;
        while (slots < 4ull * count[g]) slots <<= 1;  // load factor <= 1/4: about one key per 4-slot bucket
        table_off[g + 1] = table_off[g] + slots;
    }
    if (d_kq_.reserve(seed_off[n] + 8, st) ||
        d_cso_.reserve(cso_off[n] + 1, st) || d_table_.reserve(table_off[n] + 1, st) ||
        d_seed_off_.reserve(g0 + n + 2, st) || d_cso_off_.reserve(g0 + n + 2, st) ||
        d_table_off_.reserve(g0 + n + 2, st) || d_n_chunks_.reserve(g0 + n + 1, st))
        return 2;
    GB_CUDA(cudaMemsetAsync(d_table_.p + table_off[0], 0xFF, (table_off[n] - table_off[0]) * 8, st));
    if (d_contig_start.upload(contig_start, st) || d_chunk_base.upload(chunk_base, st) ||
        d_contig_off.upload(contig_off, st) || d_seed_off_b.upload(seed_off, st) ||
        d_cso_off_b.upload(cso_off, st) || d_table_off_b.upload(table_off, st) || d_nch.upload(n_chunks, st) ||
        d_chunk_tmp.alloc(seed_off[n] - seed_off[0] + 1))
        return 2;
    EmitParams e;
    e.seq2 = d_seq2; e.valid = d_valid; e.sel = sel_p; e.base_off = d_base_off; e.first_base = first;
    e.contig_off = d_contig_off.p; e.contig_start = d_contig_start.p; e.contig_chunk_base = d_chunk_base.p;
    e.seed_off = d_seed_off_b.p; e.cso_off = d_cso_off_b.p; e.n_chunks = d_nch.p; e.table_off = d_table_off_b.p;
    e.kq = d_kq_.p; e.cso = d_cso_.p;
    e.chunk_tmp = d_chunk_tmp.p - seed_off[0];  // indexed with absolute seed offsets
    e.table = d_table_.p;
    e.mismatch = scratch_->mismatch.p;
    ani_emit_kernel<<<(uint32_t)n, kEmitThreads, 0, st>>>(e);
    GB_LAUNCH_CHECK();
    // persistent per-genome metadata (device copies hold n+1 offsets: entry g0+n is the running end)
    GB_CUDA(cudaMemcpyAsync(d_seed_off_.p + g0, seed_off.data(), (n + 1) * 8, cudaMemcpyHostToDevice, st));
    GB_CUDA(cudaMemcpyAsync(d_cso_off_.p + g0, cso_off.data(), (n + 1) * 8, cudaMemcpyHostToDevice, st));
    GB_CUDA(cudaMemcpyAsync(d_table_off_.p + g0, table_off.data(), (n + 1) * 8, cudaMemcpyHostToDevice, st));
    GB_CUDA(cudaMemcpyAsync(d_n_chunks_.p + g0, n_chunks.data(), n * 4, cudaMemcpyHostToDevice, st));
    uint32_t mismatch = 0;
    GB_CUDA(cudaMemcpyAsync(&mismatch, scratch_->mismatch.p, 4, cudaMemcpyDeviceToHost, st));
    GB_CUDA(cudaEventRecord(ev_[1], st));
    GB_CUDA(cudaStreamSynchronize(st));  // host staging vectors die here
    if (mismatch) {
        set_error("ani index: seed count of genome " + std::to_string(mismatch - 1) + " of the batch disagrees with its selection bits");
        return 3;
    }
    float ms = 0.f;
    GB_CUDA(cudaEventElapsedTime(&ms, ev_[0], ev_[1]));
    last_build_ms = ms;
    d_kq_.n = seed_off[n]; d_cso_.n = cso_off[n]; d_table_.n = table_off[n];
    d_seed_off_.n = d_cso_off_.n = d_table_off_.n = g0 + n + 1; d_n_chunks_.n = g0 + n;
    for (size_t g = 0; g < n; g++) {
        seed_off_.push_back(seed_off[g + 1]); cso_off_.push_back(cso_off[g + 1]);
        table_off_.push_back(table_off[g + 1]); n_chunks_.push_back(n_chunks[g]);
        total_len_.push_back(total_len[g]);
    }
    return 0;
}

// Capacity hint: the index will hold n_total genomes like the ones already added.  Growing the
// device arrays batch by batch re-allocates and copies gigabytes; one reservation after the
// first batch avoids that.
int AniIndex::reserve_for(size_t n_total, cudaStream_t st) {
    const size_t have = size();
    if (have == 0 || n_total <= have) return 0;
    const double scale = 1.03 * (double)n_total / (double)have;
    if (d_kq_.reserve((size_t)(seed_off_.back() * scale) + 1024, st) ||
        d_cso_.reserve((size_t)(cso_off_.back() * scale) + 1024, st) ||
        d_table_.reserve((size_t)(table_off_.back() * scale) + 1024, st) ||
        d_seed_off_.reserve(n_total + 2, st) || d_cso_off_.reserve(n_total + 2, st) ||
        d_table_off_.reserve(n_total + 2, st) || d_n_chunks_.reserve(n_total + 1, st))
        return 2;
    return 0;
}

int AniIndex::genome_info(size_t g, uint64_t *n_seeds, uint32_t *n_chunks, uint64_t *total_len) const {
    if (g >= size()) { set_error("ani index: genome out of range"); return 3; }
    *n_seeds = seed_off_[g + 1] - seed_off_[g]; *n_chunks = n_chunks_[g]; *total_len = total_len_[g];
    return 0;
}

int AniIndex::genome_seeds(size_t g, uint32_t *ks, uint32_t *spread, uint32_t *chunk_of_seed, size_t cap,
                           cudaStream_t st) const {
    if (g >= size()) { set_error("ani index: genome out of range"); return 3; }
    const size_t ns = seed_off_[g + 1] - seed_off_[g];
    if (cap < ns) { set_error("ani index: seed buffer too small"); return 3; }
    std::vector<uint32_t> cso(n_chunks_[g] + 1);
    std::vector<uint2> kq(ns);
    GB_CUDA(cudaMemcpyAsync(kq.data(), d_kq_.p + seed_off_[g], ns * sizeof(uint2), cudaMemcpyDeviceToHost, st));
    GB_CUDA(cudaMemcpyAsync(cso.data(), d_cso_.p + cso_off_[g], cso.size() * 4, cudaMemcpyDeviceToHost, st));
    GB_CUDA(cudaStreamSynchronize(st));
    for (size_t x = 0; x < ns; x++) { ks[x] = kq[x].x; spread[x] = kq[x].y; }
    for (uint32_t t = 0; t < n_chunks_[g]; t++)
        for (uint32_t x = cso[t]; x < cso[t + 1]; x++) chunk_of_seed[x] = t;
    return 0;
}

int AniIndex::pairs(const uint32_t *pairs, size_t n_pairs, float min_af_pct, bool individual_contigs,
                    AniPairResult *out, cudaStream_t st) {
    if (n_pairs == 0) return 0;
    const bool dbg = getenv("GALAH_B200_DEBUG") != nullptr;
    auto now = [] { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    const double td0 = now();
    if (!ev_[0]) { GB_CUDA(cudaEventCreate(&ev_[0])); GB_CUDA(cudaEventCreate(&ev_[1])); }
    // the query is the pair's FIRST genome, as given (skani dist -q fasta1 -r fasta2, src/skani.rs:733-744)
    // reference ids >= size() name genomes of attached peers (attach_peer)
    const size_t n_local = size(), n_all = n_local + n_peer_genomes();
    std::vector<unsigned long long> h_ptab(n_pairs);
    std::vector<uint32_t> h_pmask(n_pairs);
    for (size_t x = 0; x < n_pairs; x++) {
        const uint32_t q = pairs[2 * x], r = pairs[2 * x + 1];
        if (q >= n_local || r >= n_all) { set_error("ani pairs: genome index out of range (the query must be local)"); return 3; }
        if (r < n_local) {
            h_ptab[x] = (unsigned long long)(uintptr_t)(d_table_.p + table_off_[r]);
            h_pmask[x] = (uint32_t)(table_off_[r + 1] - table_off_[r]) - 1;
        } else {
            const uint32_t pid = r - (uint32_t)n_local;
            const size_t g = std::upper_bound(peer_first_.begin(), peer_first_.end(), pid) - peer_first_.begin() - 1;
            const PeerGroup &pg = peers_[g];
            const uint32_t l = pid - peer_first_[g];
            h_ptab[x] = (unsigned long long)(uintptr_t)(pg.base + pg.table_off[l]);
            h_pmask[x] = (uint32_t)(pg.table_off[l + 1] - pg.table_off[l]) - 1;
        }
    }
    if (!scratch_) scratch_ = new AniScratch();
    if (scratch_->pair_table.upload(h_ptab, st) || scratch_->pair_mask.upload(h_pmask, st)) return 2;
    TmpBuf<uint32_t> &d_pairs = scratch_->pairs, &d_acc = scratch_->acc, &d_over = scratch_->overflow;
    TmpBuf<unsigned long long> &d_fx = scratch_->acc_fx;
    const uint32_t kOverflowCap = 1u << 16;
    if (d_pairs.alloc(2 * n_pairs) || d_acc.alloc((size_t)kAccWords * n_pairs) || d_fx.alloc(n_pairs) ||
        d_over.alloc(1 + 3 * (size_t)kOverflowCap))
        return 2;
    GB_CUDA(cudaMemcpyAsync(d_pairs.p, pairs, 8 * n_pairs, cudaMemcpyHostToDevice, st));
    const unsigned long long *d_idtab_p = nullptr;
    if (int rc = identity_table(&d_idtab_p, st)) return rc;
    GB_CUDA(cudaMemsetAsync(d_acc.p, 0, 4 * (size_t)kAccWords * n_pairs, st));
    GB_CUDA(cudaMemsetAsync(d_fx.p, 0, 8 * n_pairs, st));
    GB_CUDA(cudaMemsetAsync(d_over.p, 0, 4, st));
    const size_t smem = (size_t)kAniH * kRingFields * kChainThreads * sizeof(int) + (size_t)4 * 32 * kChainThreads;
    GB_CUDA(cudaFuncSetAttribute(ani_chain_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    GB_CUDA(cudaEventRecord(ev_[0], st));
    const double td1 = now();
    // work units of the whole call: (pair, query chunk), flattened so that every thread of the
    // grid owns one chunk whatever the genomes' chunk counts are
    std::vector<uint32_t> unit_prefix(n_pairs + 1, 0);
    const size_t kMaxUnits = 1u << 30;
    for (size_t b0 = 0; b0 < n_pairs;) {
        size_t b1 = b0;
        uint64_t units = 0;
        unit_prefix[b0] = 0;
        while (b1 < n_pairs && units + n_chunks_[pairs[2 * b1]] <= kMaxUnits) {
            units += n_chunks_[pairs[2 * b1]];
            unit_prefix[b1 + 1] = (uint32_t)units;
            b1++;
        }
        if (b1 == b0) { set_error("ani pairs: a genome has more than 2^30 chunks"); return 3; }
        if (units) {
            TmpBuf<uint32_t> &d_prefix = scratch_->unit_prefix;
            if (d_prefix.alloc(b1 - b0 + 1)) return 2;
            // (the stream is idle between batches only when a call needs more than one, i.e. > 2^30 chunks)
            if (b0) GB_CUDA(cudaStreamSynchronize(st));
            GB_CUDA(cudaMemcpyAsync(d_prefix.p, unit_prefix.data() + b0, (b1 - b0 + 1) * 4, cudaMemcpyHostToDevice, st));
            ChainParams p;
            p.pairs = d_pairs.p + 2 * b0; p.kq = d_kq_.p; p.cso = d_cso_.p;
            p.pair_table = reinterpret_cast<const unsigned long long *const *>(scratch_->pair_table.p) + b0;
            p.pair_mask = scratch_->pair_mask.p + b0;
            p.seed_off = d_seed_off_.p; p.cso_off = d_cso_off_.p;
            p.n_chunks = d_n_chunks_.p; p.acc = d_acc.p + (size_t)kAccWords * b0; p.acc_fx = d_fx.p + b0;
            p.idtab = d_idtab_p; p.overflow = d_over.p; p.overflow_cap = kOverflowCap;
            p.unit_prefix = d_prefix.p; p.n_pairs = (uint32_t)(b1 - b0); p.n_units = (uint32_t)units;
            ani_chain_kernel<<<(uint32_t)((units + kChainThreads - 1) / kChainThreads), kChainThreads, smem, st>>>(p);
            GB_LAUNCH_CHECK();
            // chunks too long for the table (N >= kIdTabN; never at c = 125 / 30 on 20 kb chunks of
            // ordinary sequence): finished on the host, batch-relative pair ids
            uint32_t n_over = 0;
            GB_CUDA(cudaMemcpyAsync(&n_over, d_over.p, 4, cudaMemcpyDeviceToHost, st));
            GB_CUDA(cudaStreamSynchronize(st));
            if (n_over > kOverflowCap) { set_error("ani pairs: too many over-long chunks"); return 3; }
            if (n_over) {
                std::vector<uint32_t> ov(3 * (size_t)n_over);
                GB_CUDA(cudaMemcpy(ov.data(), d_over.p + 1, 12 * (size_t)n_over, cudaMemcpyDeviceToHost));
                std::vector<unsigned long long> add(b1 - b0, 0);
                for (uint32_t x = 0; x < n_over; x++) add[ov[3 * x]] += chunk_identity_fx(ov[3 * x + 1], ov[3 * x + 2]);
                std::vector<unsigned long long> cur(b1 - b0);
                GB_CUDA(cudaMemcpy(cur.data(), d_fx.p + b0, 8 * (b1 - b0), cudaMemcpyDeviceToHost));
                for (size_t x = 0; x < b1 - b0; x++) cur[x] += add[x];
                GB_CUDA(cudaMemcpy(d_fx.p + b0, cur.data(), 8 * (b1 - b0), cudaMemcpyHostToDevice));
                GB_CUDA(cudaMemsetAsync(d_over.p, 0, 4, st));
            }
        }
        b0 = b1;
    }
    GB_CUDA(cudaEventRecord(ev_[1], st));
    const double td2 = now();
    if (scratch_->h_acc.alloc((size_t)kAccWords * n_pairs) || scratch_->h_fx.alloc(n_pairs)) return 2;
    const uint32_t *acc = scratch_->h_acc.p;
    const unsigned long long *fx = scratch_->h_fx.p;
    GB_CUDA(cudaMemcpyAsync(scratch_->h_acc.p, d_acc.p, 4 * (size_t)kAccWords * n_pairs, cudaMemcpyDeviceToHost, st));
    GB_CUDA(cudaMemcpyAsync(scratch_->h_fx.p, d_fx.p, 8 * n_pairs, cudaMemcpyDeviceToHost, st));
    GB_CUDA(cudaStreamSynchronize(st));
    float ms = 0.f;
    GB_CUDA(cudaEventElapsedTime(&ms, ev_[0], ev_[1]));
    last_chain_ms = ms;
    const double td3 = now();
    // host finish (a division or pow + the exact two-decimal rounding per pair): split over the host
    // threads when the call is large, so that it does not become the serial tail of the stage
    auto finish_range = [&](size_t x0, size_t x1) {
        for (size_t x = x0; x < x1; x++) {
            const uint32_t *a = &acc[(size_t)kAccWords * x];
            AniPairInts v;
            v.sum_fx = fx[x]; v.n_chunks = a[0]; v.cov_q = a[1]; v.cov_r = a[2]; v.sum_m = a[3];
            v.span_m = a[4]; v.span_n = a[5]; v.n_chains = a[6];
            const uint32_t r = pairs[2 * x + 1];
            const uint64_t len_r = r < n_local ? total_len_[r] : peer_total_len_[r - n_local];
            out[x] = ani_finish(v, total_len_[pairs[2 * x]], len_r, c_, individual_contigs, min_af_pct);
        }
    };
    const size_t hw = std::max<size_t>(1, std::thread::hardware_concurrency());
    const size_t nt = n_pairs >= 8192 ? std::min<size_t>(hw, 32) : 1;
    if (nt == 1) {
        finish_range(0, n_pairs);
    } else {
        std::vector<std::thread> th;
        const size_t per = (n_pairs + nt - 1) / nt;
        for (size_t t = 0; t < nt; t++) {
            const size_t x0 = t * per, x1 = std::min(n_pairs, x0 + per);
            if (x0 < x1) th.emplace_back(finish_range, x0, x1);
        }
        for (auto &t : th) t.join();
    }
    if (dbg)
        fprintf(stderr, "[ani pairs] %zu pairs: prepare + upload %.2f ms, launch loop (incl. kernel wait) %.2f ms, D2H %.2f ms, "
                "host finish %.2f ms (kernel %.2f ms)\n", n_pairs, td1 - td0, td2 - td1, td3 - td2, now() - td3, ms);
    return 0;
}

template struct DevVec<uint32_t>;
template struct DevVec<uint2>;
template struct DevVec<uint64_t>;
template struct DevVec<unsigned long long>;

}  // namespace gb200
