// Stage 1b (K2), mode 0: all-pairs sketch intersection as a BLOCK-LIST JOIN on sm_100a.
//
// Replaces the serial loop at /root/reference/src/finch.rs:75-95 (for every i<j:
// finch::distance::distance(s_i, s_j, false) -> raw_distance -> common, total).  The reference
// walks 2s merge steps per pair.  Here the table is cut into blocks of R = kShardRows = 128
// consecutive sketches and each block is merged ONCE into a single ascending list of
// (value, row-tag) entries ("block list", R*s entries).  A work item (rb, cb) then merge-
// intersects block list rb with block list cb: every equal (a, b) adds 1 to
// cnt[tag(a)][tag(b)], an R x R matrix of 16-bit counters in shared memory, so ONE merge of
// 2*R*s entries yields the exact |A n B| of R*R pairs -- 2s/R merge steps and 2*s*4/R bytes per
// pair.  `total` follows from one rank query per surviving pair (see prefilter.cu, "Exactness
// notes").
//
// Kernels
//   bl_len_kernel     valid entries per block (sum of counts) and the table-wide largest hash.
//   bl_seed_kernel    one CTA takes 8 table rows, forms their level-0 entries (value, tag = row % R;
//                     padding = (2^64-1, 0xFF)) in shared memory and runs the first three merge
//                     levels there (stride <= 1024; bl_init_kernel + global levels otherwise).
//   bl_partition_kernel / bl_merge_tma_kernel
//                     one level of the merge tree (runs of m entries -> runs of 2m), ping-ponging
//                     two buffers.  The merge-path splits of all tile boundaries come from a
//                     pre-pass (one warp per boundary, 32-ary search); a merge CTA owns one
//                     2048-entry output tile: its A and B slices arrive by TMA bulk copies, every
//                     thread merges 8 entries, the tile leaves by TMA bulk stores
//                     (bl_merge_kernel is the same level with staged loads / stores, for run
//                     lengths that are not multiples of 16).  Keys are ordered by (value, tag), so
//                     padding sorts strictly after a genuine 2^64-1 hash and the first bl_len
//                     entries of a block list are exactly its valid entries.  The LAST level
//                     writes the lists as structure-of-arrays: with sh = clz(largest valid hash
//                     of the table), w = value << sh is lossless, hi = w >> 32 is an
//                     order-preserving 32-bit key and (hi, lo) equality is value equality.
//   prefilter_join_kernel  persistent CTAs (3 per SM) pull items from an atomic counter, longest
//                     first (diagonal, adjacent blocks, the rest; or an explicit list).  The two
//                     lists are cut into segments of D = 9216 merged keys by merge-path splits
//                     (binary searches through L2, all segments of the item in parallel); a
//                     segment's A and B key slices are brought in by two 1-D TMA bulk copies
//                     (cp.async.bulk + mbarrier complete_tx, SASS UBLKCP); every thread then owns
//                     two chains of 18 consecutive merge steps, each starting at its own
//                     merge-path split in shared memory.  The merge runs on the dense 32-bit hi
//                     keys only (one LDS.32 per step, branch-free, fully unrolled when the
//                     segment touches no list end).  Ties are ordered B-first, so when a thread
//                     takes b every a with the same key lies at or after its A cursor; tie steps
//                     are recorded in a bit mask, queued, and resolved by the whole CTA at once
//                     (lo words and tags from global memory / L2).  Diagonal items (rb == cb)
//                     need no merge: equal values are adjacent in the one list (hi, lo and tags
//                     staged; lock-step run walks).  After the last segment the CTA scans cnt,
//                     applies the conservative integer thresholds and appends survivors
//                     {i, j, common, total}.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <string>

#include "common.cuh"
#include "prefilter.cuh"
#include "prefilter_dev.cuh"

namespace gb200 {

constexpr int kJR = kShardRows;            // sketches per block list
constexpr int kJLevels = 7;                // log2(kJR)
static_assert((1 << kJLevels) == kJR, "kShardRows must be a power of two");
constexpr int kJThreads = 256;
constexpr int kJE = 18;                    // merge steps per thread per segment (kJE/2 odd: the
                                           // expected A/B cursors of adjacent threads then fall
                                           // in different shared-memory bank pairs)
constexpr int kJChains = 2;                // independent merge chains per thread (ILP: each step
                                           // is a dependent LDS -> compare -> select -> LDS)
constexpr int kJD = kJThreads * kJE * kJChains;  // merged entries per segment
constexpr int kJCap = kJD + 16 + 80;       // staged entries incl. short-run overflow + alignment slack
constexpr int kJDiag = 3840;               // entries per diagonal-item segment (hi, lo, tags staged)
static_assert(2 * (kJDiag + kJR + 16) * 4 + (kJDiag + kJR + 16) <= kJCap * 4, "diagonal staging must fit");
constexpr int kJCtasPerSm = 3;
constexpr int kDR = 5;                     // diagonal items: entries per thread whose run walks are interleaved
static_assert(kJDiag % (kDR * kJThreads) == 0, "a full diagonal segment is whole passes");
constexpr int kTQ = 1024;                  // deferred-tie queue entries per CTA
constexpr uint32_t kDenseTies = 64;        // ties in an item after which its segments track B-takes in the hot loop
constexpr int kTQSegs = 16;                // segments whose (i0, j0) bases the queue can refer to
constexpr uint8_t kPadTag = 0xFF;
constexpr size_t kBlSlack = 1024;          // entries of over-read slack behind the last list

constexpr int kMThreads = 256;
constexpr int kME = 8;
constexpr int kMTile = kMThreads * kME;

// ------------------------------------------------------------------------------------------
// block-list build
// ------------------------------------------------------------------------------------------
// Level-0 lists of the blocks [row0 / R, ...): `total` entries starting at table row row0.
__global__ void __launch_bounds__(256) bl_init_kernel(const uint64_t *__restrict__ hashes,
                                                      const uint32_t *__restrict__ counts, uint32_t n,
                                                      uint32_t stride, uint64_t row0, uint64_t total,
                                                      uint64_t *__restrict__ vals,
                                                      uint8_t *__restrict__ tags) {
    const uint64_t step = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t e = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += step) {
        const uint64_t row = row0 + e / stride;
        const uint32_t col = (uint32_t)(e % stride);
        const bool valid = row < n && col < min(counts[row], stride);
        vals[e] = valid ? hashes[row * stride + col] : kPad;
        tags[e] = valid ? (uint8_t)(row % kJR) : kPadTag;
    }
}

// One warp per block of the range [pb0, pb1) (the WHOLE table unless the caller already knows
// gmax): largest valid hash -> gmax (every rank needs the same shift), and for the blocks
// [b0, b1) being built the number of valid entries -> bl_len[b - b0].
__global__ void __launch_bounds__(256) bl_len_kernel(const uint64_t *__restrict__ hashes,
                                                     const uint32_t *__restrict__ counts, uint32_t n,
                                                     uint32_t stride, uint32_t pb0, uint32_t pb1, uint32_t b0,
                                                     uint32_t b1, uint32_t *__restrict__ bl_len,
                                                     unsigned long long *__restrict__ gmax) {
    const uint32_t b = pb0 + ((blockIdx.x * blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
    if (b >= pb1) return;
    uint32_t sum = 0;
    unsigned long long mx = 0;
    for (uint32_t r = b * kJR + lane; r < min(n, (b + 1) * kJR); r += 32) {
        const uint32_t c = min(counts[r], stride);
        sum += c;
        if (c) mx = max(mx, (unsigned long long)hashes[(size_t)r * stride + c - 1]);  // rows ascend
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        sum += __shfl_xor_sync(0xffffffffu, sum, o);
        mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    }
    if (lane == 0) {
        if (b >= b0 && b < b1) bl_len[b - b0] = sum;
        if (mx && gmax) atomicMax(gmax, mx);
    }
}

// Seed of the merge tree: one CTA takes kSeedRows consecutive table rows, forms their level-0
// entries (value, tag = row % R; padding = (2^64-1, 0xFF)) in shared memory and runs the first
// log2(kSeedRows) merge levels there, so those levels cost one pass over HBM instead of one
// each.  Every level: thread t owns kSeedE consecutive outputs of one run pair (threads never
// span pairs: a pair holds 2m = kSeedE * threads-per-pair entries or fewer), finds its start by a
// merge-path split, merges into registers; after a barrier the registers replace the inputs.
// Values live at index i + (i >> 5): cursors of neighbouring threads are about kSeedE / 2 = 16
// entries apart, which the skew spreads over all banks.  Needs stride <= kSeedE * 256 / kSeedRows.
constexpr int kSeedRows = 8;
constexpr int kSeedThreads = 256;
constexpr int kSeedE = 32;
constexpr uint32_t kSeedMaxStride = kSeedE * kSeedThreads / kSeedRows;  // 1024
__host__ __device__ constexpr uint32_t seed_skew(uint32_t i) { return i + (i >> 5); }
static size_t seed_smem_bytes(uint32_t stride) {
    const uint32_t ne = kSeedRows * stride;
    return (size_t)(seed_skew(ne) + 1) * 8 + ((ne + 15u) & ~15u);
}

__global__ void __launch_bounds__(kSeedThreads) bl_seed_kernel(const uint64_t *__restrict__ hashes,
                                                               const uint32_t *__restrict__ counts, uint32_t n,
                                                               uint32_t stride, uint64_t row0,
                                                               uint64_t *__restrict__ vals,
                                                               uint8_t *__restrict__ tags) {
    extern __shared__ __align__(16) uint8_t seed_smem_raw[];
    const uint32_t ne = kSeedRows * stride;
    uint64_t *V = reinterpret_cast<uint64_t *>(seed_smem_raw);
    uint8_t *T = seed_smem_raw + (size_t)(seed_skew(ne) + 1) * 8;
    const uint32_t tid = threadIdx.x;
    const uint64_t first_row = row0 + (uint64_t)blockIdx.x * kSeedRows;
    {   // all kSeedRows loads of a column step are independent: they are in flight together
        uint32_t cnt[kSeedRows];
#pragma unroll
        for (int r = 0; r < kSeedRows; r++) cnt[r] = first_row + r < n ? min(counts[first_row + r], stride) : 0u;
        for (uint32_t col = tid; col < stride; col += kSeedThreads) {
            uint64_t v[kSeedRows];
#pragma unroll
            for (int r = 0; r < kSeedRows; r++) v[r] = col < cnt[r] ? hashes[(first_row + r) * stride + col] : kPad;
#pragma unroll
            for (int r = 0; r < kSeedRows; r++) {
                V[seed_skew(r * stride + col)] = v[r];
                T[r * stride + col] = col < cnt[r] ? (uint8_t)((first_row + r) % kJR) : kPadTag;
            }
        }
    }
    __syncthreads();
    uint32_t tpp = kSeedThreads / (kSeedRows / 2);  // threads per run pair
    for (uint32_t m = stride; m < ne; m *= 2, tpp *= 2) {
        const uint32_t pair = tid / tpp, lt = tid % tpp;
        const uint32_t base = pair * 2 * m;
        const uint32_t d0 = min(lt * kSeedE, 2 * m), d1 = min(d0 + (uint32_t)kSeedE, 2 * m);
        // split: smallest i in [max(0, d0 - m), min(d0, m)] with !(A[i] <= B[d0 - 1 - i]) under (value, tag)
        uint32_t lo = d0 > m ? d0 - m : 0, hi = min(d0, m);
        while (lo < hi) {
            const uint32_t mid = (lo + hi) >> 1;
            const uint32_t ia = base + mid, ib = base + m + d0 - 1 - mid;
            const uint64_t a = V[seed_skew(ia)], b = V[seed_skew(ib)];
            const bool le = a < b || (a == b && T[ia] <= T[ib]);
            if (le) lo = mid + 1; else hi = mid;
        }
        uint32_t i = lo, j = d0 - lo;  // cursors inside the pair's two runs
        uint64_t ov[kSeedE];
        uint32_t ot[kSeedE / 4];
#pragma unroll
        for (int w = 0; w < kSeedE / 4; w++) ot[w] = 0;
        uint64_t a = V[seed_skew(base + min(i, m - 1))], b = V[seed_skew(base + m + min(j, m - 1))];
        uint32_t ta = T[base + min(i, m - 1)], tb = T[base + m + min(j, m - 1)];
#pragma unroll
        for (int x = 0; x < kSeedE; x++) {
            bool take_a;
            if (j >= m) take_a = true;
            else if (i >= m) take_a = false;
            else take_a = a < b || (a == b && ta <= tb);
            ov[x] = take_a ? a : b;
            ot[x >> 2] |= (take_a ? ta : tb) << (8 * (x & 3));
            if (take_a) { i++; const uint32_t q = base + min(i, m - 1); a = V[seed_skew(q)]; ta = T[q]; }
            else { j++; const uint32_t q = base + m + min(j, m - 1); b = V[seed_skew(q)]; tb = T[q]; }
        }
        __syncthreads();
#pragma unroll
        for (int x = 0; x < kSeedE; x++)
            if (d0 + x < d1) {
                V[seed_skew(base + d0 + x)] = ov[x];
                T[base + d0 + x] = (uint8_t)(ot[x >> 2] >> (8 * (x & 3)));
            }
        __syncthreads();
    }
    const uint64_t out0 = (uint64_t)blockIdx.x * ne;
    for (uint32_t e = tid; e < ne; e += kSeedThreads) {
        vals[out0 + e] = V[seed_skew(e)];
        tags[out0 + e] = T[e];
    }
}

__device__ __forceinline__ bool key_le(uint64_t a, uint8_t ta, uint64_t b, uint8_t tb) {
    return a < b || (a == b && ta <= tb);
}

// One level of the merge tree: for every output run o, dst[o*2m .. o*2m+2m) = merge of
// src[o*2m .. +m) and src[o*2m+m .. +2m) under the (value, tag) order, first run first on ties.
// Merge-path splits of every tile boundary of one merge level in a pre-pass (so the dependent L2
// probes of a split overlap across all boundaries instead of serialising at the head of every
// merge CTA).  splits[o * (chunks + 1) + c] = A-side split of diagonal c*kMTile.
__global__ void __launch_bounds__(256) bl_partition_kernel(const uint64_t *__restrict__ sv,
                                                           const uint8_t *__restrict__ st, uint32_t m,
                                                           uint32_t chunks_per_run, uint64_t n_bounds,
                                                           uint32_t *__restrict__ splits) {
    // One WARP per boundary, 32-ary search: every round the 32 lanes probe 32 evenly spaced
    // candidate splits at once and a ballot narrows the range 33-fold, so a boundary costs
    // log_33(m) ~ 3-4 dependent L2 round trips instead of log_2(m) ~ 17.
    const uint64_t b = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint32_t lane = threadIdx.x & 31;
    if (b >= n_bounds) return;
    const uint64_t o = b / (chunks_per_run + 1);
    const uint32_t c = (uint32_t)(b % (chunks_per_run + 1));
    const uint64_t base = o * 2 * m;
    const uint64_t *A = sv + base, *B = sv + base + m;
    const uint8_t *tA = st + base, *tB = st + base + m;
    const uint32_t d = min(c * (uint32_t)kMTile, 2 * m);
    uint32_t lo = d > m ? d - m : 0, hi = min(d, m);
    // invariant: key_le holds for every index < lo and fails for every index >= hi (the
    // predicate is monotone along the diagonal); answer = smallest i with !(A[i] <= B[d-1-i])
    while (lo < hi) {
        const uint32_t span = hi - lo;
        // probe positions lo + (lane + 1) * span / 33, clamped into [lo, hi - 1]
        const uint32_t mid = min(hi - 1, lo + (uint32_t)(((uint64_t)(lane + 1) * span) / 33));
        const bool le = key_le(A[mid], tA[mid], B[d - 1 - mid], tB[d - 1 - mid]);
        const uint32_t ok = __ballot_sync(0xffffffffu, le);  // monotone: a prefix of the lanes
        const uint32_t n_le = __popc(ok);
        // lanes 0 .. n_le-1 hold: the answer lies after lane n_le-1's probe and at or before lane n_le's
        const uint32_t new_lo = n_le ? min(hi - 1, lo + (uint32_t)(((uint64_t)n_le * span) / 33)) + 1 : lo;
        const uint32_t new_hi = n_le < 32 ? min(hi - 1, lo + (uint32_t)(((uint64_t)(n_le + 1) * span) / 33)) : hi;
        lo = new_lo; hi = new_hi;
    }
    if (lane == 0) splits[b] = lo;
}

struct MergeSmem {
    uint64_t in_v[kMTile];
    uint64_t out_v[kMTile];
    uint8_t in_t[kMTile];
    uint8_t out_t[kMTile];
    uint32_t ts[kMThreads + 1];
};

template <bool kLast>
__global__ void __launch_bounds__(kMThreads) bl_merge_kernel(const uint64_t *__restrict__ sv,
                                                             const uint8_t *__restrict__ st,
                                                             uint64_t *__restrict__ dv,
                                                             uint8_t *__restrict__ dt, uint32_t m,
                                                             uint32_t chunks_per_run,
                                                             const uint32_t *__restrict__ splits,
                                                             const unsigned long long *__restrict__ gmax,
                                                             uint32_t *__restrict__ dhi,
                                                             uint32_t *__restrict__ dlo) {
    extern __shared__ __align__(16) uint8_t merge_smem_raw[];
    MergeSmem &S = *reinterpret_cast<MergeSmem *>(merge_smem_raw);
    const uint32_t tid = threadIdx.x;
    const uint32_t o = blockIdx.x / chunks_per_run, c = blockIdx.x % chunks_per_run;
    const uint64_t base = (uint64_t)o * 2 * m;
    const uint64_t *A = sv + base, *B = sv + base + m;
    const uint8_t *tA = st + base, *tB = st + base + m;
    const uint32_t d0 = c * kMTile, d1 = min(d0 + (uint32_t)kMTile, 2 * m);
    const uint64_t sb = (uint64_t)o * (chunks_per_run + 1) + c;
    const uint32_t i0 = splits[sb], i1 = splits[sb + 1], j0 = d0 - i0, j1 = d1 - i1;
    const uint32_t na = i1 - i0, nb = j1 - j0, len = na + nb;
    for (uint32_t x = tid; x < na; x += kMThreads) { S.in_v[x] = A[i0 + x]; S.in_t[x] = tA[i0 + x]; }
    for (uint32_t x = tid; x < nb; x += kMThreads) { S.in_v[na + x] = B[j0 + x]; S.in_t[na + x] = tB[j0 + x]; }
    __syncthreads();
    const uint64_t *As = S.in_v, *Bs = S.in_v + na;
    const uint8_t *At = S.in_t, *Bt = S.in_t + na;
    const uint32_t dt0 = min(tid * kME, len), dt1 = min(dt0 + (uint32_t)kME, len);
    {
        uint32_t lo = dt0 > nb ? dt0 - nb : 0, hi = min(dt0, na);
        while (lo < hi) {
            const uint32_t mid = (lo + hi) >> 1;
            const uint64_t a = As[mid], b = Bs[dt0 - 1 - mid];
            const bool le = a < b || (a == b && At[mid] <= Bt[dt0 - 1 - mid]);  // tags only matter on a tie
            if (le) lo = mid + 1; else hi = mid;
        }
        S.ts[tid] = lo;
        if (tid == 0) S.ts[kMThreads] = na;
    }
    __syncthreads();
    uint32_t i = S.ts[tid], j = dt0 - i;
    const uint32_t ie = S.ts[tid + 1], je = dt1 - ie;
    // the values under the two cursors live in registers: one shared-memory load per step
    uint64_t a = As[min(i, kMTile - 1u)], b = Bs[min(j, kMTile - 1u - na)];
    for (uint32_t x = dt0; x < dt1; x++) {
        bool take_a;
        if (j >= je) take_a = true;
        else if (i >= ie) take_a = false;
        else take_a = a < b || (a == b && At[i] <= Bt[j]);
        if (take_a) { S.out_v[x] = a; S.out_t[x] = At[i]; i++; a = As[min(i, kMTile - 1u)]; }
        else { S.out_v[x] = b; S.out_t[x] = Bt[j]; j++; b = Bs[min(j, kMTile - 1u - na)]; }
    }
    __syncthreads();
    // coalesced write-out of the merged tile
    int sh = 0;
    if (kLast) { const unsigned long long g = *gmax; sh = g ? __clzll((long long)g) : 0; }
    for (uint32_t x = tid; x < len; x += kMThreads) {
        const uint64_t v = S.out_v[x];
        const uint8_t tg = S.out_t[x];
        dt[base + d0 + x] = tg;
        if (kLast) {
            const uint64_t w = v << sh;
            dhi[base + d0 + x] = tg == kPadTag ? 0xFFFFFFFFu : (uint32_t)(w >> 32);
            dlo[base + d0 + x] = (uint32_t)w;
        } else {
            dv[base + d0 + x] = v;
        }
    }
}

// The same merge level with the tile moved by the TMA engine in both directions: the A and B
// slices (values and tags) arrive by four 1-D bulk copies, the merged tile leaves by two (three
// at the last level: hi, lo, tags), so the LSU only carries the merge itself -- the staged
// variant above spends about half of its shared-memory wavefronts on copying.  Needs every
// run start and tile length to be a multiple of 16 entries (m % 16 == 0: every level above the seed).
struct MergeTmaSmem {
    uint64_t in_v[kMTile + 4];  // A window (from an even index), then B's at the next even slot
    uint64_t out_v[kMTile];     // last level: out_hi = first half, out_lo = second half (uint32 each)
    uint8_t in_t[kMTile + 64];  // A tag window (from a multiple of 16), then B's
    uint8_t out_t[kMTile];
    uint32_t ts[kMThreads + 1];
    uint64_t bar;
};

template <bool kLast>
__global__ void __launch_bounds__(kMThreads) bl_merge_tma_kernel(const uint64_t *__restrict__ sv,
                                                                 const uint8_t *__restrict__ st,
                                                                 uint64_t *__restrict__ dv,
                                                                 uint8_t *__restrict__ dt, uint32_t m,
                                                                 uint32_t chunks_per_run,
                                                                 const uint32_t *__restrict__ splits,
                                                                 const unsigned long long *__restrict__ gmax,
                                                                 uint32_t *__restrict__ dhi,
                                                                 uint32_t *__restrict__ dlo) {
    extern __shared__ __align__(128) uint8_t merge_tma_smem_raw[];
    MergeTmaSmem &S = *reinterpret_cast<MergeTmaSmem *>(merge_tma_smem_raw);
    const uint32_t tid = threadIdx.x;
    const uint32_t o = blockIdx.x / chunks_per_run, c = blockIdx.x % chunks_per_run;
    const uint64_t base = (uint64_t)o * 2 * m;
    const uint32_t d0 = c * kMTile, d1 = min(d0 + (uint32_t)kMTile, 2 * m);
    const uint64_t sb = (uint64_t)o * (chunks_per_run + 1) + c;
    const uint32_t i0 = splits[sb], i1 = splits[sb + 1], j0 = d0 - i0, j1 = d1 - i1;
    const uint32_t na = i1 - i0, nb = j1 - j0, len = na + nb;
    // aligned source windows (readable slack behind the level buffers covers the round-up)
    const uint32_t av0 = i0 & ~1u, bv0 = j0 & ~1u, at0 = i0 & ~15u, bt0 = j0 & ~15u;
    const uint32_t avn = (i1 - av0 + 1u) & ~1u, bvn = (j1 - bv0 + 1u) & ~1u;
    const uint32_t atn = (i1 - at0 + 15u) & ~15u, btn = (j1 - bt0 + 15u) & ~15u;
    if (tid == 0) {
        mbar_init(&S.bar, 1);
        fence_mbar_init();
        mbar_arrive_expect_tx(&S.bar, (avn + bvn) * 8u + atn + btn);
        if (avn) tma_load_1d(S.in_v, sv + base + av0, avn * 8u, &S.bar);
        if (bvn) tma_load_1d(S.in_v + avn, sv + base + m + bv0, bvn * 8u, &S.bar);
        if (atn) tma_load_1d(S.in_t, st + base + at0, atn, &S.bar);
        if (btn) tma_load_1d(S.in_t + atn, st + base + m + bt0, btn, &S.bar);
    }
    __syncthreads();  // the barrier is initialised before anyone waits on it
    mbar_wait(&S.bar, 0);
    const uint64_t *As = S.in_v + (i0 - av0), *Bs = S.in_v + avn + (j0 - bv0);
    const uint8_t *At = S.in_t + (i0 - at0), *Bt = S.in_t + atn + (j0 - bt0);
    const uint32_t dt0 = min(tid * kME, len), dt1 = min(dt0 + (uint32_t)kME, len);
    // own start split; the end split is the next thread's start (shared through S.ts)
    {
        uint32_t lo = dt0 > nb ? dt0 - nb : 0, hi = min(dt0, na);
        while (lo < hi) {
            const uint32_t mid = (lo + hi) >> 1;
            const uint64_t a = As[mid], b = Bs[dt0 - 1 - mid];
            const bool le = a < b || (a == b && At[mid] <= Bt[dt0 - 1 - mid]);  // tags only matter on a tie
            if (le) lo = mid + 1; else hi = mid;
        }
        S.ts[tid] = lo;
        if (tid == 0) S.ts[kMThreads] = na;
    }
    __syncthreads();
    uint32_t i = S.ts[tid], j = dt0 - i;
    const uint32_t ie = S.ts[tid + 1], je = dt1 - ie;
    int sh = 0;
    if (kLast) { const unsigned long long g = *gmax; sh = g ? __clzll((long long)g) : 0; }
    uint32_t *out_hi = reinterpret_cast<uint32_t *>(S.out_v), *out_lo = out_hi + kMTile;
    // the values under the two cursors live in registers: one shared-memory load per step
    uint64_t a = As[min(i, na)], b = Bs[min(j, nb)];  // index na / nb is inside the windows' slack
    for (uint32_t x = dt0; x < dt1; x++) {
        bool take_a;
        if (j >= je) take_a = true;
        else if (i >= ie) take_a = false;
        else take_a = a < b || (a == b && At[i] <= Bt[j]);
        uint64_t v; uint8_t tg;
        if (take_a) { v = a; tg = At[i]; i++; a = As[min(i, na)]; }
        else { v = b; tg = Bt[j]; j++; b = Bs[min(j, nb)]; }
        S.out_t[x] = tg;
        if (kLast) {
            const uint64_t w = v << sh;
            out_hi[x] = tg == kPadTag ? 0xFFFFFFFFu : (uint32_t)(w >> 32);
            out_lo[x] = (uint32_t)w;
        } else {
            S.out_v[x] = v;
        }
    }
    fence_proxy_async();  // generic-proxy writes of the tile -> visible to the bulk-copy engine
    __syncthreads();
    if (tid == 0 && len) {
        tma_store_1d(dt + base + d0, S.out_t, len);
        if (kLast) {
            tma_store_1d(dhi + base + d0, out_hi, len * 4u);
            tma_store_1d(dlo + base + d0, out_lo, len * 4u);
        } else {
            tma_store_1d(dv + base + d0, S.out_v, len * 8u);
        }
        tma_store_commit_and_wait();  // shared memory must outlive the reads of the copy engine
    }
}

// ------------------------------------------------------------------------------------------
// block-pair join
// ------------------------------------------------------------------------------------------
// Merge-path split with B first on ties: smallest i in [max(0, d-lb), min(d, la)] such that
// B[d-i-1] <= A[i].  Then A[i-1] < B[d-i] and B[d-i-1] <= A[i].
template <typename PtrT>
__device__ __forceinline__ uint32_t split_bfirst(PtrT A, uint32_t la, PtrT B, uint32_t lb, uint32_t d) {
    uint32_t lo = d > lb ? d - lb : 0, hi = min(d, la);
    while (lo < hi) {
        const uint32_t mid = (lo + hi) >> 1;
        if (B[d - mid - 1] <= A[mid]) hi = mid; else lo = mid + 1;
    }
    return lo;
}

struct JoinSmem {
    uint32_t hi[kJCap];  // off-diagonal: the segment's A and B key slices; diagonal: hi | lo | tags
    uint32_t cnt[kJR * kJR / 2];  // 16-bit counters, two per word (count <= stride < 65536)
    uint64_t bar;
    unsigned long long item;
    uint32_t split[kJThreads + 1];
    uint32_t na[kJR], nb[kJR];
    // deferred ties: the hot loop only marks tie steps; their (A cursor, B index) pairs are queued
    // here and resolved later by the whole CTA at once (one tie per thread, loads in flight
    // together) instead of serially behind the segment's barrier
    uint32_t tq[kTQ];            // slot << 28 | j_rel << 14 | i_rel (relative to the segment's i0 / j0)
    uint32_t tq_i0[kTQSegs], tq_j0[kTQSegs];
    uint32_t tq_n;
    uint32_t item_ties;          // ties seen by the replay variant in the current item (picks the variant)
};
static_assert(3 * (sizeof(JoinSmem) + 1024) <= 233472, "three join CTAs must fit one SM");

// One block list in global memory (structure of arrays).
struct ListView {
    const uint32_t *hi, *lo;
    const uint8_t *tag;
    uint32_t len;
};

// Tie path of an off-diagonal item: B entry gj met an A key equal to its own at staged A index i.
// Walk A's run of that key (staged keys first, then global memory) and count the entries whose lo
// word also matches -- (hi, lo) equality is value equality.  lo words and tags are read from
// global memory (L2): ties are rare unless the two blocks hold related genomes.
__device__ __forceinline__ void cnt_inc(uint32_t *cnt, uint32_t idx) {
    atomicAdd(&cnt[idx >> 1], 1u << (16u * (idx & 1u)));
}
__device__ __forceinline__ uint32_t cnt_get(const uint32_t *cnt, uint32_t idx) {
    return (cnt[idx >> 1] >> (16u * (idx & 1u))) & 0xFFFFu;
}

__device__ __forceinline__ void match_run(uint32_t *cnt, const uint32_t *Ah, uint32_t a_ext, uint32_t i, uint32_t kb,
                                          const ListView &A, uint32_t i0, const ListView &B, uint32_t gj) {
    const uint32_t lob = B.lo[gj], tagb = B.tag[gj];
    for (uint32_t x = i; i0 + x < A.len; x++) {
        const uint32_t key = x < a_ext ? Ah[x] : A.hi[i0 + x];
        if (key != kb) break;
        if (A.lo[i0 + x] == lob) cnt_inc(cnt, (uint32_t)A.tag[i0 + x] * kJR + tagb);
    }
}

// Geometry of one segment of an off-diagonal item (uniform across the CTA).
struct SegGeom {
    uint32_t i0, i1, j0, j1, d0, d1, na_s, nb_s, a_ext, a_lo, a_off, a_cnt, b_lo, b_off, b_cnt;
};

// Queue the tie steps of one chain: bit t of `ties` marks a step whose B entry met an equal key
// at the A cursor, bit t of `took_b` says whether step t consumed a B entry, so the cursors at
// step t follow from a population count -- no re-walk of the chain, no shared-memory reads.
// (i0, j0) are the chain's cursors before step 0.  A queue that is full takes the tie at once.
// Out of line: the hot loop only records the two bit masks and never branches.
__device__ __noinline__ void queue_ties(JoinSmem &S, const uint32_t *Ah, const uint32_t *Bh, uint32_t a_ext,
                                        uint32_t i0c, uint32_t j0c, uint32_t ties, uint32_t took_b, uint32_t slot,
                                        const ListView A, uint32_t i0, const ListView B, uint32_t j0) {
    uint32_t q = atomicAdd(&S.tq_n, (uint32_t)__popc(ties));
    while (ties) {
        const uint32_t t = (uint32_t)__ffs((int)ties) - 1u;
        ties &= ties - 1u;
        const uint32_t nb = (uint32_t)__popc(took_b & ((1u << t) - 1u));
        const uint32_t j = j0c + nb, i = i0c + t - nb;
        if (q < (uint32_t)kTQ) S.tq[q] = (slot << 28) | (j << 14) | i;
        else match_run(S.cnt, Ah, a_ext, i, Bh[j], A, i0, B, j0 + j);
        q++;
    }
}

// The same without the took_b mask: re-walk the chain's steps from (i, j) in shared memory.  Used
// while an item has shown few ties, where keeping the second mask in the hot loop costs more
// (one instruction per step) than the occasional re-walk.
__device__ __noinline__ void queue_ties_replay(JoinSmem &S, const uint32_t *Ah, const uint32_t *Bh, uint32_t a_ext,
                                               uint32_t i, uint32_t j, uint32_t steps, uint32_t ties, uint32_t slot,
                                               const ListView A, uint32_t i0, const ListView B, uint32_t j0) {
    const uint32_t n_ties = (uint32_t)__popc(ties);
    atomicAdd(&S.item_ties, n_ties);
    uint32_t q = atomicAdd(&S.tq_n, n_ties);
    uint32_t ka = Ah[i], kb = Bh[j];
    for (uint32_t t = 0; t < steps && (ties >> t) != 0; t++) {
        const bool tb = kb <= ka;
        if ((ties >> t) & 1u) {
            if (q < (uint32_t)kTQ) S.tq[q] = (slot << 28) | (j << 14) | i;
            else match_run(S.cnt, Ah, a_ext, i, kb, A, i0, B, j0 + j);
            q++;
        }
        if (tb) { j++; kb = Bh[j]; } else { i++; ka = Ah[i]; }
    }
}

// Resolve the queued ties, one per thread: all six words a tie needs first are requested
// together (L2), so a round costs one memory latency for 256 ties.  CTA-wide; ends with the
// queue empty.  Callers synchronise before (queue complete) -- the trailing barrier is here.
__device__ __forceinline__ void flush_ties(JoinSmem &S, const ListView &A, const ListView &B, uint32_t tid) {
    const uint32_t n = min(S.tq_n, (uint32_t)kTQ);
    for (uint32_t q = tid; q < n; q += kJThreads) {
        const uint32_t e = S.tq[q], slot = e >> 28;
        uint32_t gi = S.tq_i0[slot] + (e & 0x3FFFu);
        const uint32_t gj = S.tq_j0[slot] + ((e >> 14) & 0x3FFFu);
        const uint32_t kb = B.hi[gj], lob = B.lo[gj], tagb = B.tag[gj];
        uint32_t ka = A.hi[gi], loa = A.lo[gi], taga = A.tag[gi];
        for (;;) {
            if (ka != kb) break;
            const uint32_t nx = gi + 1;
            const bool more = nx < A.len;
            const uint32_t ka2 = more ? A.hi[nx] : ~kb, loa2 = more ? A.lo[nx] : 0u, taga2 = more ? A.tag[nx] : 0u;
            if (loa == lob) cnt_inc(S.cnt, taga * kJR + tagb);
            if (!more) break;
            gi = nx; ka = ka2; loa = loa2; taga = taga2;
        }
    }
    __syncthreads();
    if (tid == 0) S.tq_n = 0;
}

__device__ __forceinline__ uint32_t lds32(uint32_t addr) {
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}

// Merge-intersect one staged segment.  kChecked = false requires that the segment is full and
// touches no list end (every key a thread can look at is a real entry): every thread then runs
// kJChains independent chains of exactly kJE branch-free steps, interleaved for ILP, each chain
// starting at its own merge-path split (no shared split table, no barrier).  kChecked = true
// bounds every step by the chain's own end split and takes ties at once (list ends only).
template <bool kChecked, bool kTrack>
__device__ __forceinline__ void join_segment(JoinSmem &S, const ListView &A, const ListView &B, const SegGeom &g,
                                             uint32_t tid, uint32_t slot) {
    const uint32_t *Ah = S.hi + g.a_off, *Bh = S.hi + g.a_cnt + g.b_off;
    const uint32_t len = g.na_s + g.nb_s;
    if (!kChecked) {
        uint32_t pa[kJChains], pb[kJChains], ka[kJChains], kb[kJChains], ties[kJChains], tookb[kJChains], is[kJChains];
        const uint32_t a_base = smem_u32(Ah), b_base = smem_u32(Bh);
#pragma unroll
        for (int c = 0; c < kJChains; c++) is[c] = split_bfirst(Ah, g.na_s, Bh, g.nb_s, (c * kJThreads + tid) * kJE);
#pragma unroll
        for (int c = 0; c < kJChains; c++) {
            const uint32_t cs = c * kJThreads + tid;
            pa[c] = a_base + 4u * is[c];
            pb[c] = b_base + 4u * (cs * kJE - is[c]);
            ka[c] = lds32(pa[c]); kb[c] = lds32(pb[c]);
            ties[c] = 0; tookb[c] = 0;
        }
#pragma unroll
        for (int t = 0; t < kJE; t++) {
#pragma unroll
            for (int c = 0; c < kJChains; c++) {
                const bool tb = kb[c] <= ka[c];
                ties[c] |= kb[c] == ka[c] ? (1u << t) : 0u;
                if (kTrack) tookb[c] |= tb ? (1u << t) : 0u;
                const uint32_t pn = (tb ? pb[c] : pa[c]) + 4u;
                const uint32_t nv = lds32(pn);
                pb[c] = tb ? pn : pb[c];
                pa[c] = tb ? pa[c] : pn;
                kb[c] = tb ? nv : kb[c];
                ka[c] = tb ? ka[c] : nv;
            }
        }
#pragma unroll
        for (int c = 0; c < kJChains; c++)
            if (ties[c]) {
                if (kTrack)
                    queue_ties(S, Ah, Bh, g.a_ext, is[c], (c * kJThreads + tid) * kJE - is[c], ties[c], tookb[c], slot, A,
                               g.i0, B, g.j0);
                else
                    queue_ties_replay(S, Ah, Bh, g.a_ext, is[c], (c * kJThreads + tid) * kJE - is[c], kJE, ties[c], slot,
                                      A, g.i0, B, g.j0);
            }
    } else {
        for (int c = 0; c < kJChains; c++) {
            const uint32_t cs = c * kJThreads + tid;
            const uint32_t dt0 = min(cs * kJE, len), dt1 = min(dt0 + (uint32_t)kJE, len);
            if (dt0 == dt1) continue;
            uint32_t i = split_bfirst(Ah, g.na_s, Bh, g.nb_s, dt0), j = dt0 - i;
            const uint32_t ie = split_bfirst(Ah, g.na_s, Bh, g.nb_s, dt1), je = dt1 - ie;
            uint32_t ka = Ah[i], kb = Bh[j];
            for (uint32_t t = dt0; t < dt1; t++) {
                const bool tb = (i >= ie) || (j < je && kb <= ka);
                if (tb && kb == ka) match_run(S.cnt, Ah, g.a_ext, i, kb, A, g.i0, B, g.j0 + j);
                j += tb ? 1u : 0u;
                i += tb ? 0u : 1u;
                const uint32_t nv = *(tb ? Bh + j : Ah + i);
                kb = tb ? nv : kb;
                ka = tb ? ka : nv;
            }
        }
    }
}

__global__ void __launch_bounds__(kJThreads, kJCtasPerSm) prefilter_join_kernel(const KernelParams p) {
    extern __shared__ __align__(128) uint8_t smem_raw[];
    JoinSmem &S = *reinterpret_cast<JoinSmem *>(smem_raw);
    const uint32_t tid = threadIdx.x;
    if (p.mode_flag && *p.mode_flag != 0) return;  // small tie-dense launch: the pairwise kernel takes it
    if (tid == 0) { mbar_init(&S.bar, 1); fence_mbar_init(); S.tq_n = 0; }
    __syncthreads();
    uint32_t phase = 0;
    // Item order, longest first: the diagonal items (ties are the rule there: started first they
    // overlap everything else instead of forming the kernel's tail), then the items of adjacent
    // blocks (rb, rb + 1) -- input that keeps relatives together puts its cross-block ties there --
    // then the remaining off-diagonal items row by row.  Diagonal rows are the last n_diag local
    // rows; rows with an adjacent item are the n_adj local rows from adj_lr0.
    const uint64_t n_items = p.items ? (uint64_t)p.n_explicit : (uint64_t)p.n_diag + p.n_adj + p.item_prefix[p.n_local_rb];

    for (;;) {
        if (tid == 0) S.item = atomicAdd(p.work_counter, 1ull);
        __syncthreads();
        const unsigned long long item = S.item;
        if (item >= n_items) break;
        uint32_t rb, cb;
        if (p.items) {
            rb = p.items[2 * item]; cb = p.items[2 * item + 1];
        } else if (item < p.n_diag) {
            rb = cb = p.local_rb[p.n_local_rb - p.n_diag + (uint32_t)item];
        } else if (item < (uint64_t)p.n_diag + p.n_adj) {
            rb = p.local_rb[p.adj_lr0 + (uint32_t)(item - p.n_diag)];
            cb = rb + 1;
        } else {
            const unsigned long long it = item - p.n_diag - p.n_adj;
            uint32_t lo = 0, hi = p.n_local_rb;  // last lr with item_prefix[lr] <= it (rows without items are skipped)
            while (hi - lo > 1) {
                const uint32_t mid = (lo + hi) >> 1;
                if (p.item_prefix[mid] <= it) lo = mid; else hi = mid;
            }
            rb = p.local_rb[lo];
            const uint32_t has_adj = lo >= p.adj_lr0 && lo < p.adj_lr0 + p.n_adj ? 1u : 0u;
            cb = max(rb + 1, p.cb_lo) + has_adj + (uint32_t)(it - p.item_prefix[lo]);
        }
        const uint32_t row0 = rb * kJR, col0 = cb * kJR;
        unsigned long long t_item0 = 0;
        if (p.dbg_buf && tid == 0) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_item0));

        for (uint32_t x = tid; x < kJR * kJR / 2; x += kJThreads) S.cnt[x] = 0;
        if (tid == 0) S.item_ties = 0;
        for (uint32_t x = tid; x < kJR; x += kJThreads) {
            S.na[x] = row0 + x < p.n ? min(p.counts[row0 + x], p.stride) : 0;
            S.nb[x] = col0 + x < p.n ? min(p.counts[col0 + x], p.stride) : 0;
        }
        ListView A, B;
        A.hi = p.bl_hi + (uint64_t)rb * p.bl_cap; A.lo = p.bl_lo + (uint64_t)rb * p.bl_cap;
        A.tag = p.bl_tags + (uint64_t)rb * p.bl_cap; A.len = p.bl_len[rb];
        B.hi = p.bl_hi + (uint64_t)cb * p.bl_cap; B.lo = p.bl_lo + (uint64_t)cb * p.bl_cap;
        B.tag = p.bl_tags + (uint64_t)cb * p.bl_cap; B.len = p.bl_len[cb];
        const uint32_t la = A.len, lb = B.len;
        __syncthreads();
        // every sketch of both blocks is long enough that a pair without a common hash cannot pass:
        // the threshold pass may then skip the zero counters word by word (almost all of them)
        const bool skip_zero = __syncthreads_and(tid >= kJR || (S.na[tid] >= p.zero_fails_from && S.nb[tid] >= p.zero_fails_from)) != 0;

        if (rb == cb) {
            // ---- diagonal item: equal values are adjacent in the single list (ordered by tag
            // inside a run); hi, lo and tags are all staged.  Thread t owns entries t, t + 256, ..
            // of the segment and the warp walks the runs in lock step: at distance d every lane
            // compares its entry with the one d places on (conflict-free shared-memory reads,
            // no divergent per-lane loops); the walk ends when no lane's key run goes on.
            constexpr uint32_t kSlice = kJDiag + kJR + 16;  // staged entries per array
            uint32_t *Th = S.hi, *Tl = S.hi + kSlice;
            uint8_t *Tt = reinterpret_cast<uint8_t *>(S.hi + 2 * kSlice);
            for (uint32_t x0 = 0; x0 < la; x0 += kJDiag) {
                const uint32_t x1 = min(la, x0 + (uint32_t)kJDiag);
                const uint32_t ext = min(la, x1 + (uint32_t)kJR) - x0;
                const uint32_t a_cnt = (ext + 15u) & ~15u;
                if (tid == 0) {
                    fence_proxy_async();
                    mbar_arrive_expect_tx(&S.bar, a_cnt * 9u);
                    tma_load_1d(Th, A.hi + x0, a_cnt * 4u, &S.bar);
                    tma_load_1d(Tl, A.lo + x0, a_cnt * 4u, &S.bar);
                    tma_load_1d(Tt, A.tag + x0, a_cnt, &S.bar);
                }
                mbar_wait(&S.bar, phase); phase ^= 1;
                const uint32_t nx = x1 - x0;
                // kDR entries per thread in flight: their run walks are independent, so the
                // shared-memory loads of one step overlap instead of forming one dependent chain
                for (uint32_t xb = 0; xb < nx; xb += kDR * kJThreads) {  // uniform trip count across the warp
                    uint32_t h[kDR], l[kDR], tx[kDR];
                    bool run[kDR];
                    uint32_t hit_end = 0;
#pragma unroll
                    for (int r = 0; r < kDR; r++) {
                        const uint32_t x = xb + r * kJThreads + tid;
                        run[r] = x < nx;
                        h[r] = run[r] ? Th[x] : 0u; l[r] = run[r] ? Tl[x] : 0u; tx[r] = run[r] ? Tt[x] : 0u;
                    }
                    for (uint32_t d = 1;; d++) {
                        bool any = false;
#pragma unroll
                        for (int r = 0; r < kDR; r++) {
                            const uint32_t y = xb + r * kJThreads + tid + d;
                            if (run[r] && y >= ext) { run[r] = false; hit_end |= 1u << r; }
                            const uint32_t hy = run[r] ? Th[y] : 0u, ly = run[r] ? Tl[y] : 0u, ty = run[r] ? Tt[y] : 0u;
                            run[r] = run[r] && hy == h[r];
                            if (run[r] && ly == l[r]) cnt_inc(S.cnt, min(tx[r], ty) * kJR + max(tx[r], ty));
                            any = any || run[r];
                        }
                        if (!__any_sync(0xffffffffu, any)) break;
                    }
                    // a key run that outlasts the staged extension (>= kJR equal keys) goes on in global memory
#pragma unroll
                    for (int r = 0; r < kDR; r++)
                        if (hit_end & (1u << r))
                            for (uint32_t gy = x0 + ext; gy < la && A.hi[gy] == h[r]; gy++)
                                if (A.lo[gy] == l[r]) {
                                    const uint32_t ty = A.tag[gy];
                                    cnt_inc(S.cnt, min(tx[r], ty) * kJR + max(tx[r], ty));
                                }
                }
                __syncthreads();
            }
        } else if (la != 0 && lb != 0) {
            // ---- off-diagonal item: CTA-wide merge-path intersection of two block lists
            const uint32_t total = la + lb;
            const uint32_t nseg = (total + kJD - 1) / kJD;
            bool done = false, dense = false;
            uint32_t slot = 0;  // segments since the last flush of the tie queue
            for (uint32_t kb = 0; kb < nseg && !done; kb += kJThreads - 1) {
                const uint32_t nsb = min((uint32_t)(kJThreads - 1), nseg - kb);
                if (tid <= nsb) {
                    const uint32_t d = (uint32_t)min((uint64_t)(kb + tid) * kJD, (uint64_t)total);
                    S.split[tid] = split_bfirst(A.hi, la, B.hi, lb, d);
                }
                __syncthreads();
                for (uint32_t s = 0; s < nsb; s++) {
                    SegGeom g;
                    g.d0 = (kb + s) * kJD; g.d1 = min(g.d0 + (uint32_t)kJD, total);
                    g.i0 = S.split[s]; g.i1 = S.split[s + 1];
                    g.j0 = g.d0 - g.i0; g.j1 = g.d1 - g.i1;
                    if (g.i0 >= la || g.j0 >= lb) { done = true; break; }  // one list is exhausted
                    g.na_s = g.i1 - g.i0; g.nb_s = g.j1 - g.j0;
                    if (g.nb_s == 0) continue;
                    g.a_ext = min(la, g.i1 + 16u) - g.i0;  // a few keys past the slice for short runs
                    g.a_lo = g.i0 & ~15u; g.a_off = g.i0 - g.a_lo;
                    g.a_cnt = (g.a_off + g.a_ext + 1u + 15u) & ~15u;
                    g.b_lo = g.j0 & ~15u; g.b_off = g.j0 - g.b_lo;
                    g.b_cnt = (g.b_off + g.nb_s + 1u + 15u) & ~15u;
                    if (tid == 0) {
                        S.tq_i0[slot] = g.i0; S.tq_j0[slot] = g.j0;
                        fence_proxy_async();
                        mbar_arrive_expect_tx(&S.bar, (g.a_cnt + g.b_cnt) * 4u);
                        tma_load_1d(S.hi, A.hi + g.a_lo, g.a_cnt * 4u, &S.bar);
                        tma_load_1d(S.hi + g.a_cnt, B.hi + g.b_lo, g.b_cnt * 4u, &S.bar);
                    }
                    mbar_wait(&S.bar, phase); phase ^= 1;
                    // a full segment that ends before either list does holds only real entries
                    if (g.i1 < la && g.j1 < lb && g.d1 - g.d0 == (uint32_t)kJD) {
                        if (dense) join_segment<false, true>(S, A, B, g, tid, slot);
                        else join_segment<false, false>(S, A, B, g, tid, slot);
                    } else {
                        join_segment<true, false>(S, A, B, g, tid, slot);
                    }
                    // the staged slices are free for the next TMA and the queue is complete; the thread
                    // whose enqueue came last sees the final fill, so the OR is the same for all
                    const bool half_full =
                        __syncthreads_or(*reinterpret_cast<volatile uint32_t *>(&S.tq_n) >= (uint32_t)(kTQ / 2));
                    // an item whose segments tie often (relatives in the two blocks) switches to the
                    // variant that tracks the B-takes in the hot loop instead of re-walking chains
                    if (!dense)
                        dense = __syncthreads_or(*reinterpret_cast<volatile uint32_t *>(&S.item_ties) >= kDenseTies);
                    if (++slot == (uint32_t)kTQSegs || half_full) {
                        flush_ties(S, A, B, tid);
                        slot = 0;
                    }
                }
                __syncthreads();
            }
            __syncthreads();
            flush_ties(S, A, B, tid);
        }
        __syncthreads();

        // ---- threshold the count matrix; survivors get their exact `total` and are appended
        if (skip_zero && rb != cb) {
            for (uint32_t w = tid; w < kJR * kJR / 2; w += kJThreads) {
                const uint32_t word = S.cnt[w];
                if (word == 0u) continue;
#pragma unroll
                for (uint32_t h = 0; h < 2; h++) {
                    const uint32_t common = (word >> (16u * h)) & 0xFFFFu;
                    if (common == 0u) continue;
                    const uint32_t e = 2u * w + h, r = e / kJR, c = e % kJR;
                    const uint32_t gi = row0 + r, gj = col0 + c;
                    if (gi >= p.n || gj >= p.n) continue;
                    finish_pair(p, gi, gj, p.hashes + (size_t)gi * p.stride, S.na[r], p.hashes + (size_t)gj * p.stride, S.nb[c], common);
                }
            }
        } else
        for (uint32_t e = tid; e < kJR * kJR; e += kJThreads) {
            const uint32_t r = e / kJR, c = e % kJR;
            const uint32_t gi = row0 + r, gj = col0 + c;
            if (gi >= p.n || gj >= p.n || gj <= gi) continue;
            finish_pair(p, gi, gj, p.hashes + (size_t)gi * p.stride, S.na[r],
                        p.hashes + (size_t)gj * p.stride, S.nb[c], cnt_get(S.cnt, e));
        }
        if (p.dbg_buf && tid == 0) {
            unsigned long long t1;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
            uint32_t smid;
            asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
            p.dbg_buf[item * 4 + 0] = t_item0; p.dbg_buf[item * 4 + 1] = t1;
            p.dbg_buf[item * 4 + 2] = smid; p.dbg_buf[item * 4 + 3] = ((unsigned long long)rb << 32) | cb;
        }
    }
}

// ------------------------------------------------------------------------------------------
// host
// ------------------------------------------------------------------------------------------
bool join_supported(size_t stride) { return stride < 65536; }  // 16-bit pair counters

// Largest valid hash of a (slice of a) sketch table -> *d_max (device scalar), no host sync.
int table_max_enqueue(const uint64_t *d_hashes, const uint32_t *d_counts, size_t n, size_t stride,
                      unsigned long long *d_max, cudaStream_t stream) {
    GB_CUDA(cudaMemsetAsync(d_max, 0, sizeof(unsigned long long), stream));
    const uint32_t nb = (uint32_t)((n + kJR - 1) / kJR);
    if (nb == 0) return 0;
    bl_len_kernel<<<(uint32_t)(((uint64_t)nb * 32 + 255) / 256), 256, 0, stream>>>(
        d_hashes, d_counts, (uint32_t)n, (uint32_t)stride, 0, nb, 0, 0, nullptr, d_max);
    GB_LAUNCH_CHECK();
    return 0;
}

// Block lists of a LOCAL slice of rows (the slice starts at a multiple of kShardRows rows of the
// global table, so local and global row tags agree) with the table-wide largest hash supplied by
// the caller (device scalar; at G ranks: an all-reduce MAX of table_max_enqueue's results).
int blocklist_build_local(PrefilterWorkspace &ws, const uint64_t *d_rows, const uint32_t *d_counts, size_t n_rows,
                          size_t stride, const unsigned long long *d_gmax, uint32_t *d_hi, uint32_t *d_lo,
                          uint8_t *d_tags, uint32_t *d_len, uint32_t n_blocks_out, cudaStream_t stream) {
    if (!ws.d_gmax) GB_CUDA(cudaMalloc(&ws.d_gmax, sizeof(unsigned long long)));
    GB_CUDA(cudaMemcpyAsync(ws.d_gmax, d_gmax, sizeof(unsigned long long), cudaMemcpyDeviceToDevice, stream));
    return blocklist_build(ws, d_rows, d_counts, n_rows, stride, 0, n_blocks_out, d_hi, d_lo, d_tags, d_len, stream,
                           /*gmax_ready=*/true);
}

void blocklist_layout(size_t n, size_t stride, size_t *n_blocks, size_t *entries_per_block, size_t *slack) {
    *n_blocks = (n + kJR - 1) / kJR;
    *entries_per_block = (size_t)kJR * stride;
    *slack = kBlSlack;
}

// Builds the block lists of blocks [b0, b1) from the full table into d_hi / d_lo / d_tags /
// d_len, which are SLICE-based: block b lands at offset (b - b0) * entries_per_block.
int blocklist_build(PrefilterWorkspace &ws, const uint64_t *d_hashes, const uint32_t *d_counts, size_t n,
                    size_t stride, uint32_t b0, uint32_t b1, uint32_t *d_hi, uint32_t *d_lo, uint8_t *d_tags,
                    uint32_t *d_len, cudaStream_t stream, bool gmax_ready) {
    const uint32_t nb = (uint32_t)((n + kJR - 1) / kJR);
    if (b0 > b1 || b1 > nb + 64) { set_error("blocklist_build: bad block range"); return 3; }
    if (!ws.d_gmax) GB_CUDA(cudaMalloc(&ws.d_gmax, sizeof(unsigned long long)));
    if (!gmax_ready) GB_CUDA(cudaMemsetAsync(ws.d_gmax, 0, sizeof(unsigned long long), stream));
    if (nb == 0) return 0;
    int dev = 0, sms = kNumSMsFallback;
    GB_CUDA(cudaGetDevice(&dev));
    GB_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    {   // gmax already in *ws.d_gmax (streamed path: the rest of the table may not be resident yet)
        const uint32_t pb0 = gmax_ready ? b0 : 0, pb1 = gmax_ready ? b1 : std::max(nb, b1);
        if (pb1 > pb0) {
            bl_len_kernel<<<(uint32_t)(((uint64_t)(pb1 - pb0) * 32 + 255) / 256), 256, 0, stream>>>(
                d_hashes, d_counts, (uint32_t)n, (uint32_t)stride, pb0, pb1, b0, b1, d_len,
                gmax_ready ? nullptr : ws.d_gmax);
            GB_LAUNCH_CHECK();
        }
    }
    if (b1 == b0) return 0;

    const uint64_t bl_cap = (uint64_t)kJR * stride;
    const uint64_t total = (uint64_t)(b1 - b0) * bl_cap;
    if (ws.cap_bl < total) {
        for (int x = 0; x < 2; x++) {
            if (ws.d_bl_vals[x]) GB_CUDA(cudaFree(ws.d_bl_vals[x]));
            if (ws.d_bl_tags[x]) GB_CUDA(cudaFree(ws.d_bl_tags[x]));
            ws.d_bl_vals[x] = nullptr; ws.d_bl_tags[x] = nullptr;
        }
        ws.cap_bl = 0;
        for (int x = 0; x < 2; x++) {
            // + slack: the bulk copies of the merge levels round their windows up to 16 bytes
            GB_CUDA(cudaMalloc(&ws.d_bl_vals[x], (total + 64) * 8));
            GB_CUDA(cudaMalloc(&ws.d_bl_tags[x], total + 64));
        }
        ws.cap_bl = total;
    }
    uint64_t m0 = stride;  // run length the global merge levels start from
    if (stride <= kSeedMaxStride) {
        const size_t smem = seed_smem_bytes((uint32_t)stride);
        GB_CUDA(cudaFuncSetAttribute(bl_seed_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        bl_seed_kernel<<<(uint32_t)((b1 - b0) * (kJR / kSeedRows)), kSeedThreads, smem, stream>>>(
            d_hashes, d_counts, (uint32_t)n, (uint32_t)stride, (uint64_t)b0 * kJR, ws.d_bl_vals[0], ws.d_bl_tags[0]);
        GB_LAUNCH_CHECK();
        m0 = (uint64_t)kSeedRows * stride;
    } else {
        const uint32_t grid = (uint32_t)std::min<uint64_t>((total + 255) / 256, (uint64_t)sms * 16);
        bl_init_kernel<<<grid, 256, 0, stream>>>(d_hashes, d_counts, (uint32_t)n, (uint32_t)stride,
                                                 (uint64_t)b0 * kJR, total, ws.d_bl_vals[0], ws.d_bl_tags[0]);
        GB_LAUNCH_CHECK();
    }
    GB_CUDA(cudaFuncSetAttribute(bl_merge_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(MergeSmem)));
    GB_CUDA(cudaFuncSetAttribute(bl_merge_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(MergeSmem)));
    GB_CUDA(cudaFuncSetAttribute(bl_merge_tma_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(MergeTmaSmem)));
    GB_CUDA(cudaFuncSetAttribute(bl_merge_tma_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(MergeTmaSmem)));
    int src = 0;
    {   // tile-boundary splits: at most (chunks + 1) per output run, most at the first level
        const uint64_t need = total / (2 * stride) * ((2 * stride + kMTile - 1) / kMTile + 1) + 16;
        if (ws_ensure(ws.d_splits, ws.cap_splits, need)) return 2;
    }
    for (uint64_t m = m0; m < bl_cap; m *= 2) {
        const uint32_t chunks = (uint32_t)((2 * m + kMTile - 1) / kMTile);
        const uint64_t grid = total / (2 * m) * chunks;
        if (grid > 0x7FFFFFFFull) { set_error("prefilter: table too large for the merge grid"); return 3; }
        const uint64_t n_bounds = total / (2 * m) * (chunks + 1);
        bl_partition_kernel<<<(uint32_t)((n_bounds * 32 + 255) / 256), 256, 0, stream>>>(
            ws.d_bl_vals[src], ws.d_bl_tags[src], (uint32_t)m, chunks, n_bounds, ws.d_splits);
        GB_LAUNCH_CHECK();
        const bool last = 2 * m >= bl_cap;  // last level: structure-of-arrays output into the caller's buffers
        const bool tma = m % 16 == 0;  // run starts and tile lengths are then multiples of 16 entries (bulk-copy alignment)
        uint64_t *dvv = last ? nullptr : ws.d_bl_vals[src ^ 1];
        uint8_t *dtt = last ? d_tags : ws.d_bl_tags[src ^ 1];
        uint32_t *dh = last ? d_hi : nullptr, *dl = last ? d_lo : nullptr;
        if (tma) {
            if (last)
                bl_merge_tma_kernel<true><<<(uint32_t)grid, kMThreads, sizeof(MergeTmaSmem), stream>>>(
                    ws.d_bl_vals[src], ws.d_bl_tags[src], dvv, dtt, (uint32_t)m, chunks, ws.d_splits, ws.d_gmax, dh, dl);
            else
                bl_merge_tma_kernel<false><<<(uint32_t)grid, kMThreads, sizeof(MergeTmaSmem), stream>>>(
                    ws.d_bl_vals[src], ws.d_bl_tags[src], dvv, dtt, (uint32_t)m, chunks, ws.d_splits, ws.d_gmax, dh, dl);
        } else if (last) {
            bl_merge_kernel<true><<<(uint32_t)grid, kMThreads, sizeof(MergeSmem), stream>>>(
                ws.d_bl_vals[src], ws.d_bl_tags[src], dvv, dtt, (uint32_t)m, chunks, ws.d_splits, ws.d_gmax, dh, dl);
        } else {
            bl_merge_kernel<false><<<(uint32_t)grid, kMThreads, sizeof(MergeSmem), stream>>>(
                ws.d_bl_vals[src], ws.d_bl_tags[src], dvv, dtt, (uint32_t)m, chunks, ws.d_splits, ws.d_gmax, dh, dl);
        }
        GB_LAUNCH_CHECK();
        src ^= 1;
    }
    return 0;
}

// Launches the join of one shard over block lists that cover the WHOLE table (block b at offset
// b * entries_per_block; at least kBlSlack entries of readable slack behind the last list).
int join_launch(PrefilterWorkspace &ws, KernelParams &p, const uint32_t *d_hi, const uint32_t *d_lo,
                const uint8_t *d_tags, const uint32_t *d_len, uint32_t shard, uint32_t n_shards,
                cudaStream_t stream) {
    const uint32_t nb = (p.n + kJR - 1) / kJR;
    p.bl_hi = d_hi; p.bl_lo = d_lo; p.bl_tags = d_tags; p.bl_len = d_len; p.bl_cap = (uint64_t)kJR * p.stride;
    int dev = 0, sms = kNumSMsFallback;
    GB_CUDA(cudaGetDevice(&dev));
    GB_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    if (int rc = upload_join_work_list(ws, p.n, shard, n_shards, stream, p)) return rc;
    if (p.n_local_rb == 0) return 0;
    uint64_t n_items = 0;
    for (uint32_t rb = 0; rb < nb; rb++) if (shard_of_group(rb, n_shards) == shard) n_items += nb - rb;
    const size_t smem = sizeof(JoinSmem);
    GB_CUDA(cudaFuncSetAttribute(prefilter_join_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const uint32_t grid = (uint32_t)std::min<uint64_t>(n_items, (uint64_t)sms * kJCtasPerSm);
    if (ws.record(1, stream)) return 2;
    prefilter_join_kernel<<<grid, kJThreads, smem, stream>>>(p);
    GB_LAUNCH_CHECK();
    if (ws.record(2, stream)) return 2;
    return 0;
}

// Join of an explicit list of block pairs over lists that cover the whole table (multi-GPU ring:
// one launch per peer slice as its lists arrive).  Candidates are appended to the caller's list.
int join_launch_items(PrefilterWorkspace &ws, KernelParams &p, const uint32_t *d_hi, const uint32_t *d_lo,
                      const uint8_t *d_tags, const uint32_t *d_len, const uint32_t *d_items, size_t n_items,
                      cudaStream_t stream) {
    if (n_items == 0) return 0;
    if (n_items >= 0xFFFFFFFFull) { set_error("prefilter_join: too many explicit items"); return 3; }
    constexpr uint32_t kPool = 64;
    if (!ws.d_item_counters) GB_CUDA(cudaMalloc(&ws.d_item_counters, kPool * sizeof(unsigned long long)));
    unsigned long long *ctr = ws.d_item_counters + (ws.item_counter_next++ % kPool);
    GB_CUDA(cudaMemsetAsync(ctr, 0, sizeof(unsigned long long), stream));
    p.bl_hi = d_hi; p.bl_lo = d_lo; p.bl_tags = d_tags; p.bl_len = d_len; p.bl_cap = (uint64_t)kJR * p.stride;
    p.n_row_blocks = (p.n + kJR - 1) / kJR;
    p.items = d_items; p.n_explicit = (uint32_t)n_items; p.work_counter = ctr;
    int dev = 0, sms = kNumSMsFallback;
    GB_CUDA(cudaGetDevice(&dev));
    GB_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    const size_t smem = sizeof(JoinSmem);
    GB_CUDA(cudaFuncSetAttribute(prefilter_join_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const uint32_t grid = (uint32_t)std::min<uint64_t>(n_items, (uint64_t)sms * kJCtasPerSm);
    if (ws.record(1, stream)) return 2;
    prefilter_join_kernel<<<grid, kJThreads, smem, stream>>>(p);
    GB_LAUNCH_CHECK();
    if (ws.record(2, stream)) return 2;
    return 0;
}

// Workspace-owned finished lists for `total` entries (+ readable slack behind the last list).
static int ensure_fin(PrefilterWorkspace &ws, uint64_t total, cudaStream_t stream) {
    if (ws.cap_fin >= total + kBlSlack) return 0;
    if (ws.d_fin_hi) GB_CUDA(cudaFree(ws.d_fin_hi));
    if (ws.d_fin_lo) GB_CUDA(cudaFree(ws.d_fin_lo));
    if (ws.d_fin_tags) GB_CUDA(cudaFree(ws.d_fin_tags));
    ws.d_fin_hi = ws.d_fin_lo = nullptr; ws.d_fin_tags = nullptr; ws.cap_fin = 0;
    GB_CUDA(cudaMalloc(&ws.d_fin_hi, (total + kBlSlack) * 4));
    GB_CUDA(cudaMalloc(&ws.d_fin_lo, (total + kBlSlack) * 4));
    GB_CUDA(cudaMalloc(&ws.d_fin_tags, total + kBlSlack));
    GB_CUDA(cudaMemsetAsync(ws.d_fin_hi + total, 0xFF, kBlSlack * 4, stream));
    GB_CUDA(cudaMemsetAsync(ws.d_fin_lo + total, 0xFF, kBlSlack * 4, stream));
    GB_CUDA(cudaMemsetAsync(ws.d_fin_tags + total, 0xFF, kBlSlack, stream));
    ws.cap_fin = total + kBlSlack;
    return 0;
}

// How tie-dense are the block lists?  Counts the entries whose 64-bit key equals the key 16 places
// on in the same list: zero for families of up to 16 relatives per block, most entries when a
// block is one clade.  The last CTA to finish sets flag[0] = 1 if more than a quarter qualify.
__global__ void __launch_bounds__(256) bl_dense_stat_kernel(const uint32_t *__restrict__ hi, const uint32_t *__restrict__ lo,
                                                            const uint32_t *__restrict__ len, uint32_t nb, uint64_t cap,
                                                            uint32_t *__restrict__ flag, uint32_t *__restrict__ done) {
    uint32_t eq = 0, seen = 0;
    for (uint32_t b = blockIdx.y; b < nb; b += gridDim.y) {
        const uint32_t l = len[b];
        const uint32_t *h = hi + (uint64_t)b * cap, *w = lo + (uint64_t)b * cap;
        for (uint32_t x = blockIdx.x * 256 + threadIdx.x; x + 16 < l; x += gridDim.x * 256) {
            eq += (h[x] == h[x + 16] && w[x] == w[x + 16]) ? 1u : 0u;
            seen++;
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { eq += __shfl_xor_sync(0xffffffffu, eq, o); seen += __shfl_xor_sync(0xffffffffu, seen, o); }
    if ((threadIdx.x & 31) == 0) { atomicAdd(&flag[1], eq); atomicAdd(&flag[2], seen); }
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        if (atomicAdd(done, 1u) + 1 == gridDim.x * gridDim.y) {
            const uint32_t e = atomicAdd(&flag[1], 0u), sn = atomicAdd(&flag[2], 0u);
            flag[0] = (sn > 0 && 4ull * e > sn) ? 1u : 0u;
        }
    }
}

// A launch of at most this many items runs every item on its own CTA at once; if the lists are
// tie-dense each item is a long serial tie resolution (milliseconds) and the pairwise kernel,
// which spreads the same pairs over all SMs, finishes first (crossover measured at ~2,000 items).
constexpr uint64_t kSmallJoinItems = 2048;

// Single-device path: build every list into workspace-owned arrays, then join.
int join_build_and_launch(PrefilterWorkspace &ws, KernelParams &p, uint32_t shard, uint32_t n_shards,
                          cudaStream_t stream) {
    const uint32_t nb = (p.n + kJR - 1) / kJR;
    const uint64_t total = (uint64_t)nb * kJR * p.stride;
    if (int rc = ensure_fin(ws, total, stream)) return rc;
    if (ws_ensure(ws.d_bl_len, ws.cap_bl_len, nb)) return 2;
    if (int rc = blocklist_build(ws, p.hashes, p.counts, p.n, p.stride, 0, nb, ws.d_fin_hi, ws.d_fin_lo,
                                 ws.d_fin_tags, ws.d_bl_len, stream))
        return rc;
    uint64_t n_items = 0;
    for (uint32_t rb = 0; rb < nb; rb++) if (shard_of_group(rb, n_shards) == shard) n_items += nb - rb;
    const bool guard = n_items <= kSmallJoinItems && n_items > 0;
    if (guard) {
        if (!ws.d_dense_flag) GB_CUDA(cudaMalloc(&ws.d_dense_flag, 4 * sizeof(uint32_t)));
        GB_CUDA(cudaMemsetAsync(ws.d_dense_flag, 0, 4 * sizeof(uint32_t), stream));
        const dim3 grid(8, std::min<uint32_t>(nb, 64));
        bl_dense_stat_kernel<<<grid, 256, 0, stream>>>(ws.d_fin_hi, ws.d_fin_lo, ws.d_bl_len, nb, (uint64_t)kJR * p.stride,
                                                     ws.d_dense_flag, ws.d_dense_flag + 3);
        GB_LAUNCH_CHECK();
        p.mode_flag = ws.d_dense_flag;
    }
    if (int rc = join_launch(ws, p, ws.d_fin_hi, ws.d_fin_lo, ws.d_fin_tags, ws.d_bl_len, shard, n_shards, stream)) return rc;
    if (guard) {
        if (int rc = pairwise_launch(ws, p, shard, n_shards, stream, true)) return rc;
        if (ws.record(2, stream)) return 2;  // the main-kernel time covers whichever kernel did the work
    }
    return 0;
}

// Host-buffer path, pipelined against the PCIe upload.  The table goes up in `chunks` equal slices
// of whole blocks on `copy`; as soon as a slice is resident, `compute` builds its block lists, so
// the build hides under the transfer.  The join runs in waves: wave w covers the items (rb <= cb)
// whose column block lies between the previous wave's end and this one's -- every list such an
// item needs is finished by then.  A wave costs at least one diagonal item's latency, so there
// are few of them: an early one at `wave_frac` of the blocks fills the time the GPU would idle
// while the rest of the table is still crossing PCIe (0 = none), and the final one.  The shift of
// the order-preserving keys needs the largest valid hash of the WHOLE table before the first
// slice is built: the host reads it off the row ends while the first slice is in flight.
// p must come from prefilter_prepare(.., d_table, d_counts, ..) on `compute`; d_counts resident.
int join_streamed_from_host(PrefilterWorkspace &ws, KernelParams &p, const uint64_t *h_hashes,
                            const uint32_t *h_counts, uint64_t *d_table, cudaStream_t compute, cudaStream_t copy,
                            int chunks, double wave_frac) {
    const uint32_t nb = (p.n + kJR - 1) / kJR;
    const uint32_t stride = p.stride;
    const uint64_t bl_cap = (uint64_t)kJR * stride;
    chunks = std::max(1, std::min<int>(chunks, (int)nb));
    std::vector<uint32_t> bound(chunks + 1, 0);
    for (int c = 1; c <= chunks; c++) bound[c] = (uint32_t)((uint64_t)nb * c / chunks);
    // waves end after the slices listed in wave_end (ascending, last = chunks - 1)
    std::vector<int> wave_end;
    if (wave_frac > 0.0 && wave_frac < 1.0) {
        const int c = (int)std::llround(wave_frac * chunks) - 1;
        if (c >= 0 && c < chunks - 1) wave_end.push_back(c);
    }
    wave_end.push_back(chunks - 1);
    const int waves = (int)wave_end.size();

    const bool debug = getenv("GALAH_B200_STREAM_DEBUG") != nullptr;
    std::vector<cudaEvent_t> dbg;
    std::vector<std::string> dbg_what;
    auto mark = [&](const std::string &what) {
        if (!debug) return;
        cudaEvent_t e; cudaEventCreate(&e); cudaEventRecord(e, compute); dbg.push_back(e); dbg_what.push_back(what);
    };
    if (int rc = ensure_fin(ws, (uint64_t)nb * bl_cap, compute)) return rc;
    if (ws_ensure(ws.d_bl_len, ws.cap_bl_len, nb)) return 2;
    while ((int)ws.chunk_ev.size() < chunks) {
        cudaEvent_t e;
        GB_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        ws.chunk_ev.push_back(e);
    }
    // work lists of all waves, one upload: wave w = { (rb, cb) : rb <= cb, w0 <= cb < w1 }; the
    // diagonal items belong to its last (w1 - w0) rows
    std::vector<uint32_t> local;
    std::vector<uint64_t> prefix;
    std::vector<size_t> off_local(waves), off_prefix(waves);
    for (int w = 0; w < waves; w++) {
        const uint32_t w0 = w ? bound[wave_end[w - 1] + 1] : 0, w1 = bound[wave_end[w] + 1];
        off_local[w] = local.size(); off_prefix[w] = prefix.size();
        prefix.push_back(0);
        for (uint32_t rb = 0; rb < w1; rb++) {
            local.push_back(rb);
            const uint32_t first = std::max(rb + 1, w0);              // first off-diagonal column block in the window
            const uint32_t has_adj = first == rb + 1 && first < w1;  // (rb, rb + 1) is scheduled ahead of the row
            prefix.push_back(prefix.back() + (w1 - std::min(w1, first + has_adj)));
        }
    }
    if (ws_ensure(ws.d_local_rb, ws.cap_local_rb, local.size())) return 2;
    if (ws_ensure(ws.d_item_prefix, ws.cap_prefix, prefix.size())) return 2;
    if (ws_ensure(ws.d_wave_counters, ws.cap_wave_counters, (size_t)waves)) return 2;
    // the copy stream may not touch d_table before earlier work on `compute` is done with it
    GB_CUDA(cudaEventRecord(ws.chunk_ev[0], compute));
    GB_CUDA(cudaStreamWaitEvent(copy, ws.chunk_ev[0], 0));
    for (int c = 0; c < chunks; c++) {
        const size_t r0 = (size_t)bound[c] * kJR, r1 = std::min<size_t>((size_t)bound[c + 1] * kJR, p.n);
        if (r1 > r0)
            GB_CUDA(cudaMemcpyAsync(d_table + r0 * stride, h_hashes + r0 * stride, (r1 - r0) * stride * 8,
                                    cudaMemcpyHostToDevice, copy));
        GB_CUDA(cudaEventRecord(ws.chunk_ev[c], copy));
    }
    unsigned long long gmax = 0;  // overlaps the first slice's DMA
    for (size_t r = 0; r < p.n; r++) {
        const uint32_t cnt = std::min(h_counts[r], stride);
        if (cnt) gmax = std::max<unsigned long long>(gmax, h_hashes[r * stride + cnt - 1]);
    }
    if (!ws.d_gmax) GB_CUDA(cudaMalloc(&ws.d_gmax, sizeof(unsigned long long)));
    GB_CUDA(cudaMemcpyAsync(ws.d_gmax, &gmax, sizeof(gmax), cudaMemcpyHostToDevice, compute));
    GB_CUDA(cudaMemcpyAsync(ws.d_local_rb, local.data(), local.size() * 4, cudaMemcpyHostToDevice, compute));
    GB_CUDA(cudaMemcpyAsync(ws.d_item_prefix, prefix.data(), prefix.size() * 8, cudaMemcpyHostToDevice, compute));
    GB_CUDA(cudaMemsetAsync(ws.d_wave_counters, 0, (size_t)waves * sizeof(unsigned long long), compute));
    int dev = 0, sms = kNumSMsFallback;
    GB_CUDA(cudaGetDevice(&dev));
    GB_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    const size_t smem = sizeof(JoinSmem);
    GB_CUDA(cudaFuncSetAttribute(prefilter_join_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    p.bl_hi = ws.d_fin_hi; p.bl_lo = ws.d_fin_lo; p.bl_tags = ws.d_fin_tags; p.bl_len = ws.d_bl_len; p.bl_cap = bl_cap;
    p.n_row_blocks = nb;
    unsigned long long *d_dbg = nullptr;
    uint64_t dbg_items = 0;
    if (getenv("GALAH_B200_ITEMLOG")) {
        dbg_items = (uint64_t)nb * (nb + 1) / 2;
        cudaMalloc(&d_dbg, dbg_items * 32); cudaMemset(d_dbg, 0, dbg_items * 32);
    }
    mark("start");
    int w = 0;
    for (int c = 0; c < chunks; c++) {
        const uint32_t b0 = bound[c], b1 = bound[c + 1];
        GB_CUDA(cudaStreamWaitEvent(compute, ws.chunk_ev[c], 0));
        if (b1 > b0) {
            if (int rc = blocklist_build(ws, p.hashes, p.counts, p.n, stride, b0, b1, ws.d_fin_hi + b0 * bl_cap,
                                         ws.d_fin_lo + b0 * bl_cap, ws.d_fin_tags + b0 * bl_cap, ws.d_bl_len + b0,
                                         compute, /*gmax_ready=*/true))
                return rc;
            mark("build of blocks [" + std::to_string(b0) + "," + std::to_string(b1) + ")");
        }
        if (c != wave_end[w]) continue;
        const uint32_t w0 = w ? bound[wave_end[w - 1] + 1] : 0;
        p.cb_lo = w0;
        p.n_local_rb = b1;
        p.n_diag = b1 - w0;
        p.adj_lr0 = std::max(w0, 1u) - 1;                    // rows w0 - 1 .. b1 - 2 have (rb, rb + 1) in the window
        p.n_adj = b1 >= 2 && b1 - 1 > p.adj_lr0 ? b1 - 1 - p.adj_lr0 : 0;
        p.local_rb = ws.d_local_rb + off_local[w];
        p.item_prefix = ws.d_item_prefix + off_prefix[w];
        p.work_counter = ws.d_wave_counters + w;
        p.dbg_buf = (d_dbg && w == waves - 1) ? d_dbg : nullptr;
        const uint64_t n_items = (uint64_t)p.n_diag + p.n_adj + prefix[off_prefix[w] + b1];
        if (w == waves - 1 && ws.record(1, compute)) return 2;
        if (n_items) {
            prefilter_join_kernel<<<(uint32_t)std::min<uint64_t>(n_items, (uint64_t)sms * kJCtasPerSm), kJThreads, smem,
                                    compute>>>(p);
            GB_LAUNCH_CHECK();
            mark("join of column blocks [" + std::to_string(w0) + "," + std::to_string(b1) + "), " +
                 std::to_string(n_items) + " items");
        }
        w++;
    }
    if (ws.record(2, compute)) return 2;
    if (debug) {
        cudaStreamSynchronize(compute);
        for (size_t x = 1; x < dbg.size(); x++) {
            float t = 0, d = 0;
            cudaEventElapsedTime(&t, dbg[0], dbg[x]);
            cudaEventElapsedTime(&d, dbg[x - 1], dbg[x]);
            fprintf(stderr, "[stream] +%.3f ms (%.3f ms since the previous mark, waits for the upload included): %s\n", t, d,
                    dbg_what[x].c_str());
        }
        for (cudaEvent_t e : dbg) cudaEventDestroy(e);
    }
    if (d_dbg) {
        cudaStreamSynchronize(compute);
        const uint64_t n_items = (uint64_t)p.n_diag + p.n_adj + prefix[off_prefix[waves - 1] + nb];
        std::vector<unsigned long long> h(n_items * 4);
        cudaMemcpy(h.data(), d_dbg, n_items * 32, cudaMemcpyDeviceToHost);
        cudaFree(d_dbg);
        FILE *f = fopen(getenv("GALAH_B200_ITEMLOG"), "w");
        unsigned long long t0 = ~0ull;
        for (uint64_t x = 0; x < n_items; x++) t0 = std::min(t0, h[x * 4]);
        for (uint64_t x = 0; x < n_items && f; x++)
            fprintf(f, "%llu %llu %llu %llu %llu %llu\n", (unsigned long long)x, h[x * 4] - t0, h[x * 4 + 1] - t0, h[x * 4 + 2],
                    h[x * 4 + 3] >> 32, h[x * 4 + 3] & 0xFFFFFFFFull);
        if (f) fclose(f);
    }
    return 0;
}

}  // namespace gb200
