// Stage 1b (K2): all-pairs sketch intersection on sm_100a.
//
// Replaces the serial loop at /root/reference/src/finch.rs:75-95, i.e. for every i<j
//   finch::distance::distance(s_i, s_j, false)  ->  raw_distance: two-pointer merge while both
//   sketches have elements, common = #equal, total = i + j - common
// followed by the Mash-ANI threshold.  The kernels here produce the INTEGERS (common, total);
// the f64 formula and the `>= min_ani as f64` comparison are finished on the host (abi.cu) so
// that `ln` rounds exactly like the reference's.
//
// Exactness notes
//   * raw_distance's loop ends when either list is exhausted, so with A the list whose maximum
//     is smaller (or equal):  i = |A|,  j = #{b in B : b <= max A},  common = |A n B|.
//     (Every common element is <= min(max A, max B), hence counted before the loop ends.)
//     `common` can therefore be computed by any intersection method; `total` follows from one
//     rank query.  Verified against the sequential merge in tests/test_prefilter_gpu.py.
//   * a pair can only pass if common >= cmin_by_tmin[min(|A|,|B|)] (total >= min(|A|,|B|) and
//     ANI is monotone in common/total); the tables are built on the host from the same f64
//     formula with one unit of slack, and every survivor is re-evaluated exactly on the host.
//
// Two paths, identical results (both compute |A n B| exactly for EVERY pair i < j):
//
//  mode 0 "block join" (default, prefilter_join.cu): the table is cut into blocks of kShardRows
//    consecutive sketches; each block is merged once into one sorted (value, row-tag) list, and a
//    work item intersects TWO BLOCK LISTS with a CTA-wide merge path, scattering matches into a
//    kShardRows x kShardRows count matrix in shared memory.  One merge of 2*64*s elements yields
//    the exact `common` of 64*64 pairs, i.e. 2s/64 merge steps per pair instead of 2s.
//
//  mode 1 "pairwise" (this file): prefilter_tiled_kernel -- a work item = one block of kRowBlock
//    row sketches resident in shared memory + up to kColChunk column blocks streamed through a
//    2-stage ring.  Every block of sketches is contiguous in the table, so each stage is ONE 1-D
//    TMA bulk copy (cp.async.bulk ... mbarrier::complete_tx, SASS UBLKCP) of kColBlock*stride*8
//    bytes.  Each warp intersects one pair at a time with a merge-path split across its lanes.
//    Also the fallback for sketches too large for the join's shared-memory tiles.
#include "prefilter.cuh"

#include <math.h>

#include <algorithm>

#include "common.cuh"
#include "prefilter_dev.cuh"

namespace gb200 {

// ------------------------------------------------------------------------------------------
// host: exact f64 formula + conservative integer thresholds
// ------------------------------------------------------------------------------------------
double mash_ani_f64(uint64_t common, uint64_t total, int k) {
    // finch 0.6 distance(): jaccard = common / total; mash_distance = -ln(2j/(1+j))/k clamped by
    // f64::min(1, f64::max(0, d)) (NaN-ignoring, like fmin/fmax); galah: 1.0 - mash_distance.
    double jaccard = (double)common / (double)total;
    double md = -1.0 * log((2.0 * jaccard) / (1.0 + jaccard)) / (double)k;
    md = fmin(1.0, fmax(0.0, md));
    return 1.0 - md;
}

PrefilterThresholds make_thresholds(uint32_t s_max, int k, float min_ani) {
    PrefilterThresholds t;
    const double thr = (double)min_ani;
    // cmin_by_total[T] = (smallest c in [0, T] with ani(c, T) >= thr) - 1 slack, T+1 if none.
    // ani is monotone non-decreasing in c for fixed T, so a binary search would do; the table is
    // tiny, so scan exactly.
    t.cmin_by_total.resize(2 * (size_t)s_max + 1);
    for (uint32_t T = 0; T <= 2 * s_max; T++) {
        uint32_t c = 0;
        while (c <= T && !(mash_ani_f64(c, T, k) >= thr)) c++;
        t.cmin_by_total[T] = c > 0 && c <= T ? c - 1 : c;
    }
    // cmin_by_tmin[m]: total >= m and common <= m, so the pair needs
    // common >= min over T >= m of cmin_by_total[T]; cmin_by_total is non-decreasing in T except
    // for the "none" sentinel, so take the running minimum from the right to stay conservative.
    t.cmin_by_tmin.resize((size_t)s_max + 1);
    uint32_t run = 0xFFFFFFFFu;
    std::vector<uint32_t> suffix_min(2 * (size_t)s_max + 2, 0xFFFFFFFFu);
    for (int64_t T = 2 * (int64_t)s_max; T >= 0; T--) {
        run = std::min(run, t.cmin_by_total[T]);
        suffix_min[T] = run;
    }
    for (uint32_t m = 0; m <= s_max; m++) t.cmin_by_tmin[m] = suffix_min[m];
    return t;
}

PrefilterThresholds make_containment_thresholds(uint32_t s_max, double frac, uint32_t bypass_below) {
    PrefilterThresholds t;
    t.cmin_by_tmin.resize((size_t)s_max + 1);
    for (uint32_t m = 0; m <= s_max; m++) {
        const double need = ceil(frac * (double)m);
        t.cmin_by_tmin[m] = need < 1.0 ? 1u : (uint32_t)need;
        // skani compares units with fewer than 20 markers against everything unless --faster-small
        if (m < bypass_below) t.cmin_by_tmin[m] = 0u;
    }
    t.cmin_by_total.assign(2 * (size_t)s_max + 1, 0u);
    return t;
}

int PrefilterWorkspace::record(int which, cudaStream_t stream) {
    if (!ev[which]) GB_CUDA(cudaEventCreate(&ev[which]));
    GB_CUDA(cudaEventRecord(ev[which], stream));
    if (which == 2) ev_recorded = true;
    return 0;
}

int PrefilterWorkspace::last_timing(float *build_ms, float *main_ms) {
    *build_ms = 0.f; *main_ms = 0.f;
    if (!ev_recorded) { set_error("prefilter: no completed launch to time"); return 3; }
    GB_CUDA(cudaEventSynchronize(ev[2]));
    GB_CUDA(cudaEventElapsedTime(build_ms, ev[0], ev[1]));
    GB_CUDA(cudaEventElapsedTime(main_ms, ev[1], ev[2]));
    return 0;
}

int PrefilterWorkspace::release() {
    for (int x = 0; x < 3; x++) if (ev[x]) cudaEventDestroy(ev[x]);
    cudaFree(d_cmin_by_tmin); cudaFree(d_cmin_by_total); cudaFree(d_item_prefix);
    cudaFree(d_local_rb); cudaFree(d_work_counter); cudaFree(d_bl_len); cudaFree(d_gmax);
    cudaFree(d_fin_hi); cudaFree(d_fin_lo); cudaFree(d_fin_tags); cudaFree(d_splits);
    for (int x = 0; x < 2; x++) { cudaFree(d_bl_vals[x]); cudaFree(d_bl_tags[x]); }
    cudaFree(d_wave_counters); cudaFree(d_item_counters);
    cudaFree(d_alt_local_rb); cudaFree(d_alt_item_prefix); cudaFree(d_alt_work_counter); cudaFree(d_dense_flag);
    for (cudaEvent_t e : chunk_ev) cudaEventDestroy(e);
    *this = PrefilterWorkspace();
    return 0;
}

// ------------------------------------------------------------------------------------------
// tiled kernel: TMA-staged shared-memory tiles, warp per pair
// ------------------------------------------------------------------------------------------
constexpr int kThreads = 512;
constexpr int kWarps = kThreads / 32;

__global__ void __launch_bounds__(kThreads, 1) prefilter_tiled_kernel(const KernelParams p) {
    extern __shared__ __align__(128) uint8_t smem_raw[];
    const uint32_t stride = p.stride;
    const uint32_t blk_elems = kColBlock * stride;           // u64 elements per block of 8 sketches
    uint64_t *rows = reinterpret_cast<uint64_t *>(smem_raw);  // kRowBlock * stride
    uint64_t *cols0 = rows + (size_t)kRowBlock * stride;      // 2 stages * kColBlock * stride
    uint64_t *bars = cols0 + 2 * (size_t)blk_elems;           // [0]=rows, [1..2]=col stages
    __shared__ unsigned long long s_item;
    __shared__ uint32_t s_row_cnt[kRowBlock];
    __shared__ uint32_t s_col_cnt[2][kColBlock];

    const uint32_t tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (p.run_if_flag && *p.mode_flag == 0) return;  // fallback launch of a table the join handles itself
    if (tid == 0) {
        mbar_init(&bars[0], 1); mbar_init(&bars[1], 1); mbar_init(&bars[2], 1);
        fence_mbar_init();
    }
    __syncthreads();
    uint32_t ph_rows = 0, ph_col[2] = {0, 0};
    const uint64_t n_items = p.item_prefix[p.n_local_rb];

    for (;;) {
        if (tid == 0) s_item = atomicAdd(p.work_counter, 1ull);
        __syncthreads();
        const unsigned long long item = s_item;
        if (item >= n_items) break;
        // item -> (local row block, column chunk)
        uint32_t lo = 0, hi = p.n_local_rb;  // last lr with item_prefix[lr] <= item
        while (hi - lo > 1) {
            uint32_t mid = (lo + hi) >> 1;
            if (p.item_prefix[mid] <= item) lo = mid; else hi = mid;
        }
        const uint32_t rb = p.local_rb[lo];
        const uint32_t chunk = (uint32_t)(item - p.item_prefix[lo]);
        const uint32_t cb0 = rb + chunk * kColChunk;
        const uint32_t cb1 = min(cb0 + (uint32_t)kColChunk, p.n_row_blocks);
        const uint32_t row0 = rb * kRowBlock;
        const uint32_t nrows = min((uint32_t)kRowBlock, p.n - row0);

        if (tid == 0) {
            const uint32_t bytes = nrows * stride * 8u;
            mbar_arrive_expect_tx(&bars[0], bytes);
            tma_load_1d(rows, p.hashes + (size_t)row0 * stride, bytes, &bars[0]);
            const uint32_t c0 = cb0 * kColBlock, nc = min((uint32_t)kColBlock, p.n - c0);
            mbar_arrive_expect_tx(&bars[1], nc * stride * 8u);
            tma_load_1d(cols0, p.hashes + (size_t)c0 * stride, nc * stride * 8u, &bars[1]);
        }
        if (tid < nrows) s_row_cnt[tid] = p.counts[row0 + tid];
        {
            const uint32_t c0 = cb0 * kColBlock, nc = min((uint32_t)kColBlock, p.n - c0);
            if (tid >= 32 && tid < 32 + nc) s_col_cnt[0][tid - 32] = p.counts[c0 + tid - 32];
        }
        mbar_wait(&bars[0], ph_rows); ph_rows ^= 1;

        for (uint32_t cb = cb0; cb < cb1; cb++) {
            const uint32_t st = (cb - cb0) & 1;
            const uint32_t col0 = cb * kColBlock;
            const uint32_t ncols = min((uint32_t)kColBlock, p.n - col0);
            if (cb + 1 < cb1) {  // prefetch next column block into the other stage
                const uint32_t c0 = (cb + 1) * kColBlock, nc = min((uint32_t)kColBlock, p.n - c0);
                if (tid == 0) {
                    mbar_arrive_expect_tx(&bars[1 + (st ^ 1)], nc * stride * 8u);
                    tma_load_1d(cols0 + (size_t)(st ^ 1) * blk_elems, p.hashes + (size_t)c0 * stride,
                                nc * stride * 8u, &bars[1 + (st ^ 1)]);
                }
                if (tid >= 32 && tid < 32 + nc) s_col_cnt[st ^ 1][tid - 32] = p.counts[c0 + tid - 32];
            }
            mbar_wait(&bars[1 + st], ph_col[st]); ph_col[st] ^= 1;
            __syncthreads();  // counts visible
            const uint64_t *cols = cols0 + (size_t)st * blk_elems;
            for (uint32_t q = warp; q < kRowBlock * kColBlock; q += kWarps) {
                const uint32_t r = q >> 3, c = q & 7;
                const uint32_t gi = row0 + r, gj = col0 + c;
                if (r >= nrows || c >= ncols || gj <= gi) continue;
                const uint64_t *A = rows + (size_t)r * stride;
                const uint64_t *B = cols + (size_t)c * stride;
                const uint32_t na = s_row_cnt[r], nb = s_col_cnt[st][c];
                const uint32_t common = warp_intersect(A, na, B, nb);
                if (lane == 0) finish_pair(p, gi, gj, A, na, B, nb, common);
            }
            __syncthreads();  // stage `st` (and rows, at the last cb) free for the next TMA
        }
    }
}

// ------------------------------------------------------------------------------------------
// generic kernel for sketches too large for the shared-memory tiles: warp per pair from global
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) prefilter_generic_kernel(const KernelParams p) {
    const uint32_t lane = threadIdx.x & 31;
    const uint64_t n_items = p.item_prefix[p.n_local_rb];
    const uint32_t warps_per_cta = blockDim.x >> 5, warp = threadIdx.x >> 5;
    for (;;) {
        unsigned long long item = 0;
        if (lane == 0) item = atomicAdd(p.work_counter, 1ull);
        item = __shfl_sync(0xffffffffu, item, 0);
        if (item >= n_items) break;
        uint32_t lo = 0, hi = p.n_local_rb;
        while (hi - lo > 1) {
            uint32_t mid = (lo + hi) >> 1;
            if (p.item_prefix[mid] <= item) lo = mid; else hi = mid;
        }
        const uint32_t rb = p.local_rb[lo];
        const uint32_t chunk = (uint32_t)(item - p.item_prefix[lo]);
        const uint32_t cb0 = rb + chunk * kColChunk;
        const uint32_t cb1 = min(cb0 + (uint32_t)kColChunk, p.n_row_blocks);
        const uint32_t row0 = rb * kRowBlock;
        for (uint32_t gi = row0; gi < min(row0 + kRowBlock, p.n); gi++) {
            for (uint32_t gj = max(gi + 1, cb0 * kColBlock); gj < min(cb1 * kColBlock, p.n); gj++) {
                const uint64_t *A = p.hashes + (size_t)gi * p.stride;
                const uint64_t *B = p.hashes + (size_t)gj * p.stride;
                const uint32_t na = p.counts[gi], nb = p.counts[gj];
                const uint32_t common = warp_intersect(A, na, B, nb);
                if (lane == 0) finish_pair(p, gi, gj, A, na, B, nb, common);
            }
        }
    }
    (void)warps_per_cta; (void)warp;
}

// ------------------------------------------------------------------------------------------
// host launcher
// ------------------------------------------------------------------------------------------
template <typename T>
int ws_ensure(T *&ptr, size_t &cap, size_t need) {
    if (need <= cap) return 0;
    if (ptr) GB_CUDA(cudaFree(ptr));
    ptr = nullptr; cap = 0;
    GB_CUDA(cudaMalloc(&ptr, need * sizeof(T)));
    cap = need;
    return 0;
}
template int ws_ensure<uint32_t>(uint32_t *&, size_t &, size_t);
template int ws_ensure<uint64_t>(uint64_t *&, size_t &, size_t);
template int ws_ensure<uint8_t>(uint8_t *&, size_t &, size_t);
template int ws_ensure<uint4>(uint4 *&, size_t &, size_t);
template int ws_ensure<unsigned long long>(unsigned long long *&, size_t &, size_t);

static size_t tiled_smem_bytes(size_t stride) {
    return (size_t)(kRowBlock + 2 * kColBlock) * stride * 8 + 3 * 8;
}

// Work list of one shard: the shard owns the kShardRows-row groups that shard_of_group() gives it
// (boustrophedon, so the triangular pair area is balanced); a row block of `block_rows` rows
// (block_rows divides kShardRows) belongs to the group it lies in.  An item is `chunk_blocks`
// consecutive column blocks starting at the row block's own index.
static int upload_work_list(PrefilterWorkspace &ws, size_t n, uint32_t block_rows, uint32_t chunk_blocks,
                            uint32_t shard, uint32_t n_shards, cudaStream_t stream, KernelParams &p,
                            bool diag_separate = false) {
    const uint32_t nrb = (uint32_t)((n + block_rows - 1) / block_rows);
    const uint32_t per_group = kShardRows / block_rows;
    std::vector<uint32_t> local;
    std::vector<uint64_t> prefix(1, 0);
    for (uint32_t rb = 0; rb < nrb; rb++) {
        if (shard_of_group(rb / per_group, n_shards) != shard) continue;
        local.push_back(rb);
        // join: the diagonal item and the adjacent item (rb, rb + 1) of every local row are
        // scheduled ahead of all other items (the kernel takes the longest items first)
        const uint32_t ahead = diag_separate ? 1u + (rb + 1 < nrb ? 1u : 0u) : 0u;
        prefix.push_back(prefix.back() + (nrb - rb - ahead + chunk_blocks - 1) / chunk_blocks);
    }
    p.n_row_blocks = nrb;
    p.n_local_rb = (uint32_t)local.size();
    p.n_diag = diag_separate ? (uint32_t)local.size() : 0;
    // every local row but the table's last block has an adjacent item: a prefix of the local rows
    p.adj_lr0 = 0;
    p.n_adj = diag_separate ? (uint32_t)local.size() - (!local.empty() && local.back() + 1 == nrb ? 1u : 0u) : 0;
    if (local.empty()) return 0;
    if (ws_ensure(ws.d_local_rb, ws.cap_local_rb, local.size())) return 2;
    if (ws_ensure(ws.d_item_prefix, ws.cap_prefix, prefix.size())) return 2;
    if (!ws.d_work_counter) GB_CUDA(cudaMalloc(&ws.d_work_counter, sizeof(unsigned long long)));
    // pageable sources: cudaMemcpyAsync stages them before returning, so the vectors may die
    GB_CUDA(cudaMemcpyAsync(ws.d_local_rb, local.data(), local.size() * 4, cudaMemcpyHostToDevice, stream));
    GB_CUDA(cudaMemcpyAsync(ws.d_item_prefix, prefix.data(), prefix.size() * 8, cudaMemcpyHostToDevice, stream));
    GB_CUDA(cudaMemsetAsync(ws.d_work_counter, 0, sizeof(unsigned long long), stream));
    p.local_rb = ws.d_local_rb; p.item_prefix = ws.d_item_prefix; p.work_counter = ws.d_work_counter;
    return 0;
}

// Launch of the pairwise (mode 1) kernel over one shard.  alt: use the workspace's second work list
// and run only if the device flag says so (fallback of a small tie-dense launch of the join).
int pairwise_launch(PrefilterWorkspace &ws, KernelParams p, uint32_t shard, uint32_t n_shards, cudaStream_t stream,
                    bool alt) {
    const size_t n = p.n, stride = p.stride;
    if (alt) {
        std::swap(ws.d_local_rb, ws.d_alt_local_rb); std::swap(ws.cap_local_rb, ws.cap_alt_local_rb);
        std::swap(ws.d_item_prefix, ws.d_alt_item_prefix); std::swap(ws.cap_prefix, ws.cap_alt_prefix);
        std::swap(ws.d_work_counter, ws.d_alt_work_counter);
    }
    int rc = upload_work_list(ws, n, kRowBlock, kColChunk, shard, n_shards, stream, p);
    if (alt) {
        std::swap(ws.d_local_rb, ws.d_alt_local_rb); std::swap(ws.cap_local_rb, ws.cap_alt_local_rb);
        std::swap(ws.d_item_prefix, ws.d_alt_item_prefix); std::swap(ws.cap_prefix, ws.cap_alt_prefix);
        std::swap(ws.d_work_counter, ws.d_alt_work_counter);
    }
    if (rc) return rc;
    if (p.n_local_rb == 0) return 0;
    p.run_if_flag = alt ? 1u : 0u;
    if (alt) p.mode_flag = ws.d_dense_flag;
    p.items = nullptr; p.n_explicit = 0;
    int dev = 0, sms = kNumSMsFallback, max_smem = 0;
    GB_CUDA(cudaGetDevice(&dev));
    GB_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    GB_CUDA(cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
    uint64_t n_items = 0;
    {
        const uint32_t per_group = kShardRows / kRowBlock;
        for (uint32_t rb = 0; rb < p.n_row_blocks; rb++)
            if (shard_of_group(rb / per_group, n_shards) == shard) n_items += (p.n_row_blocks - rb + kColChunk - 1) / kColChunk;
    }
    const size_t smem = tiled_smem_bytes(stride);
    if (smem + 1024 <= (size_t)max_smem) {
        GB_CUDA(cudaFuncSetAttribute(prefilter_tiled_kernel,
                                     cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        const uint32_t grid = (uint32_t)std::min<uint64_t>(n_items, (uint64_t)sms);
        prefilter_tiled_kernel<<<grid, kThreads, smem, stream>>>(p);
    } else {
        if (alt) return 0;  // sketches too large for the tiles: the join keeps the launch
        const uint32_t grid = (uint32_t)std::min<uint64_t>((n_items + 7) / 8, (uint64_t)sms * 8);
        prefilter_generic_kernel<<<std::max(grid, 1u), 256, 0, stream>>>(p);
    }
    GB_LAUNCH_CHECK();
    return 0;
}

int prefilter_prepare(PrefilterWorkspace &ws, const uint64_t *d_hashes, const uint32_t *d_counts, size_t n,
                      size_t stride, int k, float min_ani, uint32_t shard, uint32_t n_shards, cudaStream_t stream,
                      uint4 *d_cand, size_t cand_cap, unsigned long long *d_n_cand, KernelParams &p, int rule,
                      double rule_param, bool reset_counter) {
    if (n_shards == 0 || shard >= n_shards) { set_error("prefilter: bad shard"); return 3; }
    if (stride == 0 || (stride & 1)) { set_error("prefilter: stride must be even and > 0"); return 3; }
    if (n >= 0x7FFFFFFFull) { set_error("prefilter: n too large"); return 3; }
    if (stride >= (1ull << 30)) { set_error("prefilter: stride too large"); return 3; }
    if (reset_counter) GB_CUDA(cudaMemsetAsync(d_n_cand, 0, sizeof(unsigned long long), stream));
    if (!ws.th_valid || ws.th_s != (uint32_t)stride || ws.th_k != k || ws.th_min_ani != min_ani ||
        ws.th_rule != rule || ws.th_param != rule_param) {
        PrefilterThresholds th = rule != kRuleMashAni ? make_containment_thresholds((uint32_t)stride, rule_param, rule == kRuleContainmentBypassSmall ? kMarkerBypassBelow : 0u)
                                                          : make_thresholds((uint32_t)stride, k, min_ani);
        if (ws_ensure(ws.d_cmin_by_tmin, ws.cap_tmin, th.cmin_by_tmin.size())) return 2;
        if (ws_ensure(ws.d_cmin_by_total, ws.cap_total, th.cmin_by_total.size())) return 2;
        GB_CUDA(cudaMemcpyAsync(ws.d_cmin_by_tmin, th.cmin_by_tmin.data(), th.cmin_by_tmin.size() * 4,
                                cudaMemcpyHostToDevice, stream));
        GB_CUDA(cudaMemcpyAsync(ws.d_cmin_by_total, th.cmin_by_total.data(), th.cmin_by_total.size() * 4,
                                cudaMemcpyHostToDevice, stream));
        ws.th_s = (uint32_t)stride; ws.th_k = k; ws.th_min_ani = min_ani; ws.th_valid = true;
        ws.th_rule = rule; ws.th_param = rule_param;
        ws.th_zero_from = th.zero_fails_from();
    }
    p = KernelParams{};
    p.hashes = d_hashes; p.counts = d_counts; p.n = (uint32_t)n; p.stride = (uint32_t)stride;
    p.cmin_by_tmin = ws.d_cmin_by_tmin; p.cmin_by_total = ws.d_cmin_by_total;
    p.zero_fails_from = ws.th_zero_from;
    p.cand = d_cand; p.cand_cap = cand_cap; p.n_cand = d_n_cand;
    ws.ev_recorded = false;
    if (ws.record(0, stream)) return 2;
    return 0;
}

int prefilter_enqueue(PrefilterWorkspace &ws, const uint64_t *d_hashes, const uint32_t *d_counts,
                      size_t n, size_t stride, int k, float min_ani, uint32_t shard,
                      uint32_t n_shards, int mode, cudaStream_t stream, uint4 *d_cand,
                      size_t cand_cap, unsigned long long *d_n_cand, int rule, double rule_param) {
    if (mode != 0 && mode != 1) { set_error("prefilter: mode must be 0 or 1"); return 3; }
    KernelParams p;
    if (int rc = prefilter_prepare(ws, d_hashes, d_counts, n, stride, k, min_ani, shard, n_shards, stream, d_cand,
                                   cand_cap, d_n_cand, p, rule, rule_param))
        return rc;
    if (n < 2) return 0;
    if (mode == 0 && join_supported(stride)) return join_build_and_launch(ws, p, shard, n_shards, stream);

    if (ws.record(1, stream)) return 2;
    if (int rc = pairwise_launch(ws, p, shard, n_shards, stream, false)) return rc;
    if (ws.record(2, stream)) return 2;
    return 0;
}

int upload_join_work_list(PrefilterWorkspace &ws, size_t n, uint32_t shard, uint32_t n_shards,
                          cudaStream_t stream, KernelParams &p) {
    return upload_work_list(ws, n, kShardRows, 1, shard, n_shards, stream, p, /*diag_separate=*/true);
}

}  // namespace gb200
