// Stage 1b (K2): all-pairs sketch intersection.  Host-side launch interface shared by
// prefilter.cu (thresholds, pairwise kernels, dispatcher) and prefilter_join.cu (block-list
// build + block-pair join kernels).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <vector>

namespace gb200 {

// finch raw_distance + Mash ANI in f64, exactly as the reference evaluates it
// (/root/reference/src/finch.rs:78-92 on top of finch 0.6 distance()).  Host only: the GPU
// emits integers, the f64 `ln` is always evaluated by glibc so it rounds like the reference.
double mash_ani_f64(uint64_t common, uint64_t total, int k);

// Conservative integer thresholds derived from min_ani (see prefilter.cu).
struct PrefilterThresholds {
    std::vector<uint32_t> cmin_by_tmin;   // [s_max + 1]
    // smallest m0 such that cmin_by_tmin[m] >= 1 for every m >= m0: a pair of sketches with at least m0
    // hashes each and NO common hash cannot pass (lets the join skip the zero counters of a block pair)
    uint32_t zero_fails_from() const {
        uint32_t m0 = (uint32_t)cmin_by_tmin.size();
        while (m0 > 0 && cmin_by_tmin[m0 - 1] >= 1u) m0--;
        return m0;
    }
    std::vector<uint32_t> cmin_by_total;  // [2 * s_max + 1]
};
PrefilterThresholds make_thresholds(uint32_t s_max, int k, float min_ani);
// Containment screen (skani-style marker screen): a pair survives iff min(|A|,|B|) > 0 and
// common >= max(1, ceil(frac * min(|A|,|B|))); `total` is not used.
PrefilterThresholds make_containment_thresholds(uint32_t s_max, double frac, uint32_t bypass_below);
// kRuleContainment: marker containment for every pair (skani --faster-small, implied by --small-genomes);
// kRuleContainmentBypassSmall: pairs whose smaller marker sketch has < kMarkerBypassBelow entries always pass.
enum PrefilterRule { kRuleMashAni = 0, kRuleContainment = 1, kRuleContainmentBypassSmall = 2 };
constexpr uint32_t kMarkerBypassBelow = 20;

// Sharding granularity in rows == sketches per block list (== GALAH_B200_ROW_BLOCK).
constexpr int kShardRows = 128;
// Owner of row group `group` (kShardRows rows) among n_shards: boustrophedon order
// 0,1,..,G-1,G-1,..,1,0,0,1,.. so that the triangular pair area (row i has n-1-i pairs) is
// balanced to within one group per two rounds.
inline uint32_t shard_of_group(uint32_t group, uint32_t n_shards) {
    const uint32_t round = group / n_shards, pos = group % n_shards;
    return (round & 1u) ? n_shards - 1 - pos : pos;
}
// pairwise tiled merge kernel (mode 1)
constexpr int kRowBlock = 8;    // rows resident per work item (divides kShardRows)
constexpr int kColBlock = 8;    // columns streamed per TMA stage
constexpr int kColChunk = 16;   // column blocks per work item

struct PrefilterWorkspace {
    // thresholds, cached per (s_max, k, min_ani)
    uint32_t *d_cmin_by_tmin = nullptr;
    uint32_t *d_cmin_by_total = nullptr;
    size_t cap_tmin = 0, cap_total = 0;
    uint32_t th_s = 0; int th_k = 0; float th_min_ani = -1.f; bool th_valid = false;
    int th_rule = 0; double th_param = 0.0;
    uint32_t th_zero_from = 0;
    // work list of the current launch: local row blocks + item prefix + atomic work counter
    uint32_t *d_local_rb = nullptr;
    uint64_t *d_item_prefix = nullptr;
    size_t cap_local_rb = 0, cap_prefix = 0;
    unsigned long long *d_work_counter = nullptr;
    // pairwise fallback of a small tie-dense launch (join_build_and_launch): its own work list,
    // and the device flag {0 = join runs, 1 = the pairwise kernel runs} set by bl_dense_stat_kernel
    uint32_t *d_alt_local_rb = nullptr;
    uint64_t *d_alt_item_prefix = nullptr;
    size_t cap_alt_local_rb = 0, cap_alt_prefix = 0;
    unsigned long long *d_alt_work_counter = nullptr;
    uint32_t *d_dense_flag = nullptr;  // [0] flag, [1] equal-at-distance-16 count, [2] entries looked at
    // streamed host-buffer path: one work counter per wave, one "slice resident" event per slice
    unsigned long long *d_wave_counters = nullptr;
    size_t cap_wave_counters = 0;
    unsigned long long *d_item_counters = nullptr;  // pool of work counters for explicit-item launches
    uint32_t item_counter_next = 0;
    std::vector<cudaEvent_t> chunk_ev;
    // block lists (mode 0): two ping-pong buffers of n_blocks * kShardRows * stride entries
    uint64_t *d_bl_vals[2] = {nullptr, nullptr};
    uint8_t *d_bl_tags[2] = {nullptr, nullptr};
    uint32_t *d_bl_len = nullptr;
    unsigned long long *d_gmax = nullptr;  // largest valid hash of the table (device scalar)
    uint32_t *d_splits = nullptr;          // merge-path splits of the tile boundaries of one level
    size_t cap_splits = 0;
    size_t cap_bl = 0, cap_bl_len = 0;
    // finished lists of the single-device path (structure of arrays)
    uint32_t *d_fin_hi = nullptr, *d_fin_lo = nullptr;
    uint8_t *d_fin_tags = nullptr;
    size_t cap_fin = 0;
    // CUDA events on the launch stream: [0] before the build kernels, [1] before the main
    // (join / pairwise) kernel, [2] after it.  Read back with last_timing() after a sync.
    cudaEvent_t ev[3] = {nullptr, nullptr, nullptr};
    bool ev_recorded = false;
    int record(int which, cudaStream_t stream);
    int last_timing(float *build_ms, float *main_ms);
    int release();
};

// Shared by both modes: kernel-side view of one launch.
struct KernelParams {
    const uint64_t *hashes;
    const uint32_t *counts;
    uint32_t n, stride;
    uint32_t n_row_blocks;        // number of row blocks (of the mode's block size), global
    uint32_t n_local_rb;          // row blocks owned by this shard
    const uint32_t *local_rb;     // [n_local_rb] global row-block index of each local one
    const uint64_t *item_prefix;  // [n_local_rb + 1]
    unsigned long long *work_counter;
    const uint32_t *cmin_by_tmin;
    const uint32_t *cmin_by_total;
    uint32_t zero_fails_from;     // see PrefilterThresholds::zero_fails_from
    uint4 *cand;
    unsigned long long cand_cap;
    unsigned long long *n_cand;
    // mode 0 only
    const uint32_t *bl_hi, *bl_lo;  // block lists, structure of arrays: (value << sh) split in two
    const uint8_t *bl_tags;
    const uint32_t *bl_len;
    uint64_t bl_cap;  // entries per block list (kShardRows * stride)
    uint32_t n_diag;  // join: diagonal items, taken first; they belong to the LAST n_diag local rows
    const uint32_t *items;    // join: explicit work list of (rb, cb) pairs, rb <= cb, longest first; or null
    uint32_t n_explicit;      // number of explicit items
    uint32_t n_adj, adj_lr0;  // join: items (rb, rb + 1), taken next; local rows [adj_lr0, adj_lr0 + n_adj)
    unsigned long long *dbg_buf;  // per-item {start ns, end ns, sm, rb << 32 | cb} log (debug), or null
    uint32_t cb_lo;   // first column block of the launch's window (0 unless a streamed wave)
    const uint32_t *mode_flag;  // or null; join kernel leaves at once if *mode_flag != 0, the pairwise
    uint32_t run_if_flag;       //          kernel launched as a fallback (run_if_flag = 1) if it is 0
};

// Enqueue the prefilter for one shard.  d_cand: uint4 {i, j, common, total} candidates that
// survive the conservative integer test; *d_n_cand counts them (may exceed cand_cap, in which
// case the surplus was dropped and the caller must retry with a larger buffer).
// mode 0 = block-list join (default), mode 1 = pairwise warp merge of every pair.
int prefilter_enqueue(PrefilterWorkspace &ws, const uint64_t *d_hashes, const uint32_t *d_counts,
                      size_t n, size_t stride, int k, float min_ani, uint32_t shard,
                      uint32_t n_shards, int mode, cudaStream_t stream, uint4 *d_cand,
                      size_t cand_cap, unsigned long long *d_n_cand, int rule = kRuleMashAni,
                      double rule_param = 0.0);

// prefilter_join.cu
int join_build_and_launch(PrefilterWorkspace &ws, KernelParams &p, uint32_t shard, uint32_t n_shards,
                          cudaStream_t stream);
bool join_supported(size_t stride);
void blocklist_layout(size_t n, size_t stride, size_t *n_blocks, size_t *entries_per_block, size_t *slack);
int blocklist_build(PrefilterWorkspace &ws, const uint64_t *d_hashes, const uint32_t *d_counts, size_t n,
                    size_t stride, uint32_t b0, uint32_t b1, uint32_t *d_hi, uint32_t *d_lo, uint8_t *d_tags,
                    uint32_t *d_len, cudaStream_t stream, bool gmax_ready = false);
int table_max_enqueue(const uint64_t *d_hashes, const uint32_t *d_counts, size_t n, size_t stride,
                      unsigned long long *d_max, cudaStream_t stream);
int blocklist_build_local(PrefilterWorkspace &ws, const uint64_t *d_rows, const uint32_t *d_counts, size_t n_rows,
                          size_t stride, const unsigned long long *d_gmax, uint32_t *d_hi, uint32_t *d_lo,
                          uint8_t *d_tags, uint32_t *d_len, uint32_t n_blocks_out, cudaStream_t stream);
int join_streamed_from_host(PrefilterWorkspace &ws, KernelParams &p, const uint64_t *h_hashes,
                            const uint32_t *h_counts, uint64_t *d_table, cudaStream_t compute, cudaStream_t copy,
                            int chunks, double wave_frac);
int join_launch(PrefilterWorkspace &ws, KernelParams &p, const uint32_t *d_hi, const uint32_t *d_lo,
                const uint8_t *d_tags, const uint32_t *d_len, uint32_t shard, uint32_t n_shards,
                cudaStream_t stream);
// prefilter.cu: fills the thresholds / common fields of p for a launch (shared by both entry paths)
int prefilter_prepare(PrefilterWorkspace &ws, const uint64_t *d_hashes, const uint32_t *d_counts, size_t n,
                      size_t stride, int k, float min_ani, uint32_t shard, uint32_t n_shards, cudaStream_t stream,
                      uint4 *d_cand, size_t cand_cap, unsigned long long *d_n_cand, KernelParams &p,
                      int rule = kRuleMashAni, double rule_param = 0.0, bool reset_counter = true);
// prefilter_join.cu: join of an explicit list of block pairs (device array of (rb, cb), rb <= cb)
int join_launch_items(PrefilterWorkspace &ws, KernelParams &p, const uint32_t *d_hi, const uint32_t *d_lo,
                      const uint8_t *d_tags, const uint32_t *d_len, const uint32_t *d_items, size_t n_items,
                      cudaStream_t stream);
// prefilter.cu: work list with one item per (local row block, column block >= it), block = kShardRows
int pairwise_launch(PrefilterWorkspace &ws, KernelParams p, uint32_t shard, uint32_t n_shards, cudaStream_t stream,
                    bool alt);
int upload_join_work_list(PrefilterWorkspace &ws, size_t n, uint32_t shard, uint32_t n_shards,
                          cudaStream_t stream, KernelParams &p);

template <typename T>
int ws_ensure(T *&ptr, size_t &cap, size_t need);

}  // namespace gb200
