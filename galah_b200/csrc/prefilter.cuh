// Stage 1b (K2): all-pairs sketch intersection.  Host-side launch interface.
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
    std::vector<uint32_t> cmin_by_total;  // [2 * s_max + 1]
};
PrefilterThresholds make_thresholds(uint32_t s_max, int k, float min_ani);

struct PrefilterWorkspace {
    uint32_t *d_cmin_by_tmin = nullptr;
    uint32_t *d_cmin_by_total = nullptr;
    uint64_t *d_item_prefix = nullptr;
    unsigned long long *d_work_counter = nullptr;
    size_t cap_tmin = 0, cap_total = 0, cap_prefix = 0;
    int release();
};

constexpr int kRowBlock = 8;    // rows resident per work item (== GALAH_B200_ROW_BLOCK)
constexpr int kColBlock = 8;    // columns streamed per TMA stage
constexpr int kColChunk = 16;   // column blocks per work item

// Enqueue the prefilter for one shard.  d_cand: uint4 {i, j, common, total} candidates that
// survive the conservative integer test; *d_n_cand counts them (may exceed cand_cap, in which
// case the surplus was dropped and the caller must retry with a larger buffer).
int prefilter_enqueue(PrefilterWorkspace &ws, const uint64_t *d_hashes, const uint32_t *d_counts,
                      size_t n, size_t stride, int k, float min_ani, uint32_t shard,
                      uint32_t n_shards, int mode, cudaStream_t stream, uint4 *d_cand,
                      size_t cand_cap, unsigned long long *d_n_cand);

}  // namespace gb200
