// Host side of the clustering engine: the greedy two-stage logic of
// /root/reference/src/clusterer.rs:14-487 driven by a pair list instead of BTreeMap probes.
#pragma once
#include <stdint.h>

#include <functional>
#include <string>
#include <vector>

namespace gb200 {

struct PreclusterHit {
    uint32_t i, j;  // i < j, positions in the genome slice
    float ani;      // the `Some(ani)` value stored by the preclusterer
};

// calculate_ani(fasta1 = candidate representative, fasta2 = genome under consideration)
// -> true and *ani for Some(ani), false for None  (src/lib.rs:54).
using AniFn = std::function<bool(uint32_t rep, uint32_t genome, float *ani)>;
// The same for callers that hold their ANI values in tables parallel to `hits`: `hit` is the index
// of the precluster hit the pair belongs to (the engine only ever asks about hit pairs), so the
// value is one array read instead of a search per call.
using AniByHitFn = std::function<bool(uint32_t rep, uint32_t genome, size_t hit, float *ani)>;
// Called once, after every representative is known and before the membership sweep, with the hits
// whose REVERSE orientation (query = the higher index) that sweep will ask for; non-zero aborts.
using ReversePrefetchFn = std::function<int(const std::vector<size_t> &hits)>;

// Batched, lazy stage 2: the engine works out IN WAVES which (representative, genome) pairs the
// reference's two passes end up evaluating (src/clusterer.rs:216-300, 350-449: every hit pair of a
// representative with a non-representative, and of two representatives) and asks for them a batch
// at a time -- all preclusters together, query = the representative.  A wave holds the pairs of the
// representatives confirmed by the previous one; a collection of near-identical genomes needs
// (representatives x genomes) evaluations instead of one per precluster hit.
struct AniRequest { uint32_t rep, genome; uint32_t hit; };  // hit: index of the precluster hit of the pair
// some[x] / ani[x] for every request (sized by the engine); non-zero aborts.
using AniBatchFn = std::function<int(const std::vector<AniRequest> &reqs, uint8_t *some, float *ani)>;

struct ClusterResult {
    std::vector<uint32_t> members;   // concatenated clusters, representative first
    std::vector<uint64_t> offsets;   // n_clusters + 1
    uint64_t ani_calls = 0;          // calculate_ani invocations made
    uint32_t n_preclusters = 0, largest_precluster = 0;
    uint32_t ani_waves = 0;          // batches asked of an AniBatchFn
};

// Returns 0, or non-zero with `err` set (mirrors the reference's panics).
int cluster_from_hits(size_t n_genomes, const PreclusterHit *hits, size_t n_hits, bool skip_clusterer,
                      float ani_threshold, const AniFn &calculate_ani, ClusterResult &out,
                      std::string &err, const AniByHitFn *by_hit = nullptr,
                      const ReversePrefetchFn *prefetch_reverse = nullptr, const AniBatchFn *batch = nullptr,
                      uint32_t max_waves = 16);

}  // namespace gb200
