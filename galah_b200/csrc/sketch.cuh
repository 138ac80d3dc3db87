// Stage 1a (K1): per-genome bottom-s MinHash sketch.  Host-side launch interface.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace gb200 {

struct SketchWorkspace {
    unsigned long long *d_work_counter = nullptr;
    uint32_t *d_redo_n = nullptr;
    uint64_t *d_chunk_off = nullptr; size_t cap_chunk_off = 0;
    uint64_t *d_thr = nullptr; size_t cap_thr = 0;
    uint32_t *d_cand_n = nullptr; size_t cap_cand_n = 0;
    uint32_t *d_has_max = nullptr; size_t cap_has_max = 0;
    uint32_t *d_redo_list = nullptr; size_t cap_redo = 0;
    uint32_t *d_item_genome = nullptr; size_t cap_items = 0;
    uint64_t *d_cand = nullptr; size_t cap_cand = 0;
    int release();
};

uint32_t sketch_set_capacity(uint32_t s);

// Optional second product of the k = 21 scan: the K3 seed selection bits (one bit per base position
// relative to first_base; bit set iff mm_hash64(canonical 15-mer) < thr), written for every word of
// every genome of the batch.  AniIndex::add_packed_device takes them instead of running its own
// mark pass over the same bytes.
struct SeedSink {
    uint32_t *d_sel;
    uint64_t first_base;
    uint64_t thr;
    // optional: seeds per genome of the batch, counted by the scan itself (zeroed by the enqueue call);
    // AniIndex::add_packed_device then needs no counting pass over the selection bits
    uint32_t *d_seed_count = nullptr;
};

// Enqueue the sketch kernel over n packed genomes (layout: see sketch.cu / galah_b200.h).
// d_hashes: n rows of out_stride uint64 (>= s), padded with 2^64-1; d_counts: n.
int sketch_enqueue(SketchWorkspace &ws, const uint32_t *d_seq2, const uint32_t *d_valid,
                   const uint64_t *d_base_off, size_t n, int k, uint32_t s, uint64_t seed,
                   uint64_t *d_hashes, uint32_t *d_counts, size_t out_stride, cudaStream_t stream,
                   const SeedSink *seeds = nullptr, const uint64_t *host_span = nullptr);
// host_span: {base_off[0], base_off[n]} when the caller knows them on the host -- the call then
// enqueues without reading them back (no stream synchronisation: batches can be pipelined).

// FracMinHash marker sketches for the skani-style screen (see sketch.cu).  cap: row stride, a power
// of two in [256, kMarkerMaxCap]; rows wider than kMarkerPartCap are finished in value-range partitions.
constexpr uint32_t kMarkerPartCap = 16384;   // markers one shared-memory sort holds
constexpr uint32_t kMarkerMaxCap = 131072;   // 26 Mbp at c = 200, 131 Mbp at c = 1000
// The row stride for units of up to `longest` bases at density 1 / c_marker (1.5 x the expectation + slack).
inline uint32_t marker_row_capacity(uint64_t longest, uint32_t c_marker) {
    uint32_t cap = 256;
    while (cap < kMarkerMaxCap && cap < 1.5 * (double)longest / c_marker + 256.0) cap <<= 1;
    return cap;
}
int marker_sketch_enqueue(SketchWorkspace &ws, const uint32_t *d_seq2, const uint32_t *d_valid,
                          const uint64_t *d_base_off, size_t n, int k, uint32_t c_marker, uint32_t cap,
                          uint64_t *d_rows, uint32_t *d_counts, cudaStream_t stream, const SeedSink *seeds = nullptr,
                          const uint64_t *host_span = nullptr);

// Synthetic genomes, generated directly in packed form on the device.
int synth_enqueue(uint64_t seed, uint64_t index_begin, size_t n, uint64_t length, uint32_t *d_seq2,
                  uint32_t *d_valid, uint64_t *d_base_off, cudaStream_t stream, uint32_t family_size = 10,
                  uint32_t rate_shift = 0);

}  // namespace gb200
