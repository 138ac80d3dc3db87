// Stage 2 (K3): FracMinHash seed-and-chain ANI on sm_100a.  Host-side interface.
//
// Replaces the per-pair `skani dist` subprocess of /root/reference/src/skani.rs:718-788.  The
// algorithm (every constant and tie-break) is specified in oracle/skani_oracle.c's header; the
// kernels are integer-only and are checked bit-exactly against that file, the final
// 100 * (sumM / sumN)^(1/15) and the two-decimal print/parse are evaluated on the host.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>
#include <vector>

namespace gb200 {

constexpr int kAniK = 15;
constexpr uint32_t kAniChunk = 20000;
constexpr int kAniBand = 2500;
constexpr int kAniMaxGap = 300;
constexpr int kAniAlpha = 20;
constexpr int kAniH = 16;
constexpr int kAniMaxOcc = 8;
constexpr int kAniMinAnchors = 3;
constexpr int kAniFxBits = 40;             // per-chunk identities are summed as 2^40 fixed point
constexpr int kIdTabN = 1024;              // device table of identities for N < kIdTabN seeds per chunk span
constexpr int kAccWords = 8;               // u32 accumulators per pair (7 used)
constexpr uint32_t kAniLearnedMinC = 70;   // skani: "learned ANI used for c >= 70 and >= 150,000 bases
constexpr uint32_t kAniLearnedMinBases = 150000;  //  aligned and not on individual contigs"

// integer accumulators of one pair (what the chain kernel emits)
struct AniPairInts {
    uint64_t sum_fx;      // sum over counted query chunks of 2^40 (M/N)^(1/15)
    uint32_t n_chunks;    // query chunks with at least one qualified chain
    uint32_t sum_m;       // matched seeds over those chunks' spans (diagnostic)
    uint32_t span_m, span_n, n_chains;  // chained anchors / seeds inside chain spans / chains
    uint32_t cov_q, cov_r;
};

struct AniPairResult {    // == galah_b200_ani_result_t
    float ani;            // what galah parses from skani's TSV column 3 (0.0 = no row)
    float af_query, af_ref;
    uint32_t estimator;   // 0 = mean of per-chunk identities, 1 = chain-span ratio (learned-ANI stand-in)
    uint64_t sum_fx;
    uint32_t n_chunks, sum_m, span_m, span_n, n_chains, cov_q, cov_r, reserved;
};

// Growable device array.
template <typename T>
struct DevVec {
    T *p = nullptr;
    size_t n = 0, cap = 0;
    int reserve(size_t need, cudaStream_t st);
    void release() { if (p) cudaFree(p); p = nullptr; n = cap = 0; }
};

struct AniScratch;

class AniIndex {
public:
    explicit AniIndex(uint32_t c) : c_(c) {}
    ~AniIndex();
    // Adds genomes from packed sequence that is already on the device (layout of K1).
    // contig_off: n+1 offsets into contig_start/contig_len (positions relative to the genome's
    // first base, ascending).  All host vectors.
    int add_packed_device(const uint32_t *d_seq2, const uint32_t *d_valid, const uint64_t *d_base_off,
                          size_t n, const std::vector<uint64_t> &base_off_host,
                          const std::vector<uint64_t> &contig_off, const std::vector<uint32_t> &contig_start,
                          const std::vector<uint32_t> &contig_len, cudaStream_t st, const uint32_t *d_sel = nullptr,
                          const uint32_t *d_seed_count = nullptr);
    // d_seed_count (with d_sel): seeds per genome already counted by the scan -- no counting pass.
    // d_sel: seed selection bits of the batch already made by the fused k = 21 scan (sketch.cuh
    // SeedSink; one bit per base relative to base_off_host[0]); the mark pass is skipped then.
    uint64_t seed_threshold() const { return ~0ull / c_; }
    int reserve_for(size_t n_total, cudaStream_t st);
    // Forgets every genome (and detaches the peers) but keeps the device allocations: a caller that
    // re-indexes the same workload per call pays no cudaMalloc / cudaFree (device-wide synchronisations).
    void clear();
    // pairs: (query, reference) genome ids -- the query is the FIRST id.  individual_contigs: the call
    // stands for `skani triangle -i` (contig clustering), where skani never applies its learned ANI.
    int pairs(const uint32_t *pairs, size_t n_pairs, float min_af_pct, bool individual_contigs, AniPairResult *out,
              cudaStream_t st);
    size_t size() const { return total_len_.size(); }
    uint32_t c() const { return c_; }
    // ---- multi-GPU (one process per GPU): the hash tables of genomes indexed by a PEER process
    // are read in place over NVLink through a CUDA IPC mapping of the peer's table array.
    // export_tables: IPC handle of this index's table array + its per-genome slot offsets / lengths.
    // attach_peer: maps a peer's array; its n genomes become usable as the REFERENCE of a pair
    // under the ids first_id .. first_id + n - 1 (returned), never as the query.
    int export_tables(cudaIpcMemHandle_t *handle, std::vector<uint64_t> &table_off, std::vector<uint64_t> &total_len) const;
    int attach_peer(const cudaIpcMemHandle_t &handle, const uint64_t *table_off, const uint64_t *total_len, size_t n,
                    uint32_t *first_id);
    // The same inside ONE process that drives several devices (galah_b200_cluster_packed_multi): the
    // peer's table array is addressed directly (peer access enabled, unified addressing), no IPC.
    int attach_peer_direct(const unsigned long long *base, const uint64_t *table_off, const uint64_t *total_len, size_t n,
                           uint32_t *first_id);
    const unsigned long long *table_base() const { return d_table_.p; }
    const std::vector<uint64_t> &table_offsets() const { return table_off_; }
    const std::vector<uint64_t> &total_lengths() const { return total_len_; }
    size_t n_peer_genomes() const { return peer_total_len_.size(); }
    // parity hooks
    int genome_info(size_t g, uint64_t *n_seeds, uint32_t *n_chunks, uint64_t *total_len) const;
    int genome_seeds(size_t g, uint32_t *ks, uint32_t *spread, uint32_t *chunk_of_seed, size_t cap,
                     cudaStream_t st) const;
    float last_chain_ms = 0.f, last_build_ms = 0.f;

private:
    uint32_t c_;
    // host copies of per-genome metadata
    std::vector<uint64_t> seed_off_{0}, cso_off_{0}, table_off_{0}, total_len_;
    std::vector<uint32_t> n_chunks_;
    // device
    DevVec<uint2> d_kq_;  // per seed (kmer << 1 | strand, spread position)
    DevVec<uint32_t> d_cso_;
    DevVec<unsigned long long> d_table_;
    DevVec<uint64_t> d_seed_off_, d_cso_off_, d_table_off_;
    DevVec<uint32_t> d_n_chunks_;
    cudaEvent_t ev_[2] = {nullptr, nullptr};
    AniScratch *scratch_ = nullptr;
    // peer tables (attach_peer): group g covers peer ids [peer_first_[g], peer_first_[g + 1])
    struct PeerGroup { const unsigned long long *base; std::vector<uint64_t> table_off; bool ipc = true; };
    std::vector<PeerGroup> peers_;
    std::vector<std::pair<cudaIpcMemHandle_t, void *>> ipc_cache_;  // open IPC mappings, kept across clear()
    std::vector<uint32_t> peer_first_{0};
    std::vector<uint64_t> peer_total_len_;
};

// strtof(sprintf("%.2f", v)), computed exactly without the text for the values an ANI can take
float print2_parse_f32(double v);
// integers -> the f32 galah would parse (host; mirrors oracle/skani_oracle.c skani_oracle_finish)
AniPairResult ani_finish(const AniPairInts &v, uint64_t len_q, uint64_t len_r, uint32_t c, bool individual_contigs,
                         float min_af_pct);
// 2^40 (m/n)^(1/15) rounded to nearest (glibc pow): one chunk's identity as a fixed-point integer
uint64_t chunk_identity_fx(uint32_t m, uint32_t n);

}  // namespace gb200
