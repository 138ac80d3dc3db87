// extern "C" boundary of libgalah_b200.so (declared in include/galah_b200.h).
#include "../../include/galah_b200.h"

#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/stat.h>

#include <algorithm>
#include <chrono>
#include <atomic>
#include <memory>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "ani.cuh"
#include "common.cuh"
#include "host/cluster_engine.hpp"
#include "host/fasta.hpp"
#include "ingest.cuh"
#include "prefilter.cuh"
#include "sketch.cuh"

namespace gb200 {

static thread_local std::string t_error;
std::atomic<uint64_t> g_launch_count{0};

void set_error(const std::string &msg) { t_error = msg; }
int fail_cuda(cudaError_t e, const char *what, const char *file, int line) {
    t_error = std::string("CUDA error: ") + cudaGetErrorString(e) + " in " + what + " (" + file + ":" +
              std::to_string(line) + ")";
    return GALAH_B200_ERR_CUDA;
}

struct Context {
    int device = -1;
    cudaStream_t stream = nullptr;
    int prefilter_mode = 0;
    PrefilterWorkspace pws;
    SketchWorkspace sws;
    FastaDecoder fasta;
    uint8_t *h_raw = nullptr; size_t cap_raw = 0;  // pinned staging of raw file bytes for K0
    int device_ingest = 1;               // K0: decode FASTA bytes on the device (0 = host packer)
    // buffers re-used across host-pointer prefilter calls (cudaMalloc is a device-wide sync)
    uint64_t *d_table = nullptr; size_t cap_table = 0;
    uint32_t *d_counts = nullptr; size_t cap_counts = 0;
    uint4 *d_cand = nullptr; size_t cap_cand = 0;
    unsigned long long *d_n_cand = nullptr;
    cudaStream_t copy_stream = nullptr;  // PCIe uploads of the streamed host-buffer prefilter
    double stream_wave_frac = 0.0;       // early join wave after this fraction of the slices (0 = none)
    int stream_chunks = 8;               // slices of that upload (<= 1 disables the pipeline)
    uint4 *h_stage = nullptr;            // pinned landing zone: candidate count + first kStageCand candidates
    float host_ms[4] = {0, 0, 0, 0};     // last host-buffer prefilter: enqueue, wait, d2h, finish
    uint4 *h_cand_map = nullptr;         // zero-copy candidate list of the streamed call (mapped pinned, all-zero between calls)
    size_t cap_cand_map = 0;
    cudaEvent_t ev_done = nullptr;
    uint32_t *slot_seq2[2] = {nullptr, nullptr}, *slot_valid[2] = {nullptr, nullptr};  // upload slots of ingest_packed
    uint64_t *slot_off[2] = {nullptr, nullptr}, *slot_len[2] = {nullptr, nullptr}, *slot_ranges[2] = {nullptr, nullptr};
    size_t cap_slot_seq2[2] = {0, 0}, cap_slot_valid[2] = {0, 0}, cap_slot_off[2] = {0, 0}, cap_slot_len[2] = {0, 0},
           cap_slot_ranges[2] = {0, 0};
    uint32_t *d_sel = nullptr; size_t cap_sel = 0;  // K3 seed selection bits made by the fused k = 21 scan
    uint32_t *d_sel2 = nullptr; size_t cap_sel2 = 0;  // second buffer: batch b + 1 is scanned while batch b is indexed
    uint32_t *d_seedcnt[2] = {nullptr, nullptr}; size_t cap_seedcnt[2] = {0, 0};  // seeds per genome, counted by the scan
    cudaStream_t scan_stream = nullptr;  // ingest_packed: K1 scans of the next batch run beside the K3 index build
    cudaEvent_t ev_scan[2] = {nullptr, nullptr};
    AniIndex *pipe_index[2] = {nullptr, nullptr};  // K3 index of the one-call pipelines (c = 125 / 30), re-used across calls
    AniIndex &pipeline_index(bool small_genomes) {
        AniIndex *&p = pipe_index[small_genomes ? 1 : 0];
        if (!p) p = new AniIndex(small_genomes ? 30u : 125u);
        p->clear();
        return *p;
    }
    int release() {
        for (auto &p : pipe_index) { delete p; p = nullptr; }
        cudaFree(d_sel); d_sel = nullptr; cap_sel = 0;
        cudaFree(d_sel2); d_sel2 = nullptr; cap_sel2 = 0;
        for (int x = 0; x < 2; x++) { cudaFree(d_seedcnt[x]); d_seedcnt[x] = nullptr; cap_seedcnt[x] = 0; }
        if (scan_stream) cudaStreamDestroy(scan_stream);
        scan_stream = nullptr;
        for (auto &e : ev_scan) { if (e) cudaEventDestroy(e); e = nullptr; }
        for (int x = 0; x < 2; x++) {
            cudaFree(slot_seq2[x]); cudaFree(slot_valid[x]); cudaFree(slot_off[x]); cudaFree(slot_len[x]); cudaFree(slot_ranges[x]);
            slot_seq2[x] = slot_valid[x] = nullptr; slot_off[x] = slot_len[x] = slot_ranges[x] = nullptr;
            cap_slot_seq2[x] = cap_slot_valid[x] = cap_slot_off[x] = cap_slot_len[x] = cap_slot_ranges[x] = 0;
        }
        pws.release(); sws.release(); fasta.release();
        if (h_raw) cudaFreeHost(h_raw);
        h_raw = nullptr; cap_raw = 0;
        cudaFree(d_table); cudaFree(d_counts); cudaFree(d_cand); cudaFree(d_n_cand);
        if (h_stage) cudaFreeHost(h_stage);
        if (h_cand_map) cudaFreeHost(h_cand_map);
        if (ev_done) cudaEventDestroy(ev_done);
        h_cand_map = nullptr; cap_cand_map = 0; ev_done = nullptr;
        if (copy_stream) cudaStreamDestroy(copy_stream);
        d_table = nullptr; d_counts = nullptr; d_cand = nullptr; d_n_cand = nullptr;
        h_stage = nullptr; copy_stream = nullptr;
        cap_table = cap_counts = cap_cand = 0;
        return 0;
    }
};
// One context (stream, workspaces, pipeline index) and one lock PER DEVICE.  A thread works on the
// process's primary device (the one galah_b200_init bound last) unless it is one of the per-device
// workers of a multi-device call (galah_b200_cluster_packed_multi), which pin themselves to theirs.
constexpr int kMaxDevices = 16;
static Context g_ctxs[kMaxDevices];
static std::mutex g_mus[kMaxDevices];
static std::atomic<int> g_primary{0};
static thread_local int t_dev = -1;
static inline int cur_dev() { return t_dev >= 0 ? t_dev : g_primary.load(); }
#define g_ctx (::gb200::g_ctxs[::gb200::cur_dev()])
#define g_mu (::gb200::g_mus[::gb200::cur_dev()])

static int require_ctx() {
    if (g_ctx.device < 0) {
        set_error("galah_b200: no device bound -- call galah_b200_init(device) first "
                  "(there is no CPU fallback)");
        return GALAH_B200_ERR_NO_DEVICE;
    }
    GB_CUDA(cudaSetDevice(g_ctx.device));
    return 0;
}

template <typename T>
struct DevBuf {
    T *p = nullptr;
    ~DevBuf() { if (p) cudaFree(p); }
    int alloc(size_t n) {
        if (p) { cudaFree(p); p = nullptr; }
        GB_CUDA(cudaMalloc(&p, std::max<size_t>(n, 1) * sizeof(T)));
        return 0;
    }
};

// Seed sink of the fused k = 21 scan for a batch whose bases span [first, end): the selection bits
// live in a context buffer until AniIndex::add_packed_device has consumed them (same stream).
static int seed_sink_for(const AniIndex &index, uint64_t first, uint64_t end, SeedSink &sink) {
    if (ws_ensure(g_ctx.d_sel, g_ctx.cap_sel, (size_t)((end - first) / 32 + 2))) return GALAH_B200_ERR_CUDA;
    sink.d_sel = g_ctx.d_sel; sink.first_base = first; sink.thr = index.seed_threshold();
    return 0;
}

// Integer candidates {i, j, common, total} -> the reference's f64 formula, threshold and f32 store
// (src/finch.rs:78-93), sorted by (i, j) -- the iteration order of the reference's BTreeMap.  Host
// only.  The order comes from a counting sort on i (candidates are sparse: a few per row) and a
// small sort by j inside each row, O(candidates + rows) instead of a comparison sort of all.
struct CandFinisher {
    int k;
    double thr;
    std::vector<uint4> c;
    std::vector<float> ani;
    std::vector<uint8_t> keep;
    CandFinisher(int k_, float min_ani) : k(k_), thr((double)min_ani) {}
    void reserve(size_t n) { c.reserve(n); ani.reserve(n); keep.reserve(n); }
    void clear() { c.clear(); ani.clear(); keep.clear(); }
    void add(const uint4 &x) {  // the f64 `ln` dominates the host finish (~13 ns per candidate)
        const double a = mash_ani_f64(x.z, x.w, k);
        c.push_back(x); ani.push_back((float)a); keep.push_back(a >= thr ? 1 : 0);
    }
    void set(size_t q, const uint4 &x) {
        const double a = mash_ani_f64(x.z, x.w, k);
        c[q] = x; ani[q] = (float)a; keep[q] = a >= thr ? 1 : 0;
    }
    int finalize(galah_b200_pair_t **out, size_t *n_out) const {
        const size_t n_cand = c.size();
        uint32_t max_i = 0;
        for (size_t x = 0; x < n_cand; x++) max_i = std::max(max_i, c[x].x);
        std::vector<uint32_t> row_start((size_t)max_i + 2, 0);
        for (size_t x = 0; x < n_cand; x++)
            if (keep[x]) row_start[c[x].x + 1]++;
        for (size_t r = 0; r + 1 < row_start.size(); r++) row_start[r + 1] += row_start[r];
        const size_t n_pass = n_cand ? row_start.back() : 0;
        galah_b200_pair_t *res = (galah_b200_pair_t *)malloc(std::max<size_t>(n_pass, 1) * sizeof(galah_b200_pair_t));
        if (!res) { set_error("out of host memory"); return GALAH_B200_ERR_ARG; }
        std::vector<uint32_t> fill(row_start.begin(), row_start.end());
        for (size_t x = 0; x < n_cand; x++)
            if (keep[x]) res[fill[c[x].x]++] = galah_b200_pair_t{c[x].x, c[x].y, c[x].z, c[x].w, ani[x]};
        auto by_j = [](const galah_b200_pair_t &a, const galah_b200_pair_t &b) { return a.j < b.j; };
        // rows arrive in the order the kernel's CTAs appended them: sort each by j (rows are independent:
        // long lists -- dense collections -- are split over the host threads, 64 rows at a time)
        const size_t n_rows = row_start.size() - 1;
        std::atomic<size_t> next_row{0};
        auto sort_rows = [&] {
            for (;;) {
                const size_t r0 = next_row.fetch_add(64);
                if (r0 >= n_rows) break;
                for (size_t r = r0; r < std::min(n_rows, r0 + 64); r++) {
                    galah_b200_pair_t *b = res + row_start[r], *e = res + row_start[r + 1];
                    if (e - b > 1 && !std::is_sorted(b, e, by_j)) std::sort(b, e, by_j);
                }
            }
        };
        const size_t hw = std::max<size_t>(1, std::thread::hardware_concurrency());
        const size_t nt = n_pass >= (1u << 17) ? std::min<size_t>(hw, 16) : 1;
        std::vector<std::thread> th;
        for (size_t t = 1; t < nt; t++) th.emplace_back(sort_rows);
        sort_rows();
        for (auto &t : th) t.join();
        *out = res; *n_out = n_pass;
        return 0;
    }
};

static int finish_candidates(const uint4 *cand, size_t n_cand, int k, float min_ani, galah_b200_pair_t **out,
                             size_t *n_out) {
    CandFinisher fin(k, min_ani);
    const size_t hw = std::max<size_t>(1, std::thread::hardware_concurrency());
    const size_t nt = n_cand >= 32768 ? std::min<size_t>(hw, 16) : 1;
    if (nt == 1) {
        fin.reserve(n_cand);
        for (size_t x = 0; x < n_cand; x++) fin.add(cand[x]);
    } else {
        // the f64 logarithm per candidate is the cost (13 ns each): large lists are split over the host threads
        fin.c.resize(n_cand); fin.ani.resize(n_cand); fin.keep.resize(n_cand);
        std::vector<std::thread> th;
        const size_t per = (n_cand + nt - 1) / nt;
        for (size_t t = 0; t < nt; t++) {
            const size_t x0 = t * per, x1 = std::min(n_cand, x0 + per);
            if (x0 < x1) th.emplace_back([&fin, cand, x0, x1] { for (size_t x = x0; x < x1; x++) fin.set(x, cand[x]); });
        }
        for (auto &t : th) t.join();
    }
    return fin.finalize(out, n_out);
}

constexpr size_t kStageCand = 1 << 16;  // candidates fetched together with their count (1 MiB, pinned)
static double now_ms() {
    return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

// Candidate count + candidates -> host, one synchronisation in the common case: the count and
// the first kStageCand candidates land in pinned memory together; a longer list is fetched in a
// second copy.  Returns got > cap (overflow: the caller re-runs with a larger buffer) via *overflow.
static int fetch_candidates(cudaStream_t stream, std::vector<uint4> &cand, unsigned long long *got_out) {
    if (!g_ctx.h_stage) GB_CUDA(cudaMallocHost(&g_ctx.h_stage, (kStageCand + 1) * sizeof(uint4)));
    const size_t first = std::min(kStageCand, g_ctx.cap_cand);
    GB_CUDA(cudaMemcpyAsync(g_ctx.h_stage, g_ctx.d_n_cand, sizeof(unsigned long long), cudaMemcpyDeviceToHost, stream));
    GB_CUDA(cudaMemcpyAsync(g_ctx.h_stage + 1, g_ctx.d_cand, first * sizeof(uint4), cudaMemcpyDeviceToHost, stream));
    const double t0 = now_ms();
    GB_CUDA(cudaStreamSynchronize(stream));
    g_ctx.host_ms[1] = (float)(now_ms() - t0);
    const unsigned long long got = *reinterpret_cast<unsigned long long *>(g_ctx.h_stage);
    *got_out = got;
    if (got > g_ctx.cap_cand) return 0;
    cand.resize((size_t)got);
    memcpy(cand.data(), g_ctx.h_stage + 1, std::min<size_t>(got, first) * sizeof(uint4));
    if (got > first) {
        GB_CUDA(cudaMemcpyAsync(cand.data() + first, g_ctx.d_cand + first, ((size_t)got - first) * sizeof(uint4),
                                cudaMemcpyDeviceToHost, stream));
        GB_CUDA(cudaStreamSynchronize(stream));
    }
    g_ctx.host_ms[2] = (float)(now_ms() - t0) - g_ctx.host_ms[1];
    return 0;
}

// Run the prefilter kernels for one shard and finish the survivors on the host in f64
// (reference: src/finch.rs:78-93).  h_hashes != nullptr: the table is still on the host and
// d_hashes is its (not yet filled) device copy -- the first attempt streams it (upload pipelined
// against build + join, join_streamed_from_host); a retry after candidate overflow finds it resident.
static int run_prefilter(const uint64_t *d_hashes, const uint32_t *d_counts, size_t n, size_t stride,
                         int k, float min_ani, uint32_t shard, uint32_t n_shards, cudaStream_t stream,
                         galah_b200_pair_t **out, size_t *n_out, const uint64_t *h_hashes = nullptr,
                         const uint32_t *h_counts = nullptr) {
    *out = nullptr; *n_out = 0;
    if ((reinterpret_cast<uintptr_t>(d_hashes) & 15) != 0) {
        set_error("prefilter: sketch table must be 16-byte aligned");
        return GALAH_B200_ERR_ARG;
    }
    const double t_begin = now_ms();
    if (!g_ctx.d_n_cand) GB_CUDA(cudaMalloc(&g_ctx.d_n_cand, sizeof(unsigned long long)));
    // room for 32 survivors per genome, or for EVERY pair when that is at most 4 M entries (64 MB): a small
    // collection of near-identical genomes passes all its pairs and would otherwise run the kernels twice
    const size_t all_pairs = n * (n - (n ? 1 : 0)) / 2;
    size_t cap = std::max<size_t>(1 << 16, 32 * n);
    if (!h_hashes && n_shards <= 1 && all_pairs <= ((size_t)1 << 22)) cap = std::max<size_t>(1 << 16, all_pairs);
    std::vector<uint4> cand;
    for (;;) {
        if (ws_ensure(g_ctx.d_cand, g_ctx.cap_cand, cap)) return GALAH_B200_ERR_CUDA;
        if (h_hashes) {
            // Streamed call: the join appends its survivors straight into MAPPED pinned host
            // memory (one 16-byte store each, so a slot is either all zero or complete; j >= 1
            // marks it written) and this thread evaluates the f64 formula of every candidate
            // while the kernels are still running -- the host finish costs no wall time.
            if (g_ctx.cap_cand_map < cap) {
                if (g_ctx.h_cand_map) GB_CUDA(cudaFreeHost(g_ctx.h_cand_map));
                g_ctx.h_cand_map = nullptr; g_ctx.cap_cand_map = 0;
                GB_CUDA(cudaHostAlloc(&g_ctx.h_cand_map, cap * sizeof(uint4), cudaHostAllocMapped));
                memset(g_ctx.h_cand_map, 0, cap * sizeof(uint4));
                g_ctx.cap_cand_map = cap;
            }
            if (!g_ctx.ev_done) GB_CUDA(cudaEventCreateWithFlags(&g_ctx.ev_done, cudaEventDisableTiming));
            uint4 *d_map = nullptr;
            GB_CUDA(cudaHostGetDevicePointer(&d_map, g_ctx.h_cand_map, 0));
            KernelParams p;
            if (int rc = prefilter_prepare(g_ctx.pws, d_hashes, d_counts, n, stride, k, min_ani, 0, 1, stream, d_map,
                                           g_ctx.cap_cand_map, g_ctx.d_n_cand, p))
                return rc;
            if (int rc = join_streamed_from_host(g_ctx.pws, p, h_hashes, h_counts, const_cast<uint64_t *>(d_hashes),
                                                 stream, g_ctx.copy_stream, g_ctx.stream_chunks,
                                                 g_ctx.stream_wave_frac))
                return rc;
            h_hashes = nullptr;
            if (!g_ctx.h_stage) GB_CUDA(cudaMallocHost(&g_ctx.h_stage, (kStageCand + 1) * sizeof(uint4)));
            GB_CUDA(cudaMemcpyAsync(g_ctx.h_stage, g_ctx.d_n_cand, sizeof(unsigned long long), cudaMemcpyDeviceToHost, stream));
            GB_CUDA(cudaEventRecord(g_ctx.ev_done, stream));
            g_ctx.host_ms[0] = (float)(now_ms() - t_begin);
            const double t_wait = now_ms();
            CandFinisher fin(k, min_ani);
            fin.reserve(1 << 16);
            const volatile uint4 *hc = g_ctx.h_cand_map;
            size_t next = 0;
            auto drain = [&](size_t limit) {
                while (next < limit && hc[next].y != 0) {
                    fin.add(make_uint4(hc[next].x, hc[next].y, hc[next].z, hc[next].w));
                    next++;
                }
            };
            for (;;) {
                const cudaError_t q = cudaEventQuery(g_ctx.ev_done);
                if (q != cudaSuccess && q != cudaErrorNotReady) return fail_cuda(q, "cudaEventQuery", __FILE__, __LINE__);
                drain(g_ctx.cap_cand_map);
                if (q == cudaSuccess) break;
            }
            const unsigned long long got = *reinterpret_cast<const volatile unsigned long long *>(g_ctx.h_stage);
            g_ctx.host_ms[1] = (float)(now_ms() - t_wait);
            const size_t written = (size_t)std::min<unsigned long long>(got, g_ctx.cap_cand_map);
            if (got <= g_ctx.cap_cand_map) {
                // entries consumed while the kernels were running are compared once more with the
                // final memory: a 16-byte device store reaches host memory in one piece on every
                // platform this runs on, but nothing written down guarantees it
                for (size_t q = 0; q < next; q++) {
                    const uint4 now = make_uint4(hc[q].x, hc[q].y, hc[q].z, hc[q].w);
                    if (memcmp(&now, &fin.c[q], sizeof(uint4)) != 0) fin.set(q, now);
                }
                drain(written);  // everything is in place once the kernels have completed
                if (next != written) {
                    memset(g_ctx.h_cand_map, 0, g_ctx.cap_cand_map * sizeof(uint4));
                    set_error("prefilter: candidate list in mapped memory is incomplete");
                    return GALAH_B200_ERR_CUDA;
                }
            }
            memset(g_ctx.h_cand_map, 0, written * sizeof(uint4));  // all-zero again for the next call
            g_ctx.host_ms[2] = 0.f;
            if (got > g_ctx.cap_cand_map) { cap = (size_t)got; continue; }  // overflow: re-run on the resident table
            const double t0 = now_ms();
            const int rc = fin.finalize(out, n_out);
            g_ctx.host_ms[3] = (float)(now_ms() - t0);
            return rc;
        } else {
            int rc = prefilter_enqueue(g_ctx.pws, d_hashes, d_counts, n, stride, k, min_ani, shard, n_shards,
                                       g_ctx.prefilter_mode, stream, g_ctx.d_cand, g_ctx.cap_cand, g_ctx.d_n_cand);
            if (rc) return rc;
        }
        g_ctx.host_ms[0] = (float)(now_ms() - t_begin);
        unsigned long long got = 0;
        if (int rc = fetch_candidates(stream, cand, &got)) return rc;
        if (got > g_ctx.cap_cand) { cap = (size_t)got; continue; }
        break;
    }
    const double t0 = now_ms();
    int rc = finish_candidates(cand.data(), cand.size(), k, min_ani, out, n_out);
    g_ctx.host_ms[3] = (float)(now_ms() - t0);
    return rc;
}

// Raw bytes of a batch of files laid out for K0: every file starts at a multiple of 32.
struct RawBatch {
    uint8_t *bytes = nullptr;  // pinned staging buffer owned by the context (re-used across batches)
    std::vector<uint64_t> file_off, file_len, first_byte;
    bool all_fasta = true;  // every non-empty file starts (after blank lines) with '>'
};

static int stage_raw(const std::vector<std::vector<uint8_t>> &files, RawBatch &rb) {
    const size_t n = files.size();
    rb.file_off.assign(n + 1, 0); rb.file_len.assign(n, 0); rb.first_byte.assign(n, 0);
    for (size_t f = 0; f < n; f++) {
        rb.file_len[f] = files[f].size();
        rb.file_off[f + 1] = (rb.file_off[f] + files[f].size() + 31) / 32 * 32;
    }
    const size_t need = rb.file_off[n] + 64;
    if (g_ctx.cap_raw < need) {
        if (g_ctx.h_raw) GB_CUDA(cudaFreeHost(g_ctx.h_raw));
        g_ctx.h_raw = nullptr; g_ctx.cap_raw = 0;
        const size_t want = need + need / 4;
        GB_CUDA(cudaMallocHost(&g_ctx.h_raw, want));
        g_ctx.cap_raw = want;
    }
    rb.bytes = g_ctx.h_raw;
    rb.all_fasta = true;
    for (size_t f = 0; f < n; f++) {
        if (!files[f].empty()) memcpy(rb.bytes + rb.file_off[f], files[f].data(), files[f].size());
        memset(rb.bytes + rb.file_off[f] + files[f].size(), '\n', rb.file_off[f + 1] - rb.file_off[f] - files[f].size());
        size_t p = 0;
        while (p < files[f].size() && (files[f][p] == '\n' || files[f][p] == '\r')) p++;
        rb.first_byte[f] = rb.file_off[f] + p;
        if (p < files[f].size() && files[f][p] != '>') rb.all_fasta = false;
    }
    return 0;
}

// One pass over FASTA files feeding K1 (sketches) and/or the K3 index from the SAME upload:
// files are parsed and packed on host threads, a batch is concatenated (genome offsets multiples
// of 128), copied to the device once, and handed to the requested sinks.
// Marker sketches of all units, resident on the device in the K2 table layout (common row stride):
// the file-based skani preclusterer fills it batch by batch and hands it to the screen as it is.
struct MarkerTable {
    uint64_t *d_rows = nullptr;
    uint32_t *d_counts = nullptr;
    size_t n = 0, cap_rows = 0;
    uint32_t stride = 0;
    ~MarkerTable() { if (d_rows) cudaFree(d_rows); if (d_counts) cudaFree(d_counts); }
    // room for `rows` rows of `new_stride` (>= stride) hashes; existing rows keep their content
    int reserve(size_t rows, uint32_t new_stride, cudaStream_t st) {
        new_stride = std::max(new_stride, stride);
        if (rows <= cap_rows && new_stride == stride) return 0;
        const size_t new_cap = std::max(rows, rows <= cap_rows ? cap_rows : std::max(rows, 2 * cap_rows));
        uint64_t *nr = nullptr;
        uint32_t *nc = nullptr;
        GB_CUDA(cudaMalloc(&nr, std::max<size_t>(new_cap * (size_t)new_stride, 1) * 8));
        GB_CUDA(cudaMalloc(&nc, std::max<size_t>(new_cap, 1) * 4));
        if (n) {
            GB_CUDA(cudaMemcpy2DAsync(nr, (size_t)new_stride * 8, d_rows, (size_t)stride * 8, (size_t)stride * 8, n,
                                      cudaMemcpyDeviceToDevice, st));
            GB_CUDA(cudaMemcpyAsync(nc, d_counts, n * 4, cudaMemcpyDeviceToDevice, st));
        }
        GB_CUDA(cudaStreamSynchronize(st));
        if (d_rows) GB_CUDA(cudaFree(d_rows));
        if (d_counts) GB_CUDA(cudaFree(d_counts));
        d_rows = nr; d_counts = nc; cap_rows = new_cap; stride = new_stride;
        return 0;
    }
};

struct IngestSinks {
    bool sketch = false;
    int k = 21; uint32_t s = 1000; uint64_t seed = 0;
    uint32_t stride = 0;  // row stride of d_hashes (0: s)
    uint64_t *hashes = nullptr; uint32_t *counts = nullptr;  // host rows, stride s
    uint64_t *d_hashes = nullptr; uint32_t *d_counts = nullptr;  // or DEVICE rows, stride s: no host round trip
    AniIndex *ani = nullptr;
    // FracMinHash marker sketches (k = 21, density 1/c_marker): one ascending hash list per unit
    MarkerTable *markers = nullptr;
    uint32_t c_marker = 1000;
    bool per_record = false;  // contig mode: every FASTA record is its own unit
    size_t n_units = 0;       // out: genomes (or records) ingested
};

static int ingest_files(const char *const *paths, size_t n, int host_threads, IngestSinks &sinks) {
    if (host_threads <= 0) host_threads = (int)std::max(1u, std::thread::hardware_concurrency());
    const size_t kMaxBatchGenomes = 512;
    const uint64_t kMaxBatchBases = 2ull << 30;
    cudaStream_t st = g_ctx.stream;
    size_t done = 0;

    // Hands one batch that is resident on the device in the K1 / K3 layout to the requested sinks.
    // unit_first: index of the batch's first unit among all units ingested so far.
    auto feed = [&](const uint32_t *d_seq2, const uint32_t *d_valid, const uint64_t *d_off, size_t nb,
                    const std::vector<uint64_t> &base_off, const std::vector<uint64_t> &contig_off,
                    const std::vector<uint32_t> &cs, const std::vector<uint32_t> &cl, uint64_t longest,
                    size_t unit_first, size_t files_done) -> int {
        // one pass over the packed bases feeds the k = 21 sketch AND the K3 seed selection
        SeedSink seed_sink{nullptr, 0, 0};
        const bool fuse = sinks.ani && ((sinks.sketch && sinks.k == 21) || sinks.markers);
        if (fuse) if (int rc = seed_sink_for(*sinks.ani, base_off.front(), base_off.back(), seed_sink)) return rc;
        bool sel_ready = false;
        if (sinks.sketch && sinks.d_hashes) {
            // K1 writes the rows where K2 will read them (the table never leaves the device)
            const size_t stride = sinks.stride ? sinks.stride : sinks.s;
            int rc = sketch_enqueue(g_ctx.sws, d_seq2, d_valid, d_off, nb, sinks.k, sinks.s, sinks.seed,
                                    sinks.d_hashes + unit_first * stride, sinks.d_counts + unit_first, stride, st,
                                    fuse && sinks.k == 21 ? &seed_sink : nullptr);
            if (rc) return rc;
            sel_ready = fuse && sinks.k == 21;
        } else if (sinks.sketch) {
            DevBuf<uint64_t> d_hashes;
            DevBuf<uint32_t> d_counts;
            if (d_hashes.alloc(nb * (size_t)sinks.s) || d_counts.alloc(nb)) return GALAH_B200_ERR_CUDA;
            int rc = sketch_enqueue(g_ctx.sws, d_seq2, d_valid, d_off, nb, sinks.k, sinks.s, sinks.seed, d_hashes.p,
                                    d_counts.p, sinks.s, st);
            if (rc) return rc;
            GB_CUDA(cudaMemcpyAsync(sinks.hashes + unit_first * (size_t)sinks.s, d_hashes.p, nb * (size_t)sinks.s * 8,
                                    cudaMemcpyDeviceToHost, st));
            GB_CUDA(cudaMemcpyAsync(sinks.counts + unit_first, d_counts.p, nb * 4, cudaMemcpyDeviceToHost, st));
            GB_CUDA(cudaStreamSynchronize(st));
        }
        if (sinks.markers) {
            MarkerTable &mt = *sinks.markers;
            const uint32_t cap = marker_row_capacity(longest, sinks.c_marker);
            // rows for everything still to come (units per file as seen so far), common stride
            const double per_file = (double)sinks.n_units / (double)std::max<size_t>(files_done, 1);
            const size_t rows_hint = std::max(mt.n + nb, (size_t)(per_file * (double)n) + 1);
            if (int rc = mt.reserve(mt.n + nb > mt.cap_rows ? rows_hint : mt.n + nb, cap, st)) return rc;
            uint64_t *d_rows = mt.d_rows + mt.n * (size_t)mt.stride;
            uint32_t *d_cnt = mt.d_counts + mt.n;
            int rc = marker_sketch_enqueue(g_ctx.sws, d_seq2, d_valid, d_off, nb, 21, sinks.c_marker, mt.stride, d_rows,
                                           d_cnt, st, fuse && !sel_ready ? &seed_sink : nullptr);
            sel_ready = sel_ready || fuse;
            if (rc) return rc;
            std::vector<uint32_t> cnt(nb);
            GB_CUDA(cudaMemcpyAsync(cnt.data(), d_cnt, nb * 4, cudaMemcpyDeviceToHost, st));
            GB_CUDA(cudaStreamSynchronize(st));
            for (size_t x = 0; x < nb; x++)
                if (cnt[x] == 0xFFFFFFFFu) {
                    set_error("marker sketch: a unit holds more markers than its row of " + std::to_string(mt.stride) +
                              " (marker density 1/" + std::to_string(sinks.c_marker) + "; rows hold at most " +
                              std::to_string(kMarkerMaxCap) + " markers, about " +
                              std::to_string((uint64_t)kMarkerMaxCap * sinks.c_marker / 1500000) + " Mbp per unit); unsupported");
                    return GALAH_B200_ERR_UNSUPPORTED;
                }
            mt.n += nb;
        }
        if (sinks.ani) {
            const size_t before = sinks.ani->size();
            int rc = sinks.ani->add_packed_device(d_seq2, d_valid, d_off, nb, base_off, contig_off, cs, cl, st,
                                                  sel_ready ? seed_sink.d_sel : nullptr);
            if (rc) return rc;
            // first batch of a larger run: size the index once for everything still to come
            // (units per file as seen so far; exact for whole-genome units)
            if (before == 0 && files_done < n) {
                const double per_file = (double)sinks.n_units / (double)std::max<size_t>(files_done, 1);
                if (int rc2 = sinks.ani->reserve_for((size_t)(per_file * (double)n) + 1, st)) return rc2;
            }
        }
        GB_CUDA(cudaStreamSynchronize(st));
        return 0;
    };

    while (done < n) {
        const size_t want = std::min(kMaxBatchGenomes, n - done);
        std::vector<std::string> errs(want);
        std::vector<int> rcs(want, 0);
        std::atomic<size_t> next{0};
        const int nt = (int)std::min<size_t>((size_t)host_threads, want);

        // ---- K0 path (whole-genome units): host threads only read (and inflate) the files; the
        // raw bytes cross PCIe once and are decoded on the device (csrc/ingest.cu)
        if (g_ctx.device_ingest) {
            // Plain files are read by the host threads STRAIGHT into the pinned staging buffer (their sizes are known
            // from the probe, so every file's place is); gzip files are inflated into memory first.  The probe also
            // sniffs the first non-blank byte: anything but '>' (FASTQ) sends the batch to the host packer below.
            struct Probe { bool plain = false; uint64_t size = 0; bool fasta = true; };
            std::vector<Probe> probe(want);
            std::vector<std::vector<uint8_t>> raw(want);  // inflated gzip members / unseekable inputs only
            auto sniff = [](const uint8_t *d, size_t len, bool whole) {
                size_t p = 0;
                while (p < len && (d[p] == '\n' || d[p] == '\r')) p++;
                if (p < len) return d[p] == '>';
                return whole;  // an empty / all-blank file decodes to nothing; an undecided head goes to the host packer
            };
            auto prober = [&]() {
                for (;;) {
                    size_t x = next.fetch_add(1);
                    if (x >= want) break;
                    const char *path = paths[done + x];
                    FILE *fp = fopen(path, "rb");
                    if (!fp) { errs[x] = std::string("Failed to open fasta file ") + path; rcs[x] = 4; continue; }
                    uint8_t head[4096];
                    const size_t got = fread(head, 1, sizeof(head), fp);
                    struct stat sb;
                    const bool gz = got >= 2 && head[0] == 0x1f && head[1] == 0x8b;
                    const bool regular = fstat(fileno(fp), &sb) == 0 && S_ISREG(sb.st_mode);
                    fclose(fp);
                    if (!gz && regular) {
                        probe[x].plain = true; probe[x].size = (uint64_t)sb.st_size;
                        probe[x].fasta = sniff(head, got, got == (size_t)sb.st_size);
                    } else {
                        rcs[x] = read_file_bytes(path, raw[x], errs[x]);
                        probe[x].size = raw[x].size();
                        probe[x].fasta = sniff(raw[x].data(), raw[x].size(), true);
                    }
                }
            };
            {
                std::vector<std::thread> th;
                for (int t = 1; t < nt; t++) th.emplace_back(prober);
                prober();
                for (auto &t : th) t.join();
            }
            for (size_t x = 0; x < want; x++)
                if (rcs[x]) { set_error(errs[x]); return GALAH_B200_ERR_IO; }
            bool all_fasta = true;
            for (size_t x = 0; x < want; x++) all_fasta = all_fasta && probe[x].fasta;
            if (all_fasta) {
                size_t b0 = 0;
                while (b0 < want) {  // sub-batches of at most ~2 GiB of raw bytes
                    size_t b1 = b0; uint64_t bytes = 0;
                    while (b1 < want && (b1 == b0 || bytes + probe[b1].size + 32 <= kMaxBatchBases)) { bytes += probe[b1].size + 32; b1++; }
                    const size_t nf = b1 - b0;
                    RawBatch rb;
                    rb.file_off.assign(nf + 1, 0); rb.file_len.assign(nf, 0); rb.first_byte.assign(nf, 0);
                    for (size_t f = 0; f < nf; f++) {
                        rb.file_len[f] = probe[b0 + f].size;
                        rb.file_off[f + 1] = (rb.file_off[f] + probe[b0 + f].size + 31) / 32 * 32;
                    }
                    const size_t need = rb.file_off[nf] + 64;
                    if (g_ctx.cap_raw < need) {
                        if (g_ctx.h_raw) GB_CUDA(cudaFreeHost(g_ctx.h_raw));
                        g_ctx.h_raw = nullptr; g_ctx.cap_raw = 0;
                        const size_t grow = need + need / 4;
                        GB_CUDA(cudaMallocHost(&g_ctx.h_raw, grow));
                        g_ctx.cap_raw = grow;
                    }
                    rb.bytes = g_ctx.h_raw;
                    std::atomic<size_t> next_file{0};
                    auto filler = [&]() {
                        for (;;) {
                            const size_t f = next_file.fetch_add(1);
                            if (f >= nf) break;
                            const size_t x = b0 + f;
                            uint8_t *dst = rb.bytes + rb.file_off[f];
                            const size_t len = (size_t)probe[x].size;
                            if (probe[x].plain) {
                                FILE *fp = fopen(paths[done + x], "rb");
                                const size_t rd = fp && len ? fread(dst, 1, len, fp) : 0;
                                // one more byte must not be there: the file is as long as the probe saw it
                                const bool longer = fp && fgetc(fp) != EOF;
                                if (fp) fclose(fp);
                                if (!fp || rd != len || longer) { errs[x] = std::string("Failed to read ") + paths[done + x]; rcs[x] = 4; continue; }
                            } else if (len) {
                                memcpy(dst, raw[x].data(), len);
                                std::vector<uint8_t>().swap(raw[x]);
                            }
                            memset(dst + len, '\n', rb.file_off[f + 1] - rb.file_off[f] - len);
                            size_t p = 0;
                            while (p < len && (dst[p] == '\n' || dst[p] == '\r')) p++;
                            rb.first_byte[f] = rb.file_off[f] + p;
                        }
                    };
                    {
                        std::vector<std::thread> th;
                        for (int t = 1; t < nt; t++) th.emplace_back(filler);
                        filler();
                        for (auto &t : th) t.join();
                    }
                    for (size_t x = b0; x < b1; x++)
                        if (rcs[x]) { set_error(errs[x]); return GALAH_B200_ERR_IO; }
                    DecodedFiles dec, split;
                    if (int rc = g_ctx.fasta.decode(rb.bytes, rb.file_off, rb.file_len, rb.first_byte, dec, st)) return rc;
                    // contig mode: every record becomes its own unit (split on the device)
                    if (sinks.per_record)
                        if (int rc = g_ctx.fasta.split_records(dec, split, st)) return rc;
                    const DecodedFiles &un = sinks.per_record ? split : dec;
                    const size_t nb = un.n_bases.size();
                    std::vector<uint64_t> contig_off(nb + 1, 0);
                    std::vector<uint32_t> cs, cl;
                    uint64_t longest = 0;
                    for (size_t x = 0; x < nb; x++) {
                        for (uint64_t r = un.rec_off[x]; r < un.rec_off[x + 1]; r++) {
                            cs.push_back((uint32_t)un.rec_start[r]);
                            cl.push_back((uint32_t)(un.rec_end[r] - un.rec_start[r]));
                        }
                        contig_off[x + 1] = cs.size();
                        longest = std::max(longest, un.n_bases[x]);
                    }
                    const size_t unit_first = sinks.n_units;
                    sinks.n_units += nb;
                    if (nb)
                        if (int rc = feed(un.d_seq2, un.d_valid, un.d_base_off, nb, un.base_off, contig_off, cs, cl, longest,
                                          unit_first, done + b1))
                            return rc;
                    b0 = b1;
                }
                done += want;
                continue;
            }
            // FASTQ (or anything else the host packer understands) in this batch: pack on the host below
            next = 0;
        }

        // ---- host packer path (contig mode, FASTQ input, or device ingest switched off)
        std::vector<PackedGenome> batch(want);
        std::vector<std::vector<PackedGenome>> per_file(sinks.per_record ? want : 0);
        auto worker = [&]() {
            for (;;) {
                size_t x = next.fetch_add(1);
                if (x >= want) break;
                if (sinks.per_record) rcs[x] = pack_fasta_file_per_record(paths[done + x], per_file[x], false, errs[x]);
                else rcs[x] = pack_fasta_file(paths[done + x], batch[x], false, errs[x]);
            }
        };
        std::vector<std::thread> th;
        for (int t = 1; t < nt; t++) th.emplace_back(worker);
        worker();
        for (auto &t : th) t.join();
        for (size_t x = 0; x < want; x++)
            if (rcs[x]) { set_error(errs[x]); return GALAH_B200_ERR_IO; }
        if (sinks.per_record) {  // flatten: one unit per record, file order then record order
            batch.clear();
            for (auto &f : per_file) for (auto &g : f) batch.push_back(std::move(g));
        }
        const size_t n_batch = batch.size();
        const size_t unit0 = sinks.n_units;
        sinks.n_units += n_batch;
        size_t b0 = 0;
        while (b0 < n_batch) {  // split further if the batch is too large for one upload
            size_t b1 = b0; uint64_t bases = 0;
            while (b1 < n_batch && (b1 == b0 || (bases + batch[b1].padded_bases() <= kMaxBatchBases && b1 - b0 < 65536))) {
                bases += batch[b1].padded_bases(); b1++;
            }
            const size_t nb = b1 - b0;
            std::vector<uint64_t> base_off(nb + 1, 0), contig_off(nb + 1, 0);
            std::vector<uint32_t> cs, cl;
            uint64_t longest = 0;
            for (size_t x = 0; x < nb; x++) {
                const PackedGenome &pg = batch[b0 + x];
                base_off[x + 1] = base_off[x] + pg.padded_bases();
                for (size_t r = 0; r < pg.rec_start.size(); r++) {
                    cs.push_back((uint32_t)pg.rec_start[r]);
                    cl.push_back((uint32_t)(pg.rec_end[r] - pg.rec_start[r]));
                }
                contig_off[x + 1] = cs.size();
                longest = std::max<uint64_t>(longest, pg.n_bases);
            }
            const uint64_t total = base_off[nb];
            std::vector<uint32_t> seq2(total / 16 + 4, 0u), valid(total / 32 + 4, 0u);
            for (size_t x = 0; x < nb; x++) {
                const PackedGenome &pg = batch[b0 + x];
                memcpy(seq2.data() + base_off[x] / 16, pg.seq2.data(), pg.padded_bases() / 16 * 4);
                memcpy(valid.data() + base_off[x] / 32, pg.valid.data(), pg.padded_bases() / 32 * 4);
            }
            DevBuf<uint32_t> d_seq2, d_valid;
            DevBuf<uint64_t> d_off;
            if (d_seq2.alloc(seq2.size()) || d_valid.alloc(valid.size()) || d_off.alloc(nb + 1)) return GALAH_B200_ERR_CUDA;
            GB_CUDA(cudaMemcpyAsync(d_seq2.p, seq2.data(), seq2.size() * 4, cudaMemcpyHostToDevice, st));
            GB_CUDA(cudaMemcpyAsync(d_valid.p, valid.data(), valid.size() * 4, cudaMemcpyHostToDevice, st));
            GB_CUDA(cudaMemcpyAsync(d_off.p, base_off.data(), (nb + 1) * 8, cudaMemcpyHostToDevice, st));
            if (int rc = feed(d_seq2.p, d_valid.p, d_off.p, nb, base_off, contig_off, cs, cl, longest, unit0 + b0, done + want))
                return rc;
            b0 = b1;
        }
        done += want;
    }
    return 0;
}

// Rust's `{}` for an f32 that holds a short decimal (thresholds typed on a command line).
static std::string display_f32(float v) {
    char buf[64];
    for (int prec = 0; prec <= 12; prec++) {  // shortest fixed-point text that round-trips, no exponent
        snprintf(buf, sizeof(buf), "%.*f", prec, (double)v);
        if (strtof(buf, nullptr) == v) break;
    }
    return buf;
}

// SkaniPreclusterer::distances / distances_contigs (src/skani.rs:21-56, 109-225, 379-498): marker
// screen (K1 machinery + K2 join with the containment rule) -> K3 ANI on the survivors -> keep
// ANI >= threshold.  Index handle stays with the caller (needed for nothing else; freed by it).
// The skani-style preclusterer once the units are indexed (K3) and their marker table is on the
// device: marker-containment screen through the K2 join (containment rule), K3 ANI of every
// screened pair, `ani >= threshold` in f32 (src/skani.rs:205).  ms3 (optional): screen, ANI, total.
// variant: kSkaniTriangle  -- `skani triangle` (src/skani.rs:109-225): every screened pair, query = lower index;
//          kSkaniLowMem    -- `skani sketch` + `skani search` of everything against everything
//                             (src/skani.rs:229-377): both orientations reach the cache under one key and
//                             the later record wins; at skani's query order that is query = HIGHER index;
//          kSkaniReferences -- `skani search` of the non-references against the sketched references
//                             (src/skani.rs:502-687): only pairs with exactly one reference (is_ref),
//                             query = the non-reference genome.
enum SkaniVariant { kSkaniTriangle = 0, kSkaniLowMem = 1, kSkaniReferences = 2 };
static int skani_screen_and_ani(AniIndex &index, const uint64_t *d_table, const uint32_t *d_counts, size_t n_units,
                                size_t stride, float threshold_pct, float min_af_pct, bool individual_contigs,
                                cudaStream_t st, std::vector<galah_b200_pair_t> &out, uint64_t &n_screened, float *ms3,
                                SkaniVariant variant = kSkaniTriangle, const std::vector<uint8_t> *is_ref = nullptr) {
    const double t_begin = now_ms();
    if (!g_ctx.d_n_cand) GB_CUDA(cudaMalloc(&g_ctx.d_n_cand, sizeof(unsigned long long)));
    const double frac = pow(0.80, 21.0);  // marker containment of a pair at ~80 % identity
    size_t cap = std::max<size_t>(1 << 16, 64 * n_units);
    std::vector<uint4> cand;
    for (;;) {
        if (ws_ensure(g_ctx.d_cand, g_ctx.cap_cand, cap)) return GALAH_B200_ERR_CUDA;
        int rc = prefilter_enqueue(g_ctx.pws, d_table, d_counts, n_units, stride, 21, 0.f, 0, 1,
                                   join_supported(stride) ? 0 : 1, st, g_ctx.d_cand, g_ctx.cap_cand, g_ctx.d_n_cand,
                                   index.c() == 30u ? kRuleContainment : kRuleContainmentBypassSmall, frac);
        if (rc) return rc;
        unsigned long long got = 0;
        GB_CUDA(cudaMemcpyAsync(&got, g_ctx.d_n_cand, sizeof(got), cudaMemcpyDeviceToHost, st));
        GB_CUDA(cudaStreamSynchronize(st));
        if (got > g_ctx.cap_cand) { cap = (size_t)got; continue; }
        cand.resize((size_t)got);
        if (got) {
            GB_CUDA(cudaMemcpyAsync(cand.data(), g_ctx.d_cand, (size_t)got * sizeof(uint4), cudaMemcpyDeviceToHost, st));
            GB_CUDA(cudaStreamSynchronize(st));
        }
        break;
    }
    const double t_screen = now_ms();
    std::sort(cand.begin(), cand.end(), [](const uint4 &a, const uint4 &b) { return a.x != b.x ? a.x < b.x : a.y < b.y; });
    if (variant == kSkaniReferences) {
        size_t keep = 0;
        for (size_t x = 0; x < cand.size(); x++)
            if ((*is_ref)[cand[x].x] != (*is_ref)[cand[x].y]) cand[keep++] = cand[x];
        cand.resize(keep);
    }
    n_screened = cand.size();
    std::vector<uint32_t> pairs(2 * cand.size());
    for (size_t x = 0; x < cand.size(); x++) {
        bool q_is_j = variant == kSkaniLowMem;
        if (variant == kSkaniReferences) q_is_j = (*is_ref)[cand[x].x] != 0;  // the query is the non-reference
        pairs[2 * x] = q_is_j ? cand[x].y : cand[x].x;
        pairs[2 * x + 1] = q_is_j ? cand[x].x : cand[x].y;
    }
    std::vector<AniPairResult> res(cand.size());
    if (int rc = index.pairs(pairs.data(), cand.size(), min_af_pct, individual_contigs, res.data(), st)) return rc;
    out.clear();
    for (size_t x = 0; x < cand.size(); x++)
        if (res[x].ani >= threshold_pct)  // `if ani >= threshold` in f32, src/skani.rs:205
            out.push_back(galah_b200_pair_t{cand[x].x, cand[x].y, cand[x].z, cand[x].w, res[x].ani});
    if (ms3) { ms3[0] = (float)(t_screen - t_begin); ms3[1] = (float)(now_ms() - t_screen); ms3[2] = (float)(now_ms() - t_begin); }
    return 0;
}

static int skani_distances_impl(const char *const *paths, size_t n, float threshold_pct, float min_af_pct,
                                bool small_genomes, bool per_record, int host_threads,
                                std::vector<galah_b200_pair_t> &out, size_t &n_units, uint64_t &n_screened,
                                SkaniVariant variant = kSkaniTriangle, const std::vector<uint8_t> *is_ref = nullptr) {
    if (threshold_pct < 85.0f) {
        set_error("Error: skani produces inaccurate results with ANI less than 85%. Provided: " + display_f32(threshold_pct));
        return GALAH_B200_ERR_UNSUPPORTED;
    }
    if (small_genomes && variant == kSkaniLowMem) {  // src/skani.rs:243-245
        set_error("Error: skani does not support small genomes with low-memory preclustering");
        return GALAH_B200_ERR_UNSUPPORTED;
    }
    if (small_genomes && variant == kSkaniReferences) {  // src/skani.rs:518-520
        set_error("Error: skani does not support small genomes with reference genome preclustering");
        return GALAH_B200_ERR_UNSUPPORTED;
    }
    AniIndex index(small_genomes ? 30u : 125u);
    MarkerTable markers;  // marker sketches stay on the device, in the K2 table layout
    IngestSinks sinks;
    sinks.ani = &index; sinks.markers = &markers; sinks.c_marker = small_genomes ? 200u : 1000u;
    sinks.per_record = per_record;
    if (int rc = ingest_files(paths, n, host_threads, sinks)) return rc;
    n_units = sinks.n_units;
    out.clear(); n_screened = 0;
    if (n_units < 2) return 0;
    if (markers.n != n_units) { set_error("skani preclusterer: marker table out of step with the units"); return GALAH_B200_ERR_ARG; }
    return skani_screen_and_ani(index, markers.d_rows, markers.d_counts, n_units, markers.stride, threshold_pct,
                                min_af_pct, per_record, g_ctx.stream, out, n_screened, nullptr, variant, is_ref);
}

}  // namespace gb200

using namespace gb200;

extern "C" {

const char *galah_b200_last_error(void) { return t_error.c_str(); }
const char *galah_b200_version(void) { return "galah-b200 0.1.0 (sm_100a)"; }
void galah_b200_free(void *p) { free(p); }
uint64_t galah_b200_launch_count(void) { return g_launch_count.load(); }

int galah_b200_prefilter_mode(int mode) {
    std::lock_guard<std::mutex> lock(g_mu);
    const int prev = g_ctx.prefilter_mode;
    if (mode == 0 || mode == 1) g_ctx.prefilter_mode = mode;
    return prev;
}

int galah_b200_prefilter_last_timing(float *build_ms, float *main_ms) {
    std::lock_guard<std::mutex> lock(g_mu);
    if (int rc = require_ctx()) return rc;
    return g_ctx.pws.last_timing(build_ms, main_ms);
}

int galah_b200_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

static std::mutex g_init_mu;  // serialises (re)binding; taken before any per-device lock

static int init_locked(int device) {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
        cudaGetLastError();
        set_error("galah_b200_init: no CUDA device visible (this library has no CPU fallback)");
        return GALAH_B200_ERR_NO_DEVICE;
    }
    if (device < 0 || device >= n) { set_error("galah_b200_init: device index out of range"); return GALAH_B200_ERR_ARG; }
    cudaDeviceProp prop;
    GB_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10) {
        set_error(std::string("galah_b200_init: device '") + prop.name +
                  "' is not sm_100; kernels are built for sm_100a only");
        return GALAH_B200_ERR_NO_DEVICE;
    }
    if (device >= kMaxDevices) { set_error("galah_b200_init: device index beyond the context table"); return GALAH_B200_ERR_ARG; }
    if (t_dev < 0) {
        // the calling thread is not a pinned worker: `device` becomes the process's primary device, and
        // what an earlier primary device held is released (one resident working set per process, as before)
        const int old = g_primary.load();
        if (old != device && g_ctxs[old].device >= 0) {
            std::lock_guard<std::mutex> lock_old(g_mus[old]);
            cudaSetDevice(g_ctxs[old].device);
            g_ctxs[old].release();
            if (g_ctxs[old].stream) cudaStreamDestroy(g_ctxs[old].stream);
            g_ctxs[old].stream = nullptr;
            g_ctxs[old].device = -1;
        }
        g_primary.store(device);
    }
    std::lock_guard<std::mutex> lock(g_mus[device]);
    Context &C = g_ctxs[device];
    GB_CUDA(cudaSetDevice(device));
    if (!C.stream) {
        // highest priority: the small kernels of the K3 index build must get SM slots while a K1 scan of
        // the next batch (tens of thousands of CTAs, lowest priority, scan_stream) is resident
        int prio_lo = 0, prio_hi = 0;
        GB_CUDA(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
        GB_CUDA(cudaStreamCreateWithPriority(&C.stream, cudaStreamNonBlocking, prio_hi));
    }
    C.device = device;
    return 0;
}

int galah_b200_init(int device) {
    std::lock_guard<std::mutex> lock(g_init_mu);
    return init_locked(device);
}

// Binds devices 0 .. n_devices - 1 for galah_b200_cluster_packed_multi (device 0 stays the primary
// device of every single-device entry point) and opens peer access between every pair of them.
int galah_b200_init_devices(int n_devices) {
    std::lock_guard<std::mutex> lock(g_init_mu);
    if (n_devices < 1 || n_devices > kMaxDevices) { set_error("galah_b200_init_devices: bad device count"); return GALAH_B200_ERR_ARG; }
    if (int rc = init_locked(0)) return rc;
    for (int d = 1; d < n_devices; d++) {
        t_dev = d;  // bind slot d without moving the primary device
        const int rc = init_locked(d);
        t_dev = -1;
        if (rc) return rc;
    }
    for (int a = 0; a < n_devices; a++) {
        GB_CUDA(cudaSetDevice(a));
        for (int b = 0; b < n_devices; b++) {
            if (a == b) continue;
            int can = 0;
            GB_CUDA(cudaDeviceCanAccessPeer(&can, a, b));
            if (!can) { set_error("galah_b200_init_devices: devices without peer access"); return GALAH_B200_ERR_UNSUPPORTED; }
            const cudaError_t e = cudaDeviceEnablePeerAccess(b, 0);
            if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) GB_CUDA(e);
            cudaGetLastError();
        }
    }
    GB_CUDA(cudaSetDevice(0));
    return 0;
}

int galah_b200_sketch_packed_device(const uint32_t *d_seq2, const uint32_t *d_valid,
                                    const uint64_t *d_base_off, size_t n, uint8_t k, uint32_t s,
                                    uint64_t seed, uint64_t *d_hashes, uint32_t *d_counts, void *stream) {
    std::lock_guard<std::mutex> lock(g_mu);
    if (int rc = require_ctx()) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    return sketch_enqueue(g_ctx.sws, d_seq2, d_valid, d_base_off, n, k, s, seed, d_hashes, d_counts, s, st);
}

int galah_b200_sketch_packed(const uint32_t *seq2, const uint32_t *valid, const uint64_t *base_off,
                             size_t n, uint8_t k, uint32_t s, uint64_t seed, uint64_t *hashes,
                             uint32_t *counts) {
    std::lock_guard<std::mutex> lock(g_mu);
    if (int rc = require_ctx()) return rc;
    if (n == 0) return 0;
    for (size_t g = 0; g <= n; g++)
        if (base_off[g] % 128) { set_error("sketch_packed: base_off must be multiples of 128"); return GALAH_B200_ERR_ARG; }
    const uint64_t total = base_off[n];
    DevBuf<uint32_t> d_seq2, d_valid, d_counts;
    DevBuf<uint64_t> d_off, d_hashes;
    if (d_seq2.alloc(total / 16 + 4) || d_valid.alloc(total / 32 + 4) || d_off.alloc(n + 1) ||
        d_hashes.alloc(n * (size_t)s) || d_counts.alloc(n))
        return GALAH_B200_ERR_CUDA;
    cudaStream_t st = g_ctx.stream;
    GB_CUDA(cudaMemsetAsync(d_seq2.p, 0, (total / 16 + 4) * 4, st));
    GB_CUDA(cudaMemsetAsync(d_valid.p, 0, (total / 32 + 4) * 4, st));
    GB_CUDA(cudaMemcpyAsync(d_seq2.p, seq2, total / 16 * 4, cudaMemcpyHostToDevice, st));
    GB_CUDA(cudaMemcpyAsync(d_valid.p, valid, total / 32 * 4, cudaMemcpyHostToDevice, st));
    GB_CUDA(cudaMemcpyAsync(d_off.p, base_off, (n + 1) * 8, cudaMemcpyHostToDevice, st));
    int rc = sketch_enqueue(g_ctx.sws, d_seq2.p, d_valid.p, d_off.p, n, k, s, seed, d_hashes.p,
                            d_counts.p, s, st);
    if (rc) return rc;
    GB_CUDA(cudaMemcpyAsync(hashes, d_hashes.p, n * (size_t)s * 8, cudaMemcpyDeviceToHost, st));
    GB_CUDA(cudaMemcpyAsync(counts, d_counts.p, n * 4, cudaMemcpyDeviceToHost, st));
    GB_CUDA(cudaStreamSynchronize(st));
    return 0;
}

int galah_b200_sketch_files(const char *const *paths, size_t n, uint8_t k, uint32_t s, uint64_t seed,
                            int host_threads, uint64_t *hashes, uint32_t *counts) {
    std::lock_guard<std::mutex> lock(g_mu);
    if (int rc = require_ctx()) return rc;
    if (s & 1) { set_error("sketch_files: s must be even (row stride alignment)"); return GALAH_B200_ERR_ARG; }
    IngestSinks sinks;
    sinks.sketch = true; sinks.k = k; sinks.s = s; sinks.seed = seed; sinks.hashes = hashes; sinks.counts = counts;
    return ingest_files(paths, n, host_threads, sinks);
}

int galah_b200_prefilter_enqueue(const uint64_t *d_hashes, const uint32_t *d_counts, size_t n,
                                 size_t stride, uint8_t k, float min_ani, uint32_t shard,
                                 uint32_t n_shards, int mode, void *stream, uint32_t *d_cand,
                                 size_t cand_cap, unsigned long long *d_n_cand) {
    std::lock_guard<std::mutex> lock(g_mu);
    if (int rc = require_ctx()) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    return prefilter_enqueue(g_ctx.pws, d_hashes, d_counts, n, stride, k, min_ani, shard, n_shards, mode,
                             st, reinterpret_cast<uint4 *>(d_cand), cand_cap, d_n_cand);
}

int galah_b200_finish_candidates(const uint32_t *cand, size_t n_cand, uint8_t k, float min_ani,
                                 galah_b200_pair_t **out, size_t *n_out) {
    *out = nullptr; *n_out = 0;
    return finish_candidates(reinterpret_cast<const uint4 *>(cand), n_cand, k, min_ani, out, n_out);
}

int galah_b200_blocklist_layout(size_t n, size_t stride, size_t *n_blocks, size_t *entries_per_block, size_t *slack) {
    blocklist_layout(n, stride, n_blocks, entries_per_block, slack);
    return 0;
}

int galah_b200_blocklist_build(const uint64_t *d_hashes, const uint32_t *d_counts, size_t n, size_t stride,
                               size_t block_begin, size_t block_end, uint32_t *d_hi, uint32_t *d_lo,
                               uint8_t *d_tags, uint32_t *d_len, void *stream) {
    std::lock_guard<std::mutex> lock(g_mu);
    if (int rc = require_ctx()) return rc;
    if (stride == 0 || (stride & 1) || !join_supported(stride)) { set_error("blocklist_build: unsupported stride"); return GALAH_B200_ERR_ARG; }
    if (n >= 0x7FFFFFFFull) { set_error("blocklist_build: n too large"); return GALAH_B200_ERR_ARG; }
    return blocklist_build(g_ctx.pws, d_hashes, d_counts, n, stride, (uint32_t)block_begin, (uint32_t)block_end, d_hi,
                           d_lo, d_tags, d_len, (cudaStream_t)stream);
}

int galah_b200_table_max_device(const uint64_t *d_hashes, const uint32_t *d_counts, size_t n, size_t stride,
                                unsigned long long *d_max, void *stream) {
    std::lock_guard<std::mutex> lock(g_mu);
    if (int rc = require_ctx()) return rc;
    if (stride == 0 || n >= 0x7FFFFFFFull || !d_max) { set_error("table_max: bad arguments"); return GALAH_B200_ERR_ARG; }
    return table_max_enqueue(d_hashes, d_counts, n, stride, d_max, (cudaStream_t)stream);
}

int galah_b200_blocklist_build_local(const uint64_t *d_rows, const uint32_t *d_counts, size_t n_rows, size_t stride,
                                     const unsigned long long *d_table_max, size_t n_blocks_out, uint32_t *d_hi,
                                     uint32_t *d_lo, uint8_t *d_tags, uint32_t *d_len, void *stream) {
    std::lock_guard<std::mutex> lock(g_mu);
    if (int rc = require_ctx()) return rc;
    if (stride == 0 || (stride & 1) || !join_supported(stride)) { set_error("blocklist_build_local: unsupported stride"); return GALAH_B200_ERR_ARG; }
    if (n_rows >= 0x7FFFFFFFull || !d_table_max) { set_error("blocklist_build_local: bad arguments"); return GALAH_B200_ERR_ARG; }
    if (n_blocks_out * (size_t)GALAH_B200_ROW_BLOCK < n_rows) { set_error("blocklist_build_local: n_blocks_out does not cover the rows"); return GALAH_B200_ERR_ARG; }
    return blocklist_build_local(g_ctx.pws, d_rows, d_counts, n_rows, stride, d_table_max, d_hi, d_lo, d_tags, d_len,
                                 (uint32_t)n_blocks_out, (cudaStream_t)stream);
}

int galah_b200_prefilter_join_enqueue(const uint64_t *d_hashes, const uint32_t *d_counts, size_t n, size_t stride,
                                      uint8_t k, float min_ani, const uint32_t *d_hi, const uint32_t *d_lo,
                                      const uint8_t *d_tags, const uint32_t *d_len, uint32_t shard,
                                      uint32_t n_shards, void *stream, uint32_t *d_cand, size_t cand_cap,
                                      unsigned long long *d_n_cand) {
    std::lock_guard<std::mutex> lock(g_mu);
    if (int rc = require_ctx()) return rc;
    if (!join_supported(stride)) { set_error("prefilter_join: unsupported stride"); return GALAH_B200_ERR_ARG; }
    KernelParams p;
    cudaStream_t st = (cudaStream_t)stream;
    if (int rc = prefilter_prepare(g_ctx.pws, d_hashes, d_counts, n, stride, k, min_ani, shard, n_shards, st,
                                   reinterpret_cast<uint4 *>(d_cand), cand_cap, d_n_cand, p))
        return rc;
    if (n < 2) return 0;
    return join_launch(g_ctx.pws, p, d_hi, d_lo, d_tags, d_len, shard, n_shards, st);
}

int galah_b200_prefilter_join_enqueue_screen(const uint64_t *d_rows, const uint32_t *d_counts, size_t n, size_t stride,
                                             int faster_small, const uint32_t *d_hi, const uint32_t *d_lo,
                                             const uint8_t *d_tags, const uint32_t *d_len, uint32_t shard,
                                             uint32_t n_shards, void *stream, uint32_t *d_cand, size_t cand_cap,
                                             unsigned long long *d_n_cand) {
    std::lock_guard<std::mutex> lock(g_mu);
    if (int rc = require_ctx()) return rc;
    if (!join_supported(stride)) { set_error("prefilter_join: unsupported stride"); return GALAH_B200_ERR_ARG; }
    KernelParams p;
    cudaStream_t st = (cudaStream_t)stream;
    if (int rc = prefilter_prepare(g_ctx.pws, d_rows, d_counts, n, stride, 21, 0.f, shard, n_shards, st,
                                   reinterpret_cast<uint4 *>(d_cand), cand_cap, d_n_cand, p,
                                   faster_small ? kRuleContainment : kRuleContainmentBypassSmall, pow(0.80, 21.0)))
        return rc;
    if (n < 2) return 0;
    return join_launch(g_ctx.pws, p, d_hi, d_lo, d_tags, d_len, shard, n_shards, st);
}

int galah_b200_prefilter_join_items_enqueue(const uint64_t *d_hashes, const uint32_t *d_counts, size_t n, size_t stride,
                                            uint8_t k, float min_ani, const uint32_t *d_hi, const uint32_t *d_lo,
                                            const uint8_t *d_tags, const uint32_t *d_len, const uint32_t *d_items,
                                            size_t n_items, int reset_candidates, void *stream, uint32_t *d_cand,
                                            size_t cand_cap, unsigned long long *d_n_cand) {
    std::lock_guard<std::mutex> lock(g_mu);
    if (int rc = require_ctx()) return rc;
    if (!join_supported(stride)) { set_error("prefilter_join: unsupported stride"); return GALAH_B200_ERR_ARG; }
    KernelParams p;
    cudaStream_t st = (cudaStream_t)stream;
    if (int rc = prefilter_prepare(g_ctx.pws, d_hashes, d_counts, n, stride, k, min_ani, 0, 1, st,
                                   reinterpret_cast<uint4 *>(d_cand), cand_cap, d_n_cand, p, kRuleMashAni, 0.0,
                                   reset_candidates != 0))
        return rc;
    if (n < 2) return 0;
    return join_launch_items(g_ctx.pws, p, d_hi, d_lo, d_tags, d_len, d_items, n_items, st);
}

int galah_b200_prefilter_device(const uint64_t *d_hashes, const uint32_t *d_counts, size_t n,
                                size_t stride, uint8_t k, float min_ani, uint32_t shard,
                                uint32_t n_shards, void *stream, galah_b200_pair_t **out, size_t *n_out) {
    std::lock_guard<std::mutex> lock(g_mu);
    if (int rc = require_ctx()) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    return run_prefilter(d_hashes, d_counts, n, stride, k, min_ani, shard, n_shards, st, out, n_out);
}

int galah_b200_prefilter_shard(const uint64_t *hashes, const uint32_t *counts, size_t n, size_t stride,
                               uint8_t k, float min_ani, uint32_t shard, uint32_t n_shards,
                               galah_b200_pair_t **out, size_t *n_out) {
    std::lock_guard<std::mutex> lock(g_mu);
    if (int rc = require_ctx()) return rc;
    *out = nullptr; *n_out = 0;
    if (stride == 0 || (stride & 1)) { set_error("prefilter: stride must be even and > 0"); return GALAH_B200_ERR_ARG; }
    if (ws_ensure(g_ctx.d_table, g_ctx.cap_table, std::max<size_t>(n * stride, 2)) ||
        ws_ensure(g_ctx.d_counts, g_ctx.cap_counts, std::max<size_t>(n, 1)))
        return GALAH_B200_ERR_CUDA;
    cudaStream_t st = g_ctx.stream;
    // whole-table single-shard join: pipeline the upload against the kernels (slices of whole blocks)
    const bool streamed = g_ctx.prefilter_mode == 0 && n_shards == 1 && g_ctx.stream_chunks > 1 &&
                          join_supported(stride) && n >= 64 * (size_t)kShardRows;  // smaller tables: one copy, and the
                          // device path's guard against a tie-dense launch of few items applies
    if (n) {
        GB_CUDA(cudaMemcpyAsync(g_ctx.d_counts, counts, n * 4, cudaMemcpyHostToDevice, st));
        if (!streamed) GB_CUDA(cudaMemcpyAsync(g_ctx.d_table, hashes, n * stride * 8, cudaMemcpyHostToDevice, st));
    }
    if (streamed && !g_ctx.copy_stream) GB_CUDA(cudaStreamCreateWithFlags(&g_ctx.copy_stream, cudaStreamNonBlocking));
    return run_prefilter(g_ctx.d_table, g_ctx.d_counts, n, stride, k, min_ani, shard, n_shards, st, out, n_out,
                         streamed ? hashes : nullptr, streamed ? counts : nullptr);
}

int galah_b200_prefilter(const uint64_t *hashes, const uint32_t *counts, size_t n, size_t stride,
                         uint8_t k, float min_ani, galah_b200_pair_t **out, size_t *n_out) {
    return galah_b200_prefilter_shard(hashes, counts, n, stride, k, min_ani, 0, 1, out, n_out);
}

int galah_b200_prefilter_stream_chunks(int chunks) {
    std::lock_guard<std::mutex> lock(g_mu);
    const int prev = g_ctx.stream_chunks;
    if (chunks >= 0) g_ctx.stream_chunks = std::min(chunks, 64);
    return prev;
}

int galah_b200_prefilter_last_host_timing(float *ms4) {
    std::lock_guard<std::mutex> lock(g_mu);
    if (!ms4) { set_error("prefilter_last_host_timing: ms4 is NULL"); return GALAH_B200_ERR_ARG; }
    for (int x = 0; x < 4; x++) ms4[x] = g_ctx.host_ms[x];
    return 0;
}

// finch::distances (src/finch.rs:48-97): K1 writes the sketch rows where K2 reads them (the table
// never visits the host); any num_kmers (the row stride is rounded up to even for alignment, rows
// hold at most num_kmers hashes).  ani (optional): the same ingest also builds the K3 index.
static int finch_distances_locked(const char *const *paths, size_t n, float min_ani, uint32_t num_kmers,
                                  uint8_t kmer_length, int host_threads, AniIndex *ani, galah_b200_pair_t **out,
                                  size_t *n_out) {
    *out = nullptr; *n_out = 0;
    if (num_kmers == 0) { set_error("finch_distances: num_kmers must be > 0"); return GALAH_B200_ERR_ARG; }
    const uint32_t stride = num_kmers + (num_kmers & 1);
    if (ws_ensure(g_ctx.d_table, g_ctx.cap_table, std::max<size_t>(n, 1) * stride) ||
        ws_ensure(g_ctx.d_counts, g_ctx.cap_counts, std::max<size_t>(n, 1)))
        return GALAH_B200_ERR_CUDA;
    IngestSinks sinks;
    sinks.sketch = true; sinks.k = kmer_length; sinks.s = num_kmers; sinks.stride = stride; sinks.seed = 0;
    sinks.d_hashes = g_ctx.d_table; sinks.d_counts = g_ctx.d_counts; sinks.ani = ani;
    if (int rc = ingest_files(paths, n, host_threads, sinks)) return rc;
    return run_prefilter(g_ctx.d_table, g_ctx.d_counts, n, stride, kmer_length, min_ani, 0, 1, g_ctx.stream, out, n_out);
}

int galah_b200_finch_distances(const char *const *paths, size_t n, float min_ani, uint32_t num_kmers,
                               uint8_t kmer_length, int host_threads, galah_b200_pair_t **out,
                               size_t *n_out) {
    std::lock_guard<std::mutex> lock(g_mu);
    if (int rc = require_ctx()) return rc;
    return finch_distances_locked(paths, n, min_ani, num_kmers, kmer_length, host_threads, nullptr, out, n_out);
}

struct galah_b200_ani_index { gb200::AniIndex impl; explicit galah_b200_ani_index(uint32_t c) : impl(c) {} };

int galah_b200_ani_index_create(int small_genomes, galah_b200_ani_index_t **out) {
    std::lock_guard<std::mutex> lock(g_mu);
    if (int rc = require_ctx()) return rc;
    if (!out) { set_error("ani index: out is NULL"); return GALAH_B200_ERR_ARG; }
    *out = new galah_b200_ani_index(small_genomes ? 30u : 125u);
    return 0;
}

void galah_b200_ani_index_free(galah_b200_ani_index_t *idx) {
    std::lock_guard<std::mutex> lock(g_mu);
    if (idx && g_ctx.device >= 0) cudaSetDevice(g_ctx.device);
    delete idx;
}

int galah_b200_ani_finish(const galah_b200_ani_result_t *ints, uint64_t len_q, uint64_t len_r, int small_genomes,
                          int individual_contigs, float min_af_pct, galah_b200_ani_result_t *out) {
    if (!out || !ints) { set_error("ani_finish: NULL argument"); return GALAH_B200_ERR_ARG; }
    AniPairInts v;
    v.sum_fx = ints->sum_fx; v.n_chunks = ints->n_chunks; v.sum_m = ints->sum_m; v.span_m = ints->span_m;
    v.span_n = ints->span_n; v.n_chains = ints->n_chains; v.cov_q = ints->cov_q; v.cov_r = ints->cov_r;
    const AniPairResult r = ani_finish(v, len_q, len_r, small_genomes ? 30u : 125u, individual_contigs != 0, min_af_pct);
    memcpy(out, &r, sizeof(r));
    return 0;
}

uint64_t galah_b200_chunk_identity_fx(uint32_t m, uint32_t n) { return chunk_identity_fx(m, n); }

float galah_b200_print2_parse_f32(double v) { return print2_parse_f32(v); }

int galah_b200_ani_index_reserve(galah_b200_ani_index_t *idx, size_t n_total_genomes) {
    std::lock_guard<std::mutex> lock(g_mu);
    if (int rc = require_ctx()) return rc;
    if (!idx) { set_error("ani index: idx is NULL"); return GALAH_B200_ERR_ARG; }
    return idx->impl.reserve_for(n_total_genomes, g_ctx.stream);
}

size_t galah_b200_ani_index_size(const galah_b200_ani_index_t *idx) { return idx ? idx->impl.size() : 0; }

static int ani_add_packed_host(galah_b200_ani_index_t *idx, const uint32_t *seq2, const uint32_t *valid,
                               const std::vector<uint64_t> &base_off, const std::vector<uint64_t> &contig_off,
                               const std::vector<uint32_t> &contig_start, const std::vector<uint32_t> &contig_len) {
    const size_t n = base_off.size() - 1;
    const uint64_t first = base_off[0], total = base_off[n] - first;
    DevBuf<uint32_t> d_seq2, d_valid;
    DevBuf<uint64_t> d_off;
    if (d_seq2.alloc(total / 16 + 8) || d_valid.alloc(total / 32 + 8) || d_off.alloc(n + 1)) return GALAH_B200_ERR_CUDA;
    cudaStream_t st = g_ctx.stream;
    std::vector<uint64_t> rel(n + 1);
    for (size_t g = 0; g <= n; g++) rel[g] = base_off[g] - first;
    GB_CUDA(cudaMemsetAsync(d_seq2.p, 0, (total / 16 + 8) * 4, st));
    GB_CUDA(cudaMemsetAsync(d_valid.p, 0, (total / 32 + 8) * 4, st));
    GB_CUDA(cudaMemcpyAsync(d_seq2.p, seq2 + first / 16, total / 16 * 4, cudaMemcpyHostToDevice, st));
    GB_CUDA(cudaMemcpyAsync(d_valid.p, valid + first / 32, total / 32 * 4, cudaMemcpyHostToDevice, st));
    GB_CUDA(cudaMemcpyAsync(d_off.p, rel.data(), (n + 1) * 8, cudaMemcpyHostToDevice, st));
    return idx->impl.add_packed_device(d_seq2.p, d_valid.p, d_off.p, n, rel, contig_off, contig_start, contig_len, st);
}

int galah_b200_ani_index_add_packed(galah_b200_ani_index_t *idx, const uint32_t *seq2, const uint32_t *valid,
                                    const uint64_t *base_off, size_t n, const uint64_t *contig_off,
                                    const uint32_t *contig_start, const uint32_t *contig_len) {
    std::lock_guard<std::mutex> lock(g_mu);
    if (int rc = require_ctx()) return rc;
    if (!idx) { set_error("ani index: NULL index"); return GALAH_B200_ERR_ARG; }
    if (n == 0) return 0;
    for (size_t g = 0; g <= n; g++)
        if (base_off[g] % 128) { set_error("ani index: base_off must be multiples of 128"); return GALAH_B200_ERR_ARG; }
    std::vector<uint64_t> bo(base_off, base_off + n + 1), co(contig_off, contig_off + n + 1);
    std::vector<uint32_t> cs(contig_start, contig_start + co[n]), cl(contig_len, contig_len + co[n]);
    return ani_add_packed_host(idx, seq2, valid, bo, co, cs, cl);
}

int galah_b200_ani_index_add_packed_device(galah_b200_ani_index_t *idx, const uint32_t *d_seq2,
                                           const uint32_t *d_valid, const uint64_t *d_base_off,
                                           const uint64_t *base_off, const uint64_t *lengths, size_t n,
                                           void *stream) {
    std::lock_guard<std::mutex> lock(g_mu);
    if (int rc = require_ctx()) return rc;
    if (!idx) { set_error("ani index: NULL index"); return GALAH_B200_ERR_ARG; }
    std::vector<uint64_t> bo(base_off, base_off + n + 1), co(n + 1);
    std::vector<uint32_t> cs(n, 0), cl(n);
    for (size_t g = 0; g < n; g++) { co[g] = g; cl[g] = (uint32_t)lengths[g]; }
    co[n] = n;
    return idx->impl.add_packed_device(d_seq2, d_valid, d_base_off, n, bo, co, cs, cl, (cudaStream_t)stream);
}

int galah_b200_ani_index_add_files(galah_b200_ani_index_t *idx, const char *const *paths, size_t n,
                                   int host_threads) {
    std::lock_guard<std::mutex> lock(g_mu);
    if (int rc = require_ctx()) return rc;
    if (!idx) { set_error("ani index: NULL index"); return GALAH_B200_ERR_ARG; }
    IngestSinks sinks;
    sinks.ani = &idx->impl;
    return ingest_files(paths, n, host_threads, sinks);
}

int galah_b200_ani_index_genome(const galah_b200_ani_index_t *idx, size_t g, uint64_t *n_seeds,
                                uint32_t *n_chunks, uint64_t *total_len) {
    if (!idx) { set_error("ani index: NULL index"); return GALAH_B200_ERR_ARG; }
    return idx->impl.genome_info(g, n_seeds, n_chunks, total_len);
}

int galah_b200_ani_index_seeds(const galah_b200_ani_index_t *idx, size_t g, uint32_t *kmer_strand,
                               uint32_t *spread, uint32_t *chunk, size_t cap) {
    std::lock_guard<std::mutex> lock(g_mu);
    if (int rc = require_ctx()) return rc;
    if (!idx) { set_error("ani index: NULL index"); return GALAH_B200_ERR_ARG; }
    return idx->impl.genome_seeds(g, kmer_strand, spread, chunk, cap, g_ctx.stream);
}

int galah_b200_ani_pairs(galah_b200_ani_index_t *idx, const uint32_t *pairs, size_t n_pairs, float min_af_pct,
                         int individual_contigs, galah_b200_ani_result_t *results) {
    std::lock_guard<std::mutex> lock(g_mu);
    if (int rc = require_ctx()) return rc;
    if (!idx) { set_error("ani index: NULL index"); return GALAH_B200_ERR_ARG; }
    static_assert(sizeof(galah_b200_ani_result_t) == sizeof(AniPairResult), "result layout");
    return idx->impl.pairs(pairs, n_pairs, min_af_pct, individual_contigs != 0, reinterpret_cast<AniPairResult *>(results), g_ctx.stream);
}

int galah_b200_ani_index_clear(galah_b200_ani_index_t *idx) {
    std::lock_guard<std::mutex> lock(g_mu);
    if (int rc = require_ctx()) return rc;
    if (!idx) { set_error("ani index: NULL index"); return GALAH_B200_ERR_ARG; }
    idx->impl.clear();
    return 0;
}

int galah_b200_ani_last_timing(const galah_b200_ani_index_t *idx, float *build_ms, float *chain_ms) {
    if (!idx) { set_error("ani index: NULL index"); return GALAH_B200_ERR_ARG; }
    *build_ms = idx->impl.last_build_ms; *chain_ms = idx->impl.last_chain_ms;
    return 0;
}

static int cluster_engine_call(size_t n_genomes, const galah_b200_pair_t *hits, size_t n_hits, int skip_clusterer,
                               float ani_threshold, galah_b200_ani_fn calculate_ani, void *ctx,
                               const AniByHitFn *by_hit, galah_b200_clusters_t *out,
                               const ReversePrefetchFn *prefetch_reverse = nullptr, const AniBatchFn *batch = nullptr,
                               uint32_t max_waves = 16, uint32_t *waves_out = nullptr) {
    if (!out) { set_error("cluster: out is NULL"); return GALAH_B200_ERR_ARG; }
    memset(out, 0, sizeof(*out));
    std::vector<PreclusterHit> h(n_hits);
    for (size_t x = 0; x < n_hits; x++) h[x] = PreclusterHit{hits[x].i, hits[x].j, hits[x].ani};
    AniFn fn;
    if (calculate_ani)
        fn = [calculate_ani, ctx](uint32_t rep, uint32_t genome, float *ani) {
            return calculate_ani(ctx, rep, genome, ani) != 0;
        };
    ClusterResult res;
    std::string err;
    if (cluster_from_hits(n_genomes, h.data(), n_hits, skip_clusterer != 0, ani_threshold, fn, res, err, by_hit, prefetch_reverse,
                          batch, max_waves)) {
        set_error("cluster: " + err);
        return GALAH_B200_ERR_UNSUPPORTED;
    }
    if (waves_out) *waves_out = res.ani_waves;
    out->n_clusters = res.offsets.size() - 1;
    out->members = (uint32_t *)malloc(std::max<size_t>(res.members.size(), 1) * sizeof(uint32_t));
    out->offsets = (uint64_t *)malloc(res.offsets.size() * sizeof(uint64_t));
    if (!out->members || !out->offsets) { set_error("out of host memory"); return GALAH_B200_ERR_ARG; }
    if (!res.members.empty()) memcpy(out->members, res.members.data(), res.members.size() * sizeof(uint32_t));
    memcpy(out->offsets, res.offsets.data(), res.offsets.size() * sizeof(uint64_t));
    out->ani_calls = res.ani_calls;
    out->n_preclusters = res.n_preclusters;
    out->largest_precluster = res.largest_precluster;
    return 0;
}

int galah_b200_cluster_from_distances(size_t n_genomes, const galah_b200_pair_t *hits, size_t n_hits,
                                      int skip_clusterer, float ani_threshold,
                                      galah_b200_ani_fn calculate_ani, void *ctx,
                                      galah_b200_clusters_t *out) {
    return cluster_engine_call(n_genomes, hits, n_hits, skip_clusterer, ani_threshold, calculate_ani, ctx, nullptr, out);
}

int galah_b200_cluster_from_distances_batched(size_t n_genomes, const galah_b200_pair_t *hits, size_t n_hits,
                                              float ani_threshold, galah_b200_ani_batch_fn calculate_ani_batch, void *ctx,
                                              uint32_t max_waves, galah_b200_clusters_t *out, uint32_t *n_waves) {
    if (!calculate_ani_batch) { set_error("cluster_from_distances_batched: NULL callback"); return GALAH_B200_ERR_ARG; }
    std::vector<uint32_t> reps, genomes;
    const AniBatchFn batch = [&](const std::vector<AniRequest> &reqs, uint8_t *some, float *ani) -> int {
        reps.resize(reqs.size()); genomes.resize(reqs.size());
        for (size_t q = 0; q < reqs.size(); q++) { reps[q] = reqs[q].rep; genomes[q] = reqs[q].genome; }
        return calculate_ani_batch(ctx, reps.data(), genomes.data(), reqs.size(), some, ani);
    };
    return cluster_engine_call(n_genomes, hits, n_hits, 0, ani_threshold, nullptr, nullptr, nullptr, out, nullptr, &batch,
                               max_waves ? max_waves : 16, n_waves);
}

// 0: every precluster hit evaluated up front; 1 / -1 (default): in waves
static std::atomic<int> g_lazy_mode{-1};
int galah_b200_cluster_lazy(int mode) {
    if (mode < -1 || mode > 1) { set_error("cluster_lazy: mode must be -1 (auto), 0 (eager) or 1 (waves)"); return GALAH_B200_ERR_ARG; }
    g_lazy_mode.store(mode);
    return 0;
}

// ANI of hit pairs served from a table (hits sorted by (i, j), ani[x] belongs to hits[x]).
// calculate_ani(fasta1 = representative, fasta2 = genome) makes fasta1 the QUERY (skani dist -q
// fasta1 -r fasta2, src/skani.rs:733-744).  The forward table holds query = i (the lower index):
// what the representative search asks for (representatives precede the genome, src/clusterer.rs:
// 229-231).  The membership pass can ask for a representative with a HIGHER index than the genome
// (src/clusterer.rs:375-384): that is the reverse orientation, served from ani_rev when present
// (the forward value otherwise).  Lookups are pure reads: the engine calls them from several threads.
namespace {
struct AniTable {
    const galah_b200_pair_t *hits; size_t n; const float *ani; size_t stride;
    const float *ani_rev = nullptr;            // [n] or null
    const uint8_t *have_rev = nullptr;         // [n] or null: ani_rev[x] is valid
};
// The engine hands over the index of the hit the pair belongs to: no search.
bool ani_table_by_hit(const AniTable &t, uint32_t rep, uint32_t genome, size_t hit, float *ani) {
    if (rep > genome && t.ani_rev && (!t.have_rev || t.have_rev[hit])) { *ani = t.ani_rev[hit]; return true; }
    *ani = *(const float *)((const char *)t.ani + hit * t.stride);  // skani never yields None (src/skani.rs:760)
    return true;
}
int cluster_from_table(size_t n_genomes, const galah_b200_pair_t *hits, size_t n_hits, float ani_threshold,
                       const AniTable &table, galah_b200_clusters_t *out, const ReversePrefetchFn *prefetch_reverse = nullptr) {
    const AniByHitFn fn = [&table](uint32_t rep, uint32_t genome, size_t hit, float *ani) {
        return ani_table_by_hit(table, rep, genome, hit, ani);
    };
    return cluster_engine_call(n_genomes, hits, n_hits, 0, ani_threshold, nullptr, nullptr, &fn, out, prefetch_reverse);
}
}  // namespace

static int check_sorted_hits(const galah_b200_pair_t *hits, size_t n_hits) {
    for (size_t x = 1; x < n_hits; x++)
        if (hits[x - 1].i > hits[x].i || (hits[x - 1].i == hits[x].i && hits[x - 1].j >= hits[x].j)) {
            set_error("cluster_from_ani_table: hits must be sorted by (i, j) without duplicates");
            return GALAH_B200_ERR_ARG;
        }
    return 0;
}

int galah_b200_cluster_from_ani_table(size_t n_genomes, const galah_b200_pair_t *hits, size_t n_hits,
                                      const float *ani, float ani_threshold, galah_b200_clusters_t *out) {
    if (n_hits && (!hits || !ani)) { set_error("cluster_from_ani_table: NULL hits / ani"); return GALAH_B200_ERR_ARG; }
    if (int rc = check_sorted_hits(hits, n_hits)) return rc;
    AniTable table{hits, n_hits, ani, sizeof(float)};
    return cluster_from_table(n_genomes, hits, n_hits, ani_threshold, table, out);
}

int galah_b200_cluster_from_ani_tables(size_t n_genomes, const galah_b200_pair_t *hits, size_t n_hits,
                                       const float *ani_fwd, const float *ani_rev, float ani_threshold,
                                       galah_b200_clusters_t *out) {
    if (n_hits && (!hits || !ani_fwd || !ani_rev)) { set_error("cluster_from_ani_tables: NULL argument"); return GALAH_B200_ERR_ARG; }
    if (int rc = check_sorted_hits(hits, n_hits)) return rc;
    AniTable table{hits, n_hits, ani_fwd, sizeof(float)};
    table.ani_rev = ani_rev;
    return cluster_from_table(n_genomes, hits, n_hits, ani_threshold, table, out);
}

// Stages 2 + 3 of the whole path once the sketch table (device) and the K3 index are in place:
// K2 on the resident table -> f64 finish -> K3 on every hit -> greedy engine.  Caller holds g_mu.
static int cluster_from_resident(const uint64_t *d_table, const uint32_t *d_counts, size_t n, AniIndex &index,
                                 float precluster_min_ani, float ani_threshold_pct, float min_af_pct,
                                 galah_b200_clusters_t *out, galah_b200_cluster_stats_t *stats) {
    const uint32_t s = 1000;  // FinchPreclusterer{num_kmers: 1000, kmer_length: 21}, cluster_argument_parsing.rs:1301-1302
    const uint8_t k = 21;
    cudaStream_t st = g_ctx.stream;
    const double t0 = now_ms();
    galah_b200_pair_t *hits = nullptr;
    size_t n_hits = 0;
    if (int rc = run_prefilter(d_table, d_counts, n, s, k, precluster_min_ani, 0, 1, st, &hits, &n_hits)) return rc;
    struct HitGuard { galah_b200_pair_t *h; ~HitGuard() { free(h); } } hit_guard{hits};
    const double t1 = now_ms();
    // Stage 2 in waves (default): only the (representative, genome) pairs the reference's two passes evaluate
    // (src/clusterer.rs:216-300, 350-449), a batch per wave, all preclusters together.  A collection of
    // near-identical genomes (galah's stated use case) needs representatives x genomes evaluations instead of
    // one per hit; synthetic families of 10 need about a third of their hits' two orientations, in 3 waves.
    // galah_b200_cluster_lazy(0) evaluates every hit up front instead.  Same clusters either way.
    const int lazy_mode = g_lazy_mode.load();
    if (lazy_mode != 0) {
        double ani_ms = 0.0;
        float chain_ms = 0.f;
        size_t n_asked = 0;
        std::vector<uint32_t> qp;
        std::vector<AniPairResult> qres;
        const AniBatchFn batch = [&](const std::vector<AniRequest> &reqs, uint8_t *some, float *ani) -> int {
            const double ta = now_ms();
            qp.resize(2 * reqs.size()); qres.resize(reqs.size());
            // calculate_ani(fasta1 = representative, fasta2 = genome): the representative is the query
            for (size_t q = 0; q < reqs.size(); q++) { qp[2 * q] = reqs[q].rep; qp[2 * q + 1] = reqs[q].genome; }
            if (int rc2 = index.pairs(qp.data(), reqs.size(), min_af_pct, false, qres.data(), st)) return rc2;
            for (size_t q = 0; q < reqs.size(); q++) { some[q] = 1; ani[q] = qres[q].ani; }  // skani never yields None (src/skani.rs:760)
            chain_ms += index.last_chain_ms;
            n_asked += reqs.size();
            ani_ms += now_ms() - ta;
            return 0;
        };
        uint32_t waves = 0;
        int rc = cluster_engine_call(n, hits, n_hits, 0, ani_threshold_pct, nullptr, nullptr, nullptr, out, nullptr, &batch, 16, &waves);
        const double t3 = now_ms();
        if (getenv("GALAH_B200_DEBUG"))
            fprintf(stderr, "[cluster_from_resident] prefilter %.2f ms, ani in %u waves %.2f ms (%zu of %zu hit pairs), engine %.2f ms\n",
                    t1 - t0, waves, ani_ms, n_asked, n_hits, (t3 - t1) - ani_ms);
        if (stats) {
            stats->n_precluster_hits = n_hits; stats->n_ani_pairs = n_asked; stats->ani_chain_ms = chain_ms;
            stats->prefilter_ms = (float)(t1 - t0); stats->ani_ms = (float)ani_ms; stats->engine_ms = (float)((t3 - t1) - ani_ms);
            stats->ani_waves = waves;
        }
        return rc;
    }
    // stage 2 for every precluster hit (a superset of what the reference's find_any evaluates;
    // the greedy decisions only ever read values of hit pairs, so the clusters are the same);
    // query = the lower index: galah passes the representative first (src/clusterer.rs:262-296)
    std::vector<uint32_t> pairs(2 * n_hits);
    for (size_t x = 0; x < n_hits; x++) { pairs[2 * x] = hits[x].i; pairs[2 * x + 1] = hits[x].j; }
    std::vector<AniPairResult> res(n_hits);
    if (int rc = index.pairs(pairs.data(), n_hits, min_af_pct, false, res.data(), st)) return rc;
    const double t2 = now_ms();
    float chain_ms = index.last_chain_ms;
    // The membership sweep asks for representatives that come AFTER the genome: those pairs have the
    // representative as the query.  The engine reports them once every representative is known (its
    // search only reads forward values); exactly those pairs get one more K3 launch, then the
    // engine goes on to the memberships -- one engine pass, two K3 launches.
    AniTable table{hits, n_hits, n_hits ? &res[0].ani : nullptr, sizeof(AniPairResult)};
    std::vector<float> ani_rev(n_hits, 0.f);
    std::vector<uint8_t> have_rev(n_hits, 0);
    table.ani_rev = ani_rev.data(); table.have_rev = have_rev.data();
    double rev_ms = 0.0;
    size_t n_rev = 0;
    const ReversePrefetchFn prefetch = [&](const std::vector<size_t> &want) -> int {
        const double ta = now_ms();
        std::vector<uint32_t> rp(2 * want.size());
        for (size_t x = 0; x < want.size(); x++) { rp[2 * x] = hits[want[x]].j; rp[2 * x + 1] = hits[want[x]].i; }
        std::vector<AniPairResult> rres(want.size());
        if (int rc2 = index.pairs(rp.data(), want.size(), min_af_pct, false, rres.data(), st)) return rc2;
        chain_ms += index.last_chain_ms;
        for (size_t x = 0; x < want.size(); x++) { ani_rev[want[x]] = rres[x].ani; have_rev[want[x]] = 1; }
        n_rev = want.size();
        rev_ms = now_ms() - ta;
        return 0;
    };
    int rc = cluster_from_table(n, hits, n_hits, ani_threshold_pct, table, out, &prefetch);
    const double t3 = now_ms();
    const double engine_ms = (t3 - t2) - rev_ms;  // the reverse launch belongs to the ANI phase
    const double ani_total_ms = (t2 - t1) + rev_ms;
    if (stats) stats->n_ani_pairs = n_hits + n_rev;
    if (getenv("GALAH_B200_DEBUG"))
        fprintf(stderr, "[cluster_from_resident] prefilter %.2f ms, ani (both launches) %.2f ms, engine %.2f ms, "
                "%zu hits, %zu reverse requests\n", t1 - t0, ani_total_ms, engine_ms, n_hits, n_rev);
    if (stats) {
        stats->n_precluster_hits = n_hits;
        if (stats->n_ani_pairs == 0) stats->n_ani_pairs = n_hits;
        stats->ani_chain_ms = chain_ms;
        stats->prefilter_ms = (float)(t1 - t0); stats->ani_ms = (float)ani_total_ms; stats->engine_ms = (float)engine_ms;
    }
    return rc;
}

int galah_b200_cluster_files(const char *const *paths, size_t n, float precluster_min_ani, float ani_threshold_pct,
                             float min_af_pct, int small_genomes, int host_threads,
                             galah_b200_clusters_t *out, galah_b200_cluster_stats_t *stats) {
    if (!out) { set_error("cluster_files: out is NULL"); return GALAH_B200_ERR_ARG; }
    memset(out, 0, sizeof(*out));
    if (stats) memset(stats, 0, sizeof(*stats));
    // SkaniClusterer::initialise (src/skani.rs:696-698) asserts a percentage
    if (!(ani_threshold_pct > 1.0f)) { set_error("assertion failed: self.threshold > 1.0"); return GALAH_B200_ERR_UNSUPPORTED; }
    std::lock_guard<std::mutex> lock(g_mu);
    if (int rc = require_ctx()) return rc;
    const double t_begin = now_ms();
    const uint32_t s = 1000;
    AniIndex &index = g_ctx.pipeline_index(small_genomes != 0);
    // the sketch table is written by K1 where K2 reads it: it never visits the host
    if (ws_ensure(g_ctx.d_table, g_ctx.cap_table, std::max<size_t>(n, 1) * s) ||
        ws_ensure(g_ctx.d_counts, g_ctx.cap_counts, std::max<size_t>(n, 1)))
        return GALAH_B200_ERR_CUDA;
    IngestSinks sinks;
    sinks.sketch = true; sinks.k = 21; sinks.s = s; sinks.seed = 0;
    sinks.d_hashes = g_ctx.d_table; sinks.d_counts = g_ctx.d_counts; sinks.ani = &index;
    if (int rc = ingest_files(paths, n, host_threads, sinks)) return rc;
    const double t_ingest = now_ms();
    int rc = cluster_from_resident(g_ctx.d_table, g_ctx.d_counts, n, index, precluster_min_ani, ani_threshold_pct,
                                   min_af_pct, out, stats);
    if (stats) { stats->ingest_ms = (float)(t_ingest - t_begin); stats->total_ms = (float)(now_ms() - t_begin); }
    return rc;
}

// Genomes that are already PACKED (K1 / K3 layout, one contig per genome) -> K1 sketch rows in the
// device table d_table / d_counts (stride 1000) and the K3 index.  `device`: the arrays are resident
// in HBM (d_base_off = device copy of base_off); else host arrays, uploaded in batches of ~1 G
// bases on the copy stream while the previous batch is sketched and indexed.  Caller holds g_mu.
// marker_c != 0: the rows are FracMinHash MARKER sketches (k = 21, density 1 / marker_c, row stride
// marker_stride) for the skani-style screen instead of bottom-1000 MinHash sketches.
// Host input without a validity bitmap (valid == nullptr): every base of a genome's length is valid but the
// listed ranges (absolute base coordinates, half open, sorted by begin); the bitmap of a batch is then made
// on the device and a third of the bytes stays off PCIe.
struct SparseValidity { const uint64_t *begin, *end; size_t n; };

static int ingest_packed(const uint32_t *seq2, const uint32_t *valid, const uint64_t *d_base_off_or_null,
                         const uint64_t *base_off, const uint64_t *lengths, size_t n, bool device, uint64_t *d_table,
                         uint32_t *d_counts, AniIndex &index, float *sketch_ms_out, float *index_ms_out,
                         uint32_t marker_c = 0, uint32_t marker_stride = 0, const SparseValidity *sparse = nullptr) {
    if (!device && !valid && !sparse) { set_error("packed genomes: no validity bitmap and no invalid-range list"); return GALAH_B200_ERR_ARG; }
    for (size_t g = 0; g <= n; g++)
        if (base_off[g] % 128) { set_error("packed genomes: base_off must be multiples of 128"); return GALAH_B200_ERR_ARG; }
    const uint32_t s = 1000;
    cudaStream_t st = g_ctx.stream;
    // batches of about 1 G bases (375 MB packed): bounds the K1 / K3 scratch, and is the upload unit
    const uint64_t kBatchBases = 1ull << 30;  // (0.5 .. 4 G bases per batch measured within 1 % of each other)
    std::vector<size_t> cut{0};
    while (cut.back() < n) {
        size_t b1 = cut.back() + 1;
        while (b1 < n && base_off[b1 + 1] - base_off[cut.back()] <= kBatchBases) b1++;
        cut.push_back(b1);
    }
    const size_t n_batches = cut.size() - 1;
    // host input: two staging slots on the device, filled on the copy stream
    // (the slots live in the context and only ever grow: no cudaMalloc / cudaFree per call)
    struct Slot { uint32_t *seq2 = nullptr, *valid = nullptr; uint64_t *off = nullptr, *len = nullptr, *ranges = nullptr;
                  cudaEvent_t landed = nullptr, freed = nullptr; } slot[2];
    struct SlotGuard { Slot *s; ~SlotGuard() { for (int x = 0; x < 2; x++) {
                       if (s[x].landed) cudaEventDestroy(s[x].landed); if (s[x].freed) cudaEventDestroy(s[x].freed); } } } slot_guard{slot};
    uint64_t max_bases = 0; size_t max_n = 0;
    for (size_t b = 0; b < n_batches; b++) {
        max_bases = std::max(max_bases, base_off[cut[b + 1]] - base_off[cut[b]]);
        max_n = std::max(max_n, cut[b + 1] - cut[b]);
    }
    if (!device) {
        if (!g_ctx.copy_stream) GB_CUDA(cudaStreamCreateWithFlags(&g_ctx.copy_stream, cudaStreamNonBlocking));
        for (int x = 0; x < 2; x++) {
            if (ws_ensure(g_ctx.slot_seq2[x], g_ctx.cap_slot_seq2[x], max_bases / 16 + 8) ||
                ws_ensure(g_ctx.slot_valid[x], g_ctx.cap_slot_valid[x], max_bases / 32 + 8) ||
                ws_ensure(g_ctx.slot_off[x], g_ctx.cap_slot_off[x], max_n + 1) ||
                (sparse && ws_ensure(g_ctx.slot_len[x], g_ctx.cap_slot_len[x], max_n + 1)))
                return GALAH_B200_ERR_CUDA;
            slot[x].seq2 = g_ctx.slot_seq2[x]; slot[x].valid = g_ctx.slot_valid[x]; slot[x].off = g_ctx.slot_off[x];
            slot[x].len = g_ctx.slot_len[x];
            GB_CUDA(cudaEventCreateWithFlags(&slot[x].landed, cudaEventDisableTiming));
            GB_CUDA(cudaEventCreateWithFlags(&slot[x].freed, cudaEventDisableTiming));
        }
    }
    std::vector<std::vector<uint64_t>> rel(n_batches), rel_ranges(sparse ? n_batches : 0);
    auto upload = [&](size_t b) -> int {
        Slot &sl = slot[b & 1];
        const size_t g0 = cut[b], nb = cut[b + 1] - g0;
        const uint64_t first = base_off[g0], total = base_off[g0 + nb] - first;
        rel[b].resize(nb + 1);
        for (size_t g = 0; g <= nb; g++) rel[b][g] = base_off[g0 + g] - first;
        if (b >= 2) GB_CUDA(cudaStreamWaitEvent(g_ctx.copy_stream, sl.freed, 0));  // the slot's previous batch is consumed
        GB_CUDA(cudaMemcpyAsync(sl.seq2, seq2 + first / 16, total / 16 * 4, cudaMemcpyHostToDevice, g_ctx.copy_stream));
        GB_CUDA(cudaMemcpyAsync(sl.off, rel[b].data(), (nb + 1) * 8, cudaMemcpyHostToDevice, g_ctx.copy_stream));
        if (valid) {
            GB_CUDA(cudaMemcpyAsync(sl.valid, valid + first / 32, total / 32 * 4, cudaMemcpyHostToDevice, g_ctx.copy_stream));
        } else {
            // the batch's invalid ranges, relative to its first base (they are sorted: two binary searches)
            const uint64_t last = first + total;
            const size_t r0 = std::lower_bound(sparse->end, sparse->end + sparse->n, first + 1) - sparse->end;
            const size_t r1 = std::lower_bound(sparse->begin, sparse->begin + sparse->n, last) - sparse->begin;
            std::vector<uint64_t> &rr = rel_ranges[b];
            rr.clear();
            for (size_t x = r0; x < r1; x++) {
                const uint64_t rb = std::max(sparse->begin[x], first), re = std::min(sparse->end[x], last);
                if (rb < re) { rr.push_back(rb - first); rr.push_back(re - first); }
            }
            int which = (int)(b & 1);
            if (ws_ensure(g_ctx.slot_ranges[which], g_ctx.cap_slot_ranges[which], rr.size() + 2)) return GALAH_B200_ERR_CUDA;
            sl.ranges = g_ctx.slot_ranges[which];
            GB_CUDA(cudaMemcpyAsync(sl.len, lengths + g0, nb * 8, cudaMemcpyHostToDevice, g_ctx.copy_stream));
            if (!rr.empty()) GB_CUDA(cudaMemcpyAsync(sl.ranges, rr.data(), rr.size() * 8, cudaMemcpyHostToDevice, g_ctx.copy_stream));
            if (int rc = validity_from_ranges_enqueue(sl.valid, sl.off, sl.len, nb, sl.ranges, rr.size() / 2, g_ctx.copy_stream)) return rc;
        }
        GB_CUDA(cudaEventRecord(sl.landed, g_ctx.copy_stream));
        return 0;
    };
    // Two streams, software-pipelined over the batches: the K1 scan (+ K3 seed marks) of batch b + 1 is
    // enqueued on scan_stream BEFORE the host walks through the K3 index build of batch b on the
    // main stream (that build has host round trips: seed counts down, offsets up), so the GPU always
    // has the next scan queued while the host is busy.  Seed-selection bits are double-buffered.
    float sketch_ms = 0.f, index_ms = 0.f;
    if (!g_ctx.scan_stream) {
        int prio_lo = 0, prio_hi = 0;
        GB_CUDA(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
        GB_CUDA(cudaStreamCreateWithPriority(&g_ctx.scan_stream, cudaStreamNonBlocking, prio_lo));
    }
    for (auto &e : g_ctx.ev_scan) if (!e) GB_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    cudaStream_t sk = g_ctx.scan_stream;
    cudaEvent_t ev[2][2], ev_ix[2];  // [slot]: scan begin / end on sk; index begin / end on st
    for (auto &pr : ev) for (auto &e : pr) GB_CUDA(cudaEventCreate(&e));
    for (auto &e : ev_ix) GB_CUDA(cudaEventCreate(&e));
    struct EvGuard { cudaEvent_t (*e)[2]; cudaEvent_t *x; ~EvGuard() {
        for (int a = 0; a < 2; a++) for (int b = 0; b < 2; b++) cudaEventDestroy(e[a][b]);
        cudaEventDestroy(x[0]); cudaEventDestroy(x[1]); } } ev_guard{ev, ev_ix};
    // the main stream may still be working for the caller: the first scan starts behind it
    GB_CUDA(cudaEventRecord(ev_ix[0], st));
    GB_CUDA(cudaStreamWaitEvent(sk, ev_ix[0], 0));
    struct Batch { const uint32_t *seq2, *valid; const uint64_t *off; std::vector<uint64_t> bo; SeedSink sink; } bt[2];
    auto scan = [&](size_t b) -> int {
        Batch &B = bt[b & 1];
        const size_t g0 = cut[b], nb = cut[b + 1] - g0;
        if (device) {
            B.seq2 = seq2; B.valid = valid; B.off = d_base_off_or_null + g0;  // absolute offsets into the resident arrays
            B.bo.assign(base_off + g0, base_off + g0 + nb + 1);
        } else {
            GB_CUDA(cudaStreamWaitEvent(sk, slot[b & 1].landed, 0));
            B.seq2 = slot[b & 1].seq2; B.valid = slot[b & 1].valid; B.off = slot[b & 1].off;
            B.bo = rel[b];
        }
        // the k = 21 scan also marks the K3 seeds of the batch
        uint32_t *&d_sel = (b & 1) ? g_ctx.d_sel2 : g_ctx.d_sel;
        size_t &cap_sel = (b & 1) ? g_ctx.cap_sel2 : g_ctx.cap_sel;
        if (ws_ensure(d_sel, cap_sel, (size_t)((B.bo.back() - B.bo.front()) / 32 + 2))) return GALAH_B200_ERR_CUDA;
        if (ws_ensure(g_ctx.d_seedcnt[b & 1], g_ctx.cap_seedcnt[b & 1], nb)) return GALAH_B200_ERR_CUDA;
        B.sink = SeedSink{d_sel, B.bo.front(), index.seed_threshold(), g_ctx.d_seedcnt[b & 1]};
        const uint64_t span[2] = {B.bo.front(), B.bo.back()};
        GB_CUDA(cudaEventRecord(ev[b & 1][0], sk));
        if (marker_c) {
            if (int rc = marker_sketch_enqueue(g_ctx.sws, B.seq2, B.valid, B.off, nb, 21, marker_c, marker_stride,
                                               d_table + g0 * (size_t)marker_stride, d_counts + g0, sk, &B.sink, span))
                return rc;
        } else if (int rc = sketch_enqueue(g_ctx.sws, B.seq2, B.valid, B.off, nb, 21, s, 0, d_table + g0 * (size_t)s,
                                           d_counts + g0, s, sk, &B.sink, span))
            return rc;
        GB_CUDA(cudaEventRecord(ev[b & 1][1], sk));
        GB_CUDA(cudaEventRecord(g_ctx.ev_scan[b & 1], sk));
        return 0;
    };
    // host input: batches b + 1 and b + 2 cross PCIe behind the kernels of batch b (two slots)
    if (!device) for (size_t b = 0; b < std::min<size_t>(n_batches, 2); b++) if (int rc = upload(b)) return rc;
    if (n_batches) if (int rc = scan(0)) return rc;
    for (size_t b = 0; b < n_batches; b++) {
        const size_t g0 = cut[b], nb = cut[b + 1] - g0;
        if (b + 1 < n_batches) if (int rc = scan(b + 1)) return rc;
        Batch &B = bt[b & 1];
        std::vector<uint64_t> co(nb + 1);
        std::vector<uint32_t> cs(nb, 0), cl(nb);
        for (size_t g = 0; g < nb; g++) { co[g] = g; cl[g] = (uint32_t)lengths[g0 + g]; }
        co[nb] = nb;
        GB_CUDA(cudaStreamWaitEvent(st, g_ctx.ev_scan[b & 1], 0));
        GB_CUDA(cudaEventRecord(ev_ix[0], st));
        if (int rc = index.add_packed_device(B.seq2, B.valid, B.off, nb, B.bo, co, cs, cl, st, B.sink.d_sel, B.sink.d_seed_count)) return rc;
        if (b == 0 && n_batches > 1) if (int rc = index.reserve_for(n, st)) return rc;
        GB_CUDA(cudaEventRecord(ev_ix[1], st));
        if (!device) {
            GB_CUDA(cudaEventRecord(slot[b & 1].freed, st));  // scan(b) ended before index(b) began: the slot is consumed
            if (b + 2 < n_batches) if (int rc = upload(b + 2)) return rc;
        }
        GB_CUDA(cudaEventSynchronize(ev_ix[1]));
        float a = 0.f, c = 0.f;
        GB_CUDA(cudaEventElapsedTime(&a, ev[b & 1][0], ev[b & 1][1]));
        GB_CUDA(cudaEventElapsedTime(&c, ev_ix[0], ev_ix[1]));
        sketch_ms += a; index_ms += c;
        if (getenv("GALAH_B200_DEBUG"))
            fprintf(stderr, "[ingest_packed] batch %zu/%zu: %zu genomes, K1 %.2f ms, index %.2f ms (kernels %.2f ms)\n", b,
                    n_batches, nb, a, c, index.last_build_ms);
    }
    // the sketch rows are complete when scan_stream is: later work on the main stream waits for it
    GB_CUDA(cudaEventRecord(ev_ix[0], sk));
    GB_CUDA(cudaStreamWaitEvent(st, ev_ix[0], 0));
    if (sketch_ms_out) *sketch_ms_out = sketch_ms;
    if (index_ms_out) *index_ms_out = index_ms;
    return 0;
}

static int check_sparse(const uint64_t *begin, const uint64_t *end, size_t n_inv) {
    if (n_inv && (!begin || !end)) { set_error("invalid ranges: NULL arrays"); return GALAH_B200_ERR_ARG; }
    for (size_t x = 0; x < n_inv; x++)
        if (begin[x] >= end[x] || (x && begin[x] < end[x - 1])) {
            set_error("invalid ranges: must be non-empty, sorted and disjoint");
            return GALAH_B200_ERR_ARG;
        }
    return 0;
}

static int cluster_packed(const uint32_t *seq2, const uint32_t *valid, const uint64_t *d_base_off_or_null,
                          const uint64_t *base_off, const uint64_t *lengths, size_t n, bool device,
                          float precluster_min_ani, float ani_threshold_pct, float min_af_pct, int small_genomes,
                          galah_b200_clusters_t *out, galah_b200_cluster_stats_t *stats,
                          const SparseValidity *sparse = nullptr) {
    if (!out) { set_error("cluster_packed: out is NULL"); return GALAH_B200_ERR_ARG; }
    memset(out, 0, sizeof(*out));
    if (stats) memset(stats, 0, sizeof(*stats));
    if (!(ani_threshold_pct > 1.0f)) { set_error("assertion failed: self.threshold > 1.0"); return GALAH_B200_ERR_UNSUPPORTED; }
    std::lock_guard<std::mutex> lock(g_mu);
    if (int rc = require_ctx()) return rc;
    const double t_begin = now_ms();
    AniIndex &index = g_ctx.pipeline_index(small_genomes != 0);
    if (ws_ensure(g_ctx.d_table, g_ctx.cap_table, std::max<size_t>(n, 1) * 1000) ||
        ws_ensure(g_ctx.d_counts, g_ctx.cap_counts, std::max<size_t>(n, 1)))
        return GALAH_B200_ERR_CUDA;
    float sketch_ms = 0.f, index_ms = 0.f;
    if (int rc = ingest_packed(seq2, valid, d_base_off_or_null, base_off, lengths, n, device, g_ctx.d_table,
                               g_ctx.d_counts, index, &sketch_ms, &index_ms, 0, 0, sparse))
        return rc;
    const double t_ingest = now_ms();
    int rc = cluster_from_resident(g_ctx.d_table, g_ctx.d_counts, n, index, precluster_min_ani, ani_threshold_pct,
                                   min_af_pct, out, stats);
    if (stats) {
        stats->ingest_ms = (float)(t_ingest - t_begin); stats->total_ms = (float)(now_ms() - t_begin);
        stats->sketch_ms = sketch_ms; stats->index_ms = index_ms;
    }
    return rc;
}

int galah_b200_ingest_packed(const uint32_t *seq2, const uint32_t *valid, const uint64_t *d_base_off,
                             const uint64_t *base_off, const uint64_t *lengths, size_t n, int device,
                             uint64_t *d_hashes, uint32_t *d_counts, galah_b200_ani_index_t *idx, float *ms2) {
    std::lock_guard<std::mutex> lock(g_mu);
    if (int rc = require_ctx()) return rc;
    if (!idx || !d_hashes || !d_counts) { set_error("ingest_packed: NULL argument"); return GALAH_B200_ERR_ARG; }
    float a = 0.f, b = 0.f;
    int rc = ingest_packed(seq2, valid, d_base_off, base_off, lengths, n, device != 0, d_hashes, d_counts, idx->impl, &a, &b);
    if (ms2) { ms2[0] = a; ms2[1] = b; }
    return rc;
}

int galah_b200_ingest_packed_sparse(const uint32_t *seq2, const uint64_t *invalid_begin, const uint64_t *invalid_end,
                                    size_t n_invalid, const uint64_t *base_off, const uint64_t *lengths, size_t n,
                                    uint64_t *d_hashes, uint32_t *d_counts, galah_b200_ani_index_t *idx, float *ms2) {
    if (int rc = check_sparse(invalid_begin, invalid_end, n_invalid)) return rc;
    std::lock_guard<std::mutex> lock(g_mu);
    if (int rc = require_ctx()) return rc;
    if (!idx || !d_hashes || !d_counts) { set_error("ingest_packed_sparse: NULL argument"); return GALAH_B200_ERR_ARG; }
    const SparseValidity sv{invalid_begin, invalid_end, n_invalid};
    float a = 0.f, b = 0.f;
    int rc = ingest_packed(seq2, nullptr, nullptr, base_off, lengths, n, false, d_hashes, d_counts, idx->impl, &a, &b, 0, 0, &sv);
    if (ms2) { ms2[0] = a; ms2[1] = b; }
    return rc;
}

int galah_b200_cluster_packed_sparse(const uint32_t *seq2, const uint64_t *invalid_begin, const uint64_t *invalid_end,
                                     size_t n_invalid, const uint64_t *base_off, const uint64_t *lengths, size_t n,
                                     float precluster_min_ani, float ani_threshold_pct, float min_af_pct, int small_genomes,
                                     galah_b200_clusters_t *out, galah_b200_cluster_stats_t *stats) {
    if (int rc = check_sparse(invalid_begin, invalid_end, n_invalid)) return rc;
    const SparseValidity sv{invalid_begin, invalid_end, n_invalid};
    return cluster_packed(seq2, nullptr, nullptr, base_off, lengths, n, false, precluster_min_ani, ani_threshold_pct,
                          min_af_pct, small_genomes, out, stats, &sv);
}

int galah_b200_ingest_packed_markers(const uint32_t *seq2, const uint32_t *valid, const uint64_t *d_base_off,
                                     const uint64_t *base_off, const uint64_t *lengths, size_t n, int device,
                                     uint32_t marker_stride, uint64_t *d_rows, uint32_t *d_counts,
                                     galah_b200_ani_index_t *idx, float *ms2) {
    std::lock_guard<std::mutex> lock(g_mu);
    if (int rc = require_ctx()) return rc;
    if (!idx || !d_rows || !d_counts) { set_error("ingest_packed_markers: NULL argument"); return GALAH_B200_ERR_ARG; }
    float a = 0.f, b = 0.f;
    int rc = ingest_packed(seq2, valid, d_base_off, base_off, lengths, n, device != 0, d_rows, d_counts, idx->impl, &a, &b,
                           idx->impl.c() == 30u ? 200u : 1000u, marker_stride);
    if (ms2) { ms2[0] = a; ms2[1] = b; }
    return rc;
}

uint32_t galah_b200_marker_row_capacity(uint64_t longest_unit, int small_genomes) {
    return marker_row_capacity(longest_unit, small_genomes ? 200u : 1000u);
}

int galah_b200_ani_index_export_tables(const galah_b200_ani_index_t *idx, uint8_t handle[64], uint64_t *table_off,
                                       uint64_t *total_len) {
    std::lock_guard<std::mutex> lock(g_mu);
    if (int rc = require_ctx()) return rc;
    if (!idx) { set_error("ani index: NULL index"); return GALAH_B200_ERR_ARG; }
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    cudaIpcMemHandle_t h;
    std::vector<uint64_t> to, tl;
    if (int rc = idx->impl.export_tables(&h, to, tl)) return rc;
    memcpy(handle, &h, 64);
    memcpy(table_off, to.data(), to.size() * 8);
    memcpy(total_len, tl.data(), tl.size() * 8);
    return 0;
}

int galah_b200_ani_index_attach_peer(galah_b200_ani_index_t *idx, const uint8_t handle[64], const uint64_t *table_off,
                                     const uint64_t *total_len, size_t n_genomes, uint32_t *first_id) {
    std::lock_guard<std::mutex> lock(g_mu);
    if (int rc = require_ctx()) return rc;
    if (!idx || !first_id) { set_error("ani index: NULL argument"); return GALAH_B200_ERR_ARG; }
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, 64);
    return idx->impl.attach_peer(h, table_off, total_len, n_genomes, first_id);
}

int galah_b200_cluster_packed_device(const uint32_t *d_seq2, const uint32_t *d_valid, const uint64_t *d_base_off,
                                     const uint64_t *base_off, const uint64_t *lengths, size_t n,
                                     float precluster_min_ani, float ani_threshold_pct, float min_af_pct,
                                     int small_genomes, galah_b200_clusters_t *out, galah_b200_cluster_stats_t *stats) {
    return cluster_packed(d_seq2, d_valid, d_base_off, base_off, lengths, n, true, precluster_min_ani,
                          ani_threshold_pct, min_af_pct, small_genomes, out, stats);
}

int galah_b200_cluster_packed(const uint32_t *seq2, const uint32_t *valid, const uint64_t *base_off,
                              const uint64_t *lengths, size_t n, float precluster_min_ani, float ani_threshold_pct,
                              float min_af_pct, int small_genomes, galah_b200_clusters_t *out,
                              galah_b200_cluster_stats_t *stats) {
    return cluster_packed(seq2, valid, nullptr, base_off, lengths, n, false, precluster_min_ani, ani_threshold_pct,
                          min_af_pct, small_genomes, out, stats);
}

// ------------------------------------------------------------------------------------------
// The whole path on G devices of ONE process (what a single galah process with several GPUs calls):
// one host thread per device, pinned to that device's context.
//   slices     device r owns the genomes [g0[r], g0[r+1]) (equal counts, whole K2 row blocks)
//   K1 + index every device sketches and indexes ITS slice (host buffers uploaded in batches)
//   exchange   every device copies the other devices' sketch rows into its own table with
//              cudaMemcpyPeerAsync (copy engines over NVLink; no kernel, no host staging)
//   K2         every device builds the block lists and joins ITS row-block shard of the pair grid
//              (run_prefilter's shard / n_shards: the same boustrophedon split as the multi-process path)
//   hits       merged and sorted on the host (they are 20 bytes each)
//   K3         both orientations of every hit, each on the device that owns its QUERY genome; a
//              reference genome on another device is read IN PLACE through its hash table's device
//              pointer (peer access, unified addressing)
//   engine     thread 0
// ------------------------------------------------------------------------------------------
namespace {
struct MultiBarrier {
    std::mutex mu; std::condition_variable cv; int n, waiting = 0; uint64_t phase = 0;
    explicit MultiBarrier(int n_) : n(n_) {}
    void wait() {
        std::unique_lock<std::mutex> lk(mu);
        const uint64_t my = phase;
        if (++waiting == n) { waiting = 0; phase++; cv.notify_all(); }
        else cv.wait(lk, [&] { return phase != my; });
    }
};
}  // namespace

// ingest(r, first genome, genome count, sketch rows, counts, index, K1 ms, index ms): device r's slice -> its rows of
// its own sketch table + its K3 index; runs on the worker thread pinned to device r, which holds that device's lock
using MultiIngest = std::function<int(int, size_t, size_t, uint64_t *, uint32_t *, AniIndex &, float *, float *)>;

static int cluster_multi(size_t n, int n_devices, const MultiIngest &ingest, float precluster_min_ani,
                         float ani_threshold_pct, float min_af_pct, int small_genomes,
                         galah_b200_clusters_t *out, galah_b200_cluster_stats_t *stats) {
    if (!out) { set_error("cluster_packed_multi: out is NULL"); return GALAH_B200_ERR_ARG; }
    memset(out, 0, sizeof(*out));
    if (stats) memset(stats, 0, sizeof(*stats));
    if (!(ani_threshold_pct > 1.0f)) { set_error("assertion failed: self.threshold > 1.0"); return GALAH_B200_ERR_UNSUPPORTED; }
    const int G = n_devices;
    if (G < 1 || G > kMaxDevices) { set_error("cluster_packed_multi: bad device count"); return GALAH_B200_ERR_ARG; }
    for (int d = 0; d < G; d++)
        if (g_ctxs[d].device != d) { set_error("cluster_packed_multi: call galah_b200_init_devices(n_devices) first"); return GALAH_B200_ERR_NO_DEVICE; }
    const double t_begin = now_ms();
    const uint32_t s = 1000;
    // slices: equal genome counts, whole row blocks of the K2 grid (the last slice takes what is left)
    size_t per = (n + (size_t)G - 1) / (size_t)G;
    per = (per + GALAH_B200_ROW_BLOCK - 1) / GALAH_B200_ROW_BLOCK * GALAH_B200_ROW_BLOCK;
    std::vector<size_t> g0((size_t)G + 1);
    for (int r = 0; r <= G; r++) g0[r] = std::min(n, (size_t)r * per);

    MultiBarrier bar(G);
    std::vector<int> rcs((size_t)G, 0);
    std::vector<std::string> errs((size_t)G);
    std::atomic<int> failed{0};
    std::vector<std::vector<galah_b200_pair_t>> rank_hits((size_t)G);
    std::vector<galah_b200_pair_t> all_hits;
    std::vector<float> ani_fwd, ani_rev;
    std::vector<float> t_ingest((size_t)G, 0.f), t_sketch((size_t)G, 0.f), t_index((size_t)G, 0.f), t_chain((size_t)G, 0.f);
    double t_k2 = 0.0, t_k3 = 0.0, t_engine = 0.0;

    auto worker = [&](int r) {
        t_dev = r;
        std::lock_guard<std::mutex> lock(g_mus[r]);
        Context &C = g_ctxs[r];
        // a failing device keeps walking through the barriers (with nothing to do) so that nobody hangs
        auto fail = [&](int rc) { if (!rcs[r]) { rcs[r] = rc; errs[r] = galah_b200_last_error(); failed.store(1); } };
        int rc = require_ctx();
        if (rc) fail(rc);
        cudaStream_t st = C.stream;
        const size_t nr = g0[r + 1] - g0[r];
        AniIndex *index = nullptr;
        if (!rc) {
            index = &C.pipeline_index(small_genomes != 0);
            if (ws_ensure(C.d_table, C.cap_table, std::max<size_t>(n, 1) * s) ||
                ws_ensure(C.d_counts, C.cap_counts, std::max<size_t>(n, 1)))
                fail(GALAH_B200_ERR_CUDA);
        }
        const double ta = now_ms();
        if (!rcs[r] && nr) {
            float a = 0.f, b = 0.f;
            rc = ingest(r, g0[r], nr, C.d_table + g0[r] * (size_t)s, C.d_counts + g0[r], *index, &a, &b);
            if (rc) fail(rc);
            t_sketch[r] = a; t_index[r] = b;
        }
        if (!rcs[r] && cudaStreamSynchronize(st) != cudaSuccess) fail(GALAH_B200_ERR_CUDA);
        t_ingest[r] = (float)(now_ms() - ta);
        bar.wait();  // ---- every slice of the sketch table exists on its owner
        const double tb = now_ms();
        if (!failed.load()) {
            for (int p = 0; p < G && !rcs[r]; p++) {
                const size_t np = g0[p + 1] - g0[p];
                if (p == r || np == 0) continue;
                if (cudaMemcpyPeerAsync(C.d_table + g0[p] * (size_t)s, r, g_ctxs[p].d_table + g0[p] * (size_t)s, p,
                                        np * (size_t)s * 8, st) != cudaSuccess ||
                    cudaMemcpyPeerAsync(C.d_counts + g0[p], r, g_ctxs[p].d_counts + g0[p], p, np * 4, st) != cudaSuccess) {
                    set_error("cluster_packed_multi: peer copy of the sketch rows failed");
                    fail(GALAH_B200_ERR_CUDA);
                }
            }
            if (!rcs[r]) {
                galah_b200_pair_t *hits = nullptr;
                size_t n_hits = 0;
                rc = run_prefilter(C.d_table, C.d_counts, n, s, 21, precluster_min_ani, (uint32_t)r, (uint32_t)G, st, &hits, &n_hits);
                if (rc) fail(rc);
                else rank_hits[r].assign(hits, hits + n_hits);
                free(hits);
            }
        }
        bar.wait();  // ---- every shard's hits are on the host
        if (r == 0 && !failed.load()) {
            for (auto &v : rank_hits) all_hits.insert(all_hits.end(), v.begin(), v.end());
            std::sort(all_hits.begin(), all_hits.end(), [](const galah_b200_pair_t &a, const galah_b200_pair_t &b) {
                return a.i != b.i ? a.i < b.i : a.j < b.j; });
            ani_fwd.assign(all_hits.size(), 0.f); ani_rev.assign(all_hits.size(), 0.f);
            t_k2 = now_ms() - tb;
        }
        bar.wait();  // ---- the merged hit list is visible to everybody
        const double tc = now_ms();
        if (!failed.load() && nr) {
            auto owner = [&](uint32_t g) { return (int)std::min<size_t>((size_t)g / per, (size_t)G - 1); };
            // jobs of this device: (hit, orientation) whose query genome it owns
            std::vector<uint32_t> job_hit, job_pairs;
            std::vector<uint8_t> job_rev;
            std::vector<int64_t> peer_first((size_t)G, -1);
            for (size_t h = 0; h < all_hits.size() && !rcs[r]; h++) {
                for (int o = 0; o < 2; o++) {
                    const uint32_t q = o ? all_hits[h].j : all_hits[h].i, ref = o ? all_hits[h].i : all_hits[h].j;
                    if (owner(q) != r) continue;
                    const int pr = owner(ref);
                    uint32_t ref_id;
                    if (pr == r) ref_id = ref - (uint32_t)g0[r];
                    else {
                        if (peer_first[pr] < 0) {
                            const AniIndex *pi = g_ctxs[pr].pipe_index[small_genomes ? 1 : 0];
                            uint32_t first = 0;
                            rc = index->attach_peer_direct(pi->table_base(), pi->table_offsets().data(), pi->total_lengths().data(),
                                                           pi->size(), &first);
                            if (rc) { fail(rc); break; }
                            peer_first[pr] = first;
                        }
                        ref_id = (uint32_t)peer_first[pr] + (ref - (uint32_t)g0[pr]);
                    }
                    job_hit.push_back((uint32_t)h); job_rev.push_back((uint8_t)o);
                    job_pairs.push_back(q - (uint32_t)g0[r]); job_pairs.push_back(ref_id);
                }
            }
            if (!rcs[r] && !job_hit.empty()) {
                std::vector<AniPairResult> res(job_hit.size());
                rc = index->pairs(job_pairs.data(), job_hit.size(), min_af_pct, false, res.data(), st);
                if (rc) fail(rc);
                else {
                    for (size_t x = 0; x < job_hit.size(); x++) (job_rev[x] ? ani_rev : ani_fwd)[job_hit[x]] = res[x].ani;
                    t_chain[r] = index->last_chain_ms;
                }
            }
        }
        bar.wait();  // ---- every ANI value is in the shared tables; nobody reads a peer's index any more
        if (r == 0) t_k3 = now_ms() - tc;
        if (index) index->clear();
        if (r == 0 && !failed.load()) {
            const double td = now_ms();
            rc = galah_b200_cluster_from_ani_tables(n, all_hits.data(), all_hits.size(), ani_fwd.data(), ani_rev.data(),
                                                    ani_threshold_pct, out);
            if (rc) fail(rc);
            t_engine = now_ms() - td;
        }
        t_dev = -1;
    };
    std::vector<std::thread> th;
    for (int r = 0; r < G; r++) th.emplace_back(worker, r);
    for (auto &t : th) t.join();
    for (int r = 0; r < G; r++)
        if (rcs[r]) { set_error("device " + std::to_string(r) + ": " + errs[r]); return rcs[r]; }
    if (stats) {
        stats->n_precluster_hits = all_hits.size(); stats->n_ani_pairs = 2 * all_hits.size();
        stats->ingest_ms = *std::max_element(t_ingest.begin(), t_ingest.end());
        stats->sketch_ms = *std::max_element(t_sketch.begin(), t_sketch.end());
        stats->index_ms = *std::max_element(t_index.begin(), t_index.end());
        stats->ani_chain_ms = *std::max_element(t_chain.begin(), t_chain.end());
        stats->prefilter_ms = (float)t_k2; stats->ani_ms = (float)t_k3; stats->engine_ms = (float)t_engine;
        stats->total_ms = (float)(now_ms() - t_begin);
    }
    return 0;
}

int galah_b200_cluster_packed_multi(const uint32_t *seq2, const uint32_t *valid, const uint64_t *base_off,
                                    const uint64_t *lengths, size_t n, int n_devices, float precluster_min_ani,
                                    float ani_threshold_pct, float min_af_pct, int small_genomes,
                                    galah_b200_clusters_t *out, galah_b200_cluster_stats_t *stats) {
    for (size_t g = 0; g <= n; g++)
        if (base_off[g] % 128) { set_error("packed genomes: base_off must be multiples of 128"); return GALAH_B200_ERR_ARG; }
    const MultiIngest ingest = [&](int, size_t first, size_t count, uint64_t *d_rows, uint32_t *d_cnt, AniIndex &index,
                                   float *k1_ms, float *ix_ms) {
        return ingest_packed(seq2, valid, nullptr, base_off + first, lengths + first, count, false, d_rows, d_cnt, index,
                             k1_ms, ix_ms);
    };
    return cluster_multi(n, n_devices, ingest, precluster_min_ani, ani_threshold_pct, min_af_pct, small_genomes, out, stats);
}

int galah_b200_cluster_files_multi(const char *const *paths, size_t n, int n_devices, float precluster_min_ani,
                                   float ani_threshold_pct, float min_af_pct, int small_genomes, int host_threads,
                                   galah_b200_clusters_t *out, galah_b200_cluster_stats_t *stats) {
    if (host_threads <= 0) host_threads = (int)std::max(1u, std::thread::hardware_concurrency());
    const int per_device = std::max(1, host_threads / std::max(1, n_devices));
    const MultiIngest ingest = [&](int, size_t first, size_t count, uint64_t *d_rows, uint32_t *d_cnt, AniIndex &index,
                                   float *, float *) {
        IngestSinks sinks;
        sinks.sketch = true; sinks.k = 21; sinks.s = 1000; sinks.seed = 0;
        sinks.d_hashes = d_rows; sinks.d_counts = d_cnt; sinks.ani = &index;
        return ingest_files(paths + first, count, per_device, sinks);
    };
    return cluster_multi(n, n_devices, ingest, precluster_min_ani, ani_threshold_pct, min_af_pct, small_genomes, out, stats);
}

int galah_b200_skani_distances(const char *const *paths, size_t n, float threshold_pct, float min_af_pct,
                               int small_genomes, int per_record, int host_threads, galah_b200_pair_t **out,
                               size_t *n_out, size_t *n_units) {
    std::lock_guard<std::mutex> lock(g_mu);
    if (int rc = require_ctx()) return rc;
    *out = nullptr; *n_out = 0;
    std::vector<galah_b200_pair_t> hits;
    size_t units = 0; uint64_t screened = 0;
    if (int rc = skani_distances_impl(paths, n, threshold_pct, min_af_pct, small_genomes != 0, per_record != 0,
                                      host_threads, hits, units, screened))
        return rc;
    if (n_units) *n_units = units;
    galah_b200_pair_t *res = (galah_b200_pair_t *)malloc(std::max<size_t>(hits.size(), 1) * sizeof(galah_b200_pair_t));
    if (!res) { set_error("out of host memory"); return GALAH_B200_ERR_ARG; }
    if (!hits.empty()) memcpy(res, hits.data(), hits.size() * sizeof(galah_b200_pair_t));
    *out = res; *n_out = hits.size();
    return 0;
}

int galah_b200_skani_distances_packed_device(const uint32_t *d_seq2, const uint32_t *d_valid,
                                             const uint64_t *d_base_off, const uint64_t *base_off,
                                             const uint64_t *lengths, size_t n, float threshold_pct,
                                             float min_af_pct, int small_genomes, int individual_contigs,
                                             void *stream, galah_b200_pair_t **out, size_t *n_out,
                                             uint64_t *n_screened, float *ms5) {
    std::lock_guard<std::mutex> lock(g_mu);
    if (int rc = require_ctx()) return rc;
    *out = nullptr; *n_out = 0;
    if (threshold_pct < 85.0f) {
        set_error("Error: skani produces inaccurate results with ANI less than 85%. Provided: " + display_f32(threshold_pct));
        return GALAH_B200_ERR_UNSUPPORTED;
    }
    cudaStream_t st = (cudaStream_t)stream;
    const double t0 = now_ms();
    AniIndex index(small_genomes ? 30u : 125u);
    std::vector<uint64_t> bo(base_off, base_off + n + 1), co(n + 1);
    std::vector<uint32_t> cs(n, 0), cl(n);
    uint64_t longest = 0;
    for (size_t g = 0; g < n; g++) { co[g] = g; cl[g] = (uint32_t)lengths[g]; longest = std::max(longest, lengths[g]); }
    co[n] = n;
    const double t1 = now_ms();
    // marker sketches straight into the K2 table layout (row stride = cap): they never leave the device
    const uint32_t c_marker = small_genomes ? 200u : 1000u;
    const uint32_t cap = marker_row_capacity(longest, c_marker);
    DevBuf<uint64_t> d_rows;
    DevBuf<uint32_t> d_counts;
    if (d_rows.alloc(n * (size_t)cap) || d_counts.alloc(n)) return GALAH_B200_ERR_CUDA;
    SeedSink seed_sink{nullptr, 0, 0};  // one pass: marker sketches + K3 seed selection
    if (int rc = seed_sink_for(index, bo.front(), bo.back(), seed_sink)) return rc;
    if (int rc = marker_sketch_enqueue(g_ctx.sws, d_seq2, d_valid, d_base_off, n, 21, c_marker, cap, d_rows.p, d_counts.p, st,
                                       &seed_sink))
        return rc;
    if (int rc = index.add_packed_device(d_seq2, d_valid, d_base_off, n, bo, co, cs, cl, st, seed_sink.d_sel)) return rc;
    std::vector<uint32_t> cnt(n);
    GB_CUDA(cudaMemcpyAsync(cnt.data(), d_counts.p, n * 4, cudaMemcpyDeviceToHost, st));
    GB_CUDA(cudaStreamSynchronize(st));
    for (size_t g = 0; g < n; g++)
        if (cnt[g] == 0xFFFFFFFFu) {
            set_error("marker sketch: a unit holds more markers than its row of " + std::to_string(cap) +
                      " (marker density 1/" + std::to_string(c_marker) + "; rows hold at most " + std::to_string(kMarkerMaxCap) +
                      " markers, about " + std::to_string((uint64_t)kMarkerMaxCap * c_marker / 1500000) + " Mbp per unit); unsupported");
            return GALAH_B200_ERR_UNSUPPORTED;
        }
    const double t2 = now_ms();
    std::vector<galah_b200_pair_t> hits;
    uint64_t screened = 0;
    float ms3[3] = {0, 0, 0};
    if (n >= 2)
        if (int rc = skani_screen_and_ani(index, d_rows.p, d_counts.p, n, cap, threshold_pct, min_af_pct, individual_contigs != 0, st, hits, screened, ms3))
            return rc;
    if (n_screened) *n_screened = screened;
    if (ms5) { ms5[0] = (float)(t1 - t0); ms5[1] = (float)(t2 - t1); ms5[2] = ms3[0]; ms5[3] = ms3[1]; ms5[4] = (float)(now_ms() - t0); }
    galah_b200_pair_t *res = (galah_b200_pair_t *)malloc(std::max<size_t>(hits.size(), 1) * sizeof(galah_b200_pair_t));
    if (!res) { set_error("out of host memory"); return GALAH_B200_ERR_ARG; }
    if (!hits.empty()) memcpy(res, hits.data(), hits.size() * sizeof(galah_b200_pair_t));
    *out = res; *n_out = hits.size();
    return 0;
}

// SkaniPreclusterer::distances / ::distances_contigs (src/skani.rs:21-56, 109-225, 379-498) on packed
// units in HOST arrays over G devices of this process: the structure of cluster_multi with marker
// sketches in place of MinHash sketches, the containment screen in place of the finch rule, and
// every screened pair (i < j) evaluated ONCE, with i as the query, on the device that owns i.
// The skani preclusterer over G devices of this process.  `ingest(r, index, slice)` puts device r's units into its K3
// index and their marker rows EITHER in place into the device's copy of the global marker table (packed input: unit
// slices g0 and the row stride `cap` are known up front) OR into a table of its own (file input: the number of units
// of a slice and the longest unit are only known once its files are read; the global layout is agreed at a barrier
// and every device re-strides its rows into its slice of the global table).  From there on both forms are one code
// path: marker rows exchanged by peer copies, row-block shards of the containment screen, every screened pair
// (i < j) evaluated on the device that owns unit i, which reads unit j's hash table in place on its peer.
static int take_hits(const std::vector<galah_b200_pair_t> &hits, galah_b200_pair_t **out, size_t *n_out) {
    galah_b200_pair_t *res = (galah_b200_pair_t *)malloc(std::max<size_t>(hits.size(), 1) * sizeof(galah_b200_pair_t));
    if (!res) { set_error("out of host memory"); return GALAH_B200_ERR_ARG; }
    if (!hits.empty()) memcpy(res, hits.data(), hits.size() * sizeof(galah_b200_pair_t));
    *out = res; *n_out = hits.size();
    return 0;
}

struct MultiSlice { size_t n_units = 0; MarkerTable local; };
using MultiMarkerIngest = std::function<int(int r, AniIndex &index, MultiSlice &slice)>;

static int skani_distances_multi(int G, bool layout_known, std::vector<size_t> g0_in, uint32_t cap_in, const MultiMarkerIngest &ingest,
                                 float threshold_pct, float min_af_pct, int small_genomes, int individual_contigs,
                                 std::vector<galah_b200_pair_t> &hits_out, size_t *n_units_out, uint64_t *n_screened) {
    hits_out.clear();
    if (n_screened) *n_screened = 0;
    if (threshold_pct < 85.0f) {
        set_error("Error: skani produces inaccurate results with ANI less than 85%. Provided: " + display_f32(threshold_pct));
        return GALAH_B200_ERR_UNSUPPORTED;
    }
    if (G < 1 || G > kMaxDevices) { set_error("skani_distances_multi: bad device count"); return GALAH_B200_ERR_ARG; }
    for (int d = 0; d < G; d++)
        if (g_ctxs[d].device != d) { set_error("skani_distances_multi: call galah_b200_init_devices(n_devices) first"); return GALAH_B200_ERR_NO_DEVICE; }

    MultiBarrier bar(G);
    std::vector<int> rcs((size_t)G, 0);
    std::vector<std::string> errs((size_t)G);
    std::atomic<int> failed{0};
    std::vector<std::vector<uint4>> inbox((size_t)G);
    std::vector<std::mutex> inbox_mu((size_t)G);
    std::vector<std::vector<galah_b200_pair_t>> rank_hits((size_t)G);
    std::vector<MultiSlice> slices((size_t)G);
    std::atomic<uint64_t> screened{0};
    size_t n_total = layout_known ? g0_in.back() : 0;

    auto worker = [&](int r) {
        t_dev = r;
        std::lock_guard<std::mutex> lock(g_mus[r]);
        Context &C = g_ctxs[r];
        auto fail = [&](int rc) { if (!rcs[r]) { rcs[r] = rc; errs[r] = galah_b200_last_error(); failed.store(1); } };
        int rc = require_ctx();
        if (rc) fail(rc);
        cudaStream_t st = C.stream;
        std::vector<size_t> g0 = g0_in;  // unit slices: device p owns units [g0[p], g0[p + 1])
        uint32_t cap = cap_in;           // row stride of the global marker table
        size_t n = layout_known ? g0.back() : 0;
        AniIndex *index = nullptr;
        if (!rc) {
            index = &C.pipeline_index(small_genomes != 0);
            if (layout_known && (ws_ensure(C.d_table, C.cap_table, std::max<size_t>(n, 1) * cap) ||
                                 ws_ensure(C.d_counts, C.cap_counts, std::max<size_t>(n, 1))))
                fail(GALAH_B200_ERR_CUDA);
        }
        if (!rcs[r]) {
            rc = ingest(r, *index, slices[r]);
            if (rc) fail(rc);
        }
        if (!layout_known) {
            bar.wait();  // ---- every slice knows its unit count and its row stride: agree on the global layout
            g0.assign((size_t)G + 1, 0);
            cap = 0;
            for (int p = 0; p < G; p++) { g0[p + 1] = g0[p] + slices[p].n_units; cap = std::max(cap, slices[p].local.stride); }
            n = g0[G];
            if (r == 0) n_total = n;
            const size_t nr = g0[r + 1] - g0[r];
            if (!failed.load() && n >= 2) {
                if (ws_ensure(C.d_table, C.cap_table, n * (size_t)cap) || ws_ensure(C.d_counts, C.cap_counts, n)) fail(GALAH_B200_ERR_CUDA);
                const MarkerTable &mt = slices[r].local;
                if (!rcs[r] && nr && mt.n != nr) { set_error("skani preclusterer: marker table out of step with the units"); fail(GALAH_B200_ERR_ARG); }
                if (!rcs[r] && nr) {
                    // own rows at the common stride; whatever lies beyond a row's own stride reads as "no hash"
                    uint64_t *dst = C.d_table + g0[r] * (size_t)cap;
                    if (cudaMemsetAsync(dst, 0xFF, nr * (size_t)cap * 8, st) != cudaSuccess ||
                        cudaMemcpy2DAsync(dst, (size_t)cap * 8, mt.d_rows, (size_t)mt.stride * 8, (size_t)mt.stride * 8, nr,
                                          cudaMemcpyDeviceToDevice, st) != cudaSuccess ||
                        cudaMemcpyAsync(C.d_counts + g0[r], mt.d_counts, nr * 4, cudaMemcpyDeviceToDevice, st) != cudaSuccess ||
                        cudaStreamSynchronize(st) != cudaSuccess) {
                        set_error("skani_distances_multi: re-striding the marker rows failed"); fail(GALAH_B200_ERR_CUDA);
                    }
                }
            }
        }
        const size_t nr = g0[r + 1] - g0[r];
        auto owner = [&](uint32_t g) { return (int)(std::upper_bound(g0.begin() + 1, g0.end(), (size_t)g) - (g0.begin() + 1)); };
        if (layout_known && !rcs[r]) {
            std::vector<uint32_t> cnt(nr);
            if ((nr && cudaMemcpyAsync(cnt.data(), C.d_counts + g0[r], nr * 4, cudaMemcpyDeviceToHost, st) != cudaSuccess) ||
                cudaStreamSynchronize(st) != cudaSuccess) {
                set_error("skani_distances_multi: reading the marker counts failed"); fail(GALAH_B200_ERR_CUDA);
            }
            for (size_t x = 0; x < nr && !rcs[r]; x++)
                if (cnt[x] == 0xFFFFFFFFu) {
                    set_error("marker sketch: a unit holds more markers than its row of " + std::to_string(cap) + "; unsupported");
                    fail(GALAH_B200_ERR_UNSUPPORTED);
                }
        }
        bar.wait();  // ---- every slice of the marker table exists on its owner
        if (!failed.load() && n >= 2) {
            for (int p = 0; p < G && !rcs[r]; p++) {
                const size_t np = g0[p + 1] - g0[p];
                if (p == r || np == 0) continue;
                if (cudaMemcpyPeerAsync(C.d_table + g0[p] * (size_t)cap, r, g_ctxs[p].d_table + g0[p] * (size_t)cap, p,
                                        np * (size_t)cap * 8, st) != cudaSuccess ||
                    cudaMemcpyPeerAsync(C.d_counts + g0[p], r, g_ctxs[p].d_counts + g0[p], p, np * 4, st) != cudaSuccess) {
                    set_error("skani_distances_multi: peer copy of the marker rows failed");
                    fail(GALAH_B200_ERR_CUDA);
                }
            }
            // the containment screen on this device's row-block shard (skani_screen_and_ani, sharded)
            std::vector<uint4> cand;
            if (!rcs[r]) {
                if (!C.d_n_cand && cudaMalloc(&C.d_n_cand, sizeof(unsigned long long)) != cudaSuccess) fail(GALAH_B200_ERR_CUDA);
                const double frac = pow(0.80, 21.0);
                size_t want = std::max<size_t>(1 << 16, 64 * n / (size_t)G);
                while (!rcs[r]) {
                    if (ws_ensure(C.d_cand, C.cap_cand, want)) { fail(GALAH_B200_ERR_CUDA); break; }
                    rc = prefilter_enqueue(C.pws, C.d_table, C.d_counts, n, cap, 21, 0.f, (uint32_t)r, (uint32_t)G,
                                           join_supported(cap) ? 0 : 1, st, C.d_cand, C.cap_cand, C.d_n_cand,
                                           index->c() == 30u ? kRuleContainment : kRuleContainmentBypassSmall, frac);
                    if (rc) { fail(rc); break; }
                    unsigned long long got = 0;
                    if (cudaMemcpyAsync(&got, C.d_n_cand, sizeof(got), cudaMemcpyDeviceToHost, st) != cudaSuccess ||
                        cudaStreamSynchronize(st) != cudaSuccess) { set_error("screen: count read failed"); fail(GALAH_B200_ERR_CUDA); break; }
                    if (got > C.cap_cand) { want = (size_t)got; continue; }
                    cand.resize((size_t)got);
                    if (got && (cudaMemcpyAsync(cand.data(), C.d_cand, (size_t)got * sizeof(uint4), cudaMemcpyDeviceToHost, st) != cudaSuccess ||
                                cudaStreamSynchronize(st) != cudaSuccess)) { set_error("screen: candidate read failed"); fail(GALAH_B200_ERR_CUDA); }
                    break;
                }
            }
            // every screened pair travels once, to the device that owns its query (the lower index)
            if (!rcs[r]) {
                std::vector<std::vector<uint4>> outbox((size_t)G);
                for (const uint4 &c : cand) outbox[owner(c.x)].push_back(c);
                for (int p = 0; p < G; p++) {
                    if (outbox[p].empty()) continue;
                    std::lock_guard<std::mutex> lk(inbox_mu[p]);
                    inbox[p].insert(inbox[p].end(), outbox[p].begin(), outbox[p].end());
                }
            }
        }
        bar.wait();  // ---- every device holds the screened pairs whose query it owns
        if (!failed.load() && nr) {
            std::vector<uint4> &mine = inbox[r];
            std::sort(mine.begin(), mine.end(), [](const uint4 &a, const uint4 &b) { return a.x != b.x ? a.x < b.x : a.y < b.y; });
            screened.fetch_add(mine.size());
            std::vector<uint32_t> pairs(2 * mine.size());
            std::vector<int64_t> peer_first((size_t)G, -1);
            for (size_t x = 0; x < mine.size() && !rcs[r]; x++) {
                const int pr = owner(mine[x].y);
                uint32_t ref_id;
                if (pr == r) ref_id = mine[x].y - (uint32_t)g0[r];
                else {
                    if (peer_first[pr] < 0) {
                        const AniIndex *pi = g_ctxs[pr].pipe_index[small_genomes ? 1 : 0];
                        uint32_t first = 0;
                        rc = index->attach_peer_direct(pi->table_base(), pi->table_offsets().data(), pi->total_lengths().data(),
                                                       pi->size(), &first);
                        if (rc) { fail(rc); break; }
                        peer_first[pr] = first;
                    }
                    ref_id = (uint32_t)peer_first[pr] + (mine[x].y - (uint32_t)g0[pr]);
                }
                pairs[2 * x] = mine[x].x - (uint32_t)g0[r]; pairs[2 * x + 1] = ref_id;
            }
            if (!rcs[r] && !mine.empty()) {
                std::vector<AniPairResult> res(mine.size());
                rc = index->pairs(pairs.data(), mine.size(), min_af_pct, individual_contigs != 0, res.data(), st);
                if (rc) fail(rc);
                else
                    for (size_t x = 0; x < mine.size(); x++)
                        if (res[x].ani >= threshold_pct)  // `if ani >= threshold` in f32, src/skani.rs:205
                            rank_hits[r].push_back(galah_b200_pair_t{mine[x].x, mine[x].y, mine[x].z, mine[x].w, res[x].ani});
            }
        }
        bar.wait();  // ---- nobody reads a peer's index any more
        if (index) index->clear();
        t_dev = -1;
    };
    std::vector<std::thread> th;
    for (int r = 0; r < G; r++) th.emplace_back(worker, r);
    for (auto &t : th) t.join();
    for (int r = 0; r < G; r++)
        if (rcs[r]) { set_error("device " + std::to_string(r) + ": " + errs[r]); return rcs[r]; }
    // the devices own ascending slices of the query index: their lists concatenate into (i, j) order
    for (auto &v : rank_hits) hits_out.insert(hits_out.end(), v.begin(), v.end());
    if (n_units_out) *n_units_out = n_total;
    if (n_screened) *n_screened = screened.load();
    return 0;
}

int galah_b200_skani_distances_packed_multi(const uint32_t *seq2, const uint32_t *valid, const uint64_t *base_off,
                                            const uint64_t *lengths, size_t n, int n_devices, float threshold_pct,
                                            float min_af_pct, int small_genomes, int individual_contigs,
                                            galah_b200_pair_t **out, size_t *n_out, uint64_t *n_screened) {
    if (!out || !n_out) { set_error("skani_distances_packed_multi: NULL argument"); return GALAH_B200_ERR_ARG; }
    *out = nullptr; *n_out = 0;
    const int G = n_devices;
    if (G < 1 || G > kMaxDevices) { set_error("skani_distances_packed_multi: bad device count"); return GALAH_B200_ERR_ARG; }
    for (size_t g = 0; g <= n; g++)
        if (base_off[g] % 128) { set_error("packed units: base_off must be multiples of 128"); return GALAH_B200_ERR_ARG; }
    uint64_t longest = 0;
    for (size_t g = 0; g < n; g++) longest = std::max(longest, lengths[g]);
    const uint32_t c_marker = small_genomes ? 200u : 1000u;
    const uint32_t cap = marker_row_capacity(longest, c_marker);
    size_t per = (n + (size_t)G - 1) / (size_t)G;
    per = (per + GALAH_B200_ROW_BLOCK - 1) / GALAH_B200_ROW_BLOCK * GALAH_B200_ROW_BLOCK;
    std::vector<size_t> g0((size_t)G + 1);
    for (int r = 0; r <= G; r++) g0[r] = std::min(n, (size_t)r * per);
    const MultiMarkerIngest ingest = [&](int r, AniIndex &index, MultiSlice &slice) -> int {
        const size_t nr = g0[r + 1] - g0[r];
        slice.n_units = nr;
        if (!nr) return 0;
        Context &C = g_ctxs[r];
        return ingest_packed(seq2, valid, nullptr, base_off + g0[r], lengths + g0[r], nr, false, C.d_table + g0[r] * (size_t)cap,
                             C.d_counts + g0[r], index, nullptr, nullptr, c_marker, cap);
    };
    std::vector<galah_b200_pair_t> hits;
    if (int rc = skani_distances_multi(G, true, g0, cap, ingest, threshold_pct, min_af_pct, small_genomes, individual_contigs, hits,
                                       nullptr, n_screened))
        return rc;
    return take_hits(hits, out, n_out);
}

// The file form: device r reads, decodes (K0), marker-sketches and indexes the r-th slice of the path list; in contig
// mode every record is a unit, so a slice's unit count is known only after its files are read.
int galah_b200_skani_distances_multi(const char *const *paths, size_t n, int n_devices, float threshold_pct, float min_af_pct,
                                     int small_genomes, int per_record, int host_threads, galah_b200_pair_t **out,
                                     size_t *n_out, size_t *n_units) {
    if (!out || !n_out) { set_error("skani_distances_multi: NULL argument"); return GALAH_B200_ERR_ARG; }
    *out = nullptr; *n_out = 0;
    if (n_units) *n_units = 0;
    const int G = n_devices;
    if (G < 1 || G > kMaxDevices) { set_error("skani_distances_multi: bad device count"); return GALAH_B200_ERR_ARG; }
    if (host_threads <= 0) host_threads = (int)std::max(1u, std::thread::hardware_concurrency());
    const int per_device = std::max(1, host_threads / G);
    const size_t per = (n + (size_t)G - 1) / (size_t)G;  // files per device
    const MultiMarkerIngest ingest = [&](int r, AniIndex &index, MultiSlice &slice) -> int {
        const size_t f0 = std::min(n, (size_t)r * per), f1 = std::min(n, (size_t)(r + 1) * per);
        slice.n_units = 0;
        if (f0 == f1) return 0;
        IngestSinks sinks;
        sinks.ani = &index; sinks.markers = &slice.local; sinks.c_marker = small_genomes ? 200u : 1000u;
        sinks.per_record = per_record != 0;
        if (int rc = ingest_files(paths + f0, f1 - f0, per_device, sinks)) return rc;
        slice.n_units = sinks.n_units;
        return 0;
    };
    std::vector<galah_b200_pair_t> hits;
    if (int rc = skani_distances_multi(G, false, std::vector<size_t>(), 0, ingest, threshold_pct, min_af_pct, small_genomes, per_record, hits,
                                       n_units, nullptr))
        return rc;
    return take_hits(hits, out, n_out);
}

int galah_b200_cluster_files_skani_multi(const char *const *paths, size_t n, int n_devices, float precluster_ani_pct,
                                         float ani_threshold_pct, float min_af_pct, int small_genomes, int cluster_contigs,
                                         int host_threads, galah_b200_clusters_t *out, galah_b200_cluster_stats_t *stats) {
    if (!out) { set_error("cluster_files_skani_multi: out is NULL"); return GALAH_B200_ERR_ARG; }
    memset(out, 0, sizeof(*out));
    if (stats) memset(stats, 0, sizeof(*stats));
    if (!(ani_threshold_pct > 1.0f)) { set_error("assertion failed: self.threshold > 1.0"); return GALAH_B200_ERR_UNSUPPORTED; }
    galah_b200_pair_t *hits = nullptr;
    size_t n_hits = 0, n_units = 0;
    if (int rc = galah_b200_skani_distances_multi(paths, n, n_devices, precluster_ani_pct, min_af_pct, small_genomes, cluster_contigs,
                                                  host_threads, &hits, &n_hits, &n_units))
        return rc;
    struct HitGuard { galah_b200_pair_t *h; ~HitGuard() { free(h); } } hit_guard{hits};
    int rc = galah_b200_cluster_from_distances(n_units, hits, n_hits, 1, ani_threshold_pct, nullptr, nullptr, out);
    if (stats) { stats->n_precluster_hits = n_hits; stats->n_ani_pairs = n_hits; }
    return rc;
}

int galah_b200_cluster_files_skani(const char *const *paths, size_t n, float precluster_ani_pct, float ani_threshold_pct,
                                   float min_af_pct, int small_genomes, int cluster_contigs, int host_threads,
                                   galah_b200_clusters_t *out, galah_b200_cluster_stats_t *stats) {
    if (!out) { set_error("cluster_files_skani: out is NULL"); return GALAH_B200_ERR_ARG; }
    memset(out, 0, sizeof(*out));
    if (stats) memset(stats, 0, sizeof(*stats));
    if (!(ani_threshold_pct > 1.0f)) { set_error("assertion failed: self.threshold > 1.0"); return GALAH_B200_ERR_UNSUPPORTED; }
    galah_b200_pair_t *hits = nullptr;
    size_t n_hits = 0, n_units = 0;
    if (int rc = galah_b200_skani_distances(paths, n, precluster_ani_pct, min_af_pct, small_genomes, cluster_contigs,
                                            host_threads, &hits, &n_hits, &n_units))
        return rc;
    struct HitGuard { galah_b200_pair_t *h; ~HitGuard() { free(h); } } hit_guard{hits};
    // preclusterer and clusterer are both "skani" (or contigs are clustered): skip_clusterer, the
    // precluster ANIs are re-used against the final threshold (src/clusterer.rs:32-44)
    int rc = galah_b200_cluster_from_distances(n_units, hits, n_hits, 1, ani_threshold_pct, nullptr, nullptr, out);
    if (stats) { stats->n_precluster_hits = n_hits; stats->n_ani_pairs = n_hits; }
    return rc;
}

// ------------------------------------------------------------------------------------------
// Trait-shaped boundary: what an UNMODIFIED galah::clusterer::cluster() calls, in the order it
// calls it (src/clusterer.rs:14-152): preclusterer.distances*(paths) once on the caller's thread,
// then clusterer.calculate_ani(fasta1, fasta2) from nested rayon workers (src/clusterer.rs:262-296,
// 375).  A session carries what the two halves must share: the paths and hit list of the last
// distances call (so that the first calculate_ani can evaluate EVERY hit in one K3 launch), the
// resident K3 index keyed by path, and the cache of computed values.  calculate_ani is re-entrant:
// hits of the cache take a shared lock only.
// ------------------------------------------------------------------------------------------
}  // extern "C"

#include <shared_mutex>
#include <unordered_map>

struct galah_b200_session {
    std::shared_mutex mu;                 // guards everything below; lookups take it shared
    std::unordered_map<std::string, uint32_t> path_id;  // FASTA path -> genome id in `index`
    std::unique_ptr<gb200::AniIndex> index;
    bool small_genomes = false;
    float min_af_pct = -1.f;              // the cache holds values for this --min-af only
    std::unordered_map<uint64_t, float> cache;  // (query id << 32 | reference id) -> ANI as galah parses it
    std::vector<std::string> stash_paths;       // genomes of the last distances() call ...
    std::vector<galah_b200_pair_t> stash_hits;  // ... and its hit list, not yet evaluated by K3
    bool hint = false; bool hint_small = false; // clusterer configuration announced ahead of distances()
    uint64_t n_calls = 0, n_lookups = 0, n_pairs_computed = 0, n_launches = 0;
};

namespace {
using gb200::set_error;

// Adds the paths that are not indexed yet (one ingest pass), assigns ids.  Caller: exclusive lock + g_mu.
int session_index_paths(galah_b200_session *s, const std::vector<std::string> &paths, int host_threads) {
    std::vector<const char *> missing;
    std::vector<std::string> keep;
    for (const auto &p : paths)
        if (!s->path_id.count(p) && std::find(keep.begin(), keep.end(), p) == keep.end()) keep.push_back(p);
    if (keep.empty()) return 0;
    for (const auto &p : keep) missing.push_back(p.c_str());
    if (!s->index) s->index.reset(new gb200::AniIndex(s->small_genomes ? 30u : 125u));
    const size_t before = s->index->size();
    gb200::IngestSinks sinks;
    sinks.ani = s->index.get();
    if (int rc = gb200::ingest_files(missing.data(), missing.size(), host_threads, sinks)) return rc;
    if (s->index->size() != before + keep.size()) { set_error("session: index out of step with the paths"); return GALAH_B200_ERR_ARG; }
    for (size_t x = 0; x < keep.size(); x++) s->path_id[keep[x]] = (uint32_t)(before + x);
    return 0;
}

// Evaluates (query, reference) id pairs that are not cached yet, in one K3 launch.
int session_compute(galah_b200_session *s, std::vector<uint32_t> &pairs) {
    std::vector<uint32_t> todo;
    std::unordered_map<uint64_t, char> seen;
    for (size_t x = 0; x + 1 < pairs.size(); x += 2) {
        const uint64_t key = ((uint64_t)pairs[x] << 32) | pairs[x + 1];
        if (s->cache.count(key) || !seen.emplace(key, 1).second) continue;
        todo.push_back(pairs[x]); todo.push_back(pairs[x + 1]);
    }
    if (todo.empty()) return 0;
    std::vector<gb200::AniPairResult> res(todo.size() / 2);
    if (int rc = s->index->pairs(todo.data(), todo.size() / 2, s->min_af_pct, false, res.data(), g_ctx.stream)) return rc;
    for (size_t x = 0; x < res.size(); x++) s->cache[((uint64_t)todo[2 * x] << 32) | todo[2 * x + 1]] = res[x].ani;
    s->n_pairs_computed += res.size(); s->n_launches++;
    return 0;
}
}  // namespace

extern "C" {

int galah_b200_session_create(galah_b200_session_t **out) {
    if (!out) { set_error("session: out is NULL"); return GALAH_B200_ERR_ARG; }
    *out = new galah_b200_session();
    return 0;
}

void galah_b200_session_free(galah_b200_session_t *s) {
    if (!s) return;
    {
        std::lock_guard<std::mutex> lock(g_mu);
        if (g_ctx.device >= 0) cudaSetDevice(g_ctx.device);
        s->index.reset();
    }
    delete s;
}

int galah_b200_session_set_clusterer(galah_b200_session_t *s, int small_genomes) {
    if (!s) { set_error("session: NULL session"); return GALAH_B200_ERR_ARG; }
    std::unique_lock<std::shared_mutex> lk(s->mu);
    s->hint = true; s->hint_small = small_genomes != 0;
    return 0;
}

const char *galah_b200_finch_method_name(void) { return "finch"; }
const char *galah_b200_skani_method_name(void) { return "skani"; }

int galah_b200_session_finch_distances(galah_b200_session_t *s, const char *const *paths, size_t n, float min_ani,
                                       uint32_t num_kmers, uint8_t kmer_length, int low_memory, int host_threads,
                                       galah_b200_pair_t **out, size_t *n_out) {
    if (!s || !out || !n_out) { set_error("session: NULL argument"); return GALAH_B200_ERR_ARG; }
    *out = nullptr; *n_out = 0;
    if (low_memory) {  // src/finch.rs:14-15
        set_error("Low-memory clustering currently only supported with skani preclusterer");
        return GALAH_B200_ERR_UNSUPPORTED;
    }
    std::unique_lock<std::shared_mutex> lk(s->mu);
    std::lock_guard<std::mutex> lock(g_mu);
    if (int rc = require_ctx()) return rc;
    // with the clusterer announced, the same ingest pass builds the K3 index of every genome that is
    // not resident yet (all of them, on a fresh session)
    gb200::AniIndex *ani = nullptr;
    bool fresh = false;
    if (s->hint) {
        if (s->index && s->small_genomes != s->hint_small) { s->index.reset(); s->path_id.clear(); s->cache.clear(); }
        s->small_genomes = s->hint_small;
        fresh = s->path_id.empty();
        if (fresh) {
            std::vector<std::string> uniq(paths, paths + n);
            std::sort(uniq.begin(), uniq.end());
            fresh = std::adjacent_find(uniq.begin(), uniq.end()) == uniq.end();  // duplicate paths: index lazily instead
        }
        if (fresh) { s->index.reset(new gb200::AniIndex(s->small_genomes ? 30u : 125u)); ani = s->index.get(); }
    }
    if (int rc = finch_distances_locked(paths, n, min_ani, num_kmers, kmer_length, host_threads, ani, out, n_out)) {
        if (fresh) s->index.reset();
        return rc;
    }
    if (fresh) for (size_t g = 0; g < n; g++) s->path_id[paths[g]] = (uint32_t)g;
    s->stash_paths.assign(paths, paths + n);
    s->stash_hits.assign(*out, *out + *n_out);
    return 0;
}

int galah_b200_session_finch_distances_contigs(galah_b200_session_t *s, const char *const *paths, size_t n,
                                               const char *const *contig_names, size_t n_names,
                                               galah_b200_pair_t **out, size_t *n_out) {
    (void)s; (void)paths; (void)n; (void)contig_names; (void)n_names;
    // "Finch doesn't offer high-quality ANI with self-self comparisons": an empty cache (src/finch.rs:26-33)
    if (!out || !n_out) { set_error("session: NULL argument"); return GALAH_B200_ERR_ARG; }
    *out = (galah_b200_pair_t *)malloc(sizeof(galah_b200_pair_t));
    *n_out = 0;
    return 0;
}

int galah_b200_session_finch_distances_with_references(galah_b200_session_t *s, const char *const *paths, size_t n,
                                                       const char *const *reference_paths, size_t n_refs,
                                                       galah_b200_pair_t **out, size_t *n_out) {
    (void)s; (void)paths; (void)n; (void)reference_paths; (void)n_refs;
    if (out) *out = nullptr;
    if (n_out) *n_out = 0;
    set_error("Reference genome clustering currently only supported with skani preclusterer");  // src/finch.rs:40
    return GALAH_B200_ERR_UNSUPPORTED;
}

int galah_b200_session_skani_distances(galah_b200_session_t *s, const char *const *paths, size_t n, float threshold_pct,
                                       float min_aligned_threshold, int small_genomes, int low_memory,
                                       int host_threads, galah_b200_pair_t **out, size_t *n_out) {
    if (!s || !out || !n_out) { set_error("session: NULL argument"); return GALAH_B200_ERR_ARG; }
    *out = nullptr; *n_out = 0;
    std::unique_lock<std::shared_mutex> lk(s->mu);
    std::lock_guard<std::mutex> lock(g_mu);
    if (int rc = require_ctx()) return rc;
    std::vector<galah_b200_pair_t> hits;
    size_t units = 0; uint64_t screened = 0;
    const float min_af_pct = min_aligned_threshold * 100.0f;  // `min_aligned_threshold * 100.0` in f32, src/skani.rs:153
    if (int rc = skani_distances_impl(paths, n, threshold_pct, min_af_pct, small_genomes != 0, false, host_threads, hits,
                                      units, screened, low_memory ? kSkaniLowMem : kSkaniTriangle))
        return rc;
    s->stash_paths.clear(); s->stash_hits.clear();  // same method name on both sides: cluster() re-uses these values
    return take_hits(hits, out, n_out);
}

// Names of the records of a FASTA / FASTQ file as galah reads them for --cluster-contigs: the
// header line up to the first TAB (src/cluster_argument_parsing.rs:606-621).
static int record_names(const std::string &path, std::vector<std::string> &names) {
    std::vector<uint8_t> raw;
    std::string err;
    if (read_file_bytes(path, raw, err)) { set_error(err); return GALAH_B200_ERR_IO; }
    size_t p = 0;
    const bool fastq = !raw.empty() && raw[0] == '@';
    size_t line_no = 0;
    while (p < raw.size()) {
        size_t e = p;
        while (e < raw.size() && raw[e] != '\n') e++;
        const bool header = fastq ? (line_no % 4 == 0) : (raw[p] == '>');
        if (header) {
            size_t end = e;
            if (end > p && raw[end - 1] == '\r') end--;
            std::string name((const char *)raw.data() + p + 1, end - p - 1);
            const size_t tab = name.find('\t');
            if (tab != std::string::npos) name.resize(tab);
            names.push_back(name);
        }
        if (e > p || fastq) line_no++;
        p = e + 1;
    }
    return 0;
}

int galah_b200_contig_names(const char *const *paths, size_t n, char ***names_out, size_t *n_names) {
    if (!names_out || !n_names) { set_error("contig_names: NULL argument"); return GALAH_B200_ERR_ARG; }
    std::vector<std::string> all;
    for (size_t f = 0; f < n; f++) {
        std::vector<std::string> names;
        if (int rc = record_names(paths[f], names)) return rc;
        for (auto &nm : names) {
            if (std::find(all.begin(), all.end(), nm) != all.end()) {  // src/cluster_argument_parsing.rs:622-627
                set_error(std::string("Duplicate contig name found in file '") + paths[f] + "': " + nm);
                return GALAH_B200_ERR_UNSUPPORTED;
            }
            all.push_back(nm);
        }
    }
    char **res = (char **)malloc(std::max<size_t>(all.size(), 1) * sizeof(char *));
    for (size_t x = 0; x < all.size(); x++) res[x] = strdup(all[x].c_str());
    *names_out = res; *n_names = all.size();
    return 0;
}

void galah_b200_contig_names_free(char **names, size_t n) {
    if (!names) return;
    for (size_t x = 0; x < n; x++) free(names[x]);
    free(names);
}

int galah_b200_session_skani_distances_contigs(galah_b200_session_t *s, const char *const *paths, size_t n,
                                               const char *const *contig_names, size_t n_names, float threshold_pct,
                                               float min_aligned_threshold, int small_genomes, int host_threads,
                                               galah_b200_pair_t **out, size_t *n_out) {
    if (!s || !out || !n_out) { set_error("session: NULL argument"); return GALAH_B200_ERR_ARG; }
    *out = nullptr; *n_out = 0;
    // the reference maps skani's Ref_name / Query_name columns back to positions in contig_names
    // (src/skani.rs:460-474, tabs sanitised): unit u of the device path is record u in file order,
    // so the names of the records are checked against contig_names first
    std::vector<std::string> recs;
    for (size_t f = 0; f < n; f++) if (int rc = record_names(paths[f], recs)) return rc;
    std::unordered_map<std::string, uint32_t> pos;
    for (size_t x = 0; x < n_names; x++) {
        std::string nm(contig_names[x]);
        for (auto &ch : nm) if (ch == '\t') ch = ' ';
        if (!pos.count(nm)) pos[nm] = (uint32_t)x;  // `position()` finds the first occurrence
    }
    std::vector<uint32_t> unit_to_name(recs.size());
    for (size_t u = 0; u < recs.size(); u++) {
        std::string nm = recs[u];
        for (auto &ch : nm) if (ch == '\t') ch = ' ';
        auto it = pos.find(nm);
        if (it == pos.end()) { set_error("Failed to find contig name in contig_names: " + recs[u]); return GALAH_B200_ERR_UNSUPPORTED; }
        unit_to_name[u] = it->second;
    }
    std::unique_lock<std::shared_mutex> lk(s->mu);
    std::lock_guard<std::mutex> lock(g_mu);
    if (int rc = require_ctx()) return rc;
    std::vector<galah_b200_pair_t> hits;
    size_t units = 0; uint64_t screened = 0;
    const float min_af_pct = min_aligned_threshold * 100.0f;
    if (int rc = skani_distances_impl(paths, n, threshold_pct, min_af_pct, small_genomes != 0, true, host_threads, hits,
                                      units, screened))
        return rc;
    if (units != recs.size()) { set_error("contig mode: the device path saw a different number of records than the header scan"); return GALAH_B200_ERR_ARG; }
    for (auto &h : hits) {
        const uint32_t a = unit_to_name[h.i], b = unit_to_name[h.j];
        h.i = std::min(a, b); h.j = std::max(a, b);
    }
    std::sort(hits.begin(), hits.end(), [](const galah_b200_pair_t &a, const galah_b200_pair_t &b) { return a.i != b.i ? a.i < b.i : a.j < b.j; });
    s->stash_paths.clear(); s->stash_hits.clear();
    return take_hits(hits, out, n_out);
}

int galah_b200_session_skani_distances_with_references(galah_b200_session_t *s, const char *const *combined_paths,
                                                       size_t n, const char *const *reference_paths, size_t n_refs,
                                                       float threshold_pct, float min_aligned_threshold,
                                                       int small_genomes, int host_threads, galah_b200_pair_t **out,
                                                       size_t *n_out) {
    if (!s || !out || !n_out) { set_error("session: NULL argument"); return GALAH_B200_ERR_ARG; }
    *out = nullptr; *n_out = 0;
    std::vector<uint8_t> is_ref(n, 0);
    for (size_t g = 0; g < n; g++)
        for (size_t r = 0; r < n_refs; r++)
            if (strcmp(combined_paths[g], reference_paths[r]) == 0) { is_ref[g] = 1; break; }
    std::unique_lock<std::shared_mutex> lk(s->mu);
    std::lock_guard<std::mutex> lock(g_mu);
    if (int rc = require_ctx()) return rc;
    std::vector<galah_b200_pair_t> hits;
    size_t units = 0; uint64_t screened = 0;
    const float min_af_pct = min_aligned_threshold * 100.0f;
    if (int rc = skani_distances_impl(combined_paths, n, threshold_pct, min_af_pct, small_genomes != 0, false, host_threads,
                                      hits, units, screened, kSkaniReferences, &is_ref))
        return rc;
    s->stash_paths.clear(); s->stash_hits.clear();
    return take_hits(hits, out, n_out);
}

int galah_b200_session_calculate_ani(galah_b200_session_t *s, const char *fasta1, const char *fasta2,
                                     float min_aligned_threshold, int small_genomes, float *ani, int *is_some) {
    if (!s || !fasta1 || !fasta2 || !ani) { set_error("session: NULL argument"); return GALAH_B200_ERR_ARG; }
    const float min_af_pct = min_aligned_threshold * 100.0f;  // src/skani.rs:731-733
    if (is_some) *is_some = 1;  // SkaniClusterer always answers Some (src/skani.rs:708-715)
    const std::string p1(fasta1), p2(fasta2);
    {
        std::shared_lock<std::shared_mutex> lk(s->mu);
        if (s->index && s->small_genomes == (small_genomes != 0) && s->min_af_pct == min_af_pct) {
            auto a = s->path_id.find(p1), b = s->path_id.find(p2);
            if (a != s->path_id.end() && b != s->path_id.end()) {
                auto c = s->cache.find(((uint64_t)a->second << 32) | b->second);
                if (c != s->cache.end()) { *ani = c->second; return 0; }
            }
        }
    }
    std::unique_lock<std::shared_mutex> lk(s->mu);
    std::lock_guard<std::mutex> lock(g_mu);
    if (int rc = require_ctx()) return rc;
    s->n_calls++;
    if (s->index && s->small_genomes != (small_genomes != 0)) { s->index.reset(); s->path_id.clear(); s->cache.clear(); }
    if (s->min_af_pct != min_af_pct) s->cache.clear();
    s->small_genomes = small_genomes != 0; s->min_af_pct = min_af_pct;
    // genomes of the stashed distances() call are indexed together with the two asked for
    std::vector<std::string> need = s->stash_paths;
    need.push_back(p1); need.push_back(p2);
    if (int rc = session_index_paths(s, need, 0)) return rc;
    std::vector<uint32_t> pairs;
    for (const auto &h : s->stash_hits) {  // every hit of the preclusterer, query = the lower index (the representative)
        pairs.push_back(s->path_id[s->stash_paths[h.i]]); pairs.push_back(s->path_id[s->stash_paths[h.j]]);
    }
    s->stash_hits.clear(); s->stash_paths.clear();
    pairs.push_back(s->path_id[p1]); pairs.push_back(s->path_id[p2]);
    if (int rc = session_compute(s, pairs)) return rc;
    *ani = s->cache[((uint64_t)s->path_id[p1] << 32) | s->path_id[p2]];
    return 0;
}

int galah_b200_session_stats(galah_b200_session_t *s, uint64_t *n_indexed, uint64_t *n_pairs_computed, uint64_t *n_launches) {
    if (!s) { set_error("session: NULL session"); return GALAH_B200_ERR_ARG; }
    std::shared_lock<std::shared_mutex> lk(s->mu);
    if (n_indexed) *n_indexed = s->index ? s->index->size() : 0;
    if (n_pairs_computed) *n_pairs_computed = s->n_pairs_computed;
    if (n_launches) *n_launches = s->n_launches;
    return 0;
}

void galah_b200_clusters_free(galah_b200_clusters_t *c) {
    if (!c) return;
    free(c->members); free(c->offsets);
    memset(c, 0, sizeof(*c));
}

int galah_b200_pack_fasta_file(const char *path, uint32_t **seq2, uint32_t **valid, uint64_t *n_bases,
                               uint64_t **rec_start, uint64_t **rec_end, size_t *n_records) {
    PackedGenome pg;
    std::string err;
    if (pack_fasta_file(path, pg, false, err)) { set_error(err); return GALAH_B200_ERR_IO; }
    auto dup = [](const void *src, size_t bytes) { void *p = malloc(std::max<size_t>(bytes, 8)); if (p && bytes) memcpy(p, src, bytes); return p; };
    *seq2 = (uint32_t *)dup(pg.seq2.data(), pg.seq2.size() * 4);
    *valid = (uint32_t *)dup(pg.valid.data(), pg.valid.size() * 4);
    *rec_start = (uint64_t *)dup(pg.rec_start.data(), pg.rec_start.size() * 8);
    *rec_end = (uint64_t *)dup(pg.rec_end.data(), pg.rec_end.size() * 8);
    *n_bases = pg.n_bases; *n_records = pg.rec_start.size();
    return 0;
}

// K0 parity / measurement hook: decode a batch of in-memory FASTA files on the device and return
// the packed arrays (host copies) with the per-file metadata -- compared bit for bit with the
// host packer (galah_b200_pack_fasta_file) by the tests.
int galah_b200_decode_fasta_device(const uint8_t *const *files, const size_t *lens, size_t n, uint32_t **seq2,
                                   uint32_t **valid, uint64_t **base_off, uint64_t **n_bases, uint64_t **rec_off,
                                   uint64_t **rec_start, uint64_t **rec_end, uint64_t **n_ambiguous, uint64_t **n_N,
                                   float *device_ms) {
    std::lock_guard<std::mutex> lock(g_mu);
    if (int rc = require_ctx()) return rc;
    std::vector<std::vector<uint8_t>> raw(n);
    for (size_t f = 0; f < n; f++) raw[f].assign(files[f], files[f] + lens[f]);
    RawBatch rb;
    if (int rc = stage_raw(raw, rb)) return rc;
    if (!rb.all_fasta) { set_error("decode_fasta_device: a file does not start with '>'"); return GALAH_B200_ERR_UNSUPPORTED; }
    DecodedFiles dec;
    if (int rc = g_ctx.fasta.decode(rb.bytes, rb.file_off, rb.file_len, rb.first_byte, dec, g_ctx.stream)) return rc;
    const uint64_t total = dec.base_off[n];
    std::vector<uint32_t> h_seq2(total / 16 + 4, 0u), h_valid(total / 32 + 4, 0u);
    GB_CUDA(cudaMemcpyAsync(h_seq2.data(), dec.d_seq2, (total / 16) * 4, cudaMemcpyDeviceToHost, g_ctx.stream));
    GB_CUDA(cudaMemcpyAsync(h_valid.data(), dec.d_valid, (total / 32) * 4, cudaMemcpyDeviceToHost, g_ctx.stream));
    GB_CUDA(cudaStreamSynchronize(g_ctx.stream));
    auto dup = [](const void *src, size_t bytes) { void *p = malloc(std::max<size_t>(bytes, 8)); if (p && bytes) memcpy(p, src, bytes); return p; };
    *seq2 = (uint32_t *)dup(h_seq2.data(), h_seq2.size() * 4);
    *valid = (uint32_t *)dup(h_valid.data(), h_valid.size() * 4);
    *base_off = (uint64_t *)dup(dec.base_off.data(), (n + 1) * 8);
    *n_bases = (uint64_t *)dup(dec.n_bases.data(), n * 8);
    *rec_off = (uint64_t *)dup(dec.rec_off.data(), (n + 1) * 8);
    *rec_start = (uint64_t *)dup(dec.rec_start.data(), dec.rec_start.size() * 8);
    *rec_end = (uint64_t *)dup(dec.rec_end.data(), dec.rec_end.size() * 8);
    *n_ambiguous = (uint64_t *)dup(dec.n_ambiguous.data(), n * 8);
    *n_N = (uint64_t *)dup(dec.n_N.data(), n * 8);
    if (device_ms) *device_ms = g_ctx.fasta.last_ms;
    return 0;
}

int galah_b200_device_ingest(int enable) {
    std::lock_guard<std::mutex> lock(g_mu);
    const int prev = g_ctx.device_ingest;
    if (enable >= 0) g_ctx.device_ingest = enable ? 1 : 0;
    return prev;
}

int galah_b200_genome_stats(const char *const *paths, size_t n, int host_threads, galah_b200_genome_stats_t *out) {
    if (host_threads <= 0) host_threads = (int)std::max(1u, std::thread::hardware_concurrency());
    std::vector<std::string> errs(n);
    std::vector<int> rcs(n, 0);
    std::atomic<size_t> next{0};
    auto worker = [&]() {
        for (;;) {
            const size_t x = next.fetch_add(1);
            if (x >= n) break;
            PackedGenome pg;
            rcs[x] = pack_fasta_file(paths[x], pg, false, errs[x]);
            if (rcs[x]) continue;
            const GenomeAssemblyStats st = genome_stats(pg);
            if (!st.n50_valid) { rcs[x] = GALAH_B200_ERR_UNSUPPORTED; errs[x] = std::string("Failed to calculate n50 from ") + paths[x]; continue; }
            out[x] = galah_b200_genome_stats_t{st.num_contigs, st.num_ambiguous_bases, st.n50};
        }
    };
    std::vector<std::thread> th;
    const int nt = (int)std::min<size_t>((size_t)host_threads, std::max<size_t>(n, 1));
    for (int t = 1; t < nt; t++) th.emplace_back(worker);
    worker();
    for (auto &t : th) t.join();
    for (size_t x = 0; x < n; x++)
        if (rcs[x]) { set_error(errs[x]); return rcs[x] == GALAH_B200_ERR_UNSUPPORTED ? rcs[x] : GALAH_B200_ERR_IO; }
    return 0;
}

int galah_b200_synth_packed_device(uint64_t seed, uint64_t index_begin, size_t n, uint64_t length,
                                   uint32_t *d_seq2, uint32_t *d_valid, uint64_t *d_base_off, void *stream) {
    std::lock_guard<std::mutex> lock(g_mu);
    if (int rc = require_ctx()) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    return synth_enqueue(seed, index_begin, n, length, d_seq2, d_valid, d_base_off, st);
}

int galah_b200_synth_packed_device_ex(uint64_t seed, uint64_t index_begin, size_t n, uint64_t length,
                                      uint32_t family_size, uint32_t rate_shift, uint32_t *d_seq2,
                                      uint32_t *d_valid, uint64_t *d_base_off, void *stream) {
    std::lock_guard<std::mutex> lock(g_mu);
    if (int rc = require_ctx()) return rc;
    return synth_enqueue(seed, index_begin, n, length, d_seq2, d_valid, d_base_off, (cudaStream_t)stream, family_size,
                         rate_shift);
}

void *galah_b200_stream(void) { return (void *)g_ctx.stream; }

}  // extern "C"
