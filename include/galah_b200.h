/*
 * galah_b200.h -- C ABI of libgalah_b200.so, the B200 (sm_100a) implementation of Galah's
 * two-stage dereplication hot path.  Plain pointers and sizes only; no torch / C++ types.
 *
 * Each entry point names the reference interface it replaces (file:line into wwood/galah
 * v0.5.1).  The Rust-side binding a maintainer would add is shown in INTEGRATION.md.
 *
 * Conventions
 *   - every function returns 0 on success, non-zero on failure; galah_b200_last_error() then
 *     returns a thread-local, NUL-terminated message.  There is NO CPU fallback: with no usable
 *     sm_100 device every compute entry point fails with GALAH_B200_ERR_NO_DEVICE.
 *   - "host" entry points take host pointers and do their own H2D/D2H copies;
 *     "_device" entry points take device pointers valid on the current device and enqueue on
 *     the given cudaStream_t (passed as void*; NULL = the CUDA default stream, as usual).
 *   - a sketch table is n rows of `stride` uint64 (stride >= s, stride even), row g holding
 *     counts[g] ascending distinct hashes followed by 0xFFFFFFFFFFFFFFFF padding.
 *   - buffers returned through `**out` are owned by the library; release with galah_b200_free().
 */
#ifndef GALAH_B200_H
#define GALAH_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GALAH_B200_OK 0
#define GALAH_B200_ERR_NO_DEVICE 1
#define GALAH_B200_ERR_CUDA 2
#define GALAH_B200_ERR_ARG 3
#define GALAH_B200_ERR_IO 4
#define GALAH_B200_ERR_UNSUPPORTED 5 /* mirrors a reference panic; message carries its text */

/* One stage-1 hit.  (i, j) are positions in the caller's genome slice with i < j -- the key of
 * SortedPairGenomeDistanceCache (src/sorted_pair_genome_distance_cache.rs:22-28); `ani` is the
 * `Some(distance as f32)` value stored at src/finch.rs:92; common/total are finch's
 * raw_distance integers, exposed so parity can be checked bit-exactly. */
typedef struct galah_b200_pair {
    uint32_t i, j;
    uint32_t common, total;
    float ani;
} galah_b200_pair_t;

/* ---- library / device ------------------------------------------------------------------ */

/* Bind the calling process to CUDA device `device` (>= 0) and create the library stream.
 * Must be called before any compute entry point; may be called again to switch device. */
int galah_b200_init(int device);
int galah_b200_device_count(void);
const char *galah_b200_last_error(void);
const char *galah_b200_version(void);
void galah_b200_free(void *p);
/* Number of kernels this library has launched since load (bench.py's gpu_launches). */
uint64_t galah_b200_launch_count(void);

/* ---- stage 1a: sketching ----------------------------------------------------------------
 * Replaces `finch::sketch_files(paths, SketchParams::Mash{kmers_to_sketch: s, final_size: s,
 * no_strict: true, kmer_length: k, hash_seed: seed}, filters off)` at src/finch.rs:55-69.
 * FASTA/FASTQ, plain or gzip; all records of a file pooled; needletail normalize(false)
 * semantics.  hashes: n*s uint64 (row stride s, s even), counts: n. */
int galah_b200_sketch_files(const char *const *paths, size_t n, uint8_t k, uint32_t s,
                            uint64_t seed, int host_threads, uint64_t *hashes, uint32_t *counts);

/* Same kernel on sequence already packed on the HOST: 2 bits/base LSB-first in uint32 words
 * (A0 C1 G2 T3), plus a validity bitmap (1 bit/base, 1 = ACGT).  Genome g covers bases
 * [base_off[g], base_off[g+1]) of the concatenated arrays; base_off[g] must be a multiple of
 * 128.  Records inside a genome are separated by at least one invalid base. */
int galah_b200_sketch_packed(const uint32_t *seq2, const uint32_t *valid, const uint64_t *base_off,
                             size_t n, uint8_t k, uint32_t s, uint64_t seed, uint64_t *hashes,
                             uint32_t *counts);

/* Device-resident variant (all pointers are device pointers; d_hashes has row stride s). */
int galah_b200_sketch_packed_device(const uint32_t *d_seq2, const uint32_t *d_valid,
                                    const uint64_t *d_base_off, size_t n, uint8_t k, uint32_t s,
                                    uint64_t seed, uint64_t *d_hashes, uint32_t *d_counts,
                                    void *stream);

/* ---- stage 1b: all-pairs prefilter --------------------------------------------------------
 * Replaces the serial pair loop at src/finch.rs:75-95 (finch::distance::distance ->
 * raw_distance -> mash_distance; keep iff 1 - d >= min_ani as f64; store as f32).
 * Evaluates every pair i < j.
 * Output is sorted by (i, j).  min_ani is a FRACTION (src/finch.rs:5-6). */
int galah_b200_prefilter(const uint64_t *hashes, const uint32_t *counts, size_t n, size_t stride,
                         uint8_t k, float min_ani, galah_b200_pair_t **out, size_t *n_out);

/* Row-sharded variants for multi-GPU runs: the call covers the row blocks (of
 * GALAH_B200_ROW_BLOCK rows) that the boustrophedon rule gives `shard`: block b belongs to
 * shard p = b % n_shards in even rounds (b / n_shards) and n_shards-1-p in odd rounds, which
 * balances the triangular pair area across shards.  _shard takes host pointers, _device device pointers;
 * result pairs always land on the HOST. */
#define GALAH_B200_ROW_BLOCK 128
int galah_b200_prefilter_shard(const uint64_t *hashes, const uint32_t *counts, size_t n,
                               size_t stride, uint8_t k, float min_ani, uint32_t shard,
                               uint32_t n_shards, galah_b200_pair_t **out, size_t *n_out);
int galah_b200_prefilter_device(const uint64_t *d_hashes, const uint32_t *d_counts, size_t n,
                                size_t stride, uint8_t k, float min_ani, uint32_t shard,
                                uint32_t n_shards, void *stream, galah_b200_pair_t **out,
                                size_t *n_out);

/* Selects the kernel path of the prefilter entry points above (both are exact and return the
 * same pair list): 0 = block-list join (default), 1 = pairwise warp merge of every pair.
 * Returns the previous mode; any other value only queries. */
int galah_b200_prefilter_mode(int mode);

/* Device time of the most recent prefilter launch, from CUDA events the library records on the
 * launch stream: build_ms = block-list build kernels (0 in mode 1), main_ms = the join /
 * pairwise kernel alone.  Blocks until that launch has finished. */
int galah_b200_prefilter_last_timing(float *build_ms, float *main_ms);

/* Host-buffer calls (galah_b200_prefilter / _shard with n_shards == 1, mode 0) upload the table
 * in `chunks` slices of whole row blocks on a copy stream and build the block lists of each
 * slice as soon as it is resident, so the build hides under the PCIe transfer and only the join
 * is left when the transfer ends; the join writes its survivors into mapped pinned host memory
 * and the calling thread evaluates their f64 formula while the kernels still run.  chunks <= 1 disables the pipeline (one upload, then the
 * kernels); < 0 only queries.  Returns the previous setting (default 8).  The pair list is
 * identical either way. */
int galah_b200_prefilter_stream_chunks(int chunks);

/* Host wall-clock breakdown of the most recent host-buffer prefilter call, milliseconds:
 * ms4[0] = enqueue (uploads + launches issued), [1] = wait for the device, [2] = candidate
 * read-back beyond the first 65,536, [3] = host finish (f64 formula, threshold, ordering). */
int galah_b200_prefilter_last_host_timing(float *ms4);

/* Kernel-only timing hook used by bench.py: enqueues the prefilter kernels for one shard on
 * `stream` and leaves the candidate list on the device (d_cand, capacity cand_cap entries of
 * 4 x uint32 {i, j, common, total}; d_n_cand is a device uint64 counter).  No host sync.
 * mode: 0 = block-list join, 1 = pairwise warp merge of every pair. */
int galah_b200_prefilter_enqueue(const uint64_t *d_hashes, const uint32_t *d_counts, size_t n,
                                 size_t stride, uint8_t k, float min_ani, uint32_t shard,
                                 uint32_t n_shards, int mode, void *stream, uint32_t *d_cand,
                                 size_t cand_cap, unsigned long long *d_n_cand);

/* Host finish of device candidates (n_cand x {i, j, common, total} as 4 x uint32, copied back
 * by the caller): the reference's f64 Mash-ANI formula, the `>= min_ani as f64` test and the f32
 * store (src/finch.rs:78-93); result sorted by (i, j).  Pure host code. */
int galah_b200_finish_candidates(const uint32_t *cand, size_t n_cand, uint8_t k, float min_ani,
                                 galah_b200_pair_t **out, size_t *n_out);

/* Two-phase form of the block-list join for multi-GPU runs, so that the BUILD shards too: every
 * rank builds the lists of a contiguous slice of blocks from the (all-gathered) table, the slices
 * are all-gathered (9 bytes per table entry), and every rank joins its row-block shard.
 *   layout: n_blocks = ceil(n / 64) lists of entries_per_block = 64 * stride entries each; the
 *           gathered arrays need `slack` readable entries behind the last list.
 *   build:  lists of blocks [block_begin, block_end) -> d_hi / d_lo (uint32), d_tags (uint8),
 *           d_len (uint32), all SLICE-based (block b at offset (b - block_begin) * entries_per_block).
 *           block_end may exceed n_blocks (empty lists) so that slices are equal-sized.
 *   join:   d_hi / d_lo / d_tags / d_len cover blocks 0 .. (block b at b * entries_per_block).
 * All pointers are device pointers; everything is enqueued on `stream`. */
int galah_b200_blocklist_layout(size_t n, size_t stride, size_t *n_blocks, size_t *entries_per_block,
                                size_t *slack);
int galah_b200_blocklist_build(const uint64_t *d_hashes, const uint32_t *d_counts, size_t n,
                               size_t stride, size_t block_begin, size_t block_end, uint32_t *d_hi,
                               uint32_t *d_lo, uint8_t *d_tags, uint32_t *d_len, void *stream);
/* Join of an explicit list of block pairs: d_items holds n_items (rb, cb) pairs of uint32 with
 * rb <= cb (put the diagonal pairs first: they take longest).  Every pair's two lists and the
 * table rows of both blocks must be resident; candidates are APPENDED to d_cand (reset_candidates
 * != 0 zeroes the counter first).  The multi-GPU ring uses one call per peer slice as its lists
 * arrive, so the exchange overlaps the join. */
int galah_b200_prefilter_join_items_enqueue(const uint64_t *d_hashes, const uint32_t *d_counts, size_t n,
                                            size_t stride, uint8_t k, float min_ani, const uint32_t *d_hi,
                                            const uint32_t *d_lo, const uint8_t *d_tags, const uint32_t *d_len,
                                            const uint32_t *d_items, size_t n_items, int reset_candidates,
                                            void *stream, uint32_t *d_cand, size_t cand_cap,
                                            unsigned long long *d_n_cand);

/* Multi-GPU form that needs no gathered table for the build: every rank holds a slice of whole
 * row blocks (its first row a multiple of GALAH_B200_ROW_BLOCK).  table_max: largest valid hash
 * of the slice -> *d_max (device uint64; combine across ranks with an all-reduce MAX).
 * build_local: the lists of the slice's blocks (n_blocks_out >= ceil(n_rows / ROW_BLOCK); surplus
 * blocks come out empty) with the table-wide maximum supplied in *d_table_max, slice-based
 * outputs as above.  The table all-gather can then overlap the build. */
int galah_b200_table_max_device(const uint64_t *d_hashes, const uint32_t *d_counts, size_t n, size_t stride,
                                unsigned long long *d_max, void *stream);
int galah_b200_blocklist_build_local(const uint64_t *d_rows, const uint32_t *d_counts, size_t n_rows,
                                     size_t stride, const unsigned long long *d_table_max,
                                     size_t n_blocks_out, uint32_t *d_hi, uint32_t *d_lo, uint8_t *d_tags,
                                     uint32_t *d_len, void *stream);
int galah_b200_prefilter_join_enqueue(const uint64_t *d_hashes, const uint32_t *d_counts, size_t n,
                                      size_t stride, uint8_t k, float min_ani, const uint32_t *d_hi,
                                      const uint32_t *d_lo, const uint8_t *d_tags,
                                      const uint32_t *d_len, uint32_t shard, uint32_t n_shards,
                                      void *stream, uint32_t *d_cand, size_t cand_cap,
                                      unsigned long long *d_n_cand);

/* Whole `finch::distances(paths, min_ani, num_kmers, kmer_length)` (src/finch.rs:48-97):
 * sketch every path on the GPU, then the all-pairs prefilter. */
int galah_b200_finch_distances(const char *const *paths, size_t n, float min_ani,
                               uint32_t num_kmers, uint8_t kmer_length, int host_threads,
                               galah_b200_pair_t **out, size_t *n_out);

/* ---- stage 2: ANI of genome pairs -----------------------------------------------------------
 * Replaces ClusterDistanceFinder::calculate_ani for SkaniClusterer (src/skani.rs:689-788: one
 * `skani dist --min-af X [--small-genomes] -q fasta1 -r fasta2` subprocess per pair, ANI = TSV
 * column 3 parsed as f32, 0.0 when skani prints no row).  Genomes are indexed ONCE (FracMinHash
 * seeds + hash table, resident in HBM) and any number of pairs is evaluated per call.  The method
 * is a restatement of skani's published algorithm, specified in oracle/skani_oracle.c; numeric
 * parity with the skani binary is unpinned. */
typedef struct galah_b200_ani_index galah_b200_ani_index_t;
typedef struct galah_b200_ani_result {
    float ani;                /* percent, two decimals as skani prints; 0.0 = "no row" */
    float af_query, af_ref;   /* aligned fractions (0..1) of the query / reference genome */
    uint32_t estimator;       /* 0 = mean of the per-chunk identities; 1 = chain-span ratio (stands in
                                 for skani's learned ANI: c >= 70, >= 150 kb aligned, not contigs) */
    uint64_t sum_fx;          /* sum over counted query chunks of round(2^40 (M/N)^(1/15)) */
    uint32_t n_chunks;        /* query chunks (20 kb) holding at least one chain of >= 3 anchors */
    uint32_t sum_m;           /* matched seeds inside those chunks' chained spans */
    uint32_t span_m, span_n;  /* chained anchors / query seeds inside chain spans */
    uint32_t n_chains;        /* chains of >= 3 anchors */
    uint32_t cov_q, cov_r;    /* bases covered by chains on the query / reference */
    uint32_t reserved;
} galah_b200_ani_result_t;

/* small_genomes != 0 selects c = 30 (skani --small-genomes, src/skani.rs:735-737), else c = 125 */
int galah_b200_ani_index_create(int small_genomes, galah_b200_ani_index_t **out);
void galah_b200_ani_index_free(galah_b200_ani_index_t *idx);
/* Appends genomes; genome ids are assigned in order of addition, starting at 0. */
int galah_b200_ani_index_add_files(galah_b200_ani_index_t *idx, const char *const *paths, size_t n,
                                   int host_threads);
/* Host packed sequence (layout of galah_b200_sketch_packed) plus contig tables: genome g has
 * contigs contig_off[g]..contig_off[g+1], each [contig_start, contig_start + contig_len) relative
 * to the genome's first base, separated by at least one invalid base. */
int galah_b200_ani_index_add_packed(galah_b200_ani_index_t *idx, const uint32_t *seq2,
                                    const uint32_t *valid, const uint64_t *base_off, size_t n,
                                    const uint64_t *contig_off, const uint32_t *contig_start,
                                    const uint32_t *contig_len);
/* Device-resident packed genomes, one contig each of `lengths[g]` bases (synthetic inputs).
 * base_off and lengths are HOST arrays (n+1 and n entries); d_* are device pointers. */
int galah_b200_ani_index_add_packed_device(galah_b200_ani_index_t *idx, const uint32_t *d_seq2,
                                           const uint32_t *d_valid, const uint64_t *d_base_off,
                                           const uint64_t *base_off, const uint64_t *lengths,
                                           size_t n, void *stream);
/* Host finish of stage 2 (no device needed): the kernel's integer accumulators of one pair (the
 * integer fields of `ints`) -> what galah parses from skani's TSV (src/skani.rs:773-779): the
 * ANI (estimator 0 or 1, see the struct) printed with two decimals and parsed as f32; 0.0 when
 * max(AF) * 100 < min_af_pct (no row, src/skani.rs:760).
 * galah_b200_print2_parse_f32(v) == strtof(sprintf("%.2f", v));
 * galah_b200_chunk_identity_fx(m, n) == round(2^40 (m/n)^(1/15)), one chunk's term of sum_fx. */
int galah_b200_ani_finish(const galah_b200_ani_result_t *ints, uint64_t len_q, uint64_t len_r, int small_genomes,
                          int individual_contigs, float min_af_pct, galah_b200_ani_result_t *out);
float galah_b200_print2_parse_f32(double v);
uint64_t galah_b200_chunk_identity_fx(uint32_t m, uint32_t n);

/* Capacity hint: the index will hold n_total_genomes genomes like the ones already added (call
 * it after the first batch).  Avoids re-allocating the device arrays while the index grows. */
int galah_b200_ani_index_reserve(galah_b200_ani_index_t *idx, size_t n_total_genomes);
/* Forgets every genome (and detaches attached peers) but keeps the device allocations. */
int galah_b200_ani_index_clear(galah_b200_ani_index_t *idx);
size_t galah_b200_ani_index_size(const galah_b200_ani_index_t *idx);
int galah_b200_ani_index_genome(const galah_b200_ani_index_t *idx, size_t g, uint64_t *n_seeds,
                                uint32_t *n_chunks, uint64_t *total_len);
/* Parity hook: the seeds of genome g in position order (kmer << 1 | strand, spread position,
 * chunk id), cap >= n_seeds entries each. */
int galah_b200_ani_index_seeds(const galah_b200_ani_index_t *idx, size_t g, uint32_t *kmer_strand,
                               uint32_t *spread, uint32_t *chunk, size_t cap);
/* pairs: 2 * n_pairs genome ids, (query, reference) -- the query is the FIRST id, as in
 * `skani dist -q fasta1 -r fasta2` (src/skani.rs:733-744; galah passes the representative first,
 * src/clusterer.rs:262-296); results: n_pairs entries.  min_af_pct as skani's --min-af.
 * individual_contigs != 0: the units are FASTA records (`skani triangle -i`, src/skani.rs:413). */
int galah_b200_ani_pairs(galah_b200_ani_index_t *idx, const uint32_t *pairs, size_t n_pairs,
                         float min_af_pct, int individual_contigs, galah_b200_ani_result_t *results);
/* Device time (CUDA events) of the last index batch build and the last pair evaluation. */
int galah_b200_ani_last_timing(const galah_b200_ani_index_t *idx, float *build_ms, float *chain_ms);

/* ---- clustering engine (host logic) --------------------------------------------------------
 * Replaces the body of galah::clusterer::cluster() after the preclusterer has run
 * (src/clusterer.rs:56-151): partition_sketches (:452-487), preclusters largest first (:67-79),
 * find_precluster_cluster_representatives (:182-259), find_precluster_cluster_memberships
 * (:350-449).  Driven by the sparse hit list instead of N^2/2 BTreeMap probes; output order is
 * the reference's at --threads 1.  Pure host code: needs no device.
 *
 * hits: the preclusterer's SortedPairGenomeDistanceCache as (i, j, ani) records (common/total
 * are ignored).  skip_clusterer: src/clusterer.rs:32-44 (same method names, or contigs).
 * ani_threshold: ClusterDistanceFinder::get_ani_threshold() (src/lib.rs:52).
 * calculate_ani: ClusterDistanceFinder::calculate_ani(fasta1 = representative, fasta2 = genome)
 * (src/lib.rs:54) as indices; returns 1 and sets *ani for Some(ani), 0 for None.  May be NULL
 * when skip_clusterer != 0. */
typedef int (*galah_b200_ani_fn)(void *ctx, uint32_t representative, uint32_t genome, float *ani);
typedef struct galah_b200_clusters {
    uint32_t *members;      /* concatenated clusters; the representative is first in each */
    uint64_t *offsets;      /* n_clusters + 1 */
    size_t n_clusters;
    uint64_t ani_calls;     /* calculate_ani invocations made */
    uint32_t n_preclusters; /* "Found {} preclusters. The largest contained {} genomes" */
    uint32_t largest_precluster;
} galah_b200_clusters_t;
int galah_b200_cluster_from_distances(size_t n_genomes, const galah_b200_pair_t *hits,
                                      size_t n_hits, int skip_clusterer, float ani_threshold,
                                      galah_b200_ani_fn calculate_ani, void *ctx,
                                      galah_b200_clusters_t *out);
/* Same engine with calculate_ani served from a table: ani[x] is the clusterer's ANI (percent)
 * of hits[x]; hits sorted by (i, j).  This is how the Rust shim serves calculate_ani after one
 * batched galah_b200_ani_pairs call over every precluster hit. */
int galah_b200_cluster_from_ani_table(size_t n_genomes, const galah_b200_pair_t *hits, size_t n_hits,
                                      const float *ani, float ani_threshold, galah_b200_clusters_t *out);
/* The same with both orientations of every hit: ani_fwd[x] = ANI with hits[x].i as the query,
 * ani_rev[x] = with hits[x].j as the query.  calculate_ani(representative, genome) makes the
 * representative the query (skani dist -q fasta1, src/skani.rs:733-744); the membership pass asks
 * for representatives on either side of the genome (src/clusterer.rs:375-384). */
int galah_b200_cluster_from_ani_tables(size_t n_genomes, const galah_b200_pair_t *hits, size_t n_hits,
                                       const float *ani_fwd, const float *ani_rev, float ani_threshold,
                                       galah_b200_clusters_t *out);
/* The same engine with stage 2 asked for in BATCHES (what a GPU backend wants): the engine works out
 * in waves which (representative, genome) pairs the reference's two passes evaluate
 * (src/clusterer.rs:216-300, 350-449) -- every precluster hit of a representative with a genome that is
 * not an earlier representative -- and hands them to calculate_ani_batch, all preclusters together;
 * wave w holds the pairs of the representatives confirmed by wave w - 1.  For a collection of
 * near-identical genomes that is (representatives x genomes) evaluations instead of one per hit.
 * The callback fills some[x] (0 = None) and ani[x] for the n pairs (reps[x] is the QUERY, as
 * calculate_ani(fasta1, fasta2) makes fasta1, src/skani.rs:733-744) and returns 0, or non-zero to
 * abort.  After max_waves waves (0: 16) everything the undecided genomes can still ask for goes
 * out as one batch.  Clusters, their order and ani_calls equal galah_b200_cluster_from_distances'
 * (ani_calls as long as the wave budget holds).  *n_waves (optional): callback invocations. */
typedef int (*galah_b200_ani_batch_fn)(void *ctx, const uint32_t *reps, const uint32_t *genomes, size_t n,
                                       uint8_t *some, float *ani);
int galah_b200_cluster_from_distances_batched(size_t n_genomes, const galah_b200_pair_t *hits, size_t n_hits,
                                              float ani_threshold, galah_b200_ani_batch_fn calculate_ani_batch,
                                              void *ctx, uint32_t max_waves, galah_b200_clusters_t *out,
                                              uint32_t *n_waves);
/* How the single-device one-call pipelines (galah_b200_cluster_files / _packed*) run stage 2:
 * 1 or -1 (default) = in waves as above, 0 = K3 on every precluster hit up front (+ one launch for
 * the reverse orientations the membership pass needs).  The clusters are the same in both modes. */
int galah_b200_cluster_lazy(int mode);
void galah_b200_clusters_free(galah_b200_clusters_t *c);

/* The whole hot path: galah::clusterer::cluster(genomes, &FinchPreclusterer{min_ani:
 * precluster_min_ani (FRACTION), num_kmers: 1000, kmer_length: 21}, &SkaniClusterer{threshold:
 * ani_threshold_pct (PERCENT), min_aligned_threshold: min_af_pct / 100, small_genomes}, false,
 * None, None) (src/clusterer.rs:14-152 as called from src/cluster_argument_parsing.rs:1514-1530).
 * Every file is read and uploaded once; K1 sketches and the K3 index are built from the same
 * device buffers; K2 gives the precluster hits; the host engine does the greedy selection and
 * asks K3, a batch per wave, for the hit pairs it needs (galah_b200_cluster_lazy).  Cluster order
 * is the reference's at --threads 1. */
typedef struct galah_b200_cluster_stats {
    uint64_t n_precluster_hits;
    uint64_t n_ani_pairs;
    float ani_chain_ms;   /* device time of the K3 chain kernel(s) */
    /* host wall clock of the call's phases (ms): ingest = read/upload + K0 + K1 + K3 index;
     * sketch_ms / index_ms = device time of K1 / the K3 index build inside it (packed entries) */
    float ingest_ms, sketch_ms, index_ms, prefilter_ms, ani_ms, engine_ms, total_ms;
    uint32_t ani_waves;   /* stage-2 batches when it ran in waves (galah_b200_cluster_lazy), else 0 */
} galah_b200_cluster_stats_t;
int galah_b200_cluster_files(const char *const *paths, size_t n, float precluster_min_ani,
                             float ani_threshold_pct, float min_af_pct, int small_genomes,
                             int host_threads, galah_b200_clusters_t *out,
                             galah_b200_cluster_stats_t *stats);

/* The same call on genomes that are already packed (layout of galah_b200_sketch_packed: 2 bits per
 * base + validity bitmap, genome g at bases [base_off[g], base_off[g] + lengths[g]), base_off
 * multiples of 128, one contig per genome).  _packed takes HOST arrays and uploads them in batches
 * of ~1 G bases on a copy stream while the previous batch is sketched and indexed; _packed_device
 * takes arrays resident in HBM (d_base_off = device copy of base_off).  base_off / lengths are host
 * arrays in both.  What bench.py times (`e2e` and `value`). */
int galah_b200_cluster_packed(const uint32_t *seq2, const uint32_t *valid, const uint64_t *base_off,
                              const uint64_t *lengths, size_t n, float precluster_min_ani,
                              float ani_threshold_pct, float min_af_pct, int small_genomes,
                              galah_b200_clusters_t *out, galah_b200_cluster_stats_t *stats);
int galah_b200_cluster_packed_device(const uint32_t *d_seq2, const uint32_t *d_valid, const uint64_t *d_base_off,
                                     const uint64_t *base_off, const uint64_t *lengths, size_t n,
                                     float precluster_min_ani, float ani_threshold_pct, float min_af_pct,
                                     int small_genomes, galah_b200_clusters_t *out,
                                     galah_b200_cluster_stats_t *stats);

/* The same call on SEVERAL GPUs of one process -- what a single galah process on a multi-GPU box
 * calls in place of clusterer::cluster (src/clusterer.rs:14-152; the reference parallelises over
 * rayon threads, src/clusterer.rs:190-214, 267-293: here the unit of parallelism is a device).
 * galah_b200_init_devices(n) binds devices 0 .. n-1 (one context, stream and workspace each) and
 * opens peer access between them; device 0 stays the device of every single-GPU entry point.
 * galah_b200_cluster_packed_multi takes HOST arrays (layout of galah_b200_cluster_packed) and runs
 * one host thread per device: genome slices for K1 / the K3 index, the sketch rows exchanged by
 * peer copies (copy engines over NVLink), row-block shards of the K2 grid, K3 pairs on the device
 * that owns the query genome reading the reference's hash table in place on its peer, the greedy
 * engine on the host.  Clusters are identical to the single-GPU call's. */
int galah_b200_init_devices(int n_devices);
int galah_b200_cluster_packed_multi(const uint32_t *seq2, const uint32_t *valid, const uint64_t *base_off,
                                    const uint64_t *lengths, size_t n, int n_devices,
                                    float precluster_min_ani, float ani_threshold_pct, float min_af_pct,
                                    int small_genomes, galah_b200_clusters_t *out,
                                    galah_b200_cluster_stats_t *stats);
/* galah_b200_cluster_files on several GPUs: device r reads, decodes (K0), sketches and indexes the
 * r-th slice of the path list with host_threads / n_devices reader threads; the rest as above. */
int galah_b200_cluster_files_multi(const char *const *paths, size_t n, int n_devices, float precluster_min_ani,
                                   float ani_threshold_pct, float min_af_pct, int small_genomes,
                                   int host_threads, galah_b200_clusters_t *out,
                                   galah_b200_cluster_stats_t *stats);

/* galah_b200_cluster_packed without the validity bitmap: every base of a genome's `lengths[g]` is
 * valid except the listed ranges [invalid_begin[x], invalid_end[x]) (absolute base coordinates in
 * the packed arrays, sorted, disjoint: the N runs and record breaks a FASTA reader meets --
 * needletail's normalisation inside finch::sketch_files, src/finch.rs:55-69); the bitmap is built
 * on the device per upload batch.  The bitmap is a third of the packed bytes and almost constant:
 * a host-buffer call is bound by PCIe, so not sending it is a third less time on the wire. */
int galah_b200_cluster_packed_sparse(const uint32_t *seq2, const uint64_t *invalid_begin, const uint64_t *invalid_end,
                                     size_t n_invalid, const uint64_t *base_off, const uint64_t *lengths, size_t n,
                                     float precluster_min_ani, float ani_threshold_pct, float min_af_pct,
                                     int small_genomes, galah_b200_clusters_t *out,
                                     galah_b200_cluster_stats_t *stats);

/* First half of the two calls above, for callers that drive the stages themselves (the multi-GPU
 * pipeline, one process per GPU): packed genomes (host arrays if device == 0, else resident) ->
 * K1 sketch rows written to the DEVICE table d_hashes / d_counts (stride 1000) and the genomes
 * appended to the K3 index.  ms2 (optional): device time of K1 and of the index build. */
int galah_b200_ingest_packed(const uint32_t *seq2, const uint32_t *valid, const uint64_t *d_base_off,
                             const uint64_t *base_off, const uint64_t *lengths, size_t n, int device,
                             uint64_t *d_hashes, uint32_t *d_counts, galah_b200_ani_index_t *idx, float *ms2);
/* Host arrays without the validity bitmap (see galah_b200_cluster_packed_sparse). */
int galah_b200_ingest_packed_sparse(const uint32_t *seq2, const uint64_t *invalid_begin, const uint64_t *invalid_end,
                                    size_t n_invalid, const uint64_t *base_off, const uint64_t *lengths, size_t n,
                                    uint64_t *d_hashes, uint32_t *d_counts, galah_b200_ani_index_t *idx, float *ms2);
/* The same for the skani-style preclusterer: the rows are FracMinHash MARKER sketches (k = 21,
 * 1/1000, or 1/200 when idx was created with small_genomes) at row stride marker_stride
 * (galah_b200_marker_row_capacity of the longest unit), overflowing rows flagged 0xFFFFFFFF in
 * d_counts.  galah_b200_prefilter_join_enqueue_screen is the join with the marker-containment rule
 * (faster_small = skani --faster-small, part of --small-genomes) over such rows. */
int galah_b200_ingest_packed_markers(const uint32_t *seq2, const uint32_t *valid, const uint64_t *d_base_off,
                                     const uint64_t *base_off, const uint64_t *lengths, size_t n, int device,
                                     uint32_t marker_stride, uint64_t *d_rows, uint32_t *d_counts,
                                     galah_b200_ani_index_t *idx, float *ms2);
uint32_t galah_b200_marker_row_capacity(uint64_t longest_unit, int small_genomes);
int galah_b200_prefilter_join_enqueue_screen(const uint64_t *d_rows, const uint32_t *d_counts, size_t n, size_t stride,
                                             int faster_small, const uint32_t *d_hi, const uint32_t *d_lo,
                                             const uint8_t *d_tags, const uint32_t *d_len, uint32_t shard,
                                             uint32_t n_shards, void *stream, uint32_t *d_cand, size_t cand_cap,
                                             unsigned long long *d_n_cand);
/* Multi-GPU stage 2 without a collective: a process exports the CUDA IPC handle of its index's
 * hash-table array plus the per-genome slot offsets (size + 1 entries) and lengths (size entries);
 * a peer process on the same NVLink domain attaches it, after which the peer's genomes can be the
 * REFERENCE of a pair under ids first_id .. first_id + n_genomes - 1 (the query stays local): the
 * chain kernel reads the peer's table in place over NVLink.  The exporting index must outlive
 * every attached use (barrier before galah_b200_ani_index_free). */
int galah_b200_ani_index_export_tables(const galah_b200_ani_index_t *idx, uint8_t handle[64],
                                       uint64_t *table_off, uint64_t *total_len);
int galah_b200_ani_index_attach_peer(galah_b200_ani_index_t *idx, const uint8_t handle[64],
                                     const uint64_t *table_off, const uint64_t *total_len,
                                     size_t n_genomes, uint32_t *first_id);

/* ---- skani preclusterer / contig clustering ------------------------------------------------
 * Replaces SkaniPreclusterer::distances and ::distances_contigs (src/skani.rs:21-56; the
 * `skani triangle --sparse [-i]` subprocess of :109-225 and :379-498): FracMinHash marker sketches
 * (k = 21, 1/1000, or 1/200 with small_genomes) of every unit, an all-pairs marker-containment
 * screen on the GPU (the K2 join with the rule common >= max(1, ceil(0.8^21 * min(|A|,|B|))); without
 * small_genomes -- skani's --faster-small is part of --small-genomes -- a pair whose smaller marker
 * sketch has fewer than 20 entries is always compared), the
 * stage-2 ANI kernel on the survivors, keep ANI >= threshold_pct (f32, src/skani.rs:205).
 * per_record != 0 is contig mode: every FASTA record of every file is its own unit, numbered in
 * file order then record order (the order of `contig_names`, src/cluster_argument_parsing.rs:
 * 596-629).  *n_units receives the number of units.  threshold_pct < 85 fails with the reference's
 * panic text (src/skani.rs:116-121).  Result: (i, j, ani in PERCENT), sorted by (i, j);
 * common/total are the marker intersection integers. */
int galah_b200_skani_distances(const char *const *paths, size_t n, float threshold_pct, float min_af_pct,
                               int small_genomes, int per_record, int host_threads,
                               galah_b200_pair_t **out, size_t *n_out, size_t *n_units);
/* cluster() with SkaniPreclusterer + SkaniClusterer -- galah's CLI default
 * (--precluster-method skani --cluster-method skani) -- or, with cluster_contigs != 0,
 * `--cluster-contigs`: both run with skip_clusterer (src/clusterer.rs:32-44). */
/* The same preclusterer on units that are already packed on the device (K1 layout, one record per
 * unit of `lengths[g]` bases at base_off[g]): K3 index, marker sketches written straight into the
 * K2 table (they never leave the device), containment screen, K3 ANI.  This is the contig-mode
 * path (`--cluster-contigs --small-genomes`, BASELINE.json configs[4]) at scale; bench / tests
 * feed it synthetic contigs.  n_screened: pairs that passed the marker screen; ms5 (optional):
 * host wall clock of index build, marker sketches, screen, ANI, total. */
int galah_b200_skani_distances_packed_device(const uint32_t *d_seq2, const uint32_t *d_valid,
                                             const uint64_t *d_base_off, const uint64_t *base_off,
                                             const uint64_t *lengths, size_t n, float threshold_pct,
                                             float min_af_pct, int small_genomes, int individual_contigs,
                                             void *stream, galah_b200_pair_t **out, size_t *n_out,
                                             uint64_t *n_screened, float *ms5);
/* The same on HOST arrays over several GPUs of this process (galah_b200_init_devices first): unit
 * slices for the marker sketches and the K3 index, marker rows exchanged by peer copies, row-block
 * shards of the containment screen, every screened pair (i < j) evaluated once -- i is the query,
 * as in `skani triangle` (src/skani.rs:109-225) -- on the device that owns unit i, which reads
 * unit j's table in place on its peer.  Hits in (i, j) order, identical to the single-GPU call's;
 * cluster them with galah_b200_cluster_from_distances(skip_clusterer = 1): BASELINE.json configs[4]. */
int galah_b200_skani_distances_packed_multi(const uint32_t *seq2, const uint32_t *valid, const uint64_t *base_off,
                                            const uint64_t *lengths, size_t n, int n_devices, float threshold_pct,
                                            float min_af_pct, int small_genomes, int individual_contigs,
                                            galah_b200_pair_t **out, size_t *n_out, uint64_t *n_screened);
/* The file-based preclusterer and the whole skani + skani call over several GPUs of this process
 * (galah_b200_init_devices first): device r reads, decodes (K0), marker-sketches and indexes the r-th
 * slice of the path list; the devices agree on the unit numbering and the marker row stride once
 * every slice is read (in contig mode a file's record count is not known before), then proceed as
 * galah_b200_skani_distances_packed_multi.  Hits and clusters identical to the single-GPU calls'. */
int galah_b200_skani_distances_multi(const char *const *paths, size_t n, int n_devices, float threshold_pct,
                                     float min_af_pct, int small_genomes, int per_record, int host_threads,
                                     galah_b200_pair_t **out, size_t *n_out, size_t *n_units);
int galah_b200_cluster_files_skani_multi(const char *const *paths, size_t n, int n_devices, float precluster_ani_pct,
                                         float ani_threshold_pct, float min_af_pct, int small_genomes,
                                         int cluster_contigs, int host_threads, galah_b200_clusters_t *out,
                                         galah_b200_cluster_stats_t *stats);
int galah_b200_cluster_files_skani(const char *const *paths, size_t n, float precluster_ani_pct,
                                   float ani_threshold_pct, float min_af_pct, int small_genomes,
                                   int cluster_contigs, int host_threads, galah_b200_clusters_t *out,
                                   galah_b200_cluster_stats_t *stats);

/* ---- the trait-shaped boundary --------------------------------------------------------------
 * What an UNMODIFIED galah::clusterer::cluster() (src/clusterer.rs:14-152) calls, in the order it
 * calls it: PreclusterDistanceFinder::distances* once (src/lib.rs:29-45), then
 * ClusterDistanceFinder::calculate_ani(fasta1, fasta2) from nested rayon workers (src/lib.rs:54,
 * src/clusterer.rs:262-296, 375).  A session is the state the two trait objects of one cluster()
 * call share (the Rust shim holds an Arc of it in both):
 *   - *_distances stashes its paths and hit list;
 *   - the first calculate_ani indexes those genomes (unless set_clusterer let the distances call
 *     build the K3 index in its own ingest pass) and evaluates EVERY stashed hit in one K3 launch;
 *   - later calls are cache lookups under a shared lock: calculate_ani is re-entrant and may be
 *     called for any pair of paths (pairs that were never hits are indexed / computed on demand).
 * Arguments keep the reference's units: FinchPreclusterer.min_ani is a fraction, skani thresholds
 * are percentages, min_aligned_threshold is a FRACTION (multiplied by 100 in f32 as the reference
 * does, src/skani.rs:153, 733).  The reference's panics become GALAH_B200_ERR_UNSUPPORTED with the
 * reference's text in galah_b200_last_error() (src/finch.rs:15, 40; src/skani.rs:116-121, 243-245,
 * 518-520; src/cluster_argument_parsing.rs:622-627). */
typedef struct galah_b200_session galah_b200_session_t;
int galah_b200_session_create(galah_b200_session_t **out);
void galah_b200_session_free(galah_b200_session_t *s);
/* Optional: announce the SkaniClusterer's small_genomes before distances(), so that the
 * preclusterer's ingest pass also builds the K3 index (no second read of the files). */
int galah_b200_session_set_clusterer(galah_b200_session_t *s, int small_genomes);
const char *galah_b200_finch_method_name(void);  /* "finch", src/finch.rs:43-45 */
const char *galah_b200_skani_method_name(void);  /* "skani", src/skani.rs:71-73, 704-706 */
/* FinchPreclusterer (src/finch.rs:4-46): distances (low_memory != 0 -> the reference's panic),
 * distances_contigs (always an empty cache), distances_with_references (always the panic). */
int galah_b200_session_finch_distances(galah_b200_session_t *s, const char *const *paths, size_t n, float min_ani,
                                       uint32_t num_kmers, uint8_t kmer_length, int low_memory, int host_threads,
                                       galah_b200_pair_t **out, size_t *n_out);
int galah_b200_session_finch_distances_contigs(galah_b200_session_t *s, const char *const *paths, size_t n,
                                               const char *const *contig_names, size_t n_names,
                                               galah_b200_pair_t **out, size_t *n_out);
int galah_b200_session_finch_distances_with_references(galah_b200_session_t *s, const char *const *paths, size_t n,
                                                       const char *const *reference_paths, size_t n_refs,
                                                       galah_b200_pair_t **out, size_t *n_out);
/* SkaniPreclusterer (src/skani.rs:12-74): distances (low_memory: the sketch + search form of
 * src/skani.rs:229-377, where the later of a pair's two records wins, i.e. query = the HIGHER index),
 * distances_contigs (indices are positions in contig_names, matched by record name as
 * src/skani.rs:460-474 does), distances_with_references (src/skani.rs:502-687: only pairs of one
 * reference and one non-reference of `combined_paths`; the query is the non-reference). */
int galah_b200_session_skani_distances(galah_b200_session_t *s, const char *const *paths, size_t n, float threshold_pct,
                                       float min_aligned_threshold, int small_genomes, int low_memory,
                                       int host_threads, galah_b200_pair_t **out, size_t *n_out);
int galah_b200_session_skani_distances_contigs(galah_b200_session_t *s, const char *const *paths, size_t n,
                                               const char *const *contig_names, size_t n_names, float threshold_pct,
                                               float min_aligned_threshold, int small_genomes, int host_threads,
                                               galah_b200_pair_t **out, size_t *n_out);
int galah_b200_session_skani_distances_with_references(galah_b200_session_t *s, const char *const *combined_paths,
                                                       size_t n, const char *const *reference_paths, size_t n_refs,
                                                       float threshold_pct, float min_aligned_threshold,
                                                       int small_genomes, int host_threads, galah_b200_pair_t **out,
                                                       size_t *n_out);
/* SkaniClusterer::calculate_ani (src/skani.rs:708-715): fasta1 is the query (-q), fasta2 the
 * reference (-r); *is_some is always 1 (skani never answers None, 0.0 stands for "no row"). */
int galah_b200_session_calculate_ani(galah_b200_session_t *s, const char *fasta1, const char *fasta2,
                                     float min_aligned_threshold, int small_genomes, float *ani, int *is_some);
int galah_b200_session_stats(galah_b200_session_t *s, uint64_t *n_indexed, uint64_t *n_pairs_computed,
                             uint64_t *n_launches);
/* Record names of FASTA / FASTQ files as `galah cluster --cluster-contigs` collects them
 * (src/cluster_argument_parsing.rs:596-629): header line up to the first TAB, file order then
 * record order; duplicate names fail with the reference's panic text. */
int galah_b200_contig_names(const char *const *paths, size_t n, char ***names_out, size_t *n_names);
void galah_b200_contig_names_free(char **names, size_t n);

/* ---- quality-ordering inputs (host logic) ---------------------------------------------------
 * Replaces galah::genome_stats::calculate_genome_stats (src/genome_stats.rs:11-51), the per-genome
 * inputs of the Parks2020 / dRep quality formulas: a by-product of the same ingest pass that packs
 * the sequence.  n50 follows the reference exactly (contig lengths sorted ascending, first length at
 * which the running sum reaches total/2); num_ambiguous_bases counts literal N / n only. */
typedef struct galah_b200_genome_stats {
    uint64_t num_contigs, num_ambiguous_bases, n50;
} galah_b200_genome_stats_t;
int galah_b200_genome_stats(const char *const *paths, size_t n, int host_threads,
                            galah_b200_genome_stats_t *out);

/* Host ingest of one FASTA/FASTQ file (plain or gzip) into the packed layout (what the reference
 * gets from needletail `parse_fastx_file` + normalize(false), used via finch::sketch_files at
 * src/finch.rs:69): 2 bits/base + validity bitmap, records separated by one invalid base,
 * rec_start/rec_end in packed coordinates.  All outputs are malloc'd: release each with
 * galah_b200_free().  Pure host code (parity hook for the ingest stage). */
int galah_b200_pack_fasta_file(const char *path, uint32_t **seq2, uint32_t **valid, uint64_t *n_bases,
                               uint64_t **rec_start, uint64_t **rec_end, size_t *n_records);

/* K0: the same ingest ON THE DEVICE (csrc/ingest.cu).  The file-taking entry points
 * (galah_b200_sketch_files, _ani_index_add_files, _finch_distances, _cluster_files, ...) use it for
 * whole-genome FASTA input: host threads only read / inflate the files, the raw bytes are uploaded
 * once and three kernels classify, count and pack them into the layout K1 / K3 read; in contig
 * mode (one unit per record) a fourth kernel copies every record into its own unit.  FASTQ
 * input takes the host packer.
 * galah_b200_device_ingest(0 / 1) switches the device path off / on (< 0 only queries); returns
 * the previous setting (default 1).
 * galah_b200_decode_fasta_device is the parity / measurement hook: n in-memory FASTA files ->
 * host copies of the packed arrays (file f at bases [base_off[f], base_off[f+1]), multiples of
 * 128) and per-file n_bases, record ranges (rec_off, rec_start, rec_end: file-relative packed
 * coordinates), ambiguous and literal-N counts (src/genome_stats.rs:27-31).  All outputs are
 * malloc'd (galah_b200_free); *device_ms = device time of the decode incl. the upload. */
int galah_b200_device_ingest(int enable);
int galah_b200_decode_fasta_device(const uint8_t *const *files, const size_t *lens, size_t n, uint32_t **seq2,
                                   uint32_t **valid, uint64_t **base_off, uint64_t **n_bases,
                                   uint64_t **rec_off, uint64_t **rec_start, uint64_t **rec_end,
                                   uint64_t **n_ambiguous, uint64_t **n_N, float *device_ms);

/* ---- synthetic genomes (bench / tests) ------------------------------------- */
/* Generates genomes [index_begin, index_begin+n) of `length` bases each directly in packed
 * form on the device.  d_seq2 needs n * words_per_genome uint32 with
 * words_per_genome = round_up(length, 128) / 16; d_valid n * round_up(length,128)/32. */
int galah_b200_synth_packed_device(uint64_t seed, uint64_t index_begin, size_t n, uint64_t length,
                                   uint32_t *d_seq2, uint32_t *d_valid, uint64_t *d_base_off,
                                   void *stream);

/* Families of `family_size` genomes (member m = (index % family_size) % 10 takes the rate table
 * above shifted right by rate_shift): family_size 10, rate_shift 0 is the call above;
 * thousands / 2 gives one dense clade (every pair related, identity >= ~95 %). */
int galah_b200_synth_packed_device_ex(uint64_t seed, uint64_t index_begin, size_t n, uint64_t length,
                                      uint32_t family_size, uint32_t rate_shift, uint32_t *d_seq2,
                                      uint32_t *d_valid, uint64_t *d_base_off, void *stream);
/* The CUDA stream the library enqueues its own work on (for event timing by the caller). */
void *galah_b200_stream(void);

#ifdef __cplusplus
}
#endif
#endif /* GALAH_B200_H */
