// Stage 0 (K0): FASTA bytes -> packed 2-bit sequence + validity bitmap on sm_100a.
//
// Semantics = csrc/host/fasta.cpp (needletail 0.5 parse_fastx + normalize(false) as finch uses
// it, /root/reference/src/finch.rs:55-69): a record starts at a '>' that is the file's first
// non-blank byte or follows '\n'; its header runs to the next '\n'; every other byte of the record
// is sequence: ACGT / acgt / Uu -> 2-bit code, blank (space, tab, CR, LF) dropped, anything else an
// ambiguous (invalid) base; one invalid separator base between records so no k-mer spans them.
//
// Whether a byte lies inside a header line depends on everything before it, so the work is
// three passes over chunks of 8 kB (one CTA each, 32 bytes per thread held in registers):
//   fasta_scan_kernel   per chunk: position of its last record start and of its last '\n'
//                       (host: running maxima over a file's chunks -> "chunk starts inside a header")
//   fasta_pack_kernel<false>  per chunk: bases, record starts, ambiguous, N  (host: prefix sums ->
//                       every chunk's first base / record index, per-file totals, padded offsets)
//   fasta_pack_kernel<true>   writes: a CTA assembles its <= 8192 bases in shared memory (atomicOr
//                       on 2-bit fields) and ORs the words into the zero-filled output; record
//                       starts/ends are written by the thread that sees the '>'.
// Inside a chunk "in header" is H > N with H / N the running maxima (position + 1) of record
// starts / newlines: a block-wide exclusive max-scan over the threads' 32-byte summaries.
// Bound: HBM streaming (3 reads of the raw bytes + 0.375 B/base written); the two host round
// trips carry 8-16 bytes per 8 kB chunk.
#include "ingest.cuh"

#include <algorithm>

#include "common.cuh"

namespace gb200 {

constexpr int kDecThreads = 256;
constexpr int kDecPer = 32;                       // bytes per thread
constexpr int kDecChunk = kDecThreads * kDecPer;  // 8192 bytes per CTA

__device__ __forceinline__ uint32_t norm_code(uint32_t b) {  // 0..3 base, 4 ambiguous, 5 dropped
    if (b == ' ' || b == '\t' || b == '\r' || b == '\n') return 5u;
    const uint32_t u = b & 0xDFu;  // upper-case; no non-letter byte maps onto a letter
    return u == 'A' ? 0u : u == 'C' ? 1u : u == 'G' ? 2u : (u == 'T' || u == 'U') ? 3u : 4u;
}

struct DecParams {
    const uint8_t *bytes;
    const uint64_t *chunk_begin, *chunk_end;  // absolute byte range of every chunk (begin % 32 == 0)
    const uint32_t *chunk_file;
    const uint64_t *file_first;               // per file: absolute offset of its first non-blank byte
    const uint8_t *chunk_in_header;           // chunk starts inside a header line
    uint32_t *summ;                           // scan: [2c] last record start + 1, [2c+1] last '\n' + 1 (0 = none)
    uint32_t *counts;                         // count: [4c..] bases, record starts, ambiguous, N
    // write pass
    const uint32_t *chunk_base0, *chunk_rec0; // bases / record starts of the file before the chunk
    const uint64_t *base_off, *rec_off;       // per file
    uint32_t *seq2, *valid;
    uint64_t *rec_start, *rec_end;
};

// 32 bytes of this thread, as 8 words
struct Bytes32 {
    uint32_t w[8];
    __device__ __forceinline__ uint32_t at(int k) const { return (w[k >> 2] >> (8 * (k & 3))) & 0xFFu; }
};

__device__ __forceinline__ Bytes32 load32(const uint8_t *p) {
    Bytes32 b;
    const uint4 a = *reinterpret_cast<const uint4 *>(p), c = *reinterpret_cast<const uint4 *>(p + 16);
    b.w[0] = a.x; b.w[1] = a.y; b.w[2] = a.z; b.w[3] = a.w; b.w[4] = c.x; b.w[5] = c.y; b.w[6] = c.z; b.w[7] = c.w;
    return b;
}

// Exclusive block scan (256 threads) of two running maxima, seeded with a carry.
__device__ __forceinline__ void block_excl_max2(uint32_t h, uint32_t n, uint32_t carry_h, uint32_t carry_n,
                                                uint32_t &ex_h, uint32_t &ex_n, uint32_t *s_tmp /* >= 16 */) {
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t ih = h, in = n;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t uh = __shfl_up_sync(0xffffffffu, ih, o), un = __shfl_up_sync(0xffffffffu, in, o);
        if (lane >= (uint32_t)o) { ih = max(ih, uh); in = max(in, un); }
    }
    if (lane == 31) { s_tmp[warp] = ih; s_tmp[8 + warp] = in; }
    __syncthreads();
    uint32_t ph = carry_h, pn = carry_n;
    for (uint32_t w = 0; w < warp; w++) { ph = max(ph, s_tmp[w]); pn = max(pn, s_tmp[8 + w]); }
    const uint32_t uh = __shfl_up_sync(0xffffffffu, ih, 1), un = __shfl_up_sync(0xffffffffu, in, 1);
    ex_h = lane ? max(ph, uh) : ph;
    ex_n = lane ? max(pn, un) : pn;
    __syncthreads();
}

// Exclusive block scan (256 threads) of two counters.
__device__ __forceinline__ void block_excl_sum2(uint32_t a, uint32_t b, uint32_t &ex_a, uint32_t &ex_b,
                                                uint32_t *s_tmp /* >= 16 */) {
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t ia = a, ib = b;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t ua = __shfl_up_sync(0xffffffffu, ia, o), ub = __shfl_up_sync(0xffffffffu, ib, o);
        if (lane >= (uint32_t)o) { ia += ua; ib += ub; }
    }
    if (lane == 31) { s_tmp[warp] = ia; s_tmp[8 + warp] = ib; }
    __syncthreads();
    uint32_t pa = 0, pb = 0;
    for (uint32_t w = 0; w < warp; w++) { pa += s_tmp[w]; pb += s_tmp[8 + w]; }
    ex_a = pa + ia - a;
    ex_b = pb + ib - b;
    __syncthreads();
}

// Per-thread view of its 32 bytes as bit masks (bit k <-> byte k): computed once, every later phase
// is bit arithmetic on them instead of another pass over the bytes.
struct ByteMasks {
    uint32_t start;  // record starts: '>' that is the file's first non-blank byte or follows '\n'
    uint32_t nl;     // '\n'
    uint32_t ws;     // dropped bytes: blanks, and everything at or past the chunk's end
    uint32_t amb;    // ambiguous bases (not ACGTU, not blank)
    uint32_t nN;     // literally 'N' / 'n'
    uint64_t codes;  // 2-bit base codes (meaningful where the byte is an unambiguous base)
};

// lut: 256-entry class table in shared memory (0..3 base, 4 ambiguous, 5 dropped); lim: live bytes
// of this thread; prev: the byte before the thread's first one; fk: index (0..31) of the file's
// first non-blank byte if it is one of this thread's bytes, else >= 32.
__device__ __forceinline__ ByteMasks byte_masks(const Bytes32 &by, int lim, uint32_t prev, uint32_t fk,
                                                const uint8_t *lut) {
    ByteMasks m;
    m.start = 0; m.nl = 0; m.ws = 0; m.amb = 0; m.nN = 0; m.codes = 0;
#pragma unroll
    for (int k = 0; k < kDecPer; k++) {
        const uint32_t b = by.at(k);
        const uint32_t c = lut[b];
        const bool st = b == '>' && (prev == '\n' || (uint32_t)k == fk);
        m.start |= st ? (1u << k) : 0u;
        m.nl |= b == '\n' ? (1u << k) : 0u;
        m.ws |= c == 5u ? (1u << k) : 0u;
        m.amb |= c == 4u ? (1u << k) : 0u;
        m.nN |= (b | 0x20u) == 'n' ? (1u << k) : 0u;
        m.codes |= (uint64_t)(c & 3u) << (2 * k);
        prev = b;
    }
    const uint32_t live = lim >= 32 ? 0xFFFFFFFFu : ((1u << lim) - 1u);
    m.start &= live; m.nl &= live; m.amb &= live; m.nN &= live;
    m.ws |= ~live;
    return m;
}

// In-header mask of the thread's bytes: byte k is inside a header line iff it is a record start,
// or byte k-1 was inside one and was not the terminating '\n' (carry: the byte before the thread's
// first one was).  A set-dominant latch = a prefix scan with generate = start, propagate =
// "previous byte is no newline": five Kogge-Stone steps for all 32 bytes.
__device__ __forceinline__ uint32_t header_mask(uint32_t start, uint32_t nl, bool carry) {
    uint32_t G = start | (carry ? 1u : 0u), P = ~(nl << 1);
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        G |= P & (G << d);
        P &= P << d;
    }
    return G;
}

__device__ __forceinline__ void fill_lut(uint8_t *lut) {
    for (uint32_t x = threadIdx.x; x < 256; x += blockDim.x) lut[x] = (uint8_t)norm_code(x);
}

__global__ void __launch_bounds__(kDecThreads) fasta_scan_kernel(const DecParams q) {
    __shared__ uint32_t s_h[8], s_n[8];
    __shared__ uint8_t s_lut[256];
    const uint32_t c = blockIdx.x, tid = threadIdx.x;
    const uint64_t begin = q.chunk_begin[c], end = q.chunk_end[c];
    const uint64_t first = q.file_first[q.chunk_file[c]];
    const uint64_t p0 = begin + (uint64_t)tid * kDecPer;
    fill_lut(s_lut);
    __syncthreads();
    uint32_t last_h = 0, last_n = 0;
    if (p0 < end) {
        const Bytes32 by = load32(q.bytes + p0);
        const uint32_t prev = p0 ? q.bytes[p0 - 1] : (uint32_t)'\n';
        const int lim = (int)min((uint64_t)kDecPer, end - p0);
        const uint32_t fk = first >= p0 && first - p0 < (uint64_t)kDecPer ? (uint32_t)(first - p0) : 64u;
        const ByteMasks m = byte_masks(by, lim, prev, fk, s_lut);
        if (m.start) last_h = (uint32_t)p0 + (31u - (uint32_t)__clz((int)m.start)) + 1u;
        if (m.nl) last_n = (uint32_t)p0 + (31u - (uint32_t)__clz((int)m.nl)) + 1u;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        last_h = max(last_h, __shfl_xor_sync(0xffffffffu, last_h, o));
        last_n = max(last_n, __shfl_xor_sync(0xffffffffu, last_n, o));
    }
    if ((tid & 31) == 0) { s_h[tid >> 5] = last_h; s_n[tid >> 5] = last_n; }
    __syncthreads();
    if (tid == 0) {
        uint32_t h = 0, n = 0;
        for (int w = 0; w < 8; w++) { h = max(h, s_h[w]); n = max(n, s_n[w]); }
        q.summ[2 * c] = h; q.summ[2 * c + 1] = n;
    }
}

template <bool kWrite>
__global__ void __launch_bounds__(kDecThreads) fasta_pack_kernel(const DecParams q) {
    __shared__ uint32_t s_tmp[16];
    __shared__ uint32_t s_cnt[4];
    __shared__ uint8_t s_lut[256];
    __shared__ uint32_t s_seq[kWrite ? kDecChunk / 16 + 4 : 1];
    __shared__ uint32_t s_val[kWrite ? kDecChunk / 32 + 4 : 1];
    const uint32_t c = blockIdx.x, tid = threadIdx.x;
    const uint64_t begin = q.chunk_begin[c], end = q.chunk_end[c];
    const uint32_t f = q.chunk_file[c];
    const uint64_t first = q.file_first[f];
    const uint64_t p0 = begin + (uint64_t)tid * kDecPer;
    const bool live = p0 < end;
    const int lim = live ? (int)min((uint64_t)kDecPer, end - p0) : 0;
    Bytes32 by;
#pragma unroll
    for (int k = 0; k < 8; k++) by.w[k] = 0;
    uint32_t prev0 = '\n';
    if (live) { by = load32(q.bytes + p0); prev0 = p0 ? q.bytes[p0 - 1] : (uint32_t)'\n'; }
    fill_lut(s_lut);
    if (tid < 4) s_cnt[tid] = 0;
    if (kWrite) {
        for (uint32_t x = tid; x < kDecChunk / 16 + 4; x += kDecThreads) s_seq[x] = 0;
        for (uint32_t x = tid; x < kDecChunk / 32 + 4; x += kDecThreads) s_val[x] = 0;
    }
    __syncthreads();
    // ---- phase A: the thread's bytes as bit masks; its last record start / newline
    const uint32_t fk = live && first >= p0 && first - p0 < (uint64_t)kDecPer ? (uint32_t)(first - p0) : 64u;
    const ByteMasks m = byte_masks(by, lim, prev0, fk, s_lut);
    const uint32_t last_h = m.start ? (uint32_t)p0 + (31u - (uint32_t)__clz((int)m.start)) + 1u : 0u;
    const uint32_t last_n = m.nl ? (uint32_t)p0 + (31u - (uint32_t)__clz((int)m.nl)) + 1u : 0u;
    // ---- phase B: running maxima before this thread (carry: the chunk starts inside a header)
    uint32_t H0, N0;
    block_excl_max2(last_h, last_n, q.chunk_in_header[c] ? (uint32_t)begin : 0u, 0u, H0, N0, s_tmp);
    // ---- phase C: classify and count (bit arithmetic)
    const uint32_t hdr = header_mask(m.start, m.nl, H0 > N0);
    const uint32_t base = ~hdr & ~m.ws;          // bytes that are bases (valid or ambiguous)
    const uint32_t good = base & ~m.amb;         // unambiguous bases: the only ones that write bits
    const uint32_t nb = (uint32_t)__popc(base), nrec = (uint32_t)__popc(m.start);
    if (!kWrite) {
        uint32_t cb = nb, cr = nrec, ca = (uint32_t)__popc(base & m.amb), cn = (uint32_t)__popc(base & m.amb & m.nN);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            cb += __shfl_xor_sync(0xffffffffu, cb, o); cr += __shfl_xor_sync(0xffffffffu, cr, o);
            ca += __shfl_xor_sync(0xffffffffu, ca, o); cn += __shfl_xor_sync(0xffffffffu, cn, o);
        }
        if ((tid & 31) == 0) {
            atomicAdd(&s_cnt[0], cb); atomicAdd(&s_cnt[1], cr); atomicAdd(&s_cnt[2], ca); atomicAdd(&s_cnt[3], cn);
        }
        __syncthreads();
        if (tid < 4) q.counts[4 * c + tid] = s_cnt[tid];
        return;
    }
    // ---- phase D: positions, then write
    uint32_t eb, er;
    block_excl_sum2(nb, nrec, eb, er, s_tmp);
    const uint32_t cb0 = q.chunk_base0[c], cr0 = q.chunk_rec0[c];
    const uint64_t off = q.base_off[f], roff = q.rec_off[f];
    // every position this CTA writes is >= G0 and < G0 + kDecChunk + 1; A0 = G0 rounded down to a
    // validity word, local = position - A0 < kDecChunk + 33
    const uint64_t G0 = off + cb0 + (cr0 ? cr0 - 1u : 0u);
    const uint64_t A0 = G0 & ~31ull;
    // position of a base = off + (bases of the file before it) + (record starts up to it) - 1
    const int32_t dl = (int32_t)((int64_t)(off + cb0 + cr0) - 1 - (int64_t)A0);  // in [-1, 31]
    if (m.start == 0) {
        // no record starts among my bytes: my bases land on consecutive positions.  Compact their
        // codes into one 64-bit run and OR <= 3 sequence words and <= 2 validity words.
        if (base) {
            const uint32_t L = (uint32_t)(dl + (int32_t)(eb + er));  // local position of my first base
            uint64_t run = 0;
            uint32_t vrun = 0, cnt = 0;
#pragma unroll
            for (int k = 0; k < kDecPer; k++) {
                if ((base >> k) & 1u) {
                    if ((good >> k) & 1u) {
                        run |= ((m.codes >> (2 * k)) & 3ull) << (2u * cnt);
                        vrun |= 1u << cnt;
                    }
                    cnt++;
                }
            }
            if (vrun) {
                const uint32_t ws = L >> 4, so = (L & 15u) * 2u;
                const uint32_t r_lo = (uint32_t)run, r_hi = (uint32_t)(run >> 32);
                const uint32_t w0 = r_lo << so;
                const uint32_t w1 = so ? (r_lo >> (32u - so)) | (r_hi << so) : r_hi;
                const uint32_t w2 = so ? (r_hi >> (32u - so)) : 0u;
                if (w0) atomicOr(&s_seq[ws], w0);
                if (w1) atomicOr(&s_seq[ws + 1], w1);
                if (w2) atomicOr(&s_seq[ws + 2], w2);
                const uint32_t wv = L >> 5, vo = L & 31u;
                const uint32_t v0 = vrun << vo, v1 = vo ? vrun >> (32u - vo) : 0u;
                if (v0) atomicOr(&s_val[wv], v0);
                if (v1) atomicOr(&s_val[wv + 1], v1);
            }
        }
    } else {
        // a record starts among my bytes (rare): byte by byte, with the record bookkeeping
        uint32_t q32 = eb, r32 = er;  // bases / record starts of this chunk before the current byte
#pragma unroll 1
        for (int k = 0; k < kDecPer; k++) {
            if ((m.start >> k) & 1u) {
                const uint64_t nbase = (uint64_t)cb0 + q32, r = (uint64_t)cr0 + r32;
                if (r > 0) q.rec_end[roff + r - 1] = nbase + (r - 1);
                q.rec_start[roff + r] = nbase + r;
                r32++;
            }
            if ((base >> k) & 1u) {
                if ((good >> k) & 1u) {
                    const uint32_t L = (uint32_t)(dl + (int32_t)(q32 + r32));
                    atomicOr(&s_seq[L >> 4], (uint32_t)((m.codes >> (2 * k)) & 3ull) << (2u * (L & 15u)));
                    atomicOr(&s_val[L >> 5], 1u << (L & 31u));
                }
                q32++;
            }
        }
    }
    __syncthreads();
    const uint64_t ws0 = A0 >> 4, wv0 = A0 >> 5;
    for (uint32_t x = tid; x < kDecChunk / 16 + 4; x += kDecThreads)
        if (s_seq[x]) atomicOr(&q.seq2[ws0 + x], s_seq[x]);
    for (uint32_t x = tid; x < kDecChunk / 32 + 4; x += kDecThreads)
        if (s_val[x]) atomicOr(&q.valid[wv0 + x], s_val[x]);
}

// Contig mode: copy every record out of the packed files into its own 128-padded unit.  One thread
// per destination validity word (32 bases = two sequence words): the unit is found by binary
// search on the destination offsets, the source bits arrive by funnel shifts at an arbitrary
// bit offset, everything past the unit's length is zero.
__global__ void __launch_bounds__(256) fasta_split_kernel(const uint32_t *__restrict__ src_seq2,
                                                          const uint32_t *__restrict__ src_valid,
                                                          const uint64_t *__restrict__ unit_src,
                                                          const uint64_t *__restrict__ unit_off,
                                                          const uint64_t *__restrict__ unit_len, uint32_t n_units,
                                                          uint64_t n_words, uint32_t *__restrict__ dst_seq2,
                                                          uint32_t *__restrict__ dst_valid) {
    for (uint64_t gw = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; gw < n_words;
         gw += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t d = gw << 5;  // first destination base of this word
        uint32_t lo = 0, hi = n_units;  // last unit with unit_off[u] <= d
        while (hi - lo > 1) {
            const uint32_t mid = (lo + hi) >> 1;
            if (unit_off[mid] <= d) lo = mid; else hi = mid;
        }
        const uint64_t b = d - unit_off[lo], len = unit_len[lo];
        uint32_t v = 0, s0 = 0, s1 = 0;
        if (b < len) {
            const uint32_t rem = (uint32_t)min((uint64_t)32, len - b);  // bases of the unit in this word
            const uint64_t s = unit_src[lo] + b;
            const uint32_t vmask = rem >= 32 ? 0xFFFFFFFFu : (1u << rem) - 1u;
            v = __funnelshift_r(src_valid[s >> 5], src_valid[(s >> 5) + 1], (uint32_t)s & 31u) & vmask;
            const uint32_t r0 = min(rem, 16u), r1 = rem > 16 ? rem - 16u : 0u;
            const uint32_t m0 = r0 >= 16 ? 0xFFFFFFFFu : (1u << (2u * r0)) - 1u;
            const uint32_t m1 = r1 >= 16 ? 0xFFFFFFFFu : (1u << (2u * r1)) - 1u;
            s0 = __funnelshift_r(src_seq2[s >> 4], src_seq2[(s >> 4) + 1], 2u * ((uint32_t)s & 15u)) & m0;
            const uint64_t t = s + 16;
            s1 = r1 ? __funnelshift_r(src_seq2[t >> 4], src_seq2[(t >> 4) + 1], 2u * ((uint32_t)t & 15u)) & m1 : 0u;
        }
        dst_valid[gw] = v;
        dst_seq2[2 * gw] = s0;
        dst_seq2[2 * gw + 1] = s1;
    }
}

// ------------------------------------------------------------------------------------------
// host
// ------------------------------------------------------------------------------------------
template <typename T>
int FastaDecoder::Buf<T>::ensure(size_t n) {
    n = std::max<size_t>(n, 1);
    if (n <= cap) return 0;
    if (p) GB_CUDA(cudaFree(p));
    p = nullptr; cap = 0;
    const size_t want = n + n / 8;
    GB_CUDA(cudaMalloc(&p, want * sizeof(T)));
    cap = want;
    return 0;
}
template <typename T>
void FastaDecoder::Buf<T>::release() { if (p) cudaFree(p); p = nullptr; cap = 0; }

void FastaDecoder::release() {
    d_bytes_.release(); d_inhdr_.release(); d_seq2_.release(); d_valid_.release(); d_chunk_file_.release();
    d_summ_.release(); d_counts_.release(); d_chunk_base_.release(); d_chunk_rec_.release();
    d_chunk_begin_.release(); d_chunk_end_.release(); d_file_first_.release(); d_base_off_.release();
    d_rec_off_.release(); d_rec_start_.release(); d_rec_end_.release();
    d_useq2_.release(); d_uvalid_.release(); d_unit_src_.release(); d_unit_off_.release(); d_unit_len_.release();
    for (int x = 0; x < 2; x++) if (ev_[x]) { cudaEventDestroy(ev_[x]); ev_[x] = nullptr; }
}
FastaDecoder::~FastaDecoder() { release(); }

int FastaDecoder::decode(const uint8_t *h_bytes, const std::vector<uint64_t> &file_off,
                         const std::vector<uint64_t> &file_len, const std::vector<uint64_t> &first_byte,
                         DecodedFiles &out, cudaStream_t st) {
    const size_t nf = file_off.empty() ? 0 : file_off.size() - 1;
    out = DecodedFiles();
    out.n_bases.assign(nf, 0); out.n_ambiguous.assign(nf, 0); out.n_N.assign(nf, 0);
    out.rec_off.assign(nf + 1, 0); out.base_off.assign(nf + 1, 0);
    if (nf == 0) return 0;
    if (first_byte.size() != nf || file_len.size() != nf) { set_error("fasta decode: first_byte size mismatch"); return 3; }
    const uint64_t total = file_off[nf];
    if (total >= 0xFFFF0000ull) { set_error("fasta decode: batch larger than 4 GiB"); return 3; }
    for (size_t f = 0; f <= nf; f++)
        if (file_off[f] % 32) { set_error("fasta decode: file offsets must be multiples of 32"); return 3; }
    if (!ev_[0]) { GB_CUDA(cudaEventCreate(&ev_[0])); GB_CUDA(cudaEventCreate(&ev_[1])); }

    // chunk table: every file cut into 8 kB chunks (a chunk never spans files)
    std::vector<uint64_t> cbegin, cend;
    std::vector<uint32_t> cfile;
    std::vector<size_t> first_chunk(nf + 1, 0);
    for (size_t f = 0; f < nf; f++) {
        first_chunk[f] = cbegin.size();
        const uint64_t file_end = file_off[f] + file_len[f];
        if (file_end > file_off[f + 1]) { set_error("fasta decode: file longer than its slot"); return 3; }
        for (uint64_t b = file_off[f]; b < file_end; b += kDecChunk) {
            cbegin.push_back(b); cend.push_back(std::min<uint64_t>(b + kDecChunk, file_end)); cfile.push_back((uint32_t)f);
        }
    }
    first_chunk[nf] = cbegin.size();
    const size_t nc = cbegin.size();

    if (d_bytes_.ensure(total + 64) || d_chunk_begin_.ensure(nc) || d_chunk_end_.ensure(nc) || d_chunk_file_.ensure(nc) ||
        d_file_first_.ensure(nf) || d_summ_.ensure(2 * nc) || d_counts_.ensure(4 * nc) || d_inhdr_.ensure(nc) ||
        d_chunk_base_.ensure(nc) || d_chunk_rec_.ensure(nc) || d_base_off_.ensure(nf + 1) || d_rec_off_.ensure(nf + 1))
        return 2;
    GB_CUDA(cudaEventRecord(ev_[0], st));
    GB_CUDA(cudaMemcpyAsync(d_bytes_.p, h_bytes, total, cudaMemcpyHostToDevice, st));
    GB_CUDA(cudaMemsetAsync(d_bytes_.p + total, '\n', 64, st));
    if (nc == 0) {
        GB_CUDA(cudaMemsetAsync(d_base_off_.p, 0, (nf + 1) * 8, st));
        GB_CUDA(cudaStreamSynchronize(st));
        if (d_seq2_.ensure(16) || d_valid_.ensure(16)) return 2;
        out.d_seq2 = d_seq2_.p; out.d_valid = d_valid_.p; out.d_base_off = d_base_off_.p;
        return 0;
    }
    GB_CUDA(cudaMemcpyAsync(d_chunk_begin_.p, cbegin.data(), nc * 8, cudaMemcpyHostToDevice, st));
    GB_CUDA(cudaMemcpyAsync(d_chunk_end_.p, cend.data(), nc * 8, cudaMemcpyHostToDevice, st));
    GB_CUDA(cudaMemcpyAsync(d_chunk_file_.p, cfile.data(), nc * 4, cudaMemcpyHostToDevice, st));
    GB_CUDA(cudaMemcpyAsync(d_file_first_.p, first_byte.data(), nf * 8, cudaMemcpyHostToDevice, st));
    DecParams q{};
    q.bytes = d_bytes_.p; q.chunk_begin = d_chunk_begin_.p; q.chunk_end = d_chunk_end_.p; q.chunk_file = d_chunk_file_.p;
    q.file_first = d_file_first_.p; q.chunk_in_header = d_inhdr_.p; q.summ = d_summ_.p; q.counts = d_counts_.p;

    // ---- pass 1: last record start / newline per chunk -> does a chunk start inside a header line?
    fasta_scan_kernel<<<(uint32_t)nc, kDecThreads, 0, st>>>(q);
    GB_LAUNCH_CHECK();
    std::vector<uint32_t> summ(2 * nc);
    GB_CUDA(cudaMemcpyAsync(summ.data(), d_summ_.p, 2 * nc * 4, cudaMemcpyDeviceToHost, st));
    GB_CUDA(cudaStreamSynchronize(st));
    std::vector<uint8_t> inhdr(nc, 0);
    for (size_t f = 0; f < nf; f++) {
        uint32_t H = 0, N = 0;
        for (size_t c = first_chunk[f]; c < first_chunk[f + 1]; c++) {
            inhdr[c] = H > N;
            H = std::max(H, summ[2 * c]); N = std::max(N, summ[2 * c + 1]);
        }
    }
    GB_CUDA(cudaMemcpyAsync(d_inhdr_.p, inhdr.data(), nc, cudaMemcpyHostToDevice, st));

    // ---- pass 2: counts per chunk -> positions, totals, padded offsets
    fasta_pack_kernel<false><<<(uint32_t)nc, kDecThreads, 0, st>>>(q);
    GB_LAUNCH_CHECK();
    std::vector<uint32_t> counts(4 * nc);
    GB_CUDA(cudaMemcpyAsync(counts.data(), d_counts_.p, 4 * nc * 4, cudaMemcpyDeviceToHost, st));
    GB_CUDA(cudaStreamSynchronize(st));
    std::vector<uint32_t> cbase(nc), crec(nc);
    for (size_t f = 0; f < nf; f++) {
        uint64_t nb = 0, nr = 0, na = 0, nn = 0;
        for (size_t c = first_chunk[f]; c < first_chunk[f + 1]; c++) {
            cbase[c] = (uint32_t)nb; crec[c] = (uint32_t)nr;
            nb += counts[4 * c]; nr += counts[4 * c + 1]; na += counts[4 * c + 2]; nn += counts[4 * c + 3];
        }
        if (nb + nr >= 0xFFFFFFFFull) { set_error("fasta decode: a file holds 2^32 bases or more"); return 3; }
        out.n_bases[f] = nb + (nr ? nr - 1 : 0);
        out.n_ambiguous[f] = na; out.n_N[f] = nn;
        out.rec_off[f + 1] = out.rec_off[f] + nr;
        out.base_off[f + 1] = out.base_off[f] + (out.n_bases[f] + 127) / 128 * 128;
    }
    const uint64_t total_bases = out.base_off[nf], total_recs = out.rec_off[nf];
    if (d_seq2_.ensure(total_bases / 16 + 16) || d_valid_.ensure(total_bases / 32 + 16) ||
        d_rec_start_.ensure(total_recs) || d_rec_end_.ensure(total_recs))
        return 2;
    GB_CUDA(cudaMemsetAsync(d_seq2_.p, 0, (total_bases / 16 + 16) * 4, st));
    GB_CUDA(cudaMemsetAsync(d_valid_.p, 0, (total_bases / 32 + 16) * 4, st));
    GB_CUDA(cudaMemcpyAsync(d_chunk_base_.p, cbase.data(), nc * 4, cudaMemcpyHostToDevice, st));
    GB_CUDA(cudaMemcpyAsync(d_chunk_rec_.p, crec.data(), nc * 4, cudaMemcpyHostToDevice, st));
    GB_CUDA(cudaMemcpyAsync(d_base_off_.p, out.base_off.data(), (nf + 1) * 8, cudaMemcpyHostToDevice, st));
    GB_CUDA(cudaMemcpyAsync(d_rec_off_.p, out.rec_off.data(), (nf + 1) * 8, cudaMemcpyHostToDevice, st));
    q.chunk_base0 = d_chunk_base_.p; q.chunk_rec0 = d_chunk_rec_.p; q.base_off = d_base_off_.p; q.rec_off = d_rec_off_.p;
    q.seq2 = d_seq2_.p; q.valid = d_valid_.p; q.rec_start = d_rec_start_.p; q.rec_end = d_rec_end_.p;

    // ---- pass 3: write
    fasta_pack_kernel<true><<<(uint32_t)nc, kDecThreads, 0, st>>>(q);
    GB_LAUNCH_CHECK();
    out.rec_start.assign(total_recs, 0); out.rec_end.assign(total_recs, 0);
    if (total_recs) {
        GB_CUDA(cudaMemcpyAsync(out.rec_start.data(), d_rec_start_.p, total_recs * 8, cudaMemcpyDeviceToHost, st));
        GB_CUDA(cudaMemcpyAsync(out.rec_end.data(), d_rec_end_.p, total_recs * 8, cudaMemcpyDeviceToHost, st));
    }
    GB_CUDA(cudaEventRecord(ev_[1], st));
    GB_CUDA(cudaStreamSynchronize(st));
    GB_CUDA(cudaEventElapsedTime(&last_ms, ev_[0], ev_[1]));
    for (size_t f = 0; f < nf; f++)  // the last record of a file ends with the file
        if (out.rec_off[f + 1] > out.rec_off[f]) out.rec_end[out.rec_off[f + 1] - 1] = out.n_bases[f];
    out.d_seq2 = d_seq2_.p; out.d_valid = d_valid_.p; out.d_base_off = d_base_off_.p;
    return 0;
}

int FastaDecoder::split_records(const DecodedFiles &files, DecodedFiles &units, cudaStream_t st) {
    const size_t nf = files.n_bases.size();
    const size_t nu = files.rec_start.size();
    units = DecodedFiles();
    units.n_bases.assign(nu, 0); units.n_ambiguous.assign(nu, 0); units.n_N.assign(nu, 0);
    units.rec_off.assign(nu + 1, 0); units.rec_start.assign(nu, 0); units.rec_end.assign(nu, 0);
    units.base_off.assign(nu + 1, 0);
    std::vector<uint64_t> src(std::max<size_t>(nu, 1)), len(std::max<size_t>(nu, 1));
    size_t u = 0;
    for (size_t f = 0; f < nf; f++)
        for (uint64_t r = files.rec_off[f]; r < files.rec_off[f + 1]; r++, u++) {
            len[u] = files.rec_end[r] - files.rec_start[r];
            src[u] = files.base_off[f] + files.rec_start[r];
            units.n_bases[u] = len[u];
            units.rec_off[u + 1] = u + 1;
            units.rec_end[u] = len[u];
            units.base_off[u + 1] = units.base_off[u] + (len[u] + 127) / 128 * 128;
        }
    const uint64_t total = units.base_off[nu];
    if (d_useq2_.ensure(total / 16 + 16) || d_uvalid_.ensure(total / 32 + 16) || d_unit_src_.ensure(nu) ||
        d_unit_off_.ensure(nu + 1) || d_unit_len_.ensure(nu))
        return 2;
    GB_CUDA(cudaMemcpyAsync(d_unit_off_.p, units.base_off.data(), (nu + 1) * 8, cudaMemcpyHostToDevice, st));
    if (nu && total) {
        GB_CUDA(cudaMemcpyAsync(d_unit_src_.p, src.data(), nu * 8, cudaMemcpyHostToDevice, st));
        GB_CUDA(cudaMemcpyAsync(d_unit_len_.p, len.data(), nu * 8, cudaMemcpyHostToDevice, st));
        const uint64_t n_words = total / 32;
        const uint32_t grid = (uint32_t)std::min<uint64_t>((n_words + 255) / 256, 148ull * 32);
        fasta_split_kernel<<<grid, 256, 0, st>>>(files.d_seq2, files.d_valid, d_unit_src_.p, d_unit_off_.p, d_unit_len_.p,
                                                 (uint32_t)nu, n_words, d_useq2_.p, d_uvalid_.p);
        GB_LAUNCH_CHECK();
    }
    GB_CUDA(cudaMemsetAsync(d_useq2_.p + total / 16, 0, 16 * 4, st));
    GB_CUDA(cudaMemsetAsync(d_uvalid_.p + total / 32, 0, 16 * 4, st));
    GB_CUDA(cudaStreamSynchronize(st));  // the host staging vectors die here
    units.d_seq2 = d_useq2_.p; units.d_valid = d_uvalid_.p; units.d_base_off = d_unit_off_.p;
    return 0;
}

// ------------------------------------------------------------------------------------------
// Validity bitmap of packed genomes from a SPARSE description (galah_b200_cluster_packed_sparse):
// every base of a genome's `length` is valid except the listed ranges (N runs, record breaks); the
// padding behind a genome is invalid.  The bitmap is a third of the packed bytes and almost
// constant, so callers that hold packed genomes on the host need not send it over PCIe.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) validity_fill_kernel(uint32_t *__restrict__ valid, const uint64_t *__restrict__ off,
                                                            const uint64_t *__restrict__ lengths, uint32_t n) {
    const uint32_t g = blockIdx.x;
    if (g >= n) return;
    const uint64_t w0 = off[g] >> 5, w1 = off[g + 1] >> 5, len = lengths[g];
    for (uint64_t w = w0 + threadIdx.x; w < w1; w += 256) {
        const uint64_t bit0 = (w - w0) << 5;
        valid[w] = bit0 + 32 <= len ? 0xFFFFFFFFu : bit0 >= len ? 0u : (1u << (uint32_t)(len - bit0)) - 1u;
    }
}
// ranges: (begin, end) pairs in bases relative to the bitmap's first bit, half open; one warp per range
__global__ void __launch_bounds__(256) validity_clear_kernel(uint32_t *__restrict__ valid, const uint64_t *__restrict__ ranges,
                                                             uint32_t n_ranges) {
    const uint32_t r = (blockIdx.x * 256 + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (r >= n_ranges) return;
    const uint64_t b = ranges[2 * r], e = ranges[2 * r + 1];
    if (e <= b) return;
    const uint64_t wb = b >> 5, we = (e - 1) >> 5;
    for (uint64_t w = wb + lane; w <= we; w += 32) {
        uint32_t m = 0xFFFFFFFFu;
        if (w == wb) m &= 0xFFFFFFFFu << (uint32_t)(b & 31);
        if (w == we) m &= 0xFFFFFFFFu >> (31u - (uint32_t)((e - 1) & 31));
        atomicAnd(&valid[w], ~m);  // ranges of one genome may share a word
    }
}
int validity_from_ranges_enqueue(uint32_t *d_valid, const uint64_t *d_off, const uint64_t *d_lengths, size_t n,
                                 const uint64_t *d_ranges, size_t n_ranges, cudaStream_t st) {
    if (n == 0) return 0;
    validity_fill_kernel<<<(uint32_t)n, 256, 0, st>>>(d_valid, d_off, d_lengths, (uint32_t)n);
    GB_LAUNCH_CHECK();
    if (n_ranges) {
        validity_clear_kernel<<<(uint32_t)((n_ranges * 32 + 255) / 256), 256, 0, st>>>(d_valid, d_ranges, (uint32_t)n_ranges);
        GB_LAUNCH_CHECK();
    }
    return 0;
}

}  // namespace gb200
