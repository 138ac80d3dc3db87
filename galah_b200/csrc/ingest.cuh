// Stage 0 (K0): FASTA bytes -> packed 2-bit sequence + validity bitmap ON THE DEVICE.
//
// The reference reads FASTA through needletail inside finch::sketch_files
// (/root/reference/src/finch.rs:55-69) and again per pair inside skani
// (/root/reference/src/skani.rs:80-107); SURVEY.md 8f.2 names this ingest as the next row after
// the hot path: at 20-200 GB of input it is the end-to-end bound.  Host threads only read (and
// inflate) the files; the raw bytes cross PCIe once and three kernels classify, count and pack
// them into exactly the layout K1 / K3 read (csrc/host/fasta.cpp is the host packer with the
// same semantics; tests compare the two bit for bit).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <vector>

namespace gb200 {

struct DecodedFiles {
    // host metadata, one entry per file
    std::vector<uint64_t> n_bases;      // packed length incl. one separator per record boundary
    std::vector<uint64_t> n_ambiguous;  // bases that were not ACGT(U)
    std::vector<uint64_t> n_N;          // bases that were literally 'N' / 'n'
    std::vector<uint64_t> rec_off;      // n_files + 1 offsets into rec_start / rec_end
    std::vector<uint64_t> rec_start, rec_end;  // per record, packed coordinates relative to its file
    std::vector<uint64_t> base_off;     // n_files + 1, multiples of 128: file f covers [base_off[f], base_off[f+1])
    // device arrays (owned by the decoder, valid until its next decode): the layout K1 / K3 read
    uint32_t *d_seq2 = nullptr, *d_valid = nullptr;
    uint64_t *d_base_off = nullptr;
};

class FastaDecoder {
public:
    ~FastaDecoder();
    // h_bytes: the files' raw (already inflated) bytes, file f at [file_off[f], file_off[f] + file_len[f])
    // with every file_off a multiple of 32 (file_off[n] = size of the buffer); first_byte[f]: absolute
    // offset of the file's first byte that is neither '\n' nor '\r' (must be '>'), or the file's end
    // if there is none.  All files FASTA.
    int decode(const uint8_t *h_bytes, const std::vector<uint64_t> &file_off, const std::vector<uint64_t> &file_len,
               const std::vector<uint64_t> &first_byte, DecodedFiles &out, cudaStream_t st);
    // Contig mode (`--cluster-contigs`, src/cluster_argument_parsing.rs:573-629): every record of the
    // decoded files becomes its own unit -- bases only, no separators, each unit padded to a multiple
    // of 128 -- by one bit-shifting copy kernel on the device.  `units` gets one entry per record
    // (rec_start = 0, rec_end = length) and device arrays owned by the decoder.
    int split_records(const DecodedFiles &files, DecodedFiles &units, cudaStream_t st);
    float last_ms = 0.f;  // device time of the last decode (three kernels + the scans' round trips)
    void release();

private:
    template <typename T>
    struct Buf {
        T *p = nullptr;
        size_t cap = 0;
        int ensure(size_t n);
        void release();
    };
    Buf<uint8_t> d_bytes_, d_inhdr_;
    Buf<uint32_t> d_seq2_, d_valid_, d_chunk_file_, d_summ_, d_counts_, d_chunk_base_, d_chunk_rec_;
    Buf<uint64_t> d_chunk_begin_, d_chunk_end_, d_file_first_, d_base_off_, d_rec_off_, d_rec_start_, d_rec_end_;
    Buf<uint32_t> d_useq2_, d_uvalid_;
    Buf<uint64_t> d_unit_src_, d_unit_off_, d_unit_len_;
    cudaEvent_t ev_[2] = {nullptr, nullptr};
};

// Validity bitmap (1 bit per base, layout of K1) of n packed genomes from their lengths and a list of
// invalid ranges: d_off = n + 1 base offsets relative to the bitmap's first bit (multiples of 128),
// d_ranges = n_ranges (begin, end) pairs relative to the same origin.  Two small kernels on `st`.
int validity_from_ranges_enqueue(uint32_t *d_valid, const uint64_t *d_off, const uint64_t *d_lengths, size_t n,
                                 const uint64_t *d_ranges, size_t n_ranges, cudaStream_t st);

}  // namespace gb200
