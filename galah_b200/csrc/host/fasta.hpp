// FASTA/FASTQ ingest -> packed 2-bit sequence + validity bitmap (the layout K1 reads).
//
// Mirrors what the reference gets from needletail 0.5 (`parse_fastx_file`, used via
// finch::sketch_files at /root/reference/src/finch.rs:69 and directly at
// /root/reference/src/skani.rs:87-100): plain or gzip input, '>' / '@' records, sequence bytes
// normalised with normalize(false) semantics (ACGT kept, acgt upper-cased, u/U -> T, whitespace
// dropped, everything else is an ambiguous base).  Ambiguous bases and record boundaries
// become INVALID bases so that no k-mer spans them.
#pragma once
#include <stdint.h>

#include <string>
#include <vector>

namespace gb200 {

struct PackedGenome {
    std::vector<uint32_t> seq2;   // 16 bases per word, LSB first, A0 C1 G2 T3 (invalid bases: 0)
    std::vector<uint32_t> valid;  // 32 bases per word, 1 = ACGT
    uint64_t n_bases = 0;         // packed length incl. one separator per record boundary
    // per record: [start, end) in packed coordinates (separators excluded)
    std::vector<uint64_t> rec_start, rec_end;
    std::vector<std::string> rec_name;  // header up to first whitespace (needletail `id` prefix)
    uint64_t n_ambiguous = 0;           // bases that were not ACGT
    uint64_t n_N = 0;                   // bases that were literally 'N' or 'n' (genome_stats.rs:27-31)
    void clear();
    // pad seq2/valid so that n_bases rounds up to a multiple of 128 with invalid bases
    uint64_t padded_bases() const { return (n_bases + 127) / 128 * 128; }
};

// Returns 0 on success; on failure `err` describes the problem.
int read_file_bytes(const std::string &path, std::vector<uint8_t> &out, std::string &err);
int pack_fasta_bytes(const uint8_t *data, size_t n, PackedGenome &out, bool keep_names,
                     std::string &err);
int pack_fasta_file(const std::string &path, PackedGenome &out, bool keep_names, std::string &err);
// Contig mode (`--cluster-contigs`, /root/reference/src/cluster_argument_parsing.rs:573-629): every
// record of the file becomes its own single-record PackedGenome, appended to `out` in file order.
int pack_fasta_file_per_record(const std::string &path, std::vector<PackedGenome> &out, bool keep_names,
                               std::string &err);

// galah::genome_stats::calculate_genome_stats (/root/reference/src/genome_stats.rs:11-51) from the
// metadata the ingest pass already has: record count, literal N/n count, and the reference's N50
// (lengths sorted ASCENDING, first length at which the running sum reaches total/2).
struct GenomeAssemblyStats { uint64_t num_contigs, num_ambiguous_bases, n50; bool n50_valid; };
GenomeAssemblyStats genome_stats(const PackedGenome &g);

// Packs raw (un-normalised) records that are already in memory (tests / synthetic inputs).
void pack_records(const std::vector<std::string> &records, PackedGenome &out);

}  // namespace gb200
