// Fuzz of the host FASTA / FASTQ packer (galah_b200/csrc/host/fasta.cpp) under sanitizers
// (tests/test_engine_sanitizers.py):
//   g++ -std=c++17 -O1 -g -fsanitize=address,undefined -I galah_b200/csrc tools/fasta_fuzz.cpp \
//       galah_b200/csrc/host/fasta.cpp -lz -o fasta_fuzz && ./fasta_fuzz
// Random record structures (FASTA and FASTQ, LF / CRLF, empty records and lines, missing final newline, headers
// inside the sequence alphabet, non-ACGT bytes, truncated FASTQ) and plain byte noise go through pack_fasta_bytes; the
// result must be self-consistent: records ordered and inside the packed length, valid bits only on A/C/G/T codes the
// input can explain, counters adding up, padding clean.  Malformed input may be refused, never crash.
#include "host/fasta.hpp"

#include <cstdio>
#include <random>
using namespace gb200;

static bool check(const PackedGenome &g, const char *what, int round) {
    auto bad = [&](const char *why) { printf("round %d (%s): %s\n", round, what, why); return false; };
    if (g.rec_start.size() != g.rec_end.size()) return bad("record arrays differ in length");
    uint64_t prev_end = 0;
    for (size_t r = 0; r < g.rec_start.size(); r++) {
        if (g.rec_start[r] > g.rec_end[r] || g.rec_end[r] > g.n_bases) return bad("record outside the packed length");
        if (r && g.rec_start[r] < prev_end) return bad("records overlap");
        prev_end = g.rec_end[r];
    }
    const uint64_t padded = g.padded_bases();
    if (g.seq2.size() * 16 < padded || g.valid.size() * 32 < padded) return bad("arrays shorter than the padded length");
    uint64_t n_valid = 0, in_records = 0;
    for (size_t r = 0; r < g.rec_start.size(); r++) in_records += g.rec_end[r] - g.rec_start[r];
    for (uint64_t b = 0; b < padded; b++) {
        const bool v = (g.valid[b >> 5] >> (b & 31)) & 1u;
        if (v && b >= g.n_bases) return bad("valid bit in the padding");
        if (!v && ((g.seq2[b >> 4] >> (2 * (b & 15))) & 3u)) return bad("invalid base with a non-zero code");
        n_valid += v;
    }
    if (n_valid + g.n_ambiguous != in_records) return bad("valid + ambiguous != bases in records");
    if (g.n_N > g.n_ambiguous) return bad("more N than ambiguous bases");
    (void)genome_stats(g);
    return true;
}

int main() {
    std::mt19937_64 rng(11);
    const char alphabet[] = "ACGTacgtNnUuRYKM-.*xX@>+ \t";
    for (int round = 0; round < 4000; round++) {
        std::string data;
        const int kind = round % 4;  // 0 FASTA, 1 FASTQ, 2 noise, 3 truncated / mixed
        const char *nl = (rng() & 1) ? "\n" : "\r\n";
        if (kind == 2) {
            const size_t len = rng() % 600;
            for (size_t x = 0; x < len; x++) data.push_back((char)(rng() % 7 == 0 ? '\n' : (rng() % 256)));
            if (rng() & 1) data.insert(data.begin(), '>');
        } else {
            const int n_rec = rng() % 6;
            for (int r = 0; r < n_rec; r++) {
                const bool fq = kind == 1 || (kind == 3 && (rng() & 1));
                data += fq ? "@" : ">";
                for (size_t x = rng() % 30; x > 0; x--) data.push_back((char)('!' + rng() % 90));
                data += nl;
                const size_t len = rng() % 5 == 0 ? 0 : rng() % 400;
                std::string seq;
                for (size_t x = 0; x < len; x++) seq.push_back(rng() % 9 ? "ACGT"[rng() % 4] : alphabet[rng() % (sizeof(alphabet) - 1)]);
                if (fq) {
                    for (char &c : seq) if (c == '\n' || c == '\r') c = 'N';
                    data += seq; data += nl; data += "+"; data += nl;
                    for (size_t x = 0; x < seq.size(); x++) data.push_back((char)('!' + rng() % 60));
                    data += nl;
                } else {
                    const size_t width = 1 + rng() % 80;
                    for (size_t x = 0; x < seq.size(); x += width) { data += seq.substr(x, width); data += nl; if (rng() % 15 == 0) data += nl; }
                }
            }
            if (kind == 3 && !data.empty()) data.resize(rng() % data.size());  // cut anywhere
            else if (!data.empty() && (rng() & 3) == 0) data.pop_back();        // no final newline
        }
        PackedGenome g;
        std::string err;
        const int rc = pack_fasta_bytes(reinterpret_cast<const uint8_t *>(data.data()), data.size(), g, (rng() & 1) != 0, err);
        if (rc == 0 && !check(g, kind == 0 ? "fasta" : kind == 1 ? "fastq" : kind == 2 ? "noise" : "truncated", round)) return 1;
        if (rc != 0 && err.empty()) { printf("round %d: failure without a message\n", round); return 1; }
    }
    printf("fasta fuzz ok\n");
    return 0;
}
