#include "fasta.hpp"

#include <string.h>

#include <algorithm>

// zlib ships as a shared object in this image but without its header; these three entry points
// are ABI-stable.  gzread() passes non-gzip files through unchanged.
extern "C" {
typedef struct gzFile_s *gzFile;
gzFile gzopen(const char *path, const char *mode);
int gzread(gzFile file, void *buf, unsigned len);
int gzclose(gzFile file);
}

namespace gb200 {

namespace {
// 0..3 = base code, 4 = ambiguous base, 5 = dropped (whitespace)
struct NormTable {
    uint8_t t[256];
    NormTable() {
        for (int i = 0; i < 256; i++) t[i] = 4;
        t[(int)'A'] = t[(int)'a'] = 0;
        t[(int)'C'] = t[(int)'c'] = 1;
        t[(int)'G'] = t[(int)'g'] = 2;
        t[(int)'T'] = t[(int)'t'] = t[(int)'U'] = t[(int)'u'] = 3;
        t[(int)' '] = t[(int)'\t'] = t[(int)'\r'] = t[(int)'\n'] = 5;
    }
};
const NormTable kNorm;

struct Packer {
    PackedGenome &g;
    uint64_t n = 0;
    explicit Packer(PackedGenome &g_) : g(g_) {}
    inline void push(uint32_t code, bool ok) {
        const uint64_t w2 = n >> 4, wv = n >> 5;
        if (w2 >= g.seq2.size()) g.seq2.resize(g.seq2.size() * 2 + 64, 0u);
        if (wv >= g.valid.size()) g.valid.resize(g.valid.size() * 2 + 64, 0u);
        if (ok) {
            g.seq2[w2] |= code << (2 * (n & 15));
            g.valid[wv] |= 1u << (n & 31);
        }
        n++;
    }
    void feed(const uint8_t *p, size_t len) {
        for (size_t i = 0; i < len; i++) {
            const uint8_t c = kNorm.t[p[i]];
            if (c == 5) continue;
            if (c == 4) { g.n_ambiguous++; g.n_N += (p[i] == 'N' || p[i] == 'n'); push(0, false); }
            else push(c, true);
        }
    }
    void finish() {
        g.n_bases = n;
        const uint64_t padded = g.padded_bases();
        // + 16 bytes of padding words so kernels may read slightly past the end
        g.seq2.resize(padded / 16 + 4, 0u);
        g.valid.resize(padded / 32 + 4, 0u);
    }
};
}  // namespace

void PackedGenome::clear() {
    seq2.clear(); valid.clear(); n_bases = 0; rec_start.clear(); rec_end.clear(); rec_name.clear();
    n_ambiguous = 0; n_N = 0;
}

GenomeAssemblyStats genome_stats(const PackedGenome &g) {
    GenomeAssemblyStats st{g.rec_start.size(), g.n_N, 0, false};
    std::vector<uint64_t> len(g.rec_start.size());
    uint64_t total = 0;
    for (size_t r = 0; r < len.size(); r++) { len[r] = g.rec_end[r] - g.rec_start[r]; total += len[r]; }
    std::sort(len.begin(), len.end());
    const uint64_t cutoff = total / 2;
    uint64_t sum = 0;
    for (uint64_t l : len) {
        sum += l;
        if (sum >= cutoff) { st.n50 = l; st.n50_valid = true; break; }
    }
    return st;
}

int read_file_bytes(const std::string &path, std::vector<uint8_t> &out, std::string &err) {
    gzFile f = gzopen(path.c_str(), "rb");
    if (!f) { err = "Failed to open fasta file " + path; return 4; }
    out.clear();
    const unsigned CH = 4u << 20;
    size_t n = 0;
    for (;;) {
        if (out.size() < n + CH) out.resize(out.size() * 2 + CH);
        int got = gzread(f, out.data() + n, CH);
        if (got < 0) { gzclose(f); err = "Failed to read (corrupt gzip?) " + path; return 4; }
        if (got == 0) break;
        n += (size_t)got;
    }
    gzclose(f);
    out.resize(n);
    return 0;
}

int pack_fasta_bytes(const uint8_t *d, size_t n, PackedGenome &out, bool keep_names,
                     std::string &err) {
    out.clear();
    out.seq2.assign(n / 16 + 64, 0u);
    out.valid.assign(n / 32 + 64, 0u);
    Packer pk(out);
    size_t p = 0;
    while (p < n && (d[p] == '\n' || d[p] == '\r')) p++;
    if (p >= n) { pk.finish(); return 0; }
    const bool fasta = d[p] == '>', fastq = d[p] == '@';
    if (!fasta && !fastq) { err = "not a FASTA/FASTQ file (first byte is neither '>' nor '@')"; return 4; }
    bool first = true;
    while (p < n) {
        // header line
        size_t h0 = p + 1;
        while (p < n && d[p] != '\n') p++;
        size_t h1 = p;
        if (p < n) p++;
        if (keep_names) {
            size_t e = h0;
            while (e < h1 && d[e] != ' ' && d[e] != '\t' && d[e] != '\r') e++;
            out.rec_name.emplace_back(reinterpret_cast<const char *>(d + h0), e - h0);
        }
        if (!first) pk.push(0, false);  // record separator: k-mers never span records
        first = false;
        out.rec_start.push_back(pk.n);
        if (fasta) {
            const size_t start = p;
            while (p < n) {
                if (d[p] == '>' && (p == start || d[p - 1] == '\n')) break;
                p++;
            }
            pk.feed(d + start, p - start);
        } else {
            const size_t start = p;
            while (p < n && d[p] != '\n') p++;
            pk.feed(d + start, p - start);
            if (p < n) p++;
            while (p < n && d[p] != '\n') p++;  // '+' line
            if (p < n) p++;
            while (p < n && d[p] != '\n') p++;  // quality line
            if (p < n) p++;
            while (p < n && (d[p] == '\n' || d[p] == '\r')) p++;
        }
        out.rec_end.push_back(pk.n);
    }
    pk.finish();
    return 0;
}

int pack_fasta_file(const std::string &path, PackedGenome &out, bool keep_names, std::string &err) {
    std::vector<uint8_t> bytes;
    int rc = read_file_bytes(path, bytes, err);
    if (rc) return rc;
    rc = pack_fasta_bytes(bytes.data(), bytes.size(), out, keep_names, err);
    if (rc) err = path + ": " + err;
    return rc;
}

int pack_fasta_file_per_record(const std::string &path, std::vector<PackedGenome> &out, bool keep_names,
                               std::string &err) {
    PackedGenome whole;
    int rc = pack_fasta_file(path, whole, keep_names, err);
    if (rc) return rc;
    for (size_t r = 0; r < whole.rec_start.size(); r++) {
        out.emplace_back();
        PackedGenome &g = out.back();
        const uint64_t b0 = whole.rec_start[r], len = whole.rec_end[r] - b0;
        g.seq2.assign(len / 16 + 64, 0u);
        g.valid.assign(len / 32 + 64, 0u);
        for (uint64_t x = 0; x < len; x++) {  // bit copy at an arbitrary offset
            const uint64_t s = b0 + x;
            const uint32_t code = (whole.seq2[s >> 4] >> (2 * (s & 15))) & 3u;
            const uint32_t ok = (whole.valid[s >> 5] >> (s & 31)) & 1u;
            g.seq2[x >> 4] |= (ok ? code : 0u) << (2 * (x & 15));
            g.valid[x >> 5] |= ok << (x & 31);
        }
        g.n_bases = len;
        g.rec_start.push_back(0); g.rec_end.push_back(len);
        if (keep_names && r < whole.rec_name.size()) g.rec_name.push_back(whole.rec_name[r]);
        const uint64_t padded = g.padded_bases();
        g.seq2.resize(padded / 16 + 4, 0u);
        g.valid.resize(padded / 32 + 4, 0u);
    }
    return 0;
}

void pack_records(const std::vector<std::string> &records, PackedGenome &out) {
    out.clear();
    size_t total = 0;
    for (auto &r : records) total += r.size() + 1;
    out.seq2.assign(total / 16 + 64, 0u);
    out.valid.assign(total / 32 + 64, 0u);
    Packer pk(out);
    bool first = true;
    for (auto &r : records) {
        if (!first) pk.push(0, false);
        first = false;
        out.rec_start.push_back(pk.n);
        pk.feed(reinterpret_cast<const uint8_t *>(r.data()), r.size());
        out.rec_end.push_back(pk.n);
    }
    pk.finish();
}

}  // namespace gb200
