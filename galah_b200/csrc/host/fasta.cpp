#include "fasta.hpp"

#include <stdio.h>
#include <string.h>

#include <algorithm>

// zlib ships as a shared object in this image but without its header; these entry points
// are ABI-stable.  gzread() passes non-gzip files through unchanged.
extern "C" {
typedef struct gzFile_s *gzFile;
gzFile gzopen(const char *path, const char *mode);
int gzread(gzFile file, void *buf, unsigned len);
int gzclose(gzFile file);
const char *gzerror(gzFile file, int *errnum);
int gzeof(gzFile file);
#ifndef Z_OK
#define Z_OK 0
#define Z_STREAM_END 1
#endif
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

// Streams bases into the packed arrays.  The arrays are pre-sized from the input length (every
// base costs at least one input byte, plus one separator per record), so the hot loop carries
// no bounds checks: it accumulates 32 bases in registers and stores two sequence words and one
// validity word at a time.
struct Packer {
    PackedGenome &g;
    uint64_t n = 0;       // bases pushed so far
    uint64_t acc2 = 0;    // 2-bit codes of the current group of 32 bases
    uint32_t accv = 0;    // validity bits of the current group
    explicit Packer(PackedGenome &g_) : g(g_) {}
    inline void flush_group() {  // n is a multiple of 32 here
        const uint64_t grp = (n >> 5) - 1;
        g.seq2[2 * grp] = (uint32_t)acc2;
        g.seq2[2 * grp + 1] = (uint32_t)(acc2 >> 32);
        g.valid[grp] = accv;
        acc2 = 0; accv = 0;
    }
    inline void push(uint32_t code, bool ok) {
        const uint32_t sh = (uint32_t)n & 31u;
        if (ok) { acc2 |= (uint64_t)code << (2 * sh); accv |= 1u << sh; }
        n++;
        if ((n & 31u) == 0) flush_group();
    }
    void feed(const uint8_t *p, size_t len) {
        uint64_t nn = n, a2 = acc2;
        uint32_t av = accv;
        uint64_t amb = 0, nN = 0;
        for (size_t i = 0; i < len; i++) {
            const uint8_t ch = p[i];
            const uint8_t c = kNorm.t[ch];
            if (c == 5) continue;
            const uint32_t sh = (uint32_t)nn & 31u;
            if (c < 4) { a2 |= (uint64_t)c << (2 * sh); av |= 1u << sh; }
            else { amb++; nN += (ch == 'N' || ch == 'n'); }
            nn++;
            if ((nn & 31u) == 0) {
                const uint64_t grp = (nn >> 5) - 1;
                g.seq2[2 * grp] = (uint32_t)a2;
                g.seq2[2 * grp + 1] = (uint32_t)(a2 >> 32);
                g.valid[grp] = av;
                a2 = 0; av = 0;
            }
        }
        n = nn; acc2 = a2; accv = av;
        g.n_ambiguous += amb; g.n_N += nN;
    }
    void finish() {
        g.n_bases = n;
        if (n & 31u) {  // partial last group
            const uint64_t grp = n >> 5;
            g.seq2[2 * grp] = (uint32_t)acc2;
            g.seq2[2 * grp + 1] = (uint32_t)(acc2 >> 32);
            g.valid[grp] = accv;
        }
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
    out.clear();
    // plain files: one fread of the whole file; gzip (magic 1f 8b, as needletail sniffs it): zlib
    FILE *fp = fopen(path.c_str(), "rb");
    if (!fp) { err = "Failed to open fasta file " + path; return 4; }
    unsigned char magic[2] = {0, 0};
    const size_t got_magic = fread(magic, 1, 2, fp);
    if (!(got_magic == 2 && magic[0] == 0x1f && magic[1] == 0x8b)) {
        if (fseek(fp, 0, SEEK_END) == 0) {
            const long sz = ftell(fp);
            if (sz >= 0 && fseek(fp, 0, SEEK_SET) == 0) {
                out.resize((size_t)sz);
                const size_t rd = sz ? fread(out.data(), 1, (size_t)sz, fp) : 0;
                fclose(fp);
                if (rd != (size_t)sz) { err = "Failed to read " + path; return 4; }
                return 0;
            }
        }
        fclose(fp);  // not seekable: fall through to the streaming reader
    } else {
        fclose(fp);
    }
    gzFile f = gzopen(path.c_str(), "rb");
    if (!f) { err = "Failed to open fasta file " + path; return 4; }
    const unsigned CH = 4u << 20;
    size_t n = 0;
    for (;;) {
        if (out.size() < n + CH) out.resize(out.size() * 2 + CH);
        int got = gzread(f, out.data() + n, CH);
        if (got < 0) { gzclose(f); err = "Failed to read (corrupt gzip?) " + path; return 4; }
        if (got == 0) break;
        n += (size_t)got;
    }
    // gzread() == 0 is also what a stream cut off mid-member returns: a truncated .gz must fail as
    // it does in the reference (needletail), not yield a shorter genome
    int zerr = 0;
    const char *zmsg = gzerror(f, &zerr);
    const bool clean = gzeof(f) && (zerr == Z_OK || zerr == Z_STREAM_END);
    const std::string detail = zmsg ? zmsg : "";
    gzclose(f);
    if (!clean) { err = "Failed to read (truncated or corrupt gzip: " + detail + ") " + path; return 4; }
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
