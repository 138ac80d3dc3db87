"""K0's algorithm, restated in numpy / Python integers exactly as csrc/ingest.cu computes it (8 kB
chunks, 32 bytes per thread as bit masks, header state by a Kogge-Stone latch, positions from
prefix counts, 64-bit run compaction) and checked against the host packer on the CPU.  The GPU
tests (tests/test_ingest_gpu.py) check the kernels themselves; this one pins the bit arithmetic."""
import numpy as np

import galah_b200 as gb
from util import random_dna

CHUNK, PER = 8192, 32
M32 = 0xFFFFFFFF


def norm_code(b):
    if b in (0x20, 0x09, 0x0D, 0x0A):
        return 5
    u = b & 0xDF
    return {0x41: 0, 0x43: 1, 0x47: 2, 0x54: 3, 0x55: 3}.get(u, 4)


LUT = [norm_code(b) for b in range(256)]


def byte_masks(by, lim, prev, fk):
    start = nl = ws = amb = nN = 0
    codes = 0
    for k in range(PER):
        b = by[k]
        c = LUT[b]
        st = b == 0x3E and (prev == 0x0A or k == fk)
        start |= int(st) << k
        nl |= int(b == 0x0A) << k
        ws |= int(c == 5) << k
        amb |= int(c == 4) << k
        nN |= int((b | 0x20) == 0x6E) << k
        codes |= (c & 3) << (2 * k)
        prev = b
    live = M32 if lim >= 32 else (1 << lim) - 1
    return start & live, nl & live, (ws | (~live & M32)) & M32, amb & live, nN & live, codes


def header_mask(start, nl, carry):
    G = (start | (1 if carry else 0)) & M32
    P = ~(nl << 1) & M32
    d = 1
    while d < 32:
        G = (G | (P & (G << d))) & M32
        P = (P & (P << d)) & M32
        d <<= 1
    return G


def emulate(data):
    """-> (codes uint8 per packed base with 4 = invalid, rec_start, rec_end, n_ambiguous, n_N)"""
    n = len(data)
    buf = bytes(data) + b"\n" * 64
    p = 0
    while p < n and data[p] in (0x0A, 0x0D):
        p += 1
    first = p
    # pass 1 + 2: per thread masks, chunk carries, counts
    threads = []  # (chunk, masks, H0>N0)
    H = N = 0
    chunk_carry = {}
    per_thread = []
    for begin in range(0, n, CHUNK):
        end = min(begin + CHUNK, n)
        chunk_carry[begin] = H > N
        h0, n0 = (begin if H > N else 0), 0  # the kernel's carry: position + 1 of begin - 1
        for t in range(CHUNK // PER):
            p0 = begin + t * PER
            if p0 >= end:
                break
            lim = min(PER, end - p0)
            by = buf[p0:p0 + PER]
            prev = buf[p0 - 1] if p0 else 0x0A
            fk = first - p0 if p0 <= first < p0 + PER else 64
            start, nl, ws, amb, nN, codes = byte_masks(by, lim, prev, fk)
            per_thread.append((p0, start, nl, ws, amb, nN, codes, h0 > n0))
            if start:
                h0 = max(h0, p0 + start.bit_length())
            if nl:
                n0 = max(n0, p0 + nl.bit_length())
        # chunk summaries as the scan kernel reports them
        for (q0, start, nl, *_rest) in [x for x in per_thread if begin <= x[0] < end]:
            if start:
                H = max(H, q0 + start.bit_length())
            if nl:
                N = max(N, q0 + nl.bit_length())
    # pass 3: positions and writes
    out = {}
    rec_start, rec_end = [], []
    nbase = nrec = n_amb = n_N = 0
    for (p0, start, nl, ws, amb, nN, codes, carry) in per_thread:
        hdr = header_mask(start, nl, carry)
        base = ~hdr & ~ws & M32
        good = base & ~amb & M32
        n_amb += bin(base & amb).count("1")
        n_N += bin(base & amb & nN).count("1")
        if start == 0:
            L = nbase + nrec - 1  # position of my first base (file-relative)
            run = vrun = cnt = 0
            for k in range(PER):
                if (base >> k) & 1:
                    if (good >> k) & 1:
                        run |= ((codes >> (2 * k)) & 3) << (2 * cnt)
                        vrun |= 1 << cnt
                    cnt += 1
            for i in range(cnt):
                if (vrun >> i) & 1:
                    out[L + i] = (run >> (2 * i)) & 3
            nbase += cnt
        else:
            for k in range(PER):
                if (start >> k) & 1:
                    if nrec > 0:
                        rec_end.append(nbase + nrec - 1)
                    rec_start.append(nbase + nrec)
                    nrec += 1
                if (base >> k) & 1:
                    if (good >> k) & 1:
                        out[nbase + nrec - 1] = (codes >> (2 * k)) & 3
                    nbase += 1
    total = nbase + (nrec - 1 if nrec else 0)
    if nrec:
        rec_end.append(total)
    arr = np.full(total, 4, np.uint8)
    for pos, c in out.items():
        arr[pos] = c
    return arr, np.array(rec_start, np.uint64), np.array(rec_end, np.uint64), n_amb, n_N


def host(tmp_path, data, name):
    p = str(tmp_path / name)
    with open(p, "wb") as f:
        f.write(data)
    return gb.pack_fasta_file(p)


def test_k0_bit_arithmetic_matches_host_packer(tmp_path):
    rng = np.random.default_rng(8)
    cases = [
        b">a\nACGT\n", b">a\nACGT", b"\n\r\n>a desc\r\nACgtNnRYu\r\nUUAA\r\n>b\r\n\r\nGG\r\n", b">only_header\n",
        b">h1\n>h2\nAC\n>h3\n", b">a\nAC>GT\nA>C\n>b\nTT\n", b">a\n" + b"ACGT" * 5000 + b"\n",
        b">" + b"h" * 20000 + b"\nACGTACGT\n>x\nAC\n", b">a\n \t A C\tG T \n",
        b">a\n" + random_dna(8191, rng) + b"\n>b\n" + random_dna(8192, rng) + b"\n>c\n" + random_dna(3, rng),
    ]
    for k in range(6):
        recs = []
        for r in range(int(rng.integers(1, 8))):
            m = int(rng.integers(0, 6000))
            seq = bytearray(random_dna(m, rng))
            for _ in range(int(rng.integers(0, 4))):
                if m:
                    a = int(rng.integers(0, m)); b = min(m, a + int(rng.integers(1, 90)))
                    seq[a:b] = bytes(rng.choice(list(b"NnRYacgtu>"), size=b - a).astype(np.uint8))
            width = int(rng.choice([31, 32, 33, 60, 80]))
            body = b"".join(bytes(seq[i:i + width]) + (b"\r\n" if k % 2 else b"\n") for i in range(0, m, width))
            recs.append(b">r%d some text\n" % r + body)
        cases.append(b"".join(recs))
    for idx, data in enumerate(cases):
        codes, rs, re_, n_amb, n_N = emulate(data)
        hc, hs, he = host(tmp_path, data, f"c{idx}.fna")
        assert np.array_equal(codes, hc), idx
        assert np.array_equal(rs, hs) and np.array_equal(re_, he), idx
        n_sep = max(len(hs) - 1, 0)
        assert n_amb == int((hc == 4).sum()) - n_sep, idx
