#!/usr/bin/env python
"""The in-process multi-device entry (galah_b200_cluster_packed_multi, one host thread per GPU inside the
library) against the single-GPU call on the same host buffers: identical clusters, plus timings.

    python tools/check_multi_device.py [--genomes 2560] [--len 1000000] [--devices G] [--json out.json]
"""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--genomes", type=int, default=2560)
    ap.add_argument("--len", type=int, default=1_000_000)
    ap.add_argument("--devices", type=int, default=0, help="0 = all visible")
    ap.add_argument("--passes", type=int, default=2)
    ap.add_argument("--skip-single", action="store_true")
    ap.add_argument("--json", default=None)
    ap.add_argument("--contigs", type=int, default=20480, help="units of the skani-preclusterer check (0 = skip)")
    ap.add_argument("--contig-len", type=int, default=50_000)
    args = ap.parse_args()
    import torch
    import galah_b200 as gb
    G = args.devices or torch.cuda.device_count()
    n, L = args.genomes, args.len
    gb.init(0)
    dev = torch.device("cuda", 0)
    lay = gb.synth_layout(n, L)
    # synthesise in slabs on device 0, keep the packed genomes in pinned host memory
    h_seq = torch.empty(lay["seq2_words"], dtype=torch.int32, pin_memory=True)
    h_val = torch.empty(lay["valid_words"], dtype=torch.int32, pin_memory=True)
    slab = max(1, min(n, (8 << 30) // max(1, lay["padded"] // 4)))
    for g0 in range(0, n, slab):
        m = min(slab, n - g0)
        l2 = gb.synth_layout(m, L)
        d_seq = torch.empty(l2["seq2_words"], dtype=torch.int32, device=dev)
        d_val = torch.empty(l2["valid_words"], dtype=torch.int32, device=dev)
        d_off = torch.empty(m + 1, dtype=torch.int64, device=dev)
        gb.synth_packed_device(1, g0, m, L, d_seq.data_ptr(), d_val.data_ptr(), d_off.data_ptr(), torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        w0, v0 = g0 * lay["padded"] // 16, g0 * lay["padded"] // 32
        h_seq[w0: w0 + m * lay["padded"] // 16].copy_(d_seq[: m * lay["padded"] // 16])
        h_val[v0: v0 + m * lay["padded"] // 32].copy_(d_val[: m * lay["padded"] // 32])
        del d_seq, d_val, d_off
    torch.cuda.synchronize()
    torch.cuda.empty_cache()
    base_off = np.arange(n + 1, dtype=np.uint64) * np.uint64(lay["padded"])
    lengths = np.full(n, L, np.uint64)
    pairs = n * (n - 1) // 2
    out = {"genomes": n, "genome_bp": L, "pairs": pairs, "devices": G}
    single = None
    if not args.skip_single:
        for _ in range(args.passes):
            t0 = time.perf_counter()
            single, info1 = gb.cluster_packed(h_seq.data_ptr(), h_val.data_ptr(), base_off, lengths, device=False)
            out["single_gpu_s"] = time.perf_counter() - t0
        out["single_gpu_clusters"] = len(single)
        out["single_gpu_sha1"] = single.sha1()
    gb.init_devices(G)
    for _ in range(args.passes):
        t0 = time.perf_counter()
        multi, info = gb.cluster_packed_multi(h_seq.data_ptr(), h_val.data_ptr(), base_off, lengths, G)
        out["multi_s"] = time.perf_counter() - t0
    out["multi_clusters"] = len(multi)
    out["multi_sha1"] = multi.sha1()
    out["multi_pairs_per_s"] = pairs / out["multi_s"]
    out["multi_phases_ms"] = {k: round(float(v), 2) for k, v in info.items() if k.endswith("_ms")}
    out["counts"] = {k: int(v) for k, v in info.items() if not k.endswith("_ms")}
    ok = single is None or (single == multi)
    out["identical_to_single_gpu"] = None if single is None else bool(ok)
    # the same through FASTA files: 300 small genomes (families of 3) written here, clustered on one and on G devices
    import tempfile
    rng = np.random.default_rng(7)
    with tempfile.TemporaryDirectory() as tmp:
        paths = []
        for fam in range(100):
            founder = rng.integers(0, 4, 120_000, dtype=np.uint8)
            for member in range(3):
                seq = founder.copy()
                if member:
                    pos = rng.choice(len(seq), size=len(seq) // (60 * member), replace=False)
                    seq[pos] = (seq[pos] + rng.integers(1, 4, len(pos), dtype=np.uint8)) % 4
                text = np.frombuffer(b"ACGT", dtype=np.uint8)[seq]
                path = os.path.join(tmp, f"g{fam:03d}_{member}.fna")
                with open(path, "wb") as f:
                    f.write(b">c1\n")
                    for x in range(0, len(text), 80):
                        f.write(text[x:x + 80].tobytes() + b"\n")
                paths.append(path)
        gb.init(0)
        one, _ = gb.cluster(paths)
        gb.init_devices(G)
        many, finfo = gb.cluster_multi(paths, G)
        out["files"] = {"genomes": len(paths), "clusters": len(one), "identical_to_single_gpu": bool(one == many),
                        "multi_phases_ms": {k: round(float(v), 2) for k, v in finfo.items() if k.endswith("_ms")}}
        ok = ok and (one == many)
        # the file-based skani preclusterer (galah_b200_skani_distances_multi / _cluster_files_skani_multi): the same
        # genomes, then contig mode on multi-record files with 1..6 records each (a slice's unit count is only known
        # once its files are read), then the reference's own five-genome fixture (src/clusterer.rs:725-757)
        cfiles = []
        for fam in range(40):
            founder = rng.integers(0, 4, 30_000, dtype=np.uint8)
            path = os.path.join(tmp, f"contigs{fam:02d}.fna")
            with open(path, "wb") as f:
                for rec in range(1 + fam % 6):
                    seq = founder.copy()
                    if rec:
                        pos = rng.choice(len(seq), size=len(seq) // (40 * rec), replace=False)
                        seq[pos] = (seq[pos] + rng.integers(1, 4, len(pos), dtype=np.uint8)) % 4
                    f.write(f">f{fam}_r{rec} some description\n".encode())
                    text = np.frombuffer(b"ACGT", dtype=np.uint8)[seq]
                    for x in range(0, len(text), 70):
                        f.write(text[x:x + 70].tobytes() + b"\n")
            cfiles.append(path)
        golden = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
        fixture = [os.path.join(golden, "abisko4", f) for f in ("73.20120800_S1X.13.fna.gz", "73.20120600_S2D.19.fna.gz",
                                                                "73.20120700_S3X.12.fna.gz", "73.20110800_S2D.13.fna.gz")]
        fixture.append(os.path.join(golden, "antonio_mags", "BE_RX_R2_MAG52.fna.gz"))
        sk = {}
        for name, plist, kw in (("genomes", paths, dict(threshold=95.0, small_genomes=False, contigs=False)),
                                ("contigs", cfiles, dict(threshold=95.0, small_genomes=True, contigs=True)),
                                ("fixture", fixture, dict(threshold=90.0, min_aligned_fraction=20.0))):
            gb.init(0)
            h_one, u_one = gb.skani_distances(plist, **kw)
            gb.init_devices(G)
            h_many, u_many = gb.skani_distances_multi(plist, G, **kw)
            same = u_one == u_many and len(h_one) == len(h_many) and bool(np.all(h_one == h_many))
            sk[name] = {"files": len(plist), "units": u_many, "hits": int(len(h_many)), "identical_to_single_gpu": same}
            ok = ok and same
        gb.init_devices(G)
        fx, _ = gb.cluster_skani_multi(fixture, G, precluster_ani=90.0, ani=99.0, min_aligned_fraction=20.0)
        sk["fixture"]["clusters"] = fx.tolist()
        ok = ok and sorted(sorted(c) for c in fx.tolist()) == [[0, 1, 3], [2], [4]]
        out["skani_files"] = sk
    # ---- the skani preclusterer (contig mode) on the same buffers read as contigs of --contig-len bases
    if args.contigs:
        nc, Lc = args.contigs, args.contig_len
        layc = gb.synth_layout(nc, Lc)
        hc_seq = torch.empty(layc["seq2_words"], dtype=torch.int32, pin_memory=True)
        hc_val = torch.empty(layc["valid_words"], dtype=torch.int32, pin_memory=True)
        gb.init(0)
        dc_seq = torch.zeros(layc["seq2_words"], dtype=torch.int32, device=dev)
        dc_val = torch.zeros(layc["valid_words"], dtype=torch.int32, device=dev)
        dc_off = torch.zeros(nc + 1, dtype=torch.int64, device=dev)
        st = torch.cuda.current_stream().cuda_stream
        for c0 in range(0, nc, 1 << 16):
            m = min(1 << 16, nc - c0)
            gb.synth_packed_device(1, c0, m, Lc, dc_seq[c0 * layc["padded"] // 16:].data_ptr(),
                                   dc_val[c0 * layc["padded"] // 32:].data_ptr(), dc_off[c0:].data_ptr(), st)
        torch.cuda.synchronize()
        c_off = np.arange(nc + 1, dtype=np.uint64) * np.uint64(layc["padded"])
        dc_off.copy_(torch.from_numpy(c_off.view(np.int64)))
        hc_seq.copy_(dc_seq); hc_val.copy_(dc_val)
        torch.cuda.synchronize()
        c_len = np.full(nc, Lc, np.uint64)
        t0 = time.perf_counter()
        h1, i1 = gb.skani_distances_packed_device(dc_seq.data_ptr(), dc_val.data_ptr(), dc_off.data_ptr(), c_off, c_len, 95.0, 15.0,
                                                  small_genomes=True, stream=st)
        t_one = time.perf_counter() - t0
        del dc_seq, dc_val
        torch.cuda.empty_cache()
        gb.init_devices(G)
        for _ in range(args.passes):
            t0 = time.perf_counter()
            hm, im = gb.skani_distances_packed_multi(hc_seq.data_ptr(), hc_val.data_ptr(), c_off, c_len, G, 95.0, 15.0, small_genomes=True)
            t_many = time.perf_counter() - t0
        same = len(h1) == len(hm) and bool(np.all(h1 == hm))
        t0 = time.perf_counter()
        ccl, _ = gb.cluster_from_distances(nc, hm, 95.0, None, skip_clusterer=True)
        t_engine = time.perf_counter() - t0
        out["contigs"] = {"units": nc, "unit_bp": Lc, "pairs": nc * (nc - 1) // 2, "hits": int(len(hm)), "n_screened": im["n_screened"],
                          "single_gpu_resident_s": t_one, "multi_host_buffers_s": t_many, "engine_s": t_engine,
                          "clusters": len(ccl), "clusters_sha1": ccl.sha1(), "identical_to_single_gpu": same,
                          "pairs_per_s": (nc * (nc - 1) // 2) / (t_many + t_engine)}
        ok = ok and same
    print(json.dumps(out), flush=True)
    if args.json:
        with open(args.json, "w") as f:
            json.dump(out, f)
    print("ALL OK" if ok else "MISMATCH", flush=True)
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
