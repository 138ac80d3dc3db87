#!/usr/bin/env python
"""BASELINE.json configs[4] at scale on ONE GPU: N synthetic contigs (families of 10, 50 kbp each,
SURVEY.md 8d) through the contig-mode path `--cluster-contigs --small-genomes`
(SkaniPreclusterer::distances_contigs, reference src/skani.rs:379-498, then the skip_clusterer
greedy stage of src/clusterer.rs:32-44): K3 index, marker sketches, marker-containment screen on
the K2 join, K3 ANI on the survivors, host engine.  Prints one JSON line.

    python tools/config5_contigs.py --contigs 100000
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
    ap.add_argument("--contigs", type=int, default=100_000)
    ap.add_argument("--length", type=int, default=50_000)
    ap.add_argument("--ani", type=float, default=95.0)
    ap.add_argument("--min-af", type=float, default=15.0)
    args = ap.parse_args()
    import torch
    import galah_b200 as gb

    gb.init(0)
    dev = torch.device("cuda", 0)
    n, L = args.contigs, args.length
    lay = gb.synth_layout(n, L)
    st = torch.cuda.current_stream().cuda_stream
    d_seq = torch.zeros(lay["seq2_words"], dtype=torch.int32, device=dev)
    d_val = torch.zeros(lay["valid_words"], dtype=torch.int32, device=dev)
    d_off = torch.zeros(n + 1, dtype=torch.int64, device=dev)
    t0 = time.perf_counter()
    step = 1 << 16  # generate in slices (the generator takes one grid per call)
    for g0 in range(0, n, step):
        nb = min(step, n - g0)
        gb.synth_packed_device(1, g0, nb, L, d_seq[g0 * lay["padded"] // 16:].data_ptr(),
                               d_val[g0 * lay["padded"] // 32:].data_ptr(), d_off[g0:].data_ptr(), st)
    torch.cuda.synchronize()
    base_off = np.arange(n + 1, dtype=np.uint64) * np.uint64(lay["padded"])
    d_off.copy_(torch.from_numpy(base_off.view(np.int64)))
    torch.cuda.synchronize()
    t_synth = time.perf_counter() - t0
    t0 = time.perf_counter()
    hits, info = gb.skani_distances_packed_device(d_seq.data_ptr(), d_val.data_ptr(), d_off.data_ptr(), base_off,
                                                  np.full(n, L, np.uint64), args.ani, args.min_af, small_genomes=True,
                                                  stream=st)
    t_pre = time.perf_counter() - t0
    t0 = time.perf_counter()
    clusters, cinfo = gb.cluster_from_distances(n, hits, args.ani, None, skip_clusterer=True)
    t_greedy = time.perf_counter() - t0
    pairs = n * (n - 1) // 2
    same_family = bool(np.all(hits["i"] // 10 == hits["j"] // 10))
    line = {
        "workload": f"{n} synthetic contigs x {L} bp (families of 10, seed 1), --cluster-contigs --small-genomes, "
                    f"ANI {args.ani}, min-AF {args.min_af} (BASELINE.json configs[4] on one GPU)",
        "n_gpus": 1, "contig_pairs": pairs, "screened_pairs": info["n_screened"], "hits": int(len(hits)),
        "clusters": len(clusters), "all_hits_within_families": same_family,
        "seconds": {"synthesise": t_synth, "precluster_total": t_pre, "greedy": t_greedy},
        "precluster_ms": {k: info[k] for k in ("index_ms", "markers_ms", "screen_ms", "ani_ms", "total_ms")},
        "contig_pairs_per_s": pairs / (t_pre + t_greedy),
        "peak_device_memory_gb": torch.cuda.max_memory_allocated() / 1e9,
        "note": "peak_device_memory_gb counts torch's buffers only; the library's own allocations come on top",
    }
    print(json.dumps(line), flush=True)


if __name__ == "__main__":
    main()
