#!/usr/bin/env python
"""BASELINE.json configs[3] and configs[4] on G GPUs of one box (one process per GPU, torchrun):

  --config 3   N synthetic 2 Mbp genomes (default 102,400 = 100k rounded to whole row blocks per rank at
               G = 8), finch 0.9 prefilter + ANI 95 two-stage, full pipeline (ShardedPipeline)
  --config 4   N synthetic 50 kbp contigs (default 1,003,520), --cluster-contigs --small-genomes
               (ShardedPipeline.run_skani)

Every rank generates and keeps ITS slice of the packed units in HBM (setup, untimed).  Rank 0 prints
one JSON line with the times, the cluster count and a SHA-1 of the cluster lists: runs at different
G must print the same digest (the clusters do not depend on the sharding).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 tools/config_runs.py --config 3
"""
import argparse
import hashlib
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", type=int, required=True, choices=[3, 4])
    ap.add_argument("--units", type=int, default=0)
    ap.add_argument("--steps", type=int, default=2)
    args = ap.parse_args()
    import torch
    import torch.distributed as dist
    import galah_b200 as gb
    from galah_b200.distributed import ShardedPipeline

    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    gb.init(local)
    os.environ.setdefault("NCCL_DEBUG", "WARN")
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    st = torch.cuda.current_stream().cuda_stream
    if args.config == 3:
        n, L, small = args.units or 102_400, 2_000_000, False
    else:
        n, L, small = args.units or 1_003_520, 50_000, True
    assert n % (world * 10) == 0, "whole families per rank"
    n_local = n // world
    lay = gb.synth_layout(n_local, L)
    d_seq = torch.empty(lay["seq2_words"], dtype=torch.int32, device=dev)
    d_val = torch.empty(lay["valid_words"], dtype=torch.int32, device=dev)
    d_off = torch.empty(n_local + 1, dtype=torch.int64, device=dev)
    step = 1 << 16
    for g0 in range(0, n_local, step):  # the generator takes one grid per call
        nb = min(step, n_local - g0)
        gb.synth_packed_device(1, rank * n_local + g0, nb, L, d_seq[g0 * lay["padded"] // 16:].data_ptr(),
                               d_val[g0 * lay["padded"] // 32:].data_ptr(), d_off[g0:].data_ptr(), st)
    base_off = np.arange(n_local + 1, dtype=np.uint64) * np.uint64(lay["padded"])
    d_off.copy_(torch.from_numpy(base_off.view(np.int64)))
    torch.cuda.synchronize()
    lengths = np.full(n_local, L, np.uint64)
    stride = 1000 if args.config == 3 else gb.marker_row_capacity(L, small)
    pipe = ShardedPipeline(gb, dist, n_local, stride, dev)
    times, infos, clusters = [], [], None
    for it in range(1 + args.steps):
        dist.barrier(); torch.cuda.synchronize()
        t0 = time.perf_counter()
        if args.config == 3:
            clusters, info = pipe.step_device(d_seq.data_ptr(), d_val.data_ptr(), d_off.data_ptr(), base_off, lengths,
                                              0.9, 95.0, 15.0)
        else:
            clusters, info = pipe.run_skani(d_seq.data_ptr(), d_val.data_ptr(), d_off.data_ptr(), base_off, lengths, True,
                                            95.0, 95.0, 15.0, small_genomes=True, individual_contigs=True)
        torch.cuda.synchronize()
        dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        if it >= 1:
            times.append(float(dt.item())); infos.append(info)
    if rank == 0:
        pairs = n * (n - 1) // 2
        t = float(np.mean(times))
        keys = [k for k in infos[-1] if k.endswith("_ms") and not isinstance(infos[-1][k], dict)]
        line = {"config": f"BASELINE.json configs[{args.config}]", "n_gpus": world, "units": n, "unit_bp": L, "pairs": pairs,
                "seconds_per_pass": t, "pairs_per_s": pairs / t, "clusters": len(clusters),
                "clusters_sha1": clusters.sha1(), "phases_ms_rank0": {k: float(np.median([i[k] for i in infos])) for k in keys},
                "counts": {k: infos[-1][k] for k in infos[-1] if not k.endswith("_ms") and isinstance(infos[-1][k], (int, float))},
                "host_detail_ms_rank0": infos[-1].get("host_detail_ms"),
                "resident_gb_per_gpu": (d_seq.numel() + d_val.numel()) * 4 / 1e9}
        print(json.dumps(line), flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
