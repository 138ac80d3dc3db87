#!/usr/bin/env python
"""Multi-GPU parity check of galah_b200.distributed.ShardedPipeline (run under torchrun, one rank per
GPU): the whole two-stage path on G GPUs (genome slices, sharded K2, K3 with peer-mapped tables over
NVLink, engine on rank 0) must return the clusters of the single-GPU one-call pipeline on the same
synthetic genomes -- and ShardedPrefilter's hits must equal the CPU oracle's on a row sample.

    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/check_pipeline_sharded.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist
    import galah_b200 as gb
    from galah_b200.distributed import ShardedPipeline

    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    gb.init(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    st = torch.cuda.current_stream().cuda_stream
    ok = True
    # (genomes per rank, genome length, family size): the odd family size puts families across the
    # rank boundary, so that stage 2 has pairs whose reference table lives on a peer
    for n_local, L, fam in ((256, 200_000, 10), (384, 150_000, 7), (130, 120_000, 10)):
        n = n_local * world

        def synth(index_begin, count):
            lay = gb.synth_layout(count, L)
            d_seq = torch.empty(lay["seq2_words"], dtype=torch.int32, device=dev)
            d_val = torch.empty(lay["valid_words"], dtype=torch.int32, device=dev)
            d_off = torch.empty(count + 1, dtype=torch.int64, device=dev)
            gb.synth_packed_device_ex(7, index_begin, count, L, fam, 0, d_seq.data_ptr(), d_val.data_ptr(), d_off.data_ptr(), st)
            torch.cuda.synchronize()
            return d_seq, d_val, d_off, np.arange(count + 1, dtype=np.uint64) * np.uint64(lay["padded"]), np.full(count, L, np.uint64)

        d_seq, d_val, d_off, bo, ln = synth(rank * n_local, n_local)
        pipe = ShardedPipeline(gb, dist, n_local, 1000, dev)
        for mode in ("device", "host"):
            if mode == "device":
                clusters, info = pipe.step_device(d_seq.data_ptr(), d_val.data_ptr(), d_off.data_ptr(), bo, ln)
            else:
                h_seq, h_val = d_seq.cpu().pin_memory(), d_val.cpu().pin_memory()
                clusters, info = pipe.step_host(h_seq.data_ptr(), h_val.data_ptr(), bo, ln)
            remote = torch.tensor([info["remote_reference_pairs"]], device=dev)
            dist.all_reduce(remote)
            if rank == 0:
                a_seq, a_val, a_off, a_bo, a_ln = synth(0, n)
                exp, einfo = gb.cluster_packed(a_seq.data_ptr(), a_val.data_ptr(), a_bo, a_ln, device=True,
                                               d_base_off=a_off.data_ptr())
                same = clusters == exp and info["n_precluster_hits"] == einfo["n_precluster_hits"]
                print(f"n_local={n_local} L={L} family={fam} {mode}: {len(exp)} clusters, {einfo['n_precluster_hits']} hits, "
                      f"{int(remote.item())} stage-2 jobs read a peer's table: {'OK' if same else 'MISMATCH'}", flush=True)
                ok = ok and same
                del a_seq, a_val, a_off
    # ---- contig mode (BASELINE.json configs[4] shape): run_skani vs the single-GPU preclusterer + engine
    for n_local, L, fam in ((640, 30_000, 10), (384, 20_000, 7)):
        n = n_local * world

        def synth(index_begin, count):
            lay = gb.synth_layout(count, L)
            d_seq = torch.empty(lay["seq2_words"], dtype=torch.int32, device=dev)
            d_val = torch.empty(lay["valid_words"], dtype=torch.int32, device=dev)
            d_off = torch.empty(count + 1, dtype=torch.int64, device=dev)
            gb.synth_packed_device_ex(7, index_begin, count, L, fam, 0, d_seq.data_ptr(), d_val.data_ptr(), d_off.data_ptr(), st)
            torch.cuda.synchronize()
            return d_seq, d_val, d_off, np.arange(count + 1, dtype=np.uint64) * np.uint64(lay["padded"]), np.full(count, L, np.uint64)

        d_seq, d_val, d_off, bo, ln = synth(rank * n_local, n_local)
        pipe = ShardedPipeline(gb, dist, n_local, gb.marker_row_capacity(L, True), dev)
        clusters, info = pipe.run_skani(d_seq.data_ptr(), d_val.data_ptr(), d_off.data_ptr(), bo, ln, True, 92.0, 92.0, 15.0,
                                        small_genomes=True, individual_contigs=True)
        remote = torch.tensor([info["remote_reference_pairs"]], device=dev)
        dist.all_reduce(remote)
        if rank == 0:
            a_seq, a_val, a_off, a_bo, a_ln = synth(0, n)
            hits, _ = gb.skani_distances_packed_device(a_seq.data_ptr(), a_val.data_ptr(), a_off.data_ptr(), a_bo, a_ln, 92.0,
                                                       15.0, small_genomes=True, stream=st)
            exp, _ = gb.cluster_from_distances(n, hits, 92.0, None, skip_clusterer=True)
            same = clusters == exp and info["n_hits"] == len(hits)
            print(f"contigs n_local={n_local} L={L} family={fam}: {len(exp)} clusters, {len(hits)} hits, "
                  f"{int(remote.item())} pairs read a peer's table: {'OK' if same else 'MISMATCH'}", flush=True)
            ok = ok and same
    if rank == 0:
        print("ALL OK" if ok else "FAILED", flush=True)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok or rank != 0 else 1)


if __name__ == "__main__":
    main()
