#!/usr/bin/env python
"""Multi-GPU parity check of galah_b200.distributed.ShardedPrefilter (run under torchrun, one rank
per GPU): every rank passes its slice of the same seeded sketch table; the union of the ranks'
pair lists must equal the single-GPU result of rank 0, bit for bit.

    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/check_sharded.py
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))


def main():
    import torch
    import torch.distributed as dist
    import galah_b200 as gb
    from galah_b200.distributed import ShardedPrefilter
    from util import random_family_table

    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    gb.init(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    ok = True
    # whole-block slices (local build; ring exchange when GALAH_B200_RING=1) and not (gather-first)
    for n_local, ragged in ((256, False), (384, True), (200, False)):
        n = n_local * world
        rng = np.random.default_rng(100 + n_local)
        table, counts = random_family_table(n, 1000, rng, ragged=ragged)
        sp = ShardedPrefilter(gb, dist, n_local, 1000, dev)
        h_t = torch.from_numpy(table[rank * n_local:(rank + 1) * n_local].view(np.int64)).pin_memory()
        h_c = torch.from_numpy(counts[rank * n_local:(rank + 1) * n_local].view(np.int32)).pin_memory()
        for rep in range(2):
            mine = sp(h_t, h_c, 21, 0.9)
            gathered = [None] * world if rank == 0 else None
            dist.gather_object(mine, gathered, dst=0)
            if rank == 0:
                allp = np.concatenate(gathered)
                allp = allp[np.lexsort((allp["j"], allp["i"]))]
                exp = gb.prefilter(table, counts, 21, 0.9)
                same = len(allp) == len(exp) and all(np.array_equal(allp[f], exp[f]) for f in ("i", "j", "common", "total")) \
                    and np.array_equal(allp["ani"].view(np.uint32), exp["ani"].view(np.uint32))
                share = [len(g) for g in gathered]
                print(f"n_local={n_local} ragged={ragged} ring={sp.ring} rep={rep}: pairs {len(exp)} per-rank {share} "
                      f"{'OK' if same else 'MISMATCH'}", flush=True)
                ok = ok and same
    if rank == 0:
        print("ALL OK" if ok else "FAILED", flush=True)
    dist.destroy_process_group()
    sys.exit(0 if ok or rank != 0 else 1)


if __name__ == "__main__":
    main()
