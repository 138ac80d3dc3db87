#!/usr/bin/env python
"""Small driver for ncu captures of K2 (build + join) and K3 (emit, chain): the one-call pipeline over
N synthetic genomes resident on the device, twice (the second pass is the one to capture).

    ncu --set full --clock-control none --import-source on -k regex:"ani_chain|prefilter_join|ani_emit" \
        -o gpurun_out/k23 python tools/k23_profile.py 10000
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    import torch
    import galah_b200 as gb
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
    L = 2_000_000
    gb.init(0)
    dev = torch.device("cuda", 0)
    lay = gb.synth_layout(n, L)
    d_seq = torch.empty(lay["seq2_words"], dtype=torch.int32, device=dev)
    d_val = torch.empty(lay["valid_words"], dtype=torch.int32, device=dev)
    d_off = torch.empty(n + 1, dtype=torch.int64, device=dev)
    gb.synth_packed_device(1, 0, n, L, d_seq.data_ptr(), d_val.data_ptr(), d_off.data_ptr(), torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    base_off = np.arange(n + 1, dtype=np.uint64) * np.uint64(lay["padded"])
    for rep in range(2):
        cl, info = gb.cluster_packed(d_seq.data_ptr(), d_val.data_ptr(), base_off, np.full(n, L, np.uint64), device=True,
                                     d_base_off=d_off.data_ptr())
        print({k: round(v, 2) if isinstance(v, float) else v for k, v in info.items()}, flush=True)


if __name__ == "__main__":
    main()
