#!/usr/bin/env python
"""Small driver for ncu captures of the ingest kernels (K1 fused scan, select, K3 count / emit):
one ingest_packed pass over N synthetic genomes resident on the device.

    ncu --set full --clock-control none --import-source on -k regex:scan21 -c 1 -o gpurun_out/k1 \
        python tools/k1_profile.py 512
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    import torch
    import galah_b200 as gb
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
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
    table = torch.empty((n, 1000), dtype=torch.int64, device=dev)
    counts = torch.empty(n, dtype=torch.int32, device=dev)
    idx = gb.AniIndex()
    for rep in range(2):
        idx.clear()
        ms = idx.ingest_packed(d_seq.data_ptr(), d_val.data_ptr(), base_off, np.full(n, L, np.uint64), table.data_ptr(),
                               counts.data_ptr(), device=True, d_base_off=d_off.data_ptr())
        print(f"pass {rep}: K1 {ms[0]:.2f} ms, index {ms[1]:.2f} ms for {n} genomes", flush=True)


if __name__ == "__main__":
    main()
