import os, sys, time
sys.path.insert(0, "/root/repo")
import numpy as np, torch
import galah_b200 as gb
gb.init(0)
sys.path.insert(0, "/root/repo/tests")
from util import random_family_table
rng = np.random.default_rng(5)
n = 10000
table, counts = random_family_table(n, 1000, rng)
def run(t, c, label):
    d_t = torch.from_numpy(t.view(np.int64)).cuda(); d_c = torch.from_numpy(c.view(np.int32)).cuda()
    cap = 1 << 20
    d_cand = torch.zeros((cap, 4), dtype=torch.int32, device="cuda"); d_n = torch.zeros(1, dtype=torch.int64, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    ms = []
    for it in range(8):
        gb.prefilter_enqueue(d_t.data_ptr(), d_c.data_ptr(), n, 1000, 21, 0.9, 0, 1, 0, st, d_cand.data_ptr(), cap, d_n.data_ptr())
        torch.cuda.synchronize()
        ms.append(gb.prefilter_last_timing())
    print(label, "build %.3f join %.3f ms" % tuple(np.median(np.array(ms[2:]), axis=0)), int(d_n.item()), flush=True)
run(table, counts, "contiguous")
perm = rng.permutation(n)
run(np.ascontiguousarray(table[perm]), np.ascontiguousarray(counts[perm]), "shuffled")
