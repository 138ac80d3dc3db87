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
ht = torch.from_numpy(table.view(np.int64)).pin_memory(); hc = torch.from_numpy(counts.view(np.int32)).pin_memory()
t = ht.numpy().view(np.uint64); c = hc.numpy().view(np.uint32)
def run(chunks, label, env={}):
    gb.prefilter_stream_chunks(chunks)
    for e, v in env.items(): os.environ[e] = v
    ts = []
    for it in range(12):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        r = gb.prefilter(t, c, 21, 0.9)
        dt = time.perf_counter() - t0
        if it >= 2: ts.append(dt)
    for e in env: os.environ.pop(e)
    print(label, chunks, f"mean {np.mean(ts)*1e3:.3f} ms min {np.min(ts)*1e3:.3f}", len(r), gb.prefilter_last_host_timing(), flush=True)
for rep in range(2):
    run(1, "plain")
    run(4, "s4-map")
    run(4, "s4-nomap", {"GALAH_B200_NO_MAP": "1"})
    run(8, "s8-map")
    run(2, "s2-map")
