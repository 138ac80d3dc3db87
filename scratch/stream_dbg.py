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
    for it in range(4):
        if it == 3: os.environ["GALAH_B200_STREAM_DEBUG"] = "1"
        torch.cuda.synchronize(); t0 = time.perf_counter()
        r = gb.prefilter(t, c, 21, 0.9)
        dt = time.perf_counter() - t0
        os.environ.pop("GALAH_B200_STREAM_DEBUG", None)
    for e in env: os.environ.pop(e)
    print(label, chunks, f"{dt*1e3:.3f} ms", len(r), gb.prefilter_last_timing(), flush=True)
run(1, "plain")
run(4, "s4")
run(4, "s4-w50", {"GALAH_B200_STREAM_WAVE": "0.5"})
run(4, "s4-w75", {"GALAH_B200_STREAM_WAVE": "0.75"})
run(6, "s6")
run(6, "s6-w66", {"GALAH_B200_STREAM_WAVE": "0.667"})
run(8, "s8-w75", {"GALAH_B200_STREAM_WAVE": "0.75"})
