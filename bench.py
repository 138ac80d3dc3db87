#!/usr/bin/env python
"""bench.py -- genome-pairs/sec of the finch-prefilter hot path on B200 (BASELINE.json metric).

A "step" is one pass of the all-pairs prefilter (reference src/finch.rs:75-95) over the sketch
table of N synthetic genomes (families of 10, 2 Mbp, seed 1, sketched on the GPU by
K1 during untimed setup).  At --gpus 1 the workload is BASELINE.json configs[1]: 10,000 genomes,
s = 1000, prefilter only.  At G > 1 GPUs the genome count grows as 10,000 * sqrt(G) so that the
pair area per GPU is constant (weak scaling); every rank sketches its own genome slice, and the
timed step is: NCCL all-gather of the sketch table -> row-block-sharded prefilter kernel.

  value    = pairs / device time (CUDA events on the launch stream, max over ranks), inputs in HBM
  e2e      = same metric through the host-buffer C-ABI call (H2D table, kernel, D2H pair list,
             f64 finish + sort on the host)
  roofline = algorithmic bytes (16,000 B/pair = 2*s*8) / kernel time vs the measured HBM peak
  cpu_baseline = the oracle's serial pair loop (as the reference's) on a row sample, host cores

`--impl reference` times the CPU restatement of the reference path (oracle/) only.
"""
import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

S = 1000
K = 21
MIN_ANI = 0.9
SEED = 1
BYTES_PER_PAIR = 2 * S * 8


def n_genomes_for(gpus, base):
    n = base * math.sqrt(gpus)
    # per rank a whole number of families of 10 AND of row blocks of 128 (so a rank's slice is whole
    # block lists and its build needs no gathered table): multiples of lcm(10, 128) = 640
    q = 640 * gpus if gpus > 1 else 80
    return max(q, int(round(n / q)) * q)


def ncu_traffic(kernel, n, world):
    """dram bytes per launch of `kernel` from the committed ncu capture (profiles/ncu_traffic.json);
    only valid for the workload it was captured on (1 GPU, same N), else None."""
    p = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if world != 1 or not os.path.exists(p):
        return None
    with open(p) as f:
        t = json.load(f)
    if t.get("n_genomes") != n or kernel not in t:
        return None
    return t[kernel]["dram_bytes_read"] + t[kernel]["dram_bytes_write"]


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.device = device
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20",
                 "-i", str(self.device)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def mark(self):
        """Drop what was sampled so far, but keep the latest row so that a timed region shorter
        than the sampling interval still has a reading taken under the same load."""
        self.rows = self.rows[-1:]

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm = [float(r[1]) for r in self.rows if len(r) >= 8 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 8 and r[2].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [nm for c, nm in enumerate(names) if any(len(r) >= 8 and r[4 + c] == "Active" for r in self.rows)]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


def run_reference(args):
    """CPU restatement of the reference path (oracle/): all-core sketch of a genome sample
    (untimed, as our arm's setup), then the SERIAL pair loop exactly as src/finch.rs:75-95."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import oracle
    oracle.build()
    cores = os.cpu_count() or 1
    n_sample = args.ref_genomes
    t0 = time.time()
    table, counts = oracle.sketch_synth(SEED, 0, n_sample, args.genome_len, K, S, 0)
    t_sketch = time.time() - t0
    rows = args.ref_rows
    pairs_per_step = sum(n_sample - 1 - i for i in range(rows))
    times = []
    for it in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        oracle.prefilter(table, counts, K, MIN_ANI, row_begin=0, row_end=rows)
        dt = time.perf_counter() - t0
        if it >= args.warmup:
            times.append(dt)
    total = sum(times)
    value = pairs_per_step * len(times) / total
    # context only: the same rows with every host thread (the reference's loop itself is serial,
    # src/finch.rs:75-95, so the headline of this arm stays the one-thread figure)
    t0 = time.perf_counter()
    oracle.prefilter_count_mt(table, counts, K, MIN_ANI, row_begin=0, row_end=rows)
    all_cores_value = pairs_per_step / (time.perf_counter() - t0)
    sample = (f"rows 0..{rows} x {n_sample} columns of the synthetic sketch table "
              f"({pairs_per_step} pairs/step), serial loop as src/finch.rs:75-95; "
              f"sketch of the {n_sample}-genome sample on {cores} threads took {t_sketch:.1f}s (untimed)")
    line = {
        "impl": "reference", "metric": "genome-pairs/sec (finch prefilter, s=1000)", "value": value,
        "unit": "pairs/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * total / len(times), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u64", "data": "synthetic",
        "config": {"workload": f"{n_genomes_for(args.gpus, args.n_genomes)} synthetic 2 Mbp genomes, "
                               "s=1000 finch prefilter only (BASELINE.json configs[1])",
                   "timed_sample": sample},
        "cpu_baseline": {"value": value, "unit": "pairs/s", "cores": 1, "kind": "port", "sample": sample,
                         "all_cores_value": all_cores_value, "all_cores": cores},
        "e2e": {"value": value, "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--n-genomes", type=int, default=10000, help="genomes at 1 GPU (x sqrt(G) at G GPUs)")
    ap.add_argument("--genome-len", type=int, default=2_000_000)
    ap.add_argument("--mode", type=int, default=0, help="0 = block-list join (default), 1 = pairwise warp merge")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-other-mode", action="store_true", help="skip timing the other kernel path")
    ap.add_argument("--no-stage2", action="store_true", help="skip the stage-2 (ANI on prefilter survivors) section")
    ap.add_argument("--no-ingest", action="store_true", help="skip the FASTA-ingest (K0) section")
    ap.add_argument("--cpu-rows", type=int, default=100)
    ap.add_argument("--ref-genomes", type=int, default=2000)
    ap.add_argument("--ref-rows", type=int, default=100)
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        return run_reference(args)

    # stdout carries exactly ONE JSON line: everything else a library prints there (the NCCL
    # version banner, for one) is sent to stderr at the file-descriptor level
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)

    import torch
    import torch.distributed as dist
    import galah_b200 as gb

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}: launch with torch.distributed.run")
    torch.cuda.set_device(local_rank)
    gb.init(local_rank)
    if world > 1:
        # NCCL prints its version banner on STDOUT at NCCL_DEBUG=VERSION; keep stdout to the one JSON line
        if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
            os.environ["NCCL_DEBUG"] = "WARN"
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dev = torch.device("cuda", local_rank)
    stream = torch.cuda.current_stream()
    st = stream.cuda_stream

    n = n_genomes_for(world, args.n_genomes) if args.n_genomes >= 80 else args.n_genomes
    n_local = n // world
    L = args.genome_len

    # ---------------- setup (untimed): synthetic genomes -> K1 sketches of this rank's slice
    my_table = torch.empty((n_local, S), dtype=torch.int64, device=dev)
    my_counts = torch.empty(n_local, dtype=torch.int32, device=dev)
    batch = max(1, min(n_local, (1 << 30) // max(L, 1)))  # ~1 G bases (375 MB packed) per batch
    lay = gb.synth_layout(batch, L)
    d_seq = torch.zeros(lay["seq2_words"], dtype=torch.int32, device=dev)
    d_val = torch.zeros(lay["valid_words"], dtype=torch.int32, device=dev)
    d_off = torch.zeros(batch + 1, dtype=torch.int64, device=dev)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sketch_ms = 0.0
    synth_ms = 0.0
    stage2 = not args.no_stage2 and world == 1
    ani_index = gb.AniIndex() if stage2 else None
    ani_build_ms = 0.0
    for b0 in range(0, n_local, batch):
        nb = min(batch, n_local - b0)
        e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        e0.record(stream)
        gb.synth_packed_device(SEED, rank * n_local + b0, nb, L, d_seq.data_ptr(), d_val.data_ptr(),
                               d_off.data_ptr(), st)
        e1.record(stream)
        gb.sketch_packed_device(d_seq.data_ptr(), d_val.data_ptr(), d_off.data_ptr(), nb, K, S, 0,
                                my_table[b0:].data_ptr(), my_counts[b0:].data_ptr(), st)
        e2.record(stream)
        torch.cuda.synchronize()
        synth_ms += e0.elapsed_time(e1)
        sketch_ms += e1.elapsed_time(e2)
        if stage2:
            ani_index.add_packed_device(d_seq.data_ptr(), d_val.data_ptr(), d_off.data_ptr(),
                                        np.arange(nb + 1, dtype=np.uint64) * np.uint64(lay["padded"]),
                                        np.full(nb, L, np.uint64), st)
            ani_build_ms += ani_index.last_timing()[0]
            if b0 == 0:
                ani_index.reserve(n_local)
    del d_seq, d_val, d_off

    flush = torch.empty(256 << 20, dtype=torch.int8, device=dev)  # > 126 MB L2
    sp = None
    if world > 1:
        # multi-GPU: galah_b200.distributed.ShardedPrefilter is the product path; the timed step is
        # its device part (all-reduce of the largest hash, table all-gather overlapped with the
        # per-rank build of its own block lists, all-gather of the lists, join of the rank's shard)
        from galah_b200.distributed import ShardedPrefilter
        sp = ShardedPrefilter(gb, dist, n_local, S, dev)
        sp.my_table.copy_(my_table)
        sp.my_counts.copy_(my_counts)
        table, counts, d_cand, d_ncand, cand_cap = sp.table, sp.counts, sp.d_cand, sp.d_ncand, sp.cand_cap
    else:
        table, counts = my_table, my_counts
        cand_cap = max(1 << 20, 64 * n)
        d_cand = torch.empty((cand_cap, 4), dtype=torch.int32, device=dev)
        d_ncand = torch.zeros(1, dtype=torch.int64, device=dev)

    def timed(mode, steps, warmup):
        """`steps` timed passes of `mode`; returns (total_ms max over ranks, launches, per-kernel ms)."""
        def step():
            if world > 1 and mode == 0:
                sp.step_device(K, MIN_ANI)
                return
            if world > 1:
                dist.all_gather_into_tensor(table, my_table)
                dist.all_gather_into_tensor(counts, my_counts)
            gb.prefilter_enqueue(table.data_ptr(), counts.data_ptr(), n, S, K, MIN_ANI, rank, world,
                                 mode, st, d_cand.data_ptr(), cand_cap, d_ncand.data_ptr())
        for _ in range(warmup):
            flush.fill_(1)
            step()
        sync_all()
        if steps == 0:
            return 0.0, 0, 0.0, 0.0
        launches0 = gb.launch_count()
        step_ms, build_ms, main_ms = [], [], []
        for _ in range(steps):
            flush.fill_(1)  # L2 flush between timed iterations (outside the events)
            sync_all()
            ev0.record(stream)
            t_host = time.perf_counter()
            step()
            host_enqueue_ms.append(1e3 * (time.perf_counter() - t_host))
            ev1.record(stream)
            torch.cuda.synchronize()
            step_ms.append(ev0.elapsed_time(ev1))
            b, m = gb.prefilter_last_timing()
            build_ms.append(b); main_ms.append(m)
        sync_all()
        launches = gb.launch_count() - launches0
        tot = torch.tensor([sum(step_ms)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tot, op=dist.ReduceOp.MAX)
        return float(tot.item()), launches, float(np.mean(build_ms)), float(np.mean(main_ms))

    host_enqueue_ms = []

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()  # nvidia-smi needs a moment to start: launched ahead of the warm-up passes
    timed(args.mode, 0, args.warmup)
    if rank == 0:
        sampler.mark()   # only samples from here on (the timed region) are kept
    total_ms, launches, build_ms, main_ms = timed(args.mode, args.steps, 0)
    clocks = sampler.stop() if rank == 0 else None
    pairs = n * (n - 1) // 2
    value = pairs * args.steps / (total_ms * 1e-3)
    n_cand = int(d_ncand.item())
    # the other exact kernel path, for comparison (fewer steps: the pairwise kernel is ~100x slower)
    other = None
    if not args.no_other_mode:
        o_steps = max(1, min(args.steps, 3))
        o_total, o_launches, o_build, o_main = timed(1 - args.mode, o_steps, 1)
        other = {"mode": 1 - args.mode, "value": pairs * o_steps / (o_total * 1e-3), "unit": "pairs/s",
                 "ms_per_step": o_total / o_steps, "steps": o_steps, "main_kernel_ms": o_main,
                 "candidates": int(d_ncand.item())}

    # ---------------- e2e: host buffers through the public API (H2D + kernels + D2H + host finish)
    # 1 GPU: the host-buffer C-ABI call.  G GPUs: galah_b200.distributed.ShardedPrefilter -- every
    # rank uploads ITS slice of the table, all-gather over NVLink, sharded build + join.
    h_table = table.cpu().pin_memory()
    h_counts = counts.cpu().pin_memory()
    np_table = h_table.numpy().view(np.uint64)   # views of the pinned host buffers
    np_counts = h_counts.numpy().view(np.uint32)
    e2e_times = []
    n_pass = 0
    if world > 1:
        h_my_table = h_table[rank * n_local:(rank + 1) * n_local]
        h_my_counts = h_counts[rank * n_local:(rank + 1) * n_local]
    for it in range(2 + min(args.steps, 5)):
        sync_all()
        t0 = time.perf_counter()
        if world > 1:
            res = sp(h_my_table, h_my_counts, K, MIN_ANI)
        else:
            res = gb.prefilter(np_table, np_counts, K, MIN_ANI, shard=rank, n_shards=world)
        dt = time.perf_counter() - t0
        n_pass = len(res)
        if it >= 2:
            e2e_times.append(dt)
    e2e_host = gb.prefilter_last_host_timing() if world == 1 else None
    e2e_single_upload = None
    if world == 1:
        # the same call with the upload pipeline off (one 80 MB copy, then the kernels)
        prev_chunks = gb.prefilter_stream_chunks(1)
        ts = []
        for it in range(4):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            res1 = gb.prefilter(np_table, np_counts, K, MIN_ANI)
            if it >= 1:
                ts.append(time.perf_counter() - t0)
        gb.prefilter_stream_chunks(prev_chunks)
        assert len(res1) == len(res) and np.array_equal(res1["j"], res["j"]) and np.array_equal(
            res1["ani"].view(np.uint32), res["ani"].view(np.uint32))
        e2e_single_upload = {"value": n * (n - 1) // 2 / (sum(ts) / len(ts)), "ms": 1e3 * sum(ts) / len(ts),
                             "host_ms": gb.prefilter_last_host_timing()}
    e2e_t = torch.tensor([sum(e2e_times) / len(e2e_times)], dtype=torch.float64, device=dev)
    n_pass_t = torch.tensor([n_pass], dtype=torch.int64, device=dev)
    if world > 1:
        dist.all_reduce(e2e_t, op=dist.ReduceOp.MAX)
        dist.all_reduce(n_pass_t, op=dist.ReduceOp.SUM)
    e2e_value = pairs / float(e2e_t.item())
    h2d = h_table.numel() * 8 + h_counts.numel() * 4  # whole job: at G GPUs every rank uploads 1/G of it
    d2h = int(n_pass) * 16 + 8

    # ---------------- stage 2: ANI of every prefilter survivor (K3), then the greedy engine
    two_stage = None
    if stage2:
        hit_pairs = np.stack([res["i"], res["j"]], axis=1).astype(np.uint32)
        ani_index.pairs(hit_pairs[: min(len(hit_pairs), 1024)], 15.0)  # warm-up
        t0 = time.perf_counter()
        ani_res = ani_index.pairs(hit_pairs, 15.0)
        t_ani = time.perf_counter() - t0
        chain_ms = ani_index.last_timing()[1]
        t0 = time.perf_counter()
        clusters, cinfo = gb.cluster_from_ani_table(n, res, ani_res["ani"], 95.0)
        t_greedy = time.perf_counter() - t0
        table = {(int(a), int(b)): float(v) for (a, b), v in zip(hit_pairs, ani_res["ani"])}
        seeds_per_genome = float(np.mean([ani_index.genome(g)["n_seeds"] for g in range(0, n, max(1, n // 50))]))
        e2e_prefilter_s = float(e2e_t.item())
        two_stage = {
            "workload": f"ANI (c=125, k=15) of the {len(hit_pairs)} prefilter survivors, 95 % threshold, min-AF 15",
            "ani_pairs": int(len(hit_pairs)), "ani_pairs_per_s": len(hit_pairs) / t_ani,
            "ani_chain_kernel_ms": chain_ms, "ani_call_ms": 1e3 * t_ani,
            "index_build_ms_total": ani_build_ms, "index_genomes_per_s": n / (ani_build_ms * 1e-3) if ani_build_ms else None,
            "seeds_per_genome": seeds_per_genome,
            "chain_kernel_gbs_algorithmic": (2 * seeds_per_genome * 12 * len(hit_pairs)) / (chain_ms * 1e-3) / 1e9 if chain_ms else None,
            "greedy_engine_ms": 1e3 * t_greedy, "clusters": len(clusters),
            "genome_pairs_per_s_prefilter_plus_ani": pairs / (e2e_prefilter_s + t_ani + t_greedy),
            "note": "genome_pairs_per_s_prefilter_plus_ani = N(N-1)/2 pairs considered / (host-buffer prefilter call "
                    "+ ANI call + greedy engine, all through the C ABI)",
        }
        if not args.no_cpu_baseline:
            import oracle
            sample = hit_pairs[:: max(1, len(hit_pairs) // 60)][:60]
            gen = {}
            for g in sorted(set(int(x) for x in sample.ravel())):
                gen[g] = oracle.AniGenome(*oracle.codes_from_ascii(oracle.synth_genome(SEED, g, L)))
            t0 = time.perf_counter()
            bad = 0
            for a, b in sample:
                v = oracle.ani_pair(gen[int(a)], gen[int(b)], 15.0)[0]
                bad += int(np.float32(v) != np.float32(table[(int(a), int(b))]))
            dt = time.perf_counter() - t0
            two_stage["cpu_baseline"] = {"value": len(sample) / dt, "unit": "ANI pairs/s", "cores": 1, "kind": "port",
                                         "sample": f"{len(sample)} of the survivor pairs, seeds precomputed (the "
                                                   "reference re-sketches both genomes in a fresh skani process per pair)",
                                         "gpu_matches_oracle_on_sample": bad == 0}

    # ---------------- K0: FASTA ingest on the device vs the host packer
    ingest = None
    if rank == 0 and world == 1 and not args.no_ingest:
        import tempfile
        rng = np.random.default_rng(SEED)
        n_files, glen, width = 48, 2_000_000, 80
        acgt = np.frombuffer(b"ACGT", np.uint8)
        files = []
        for g in range(n_files):
            seq = acgt[rng.integers(0, 4, size=glen)].reshape(-1, width)
            body = np.concatenate([seq, np.full((seq.shape[0], 1), 10, np.uint8)], axis=1).tobytes()
            files.append(b">genome_%d synthetic\n" % g + body)
        raw_bytes = sum(len(f) for f in files)
        gb.decode_fasta_device(files, unpack=False)  # warm-up: sizes the decoder's device buffers
        meta, dec_ms = gb.decode_fasta_device(files, unpack=False)
        with tempfile.TemporaryDirectory(dir="/dev/shm" if os.path.isdir("/dev/shm") else None) as td:
            paths = []
            for g, data in enumerate(files):
                paths.append(os.path.join(td, f"g{g}.fna"))
                with open(paths[-1], "wb") as f:
                    f.write(data)
            times = {}
            tables = {}
            for mode in (1, 0, 1, 0):
                prev = gb.device_ingest(mode)
                t0 = time.perf_counter()
                tables[mode] = gb.sketch_files(paths, K, S)
                times[mode] = time.perf_counter() - t0  # the second round of each mode is kept (warm)
                gb.device_ingest(prev)
            same = np.array_equal(tables[1][0], tables[0][0]) and np.array_equal(tables[1][1], tables[0][1])
        ingest = {"workload": f"{n_files} synthetic FASTA files x {glen} bp, {width}-column lines ({raw_bytes} bytes)",
                  "k0_decode_ms": dec_ms, "k0_decode_gbytes_per_s": raw_bytes / (dec_ms * 1e-3) / 1e9,
                  "k0_note": "upload from pinned staging + 3 kernels + 2 host round trips of per-chunk summaries",
                  "sketch_files_s_device_ingest": times[1], "sketch_files_s_host_packer": times[0],
                  "host_threads": os.cpu_count(), "tables_identical": bool(same),
                  "all_bases_ok": all(m["n_bases"] == glen and m["n_ambiguous"] == 0 for m in meta)}

    if rank == 0:
        peak, peak_src = peaks()
        kernel_names = {0: "prefilter_join_kernel", 1: "prefilter_tiled_kernel"}
        pairs_per_launch = pairs / world
        achieved = BYTES_PER_PAIR * pairs_per_launch / (main_ms * 1e-3) / 1e9
        # bytes the join kernel itself has to read: each (rb, cb) item streams the 32-bit keys of two
        # block lists of R*s entries (R = 128 sketches) once for R*R pairs (lo words / tags only on key ties)
        own_bytes_per_pair = 2 * S * 4 / gb.ROW_BLOCK if args.mode == 0 else BYTES_PER_PAIR
        cpu = None
        if not args.no_cpu_baseline and world == 1:
            import oracle
            oracle.build()
            tab = h_table.numpy().view(np.uint64)
            cnt = h_counts.numpy().view(np.uint32)
            rows = min(args.cpu_rows, n)
            t0 = time.perf_counter()
            exp = oracle.prefilter(tab, cnt, K, MIN_ANI, row_begin=0, row_end=rows)
            dt = time.perf_counter() - t0
            sample_pairs = sum(n - 1 - i for i in range(rows))
            # the same rows from the GPU result must agree bit-exactly (the oracle as checker)
            got = res[res["i"] < rows]
            ok = len(got) == len(exp) and all(
                np.array_equal(got[f], exp[f]) for f in ("i", "j", "common", "total")) and np.array_equal(
                got["ani"].view(np.uint32), exp["ani"].view(np.uint32))
            t1 = time.perf_counter()
            oracle.prefilter_count_mt(tab, cnt, K, MIN_ANI, row_begin=0, row_end=rows)
            dt_mt = time.perf_counter() - t1
            cpu = {"value": sample_pairs / dt, "unit": "pairs/s", "cores": 1, "kind": "port",
                   "sample": f"rows 0..{rows} of the same {n}-genome table ({sample_pairs} pairs), serial "
                             "loop as src/finch.rs:75-95 (the reference's pair loop is single-threaded)",
                   "all_cores_value": sample_pairs / dt_mt, "all_cores": os.cpu_count(),
                   "gpu_matches_oracle_on_sample": bool(ok)}
        line = {
            "metric": "genome-pairs/sec (finch prefilter, s=1000)", "value": value, "unit": "pairs/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": total_ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u64",
            "data": "synthetic",
            "config": {"workload": f"{n} synthetic {L} bp genomes (families of 10, seed {SEED}), s={S} k={K} "
                                   f"finch prefilter only, min_ani {MIN_ANI} (BASELINE.json configs[1]"
                                   f"{' scaled by sqrt(G) genomes' if world > 1 else ''})",
                       "pairs_per_step": pairs, "mode": args.mode,
                       "l2": "flushed between timed iterations (256 MiB write)",
                       "sharding": "boustrophedon row blocks of 128; inside the step: 8-byte all-reduce (largest hash), "
                                   "NCCL all-gather of the sketch table overlapped with the per-rank build of its own "
                                   "block lists, NCCL all-gather of the lists, join of the rank's shard"
                                   if world > 1 else "single GPU"},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": "pairs/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "passing_pairs": int(n_pass_t.item()), "ms": 1e3 * float(e2e_t.item()),
                    "pipeline": f"upload in {gb.prefilter_stream_chunks()} slices on a copy stream, block lists of a slice "
                                "built as it lands, join after the last; survivors land in mapped pinned memory and "
                                "their f64 finish runs on the host while the kernels run" if world == 1 else "per-rank slice upload, NVLink all-gathers",
                    "host_ms": e2e_host, "single_upload": e2e_single_upload},
            "gpu_launches": launches,
            "host_enqueue_ms_per_step": float(np.median(host_enqueue_ms[: args.steps])) if host_enqueue_ms else None,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": ncu_traffic(kernel_names[args.mode], n, world),
                         "traffic_unit": "bytes per launch (ncu dram read + write, profiles/ncu_traffic.json)",
                         "kernel": kernel_names[args.mode],
                         "kernel_ms": main_ms, "build_kernels_ms": build_ms,
                         "algorithmic_bytes_per_pair": BYTES_PER_PAIR, "peak_source": peak_src,
                         "kernel_own_bytes_per_pair": own_bytes_per_pair,
                         "kernel_own_gbs": own_bytes_per_pair * pairs_per_launch / (main_ms * 1e-3) / 1e9,
                         "note": "achieved uses the reference algorithm's 2*s*8 B/pair (SURVEY.md 8d). The join "
                                 "kernel computes every pair's exact intersection from ONE merge of two 128-sketch "
                                 "block lists per 128x128 pairs, so it reads 2*s*4/128 B/pair (kernel_own_*) and the "
                                 "fraction of the 16 kB/pair roofline exceeds 1 by design (DESIGN.md K2)"
                                 if args.mode == 0 else
                                 "pairwise kernel re-uses staged sketches from shared memory (16 kB/pair is read "
                                 "from shared memory, not HBM)"},
            "other_mode": other,
            "two_stage": two_stage,
            "ingest": ingest,
            "cpu_baseline": cpu,
            "candidates": n_cand,
            "sketch": {"genomes_per_s": n_local / (sketch_ms * 1e-3) if sketch_ms else None,
                       "gbases_per_s": n_local * L / (sketch_ms * 1e-3) / 1e9 if sketch_ms else None,
                       "ms": sketch_ms, "synth_ms": synth_ms, "genomes_per_rank": n_local},
        }
        sys.stdout.flush()
        os.write(json_fd, (json.dumps(line) + "\n").encode())
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
