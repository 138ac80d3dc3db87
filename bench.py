#!/usr/bin/env python
"""bench.py -- genome-pairs/sec (prefilter+ANI) of Galah's two-stage hot path on B200 (BASELINE.json).

A "step" is ONE pass of the whole hot path (reference src/clusterer.rs:14-152 driving
src/finch.rs:48-97 and src/skani.rs:689-788) over N synthetic genomes (families of
10, 2 Mbp, seed 1): K1 MinHash sketches + K3 seed index of every genome, K2 all-pairs prefilter
(finch, min_ani 0.9), K3 ANI (skani restatement, 95 %, min-AF 15) of the prefilter hits the greedy
selection asks about -- exactly the (representative, genome) pairs the reference's two passes
evaluate, a batch per wave of the engine --, greedy representative selection.  At --gpus 1 the workload is BASELINE.json configs[2]: 50,000 genomes.
Genome synthesis is setup (untimed); everything the reference does inside cluster() is timed.

  value    = N(N-1)/2 pairs / step time, packed genomes resident in HBM when the step starts
             (galah_b200_cluster_packed_device; CUDA events on the library's stream)
  e2e      = the same through galah_b200_cluster_packed with HOST buffers: every step uploads the
             packed genomes (pinned, batches behind the previous batch's kernels) and reads the
             survivors / accumulators / clusters back
  roofline = the step's dominant kernel family against the measured HBM peak; `kernels` lists
             every family with its algorithmic bytes, device time and share of the step
  cpu_baseline = the oracle port of the same path on a bounded sample, host cores

At G > 1 GPUs (one process per GPU) the genome count grows as N * sqrt(G) (constant pair area per
GPU, "weak"): every rank sketches and indexes its own genome slice, K2 is row-block sharded, K3
pairs go to the rank that owns the query genome and read the reference genome's table through
peer-mapped memory over NVLink; the greedy engine runs on rank 0.

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
MIN_ANI = 0.9          # FinchPreclusterer min_ani (fraction), galah --precluster-ani 90
ANI_PCT = 95.0         # SkaniClusterer threshold, galah --ani 95
MIN_AF = 15.0          # galah --min-aligned-fraction 15
SEED = 1
METRIC = "genome-pairs/sec (prefilter+ANI)"


def n_genomes_for(gpus, base):
    if gpus == 1:
        return base
    # per rank a whole number of families of 10 AND of row blocks of 128: multiples of 640 per rank
    q = 640 * gpus
    return max(q, int(round(base * math.sqrt(gpus) / q)) * q)


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def ncu_traffic(kernel):
    """dram bytes per launch of `kernel` from the committed ncu capture (profiles/ncu_traffic.json)."""
    p = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if not os.path.exists(p):
        return None
    with open(p) as f:
        t = json.load(f)
    if kernel not in t:
        return None
    out = {"bytes_per_launch": t[kernel]["dram_bytes_read"] + t[kernel]["dram_bytes_write"],
           "captured_at": t[kernel].get("workload", t.get("workload"))}
    if t[kernel].get("units_in_capture"):
        out["bytes_per_unit"] = out["bytes_per_launch"] / t[kernel]["units_in_capture"]
    return out


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
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "250",
                 "-i", str(self.device)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def mark(self):
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


# ------------------------------------------------------------------------------------------------
# CPU restatement of the reference path (oracle/), on a bounded sample of the workload
# ------------------------------------------------------------------------------------------------
def cpu_path_sample(n_total, L, n_sketch, pair_target_serial, pair_target_mt, n_ani, table=None):
    """Times the three parts of the reference's cluster() on samples of the synthetic workload and
    combines them into the metric for the FULL workload of n_total genomes:
      sketching   finch::sketch_files, all host threads (src/finch.rs:55-69)      -> s per genome
      pair loop   the reference's SERIAL nested loop (src/finch.rs:75-95)          -> s per pair
      ANI         skani per prefilter hit: both genomes re-sketched per pair, as the reference's
                  one-subprocess-per-pair does (src/skani.rs:718-788), all host threads via the
                  reference's rayon find_any (src/clusterer.rs:262-296)           -> s per hit
    value = P / (n_total * t_sketch + P * t_pair + hits * t_ani),  P = n_total (n_total - 1) / 2,
    hits = 4.5 prefilter hits per family of 10 as measured on the GPU arm (passed in by the caller
    through n_ani's companion `hits_total`, else estimated from the sample)."""
    import oracle
    oracle.build()
    cores = os.cpu_count() or 1
    t0 = time.perf_counter()
    tab, cnt = oracle.sketch_synth(SEED, 0, n_sketch, L, K, S, 0)
    t_sketch = (time.perf_counter() - t0) / n_sketch            # wall per genome with all threads
    if table is not None:
        tab, cnt = table
    n_tab = len(cnt)
    rows_serial = max(1, min(n_tab - 1, int(math.ceil(pair_target_serial / max(1, n_tab - 1)))))
    pairs_serial = sum(n_tab - 1 - i for i in range(rows_serial))
    t0 = time.perf_counter()
    hits = oracle.prefilter(tab, cnt, K, MIN_ANI, row_begin=0, row_end=rows_serial)
    t_pair = (time.perf_counter() - t0) / pairs_serial
    rows_mt, acc = 0, 0
    while rows_mt < n_tab - 1 and acc < pair_target_mt:
        acc += n_tab - 1 - rows_mt
        rows_mt += 1
    t0 = time.perf_counter()
    n_hits_mt = oracle.prefilter_count_mt(tab, cnt, K, MIN_ANI, row_begin=0, row_end=rows_mt)
    t_pair_mt = (time.perf_counter() - t0) / max(acc, 1)
    # ANI: hit pairs of the sample, each with both genomes regenerated + seeded (nothing cached)
    sample = [(int(h["i"]), int(h["j"])) for h in hits[:n_ani]]
    t0 = time.perf_counter()
    anis = []
    for a, b in sample:
        ga = oracle.AniGenome(*oracle.codes_from_ascii(oracle.synth_genome(SEED, a, L)))
        gb_ = oracle.AniGenome(*oracle.codes_from_ascii(oracle.synth_genome(SEED, b, L)))
        anis.append(float(oracle.ani_pair(ga, gb_, MIN_AF)[0]))
    t_ani_serial = (time.perf_counter() - t0) / max(len(sample), 1)
    t_ani = t_ani_serial / cores                                  # rayon over pairs in the reference
    return {"t_sketch_per_genome_s": t_sketch, "t_pair_serial_s": t_pair, "t_pair_all_cores_s": t_pair_mt,
            "t_ani_per_hit_serial_s": t_ani_serial, "t_ani_per_hit_all_cores_s": t_ani, "cores": cores,
            "n_sketch": n_sketch, "pairs_serial": pairs_serial, "pairs_all_cores": acc, "n_ani": len(sample),
            "hits_in_serial_rows": int(len(hits)), "hits_in_mt_rows": int(n_hits_mt), "table_genomes": n_tab,
            "sample_anis": anis, "sample_pairs": sample}


def cpu_value(parts, n_total, hits_total):
    P = n_total * (n_total - 1) / 2
    t = n_total * parts["t_sketch_per_genome_s"] + P * parts["t_pair_serial_s"] + hits_total * parts["t_ani_per_hit_all_cores_s"]
    t_mt = n_total * parts["t_sketch_per_genome_s"] + P * parts["t_pair_all_cores_s"] + hits_total * parts["t_ani_per_hit_all_cores_s"]
    return P / t, P / t_mt


def sample_text(parts, n_total):
    return (f"sketch of {parts['n_sketch']} genomes on {parts['cores']} threads; pair loop over rows of a "
            f"{parts['table_genomes']}-genome sketch table of the same synthetic families: {parts['pairs_serial']} pairs "
            f"serial (as src/finch.rs:75-95) + {parts['pairs_all_cores']} pairs on all threads; ANI of {parts['n_ani']} "
            f"prefilter hits, both genomes re-seeded per pair (as skani per subprocess); combined for {n_total} genomes")


def run_reference(args):
    """`--impl reference`: the oracle port of the reference path on the box's host cores.  Each step
    is a bounded sample (>= 1e7 genome pairs through the pair loop, all threads where the
    reference has threads); the headline keeps the reference's SERIAL pair loop."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import oracle
    oracle.build()
    n_total = n_genomes_for(args.gpus, args.n_genomes)
    L = args.genome_len
    cores = os.cpu_count() or 1
    t0 = time.perf_counter()
    n_tab = args.ref_table_genomes
    table = oracle.sketch_synth(SEED, 0, n_tab, L, K, S, 0)   # setup of the pair-loop table (untimed)
    t_setup = time.perf_counter() - t0
    hits_total = 4.5 * n_total / 10 * 1.0                        # 45 within-family pairs / 10 genomes pass 0.9
    times, vals, vals_mt, last = [], [], [], None
    for it in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        parts = cpu_path_sample(n_total, L, n_sketch=2 * cores, pair_target_serial=2.0e5, pair_target_mt=1.0e7,
                                n_ani=4, table=table)
        dt = time.perf_counter() - t0
        if it >= args.warmup:
            v, v_mt = cpu_value(parts, n_total, hits_total)
            times.append(dt); vals.append(v); vals_mt.append(v_mt); last = parts
    value = float(np.mean(vals))
    sample = sample_text(last, n_total) + f"; table sketched once in {t_setup:.1f}s (setup)"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "pairs/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * float(np.mean(times)),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u64", "data": "synthetic",
        "config": {"workload": workload_text(n_total, L, args.gpus), "timed_sample": sample},
        "cpu_baseline": {"value": value, "unit": "pairs/s", "cores": cores, "kind": "port", "sample": sample,
                         "all_cores_pair_loop_value": float(np.mean(vals_mt)),
                         "parts": {k: v for k, v in last.items() if not k.startswith("sample_")}},
        "e2e": {"value": value, "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def workload_text(n, L, gpus):
    return (f"{n} synthetic {L} bp genomes (families of 10, seed {SEED}): finch prefilter s={S} k={K} min_ani {MIN_ANI} "
            f"+ skani-restatement ANI {ANI_PCT} %, min-AF {MIN_AF} %, greedy clustering (BASELINE.json configs[2]"
            f"{', genomes scaled by sqrt(G)' if gpus > 1 else ''})")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--n-genomes", type=int, default=50000, help="genomes at 1 GPU (x sqrt(G) at G GPUs)")
    ap.add_argument("--genome-len", type=int, default=2_000_000)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-prefilter-record", action="store_true", help="skip the prefilter-only (K2) sub-record")
    ap.add_argument("--no-dense", action="store_true", help="skip the dense-clade sub-record")
    ap.add_argument("--no-ingest", action="store_true", help="skip the FASTA-ingest (K0) sub-record")
    ap.add_argument("--dense-genomes", type=int, default=2000)
    ap.add_argument("--ref-table-genomes", type=int, default=4500, help="reference arm: genomes of its sketch table (4500 -> 1.01e7 pairs per step)")
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
        if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
            os.environ["NCCL_DEBUG"] = "WARN"
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dev = torch.device("cuda", local_rank)
    lib_stream = torch.cuda.ExternalStream(gb.stream(), device=dev)  # the stream the library launches on

    n = n_genomes_for(world, args.n_genomes)
    n_local = n // world
    L = args.genome_len
    pairs = n * (n - 1) // 2

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    # ---------------- setup (untimed): this rank's synthetic genomes, packed, resident in HBM
    lay = gb.synth_layout(n_local, L)
    d_seq = torch.empty(lay["seq2_words"], dtype=torch.int32, device=dev)
    d_val = torch.empty(lay["valid_words"], dtype=torch.int32, device=dev)
    d_off = torch.empty(n_local + 1, dtype=torch.int64, device=dev)
    t0 = time.perf_counter()
    gb.synth_packed_device(SEED, rank * n_local, n_local, L, d_seq.data_ptr(), d_val.data_ptr(), d_off.data_ptr(),
                           torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    synth_s = time.perf_counter() - t0
    base_off = np.arange(n_local + 1, dtype=np.uint64) * np.uint64(lay["padded"])
    lengths = np.full(n_local, L, np.uint64)
    resident_bytes = d_seq.numel() * 4 + d_val.numel() * 4

    pipe = None
    if world > 1:
        from galah_b200.distributed import ShardedPipeline
        pipe = ShardedPipeline(gb, dist, n_local, S, dev)

    def step():
        if world > 1:
            return pipe.step_device(d_seq.data_ptr(), d_val.data_ptr(), d_off.data_ptr(), base_off, lengths,
                                    MIN_ANI, ANI_PCT, MIN_AF)
        return gb.cluster_packed(d_seq.data_ptr(), d_val.data_ptr(), base_off, lengths, precluster_ani=MIN_ANI,
                                 ani=ANI_PCT, min_aligned_fraction=MIN_AF, device=True, d_base_off=d_off.data_ptr())

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    for _ in range(args.warmup):
        clusters, info = step()
    sync_all()
    if rank == 0:
        sampler.mark()
    launches0 = gb.launch_count()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    wall, infos = [], []
    for it in range(args.steps):
        sync_all()
        ev[it][0].record(lib_stream)
        t0 = time.perf_counter()
        clusters, info = step()
        ev[it][1].record(lib_stream)
        torch.cuda.synchronize()
        wall.append(time.perf_counter() - t0)
        infos.append(info)
    sync_all()
    launches = gb.launch_count() - launches0
    step_ms = [a.elapsed_time(b) for a, b in ev]
    tot = torch.tensor([sum(step_ms)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tot, op=dist.ReduceOp.MAX)
    total_ms = float(tot.item())
    clocks = sampler.stop() if rank == 0 else None
    value = pairs * args.steps / (total_ms * 1e-3)
    phase = {k: float(np.median([i[k] for i in infos])) for k in
             ("sketch_ms", "index_ms", "prefilter_ms", "ani_ms", "ani_chain_ms", "engine_ms", "ingest_ms", "total_ms")
             if k in infos[0]}
    n_hits = int(infos[-1]["n_precluster_hits"])
    n_clusters = len(clusters) if clusters is not None else None

    # ---------------- e2e: HOST buffers through the C ABI, every step uploads the packed genomes
    e2e = None
    if not args.no_e2e:
        import psutil
        need = resident_bytes
        vm = psutil.virtual_memory()
        avail = vm.available
        # every rank pins a host copy of ITS packed genomes: all of them together must stay well inside
        # the box's memory (a box driven out of memory is worse than a missing sub-record)
        if need * world <= 0.45 * vm.total and avail > 1.3 * need * world + (16 << 30):
            h_seq = torch.empty(d_seq.numel(), dtype=torch.int32, pin_memory=True)
            h_val = torch.empty(d_val.numel(), dtype=torch.int32, pin_memory=True)
            h_seq.copy_(d_seq); h_val.copy_(d_val)
            torch.cuda.synchronize()

            # the host-side input of the headline e2e: the 2-bit packed sequence plus the list of invalid base
            # ranges (N runs, record breaks: none in the synthetic genomes); the validity bitmap -- a third of
            # the packed bytes, almost constant -- is built on the device.  The call that uploads the bitmap
            # too is timed beside it (with_validity_bitmap).
            no_ranges = (np.zeros(0, np.uint64), np.zeros(0, np.uint64))

            def e2e_step(bitmap):
                if world > 1:
                    return pipe.step_host(h_seq.data_ptr(), h_val.data_ptr() if bitmap else no_ranges, base_off, lengths,
                                          MIN_ANI, ANI_PCT, MIN_AF)
                if bitmap:
                    return gb.cluster_packed(h_seq.data_ptr(), h_val.data_ptr(), base_off, lengths, precluster_ani=MIN_ANI,
                                             ani=ANI_PCT, min_aligned_fraction=MIN_AF, device=False)
                return gb.cluster_packed_sparse(h_seq.data_ptr(), no_ranges, base_off, lengths, precluster_ani=MIN_ANI,
                                                ani=ANI_PCT, min_aligned_fraction=MIN_AF)

            def time_e2e(bitmap, reps):
                ts, info_, cl_ = [], None, None
                for it in range(1 + reps):
                    sync_all()
                    t0 = time.perf_counter()
                    cl_, info_ = e2e_step(bitmap)
                    dt = time.perf_counter() - t0
                    if it >= 1:
                        ts.append(dt)
                t_ = torch.tensor([float(np.mean(ts))], dtype=torch.float64, device=dev)
                if world > 1:
                    dist.all_reduce(t_, op=dist.ReduceOp.MAX)
                return float(t_.item()), info_, cl_
            t_bitmap, _, cl_bitmap = time_e2e(True, min(args.steps, 2))
            t_e2e, e_info, e_clusters = time_e2e(False, min(args.steps, 3))
            e_t = torch.tensor([t_e2e], dtype=torch.float64, device=dev)
            same = (e_clusters == clusters and cl_bitmap == clusters) if rank == 0 else None
            seq_bytes = d_seq.numel() * 4
            h2d = world * (seq_bytes + (n_local + 1) * 8 + n_local * 8)
            # survivors (16 B each), the K3 accumulators of every evaluated pair (7 x u32 + one u64), counts
            d2h = n_hits * 16 + int(e_info.get("my_ani_pairs") or e_info.get("n_ani_pairs") or n_hits) * 36 * (world if world > 1 else 1) + n * 4 + 16
            e2e = {"value": pairs / float(e_t.item()), "unit": "pairs/s", "h2d_bytes_per_step": int(h2d),
                   "d2h_bytes_per_step": int(d2h), "ms": 1e3 * float(e_t.item()),
                   "phases_ms": {k: e_info[k] for k in ("ingest_ms", "sketch_ms", "index_ms", "prefilter_ms", "ani_ms", "engine_ms") if k in e_info},
                   "input": "HOST buffers: 2-bit packed sequence + invalid base ranges (none in the synthetic genomes); the "
                            "validity bitmap is built on the device (galah_b200_cluster_packed_sparse)",
                   "with_validity_bitmap": {"value": pairs / t_bitmap, "ms": 1e3 * t_bitmap,
                                            "h2d_bytes_per_step": int(world * (need + (n_local + 1) * 8)),
                                            "call": "galah_b200_cluster_packed (sequence + validity bitmap uploaded)"},
                   "upload_overlap": "batches of ~1 G bases cross PCIe on a copy stream behind the K1 / K3-index kernels of the previous batch",
                   "pcie_gbs_if_serial": seq_bytes / 1e9 / float(e_t.item()),
                   "clusters_identical_to_resident_run": same}
            del h_seq, h_val
        else:
            e2e = {"value": None, "unit": "pairs/s", "h2d_bytes_per_step": int(need), "d2h_bytes_per_step": 0,
                   "skipped": f"host has {avail >> 30} GiB available of {vm.total >> 30}; the pinned copies of the packed genomes "
                              f"need {need * world >> 30} GiB on this box (limit: 45 % of its memory)"}

    line = None
    if rank == 0:
        peak, peak_src = peaks()
        sub = {}
        if world == 1:
            sub = single_gpu_records(args, gb, torch, dev, d_seq, d_val, d_off, base_off, lengths, n, L, clusters)
        # ---- roofline: one entry per kernel family, algorithmic bytes per unit from SURVEY.md 8d
        seeds_per_genome = L / 125.0
        fam = []
        def add(name, kernels, ms, units, unit, bytes_per_unit, bound, note):
            if ms is None or ms <= 0:
                return
            gbs = bytes_per_unit * units / (ms * 1e-3) / 1e9
            fam.append({"family": name, "kernels": kernels, "ms_per_step": ms, "share_of_step": ms / (total_ms / args.steps),
                        "units_per_step": units, "unit": unit, "algorithmic_bytes_per_unit": bytes_per_unit,
                        "achieved_gbs": gbs, "frac_of_hbm_peak": gbs / peak, "bound": bound, "note": note})
        add("K1 sketch", ["scan21v2_kernel", "sketch_select_kernel"], phase.get("sketch_ms"), n_local, "genome",
            L / 4 + L / 8 + 8 * S, "integer issue slots: ALU and FMA pipes level (profiles/r2_pipe_rates_b200.txt); "
            "HBM fraction reported as asked, see instruction_roofline",
            "L/4 packed + L/8 validity read, 8 s written per genome; the time is taken while the K3 index kernels of the "
            "previous batch share the SMs (two streams), so K1 + K3 index > ingest")
        if fam and fam[-1]["family"] == "K1 sketch" and clocks and clocks.get("sm_mhz"):
            # what bounds the kernel: issue slots per k-mer position, counted in the SASS of the hot path of
            # scan21v2_kernel<0,1,1> (tools: cuobjdump -sass): 63 ALU-pipe (LOP3 / SHF / PRMT / SEL / ISETP, 0.5 per clock
            # and sub-partition), 46 IMAD + 16 IMAD.WIDE (0.5 / 0.25 per clock: 78 FMA-pipe slots), 12 IADD3, 10 other
            # = 147 instructions = 163 issue clocks (a wide multiply holds its pipe for two slots); the two pipes are level
            issue_clocks_per_kmer, sms, subparts = 163.0, 148, 4
            peak_kmers = sms * subparts * 32 * clocks["sm_mhz"] * 1e6 / issue_clocks_per_kmer
            got_kmers = n_local * float(L) / (fam[-1]["ms_per_step"] * 1e-3)
            fam[-1]["instruction_roofline"] = {
                "bound": "issue slots (FMA pipe 156 clocks, ALU pipe 126 clocks, issue 163 clocks per warp and k-mer position)",
                "alu_pipe_instructions_per_kmer": 63, "imad_per_kmer": 46, "imad_wide_per_kmer": 16, "iadd3_per_kmer": 12,
                "all_instructions_per_kmer": 147, "issue_clocks_per_kmer": issue_clocks_per_kmer,
                "peak_kmers_per_s": peak_kmers, "achieved_kmers_per_s": got_kmers, "frac": got_kmers / peak_kmers,
                "source": "SASS count of the loop + measured pipe rates (tools/pipe_bench.cu, profiles/r2_pipe_rates_b200.txt); "
                          "the time also holds the K3 index kernels of the previous batch, which share the SMs"}
        add("K3 index", ["ani_count_kernel", "ani_emit_kernel"], phase.get("index_ms"), n_local, "genome",
            L / 4 + L / 8 + seeds_per_genome * (8 + 16), "integer ALU (mm_hash64 per k-mer) + scattered table inserts",
            "packed read + 8 B seed + 2 x 8 B table slots per seed written")
        k2 = sub.get("prefilter_only") or {}
        add("K2 prefilter", ["bl_* build", "prefilter_join_kernel"], k2.get("ms_per_step"), pairs, "pair", 2 * S * 8,
            "issue slots / shared-memory wavefronts (block-list join); the 16 kB/pair figure is the reference "
            "algorithm's traffic, which the join does not move: frac_of_hbm_peak > 1 is a REUSE factor, see own_*",
            "own traffic: 2*s*4/128 = 62.5 B of L2-resident list keys per pair")
        if fam and fam[-1]["family"] == "K2 prefilter":
            own = 2 * S * 4 / gb.ROW_BLOCK
            fam[-1]["own_bytes_per_unit"] = own
            fam[-1]["own_gbs"] = own * pairs / (k2["join_kernel_ms"] * 1e-3) / 1e9 if k2.get("join_kernel_ms") else None
            fam[-1]["own_frac_of_hbm_peak"] = fam[-1]["own_gbs"] / peak if fam[-1]["own_gbs"] else None
            fam[-1]["dram_traffic"] = ncu_traffic("prefilter_join_kernel")
        n_eval = int(infos[-1].get("my_ani_pairs") or infos[-1].get("n_ani_pairs") or n_hits)
        add("K3 chain", ["ani_chain_kernel"], phase.get("ani_chain_ms"), n_eval, "pair evaluation",
            2 * seeds_per_genome * 12, "latency of table probes (L2 / HBM) + SIMT divergence of the chaining step",
            "2 * (L/c) * 12 B seed entries per pair (SURVEY.md 8d); every K3 launch of a step: one per wave of the greedy "
            "engine, only the (representative, genome) pairs the reference's two passes evaluate (sharded runs: this rank's share)")
        if fam and fam[-1]["family"] == "K3 chain":
            fam[-1]["ani_waves"] = int(infos[-1].get("ani_waves", 0))
            fam[-1]["prefilter_hits"] = n_hits
        dom = max(fam, key=lambda f: f["ms_per_step"]) if fam else None
        roofline = None
        dom_traffic = None
        if dom:
            # measured DRAM bytes of the dominant kernel, scaled to the units this line's `achieved` covers
            tr = ncu_traffic(dom["kernels"][0]) or {}
            dom_traffic = tr["bytes_per_unit"] * dom["units_per_step"] if tr.get("bytes_per_unit") else tr.get("bytes_per_launch")
            dom["dram_traffic"] = tr or None
            roofline = {"bound": "hbm", "achieved": dom["achieved_gbs"], "peak": peak, "unit": "GB/s",
                        "frac": dom["achieved_gbs"] / peak, "traffic": dom_traffic,
                        "kernel": dom["family"] + " (" + ", ".join(dom["kernels"]) + ")", "kernel_ms": dom["ms_per_step"],
                        "share_of_step": dom["share_of_step"], "algorithmic_bytes_per_unit": dom["algorithmic_bytes_per_unit"],
                        "units_per_launch": dom["units_per_step"], "peak_source": peak_src,
                        "note": "dominant kernel family of the step by device time; " + dom["bound"],
                        "kernels": fam}
        cpu = None
        if not args.no_cpu_baseline and world == 1:
            import oracle
            parts = cpu_path_sample(n, L, n_sketch=2 * (os.cpu_count() or 1), pair_target_serial=2.0e5,
                                    pair_target_mt=2.0e6, n_ani=4, table=sub.get("_table_sample"))
            v, v_mt = cpu_value(parts, n, n_hits)
            ok = None
            if sub.get("_ani_lookup") is not None:
                look = sub["_ani_lookup"]
                ok = all(np.float32(look.get(p, -1.0)) == np.float32(a) for p, a in zip(parts["sample_pairs"], parts["sample_anis"]))
            cpu = {"value": v, "unit": "pairs/s", "cores": parts["cores"], "kind": "port", "sample": sample_text(parts, n),
                   "all_cores_pair_loop_value": v_mt, "gpu_matches_oracle_on_sample": ok,
                   "parts": {k: v2 for k, v2 in parts.items() if not k.startswith("sample_")}}
        sub = {k: v for k, v in sub.items() if not k.startswith("_")}
        line = {
            "metric": METRIC, "value": value, "unit": "pairs/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u64", "data": "synthetic",
            "config": {"workload": workload_text(n, L, world), "pairs_per_step": pairs, "genomes": n,
                       "l2": f"inputs larger than L2 ({resident_bytes >> 20} MiB of packed sequence per GPU, read once per step)",
                       "timed_region": "K1 sketch + K3 index of every genome, K2 all pairs, " +
                                       "K3 ANI of the (representative, genome) hit pairs the reference's greedy passes evaluate, "
                                       "asked for in waves by the engine" + (", " if world == 1 else " (replicated on every rank; a request "
                                       "runs on the owner of its query, a wave's values are all-gathered), ") +
                                       "f64 finish + greedy engine on the host; genome synthesis is setup",
                       "sharding": "one process per GPU: genome slices for K1 / K3 index, boustrophedon row blocks for K2 "
                                   "(NCCL all-gather of the sketch table), K3 pairs on the query's rank reading the "
                                   "reference table through peer-mapped memory (NVLink), engine on rank 0"
                                   if world > 1 else "single GPU"},
            "clocks": clocks,
            "e2e": e2e,
            "gpu_launches": launches,
            "phases_ms": phase,
            "result": {"prefilter_hits": n_hits, "clusters": n_clusters, "host_wall_ms_per_step": 1e3 * float(np.mean(wall)),
                       "synth_setup_s": synth_s},
            "roofline": roofline,
            "cpu_baseline": cpu,
        }
        if infos and isinstance(infos[-1].get("host_detail_ms"), dict):
            line["host_detail_ms_rank0"] = infos[-1]["host_detail_ms"]  # multi-GPU: the host side of the ANI / engine phases
        line.update(sub)
        sys.stdout.flush()
        os.write(json_fd, (json.dumps(line) + "\n").encode())
    if world > 1:
        dist.destroy_process_group()


def single_gpu_records(args, gb, torch, dev, d_seq, d_val, d_off, base_off, lengths, n, L, clusters):
    """Sub-records of the 1-GPU line: prefilter only (K2, the configs[1] kernel at this N), a dense
    clade, FASTA ingest.  None of them is inside the headline's timed region."""
    out = {}
    st = torch.cuda.current_stream().cuda_stream
    stream = torch.cuda.current_stream()
    flush = torch.empty(256 << 20, dtype=torch.int8, device=dev)
    table = torch.empty((n, S), dtype=torch.int64, device=dev)
    counts = torch.empty(n, dtype=torch.int32, device=dev)
    gb.sketch_packed_device(d_seq.data_ptr(), d_val.data_ptr(), d_off.data_ptr(), n, K, S, 0, table.data_ptr(),
                            counts.data_ptr(), st)
    torch.cuda.synchronize()
    n_tab = min(n, args.ref_table_genomes)
    out["_table_sample"] = (table[:n_tab].cpu().numpy().view(np.uint64), counts[:n_tab].cpu().numpy().view(np.uint32))

    def time_k2(tab, cnt, nn, mode, steps, warm):
        cand_cap = max(1 << 20, 64 * nn)
        d_cand = torch.empty((cand_cap, 4), dtype=torch.int32, device=dev)
        d_ncand = torch.zeros(1, dtype=torch.int64, device=dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ms, b_ms, m_ms = [], [], []
        for it in range(warm + steps):
            flush.fill_(1)
            torch.cuda.synchronize()
            e0.record(stream)
            gb.prefilter_enqueue(tab.data_ptr(), cnt.data_ptr(), nn, S, K, MIN_ANI, 0, 1, mode, st, d_cand.data_ptr(),
                                 cand_cap, d_ncand.data_ptr())
            e1.record(stream)
            torch.cuda.synchronize()
            if it >= warm:
                ms.append(e0.elapsed_time(e1))
                b, m = gb.prefilter_last_timing()
                b_ms.append(b); m_ms.append(m)
        return float(np.mean(ms)), float(np.mean(b_ms)), float(np.mean(m_ms)), int(d_ncand.item())

    if not args.no_prefilter_record:
        ms, b, m, ncand = time_k2(table, counts, n, 0, 10, 3)
        P = n * (n - 1) // 2
        out["prefilter_only"] = {"workload": f"K2 alone on the resident {n} x {S} sketch table (BASELINE.json configs[1] kernel at this N), L2 flushed",
                                 "value": P / (ms * 1e-3), "unit": "pairs/s", "ms_per_step": ms, "build_kernels_ms": b,
                                 "join_kernel_ms": m, "candidates": ncand}
        # ANI lookup of sampled hits for the CPU check (host-buffer prefilter + K3 on a slice of genomes)
    try:
        idx = gb.AniIndex()
        n_s = min(n, 200)
        idx.add_packed_device(d_seq.data_ptr(), d_val.data_ptr(), d_off.data_ptr(), base_off[: n_s + 1], lengths[:n_s], st)
        prs = np.array([(a, b) for f in range(n_s // 10) for a in range(10 * f, 10 * f + 10) for b in range(a + 1, 10 * f + 10)], np.uint32)
        res = idx.pairs(prs, MIN_AF)
        out["_ani_lookup"] = {(int(a), int(b)): float(v) for (a, b), v in zip(prs, res["ani"])}
        idx.close()
    except Exception as e:  # the check is optional; the headline does not depend on it
        out["_ani_lookup"] = None
        out["ani_lookup_error"] = str(e)

    if not args.no_dense:
        nd = args.dense_genomes
        lay = gb.synth_layout(nd, L)
        s2 = torch.empty(lay["seq2_words"], dtype=torch.int32, device=dev)
        v2 = torch.empty(lay["valid_words"], dtype=torch.int32, device=dev)
        o2 = torch.empty(nd + 1, dtype=torch.int64, device=dev)
        gb.synth_packed_device_ex(SEED + 1, 0, nd, L, nd, 2, s2.data_ptr(), v2.data_ptr(), o2.data_ptr(), st)
        torch.cuda.synchronize()
        bo = np.arange(nd + 1, dtype=np.uint64) * np.uint64(lay["padded"])
        ln = np.full(nd, L, np.uint64)
        t2 = torch.empty((nd, S), dtype=torch.int64, device=dev)
        c2 = torch.empty(nd, dtype=torch.int32, device=dev)
        gb.sketch_packed_device(s2.data_ptr(), v2.data_ptr(), o2.data_ptr(), nd, K, S, 0, t2.data_ptr(), c2.data_ptr(), st)
        torch.cuda.synchronize()
        ms0, b0, m0, nc0 = time_k2(t2, c2, nd, 0, 5, 2)
        ms1, b1, m1, nc1 = time_k2(t2, c2, nd, 1, 3, 1)
        P = nd * (nd - 1) // 2
        def two_stage(mode):
            gb.cluster_lazy(mode)
            try:
                t0 = time.perf_counter()
                c, i = gb.cluster_packed(s2.data_ptr(), v2.data_ptr(), bo, ln, precluster_ani=MIN_ANI, ani=ANI_PCT,
                                         min_aligned_fraction=MIN_AF, device=True, d_base_off=o2.data_ptr())
                return c, i, time.perf_counter() - t0
            finally:
                gb.cluster_lazy(-1)
        two_stage(-1)  # warm-up (workspaces sized for this input)
        cl, info, t_full = two_stage(-1)   # default: stage 2 in waves
        cl_e, info_e, t_eager = two_stage(0)  # K3 on every precluster hit up front
        out["dense"] = {"workload": f"ONE clade: {nd} genomes x {L} bp derived from one founder at 0..2.5 % substitutions "
                                    "(every pair related; every block pair of the join is tie-dense)",
                        "pairs": P, "join_mode_ms": ms0, "join_kernel_ms": m0, "pairwise_mode_ms": ms1,
                        "prefilter_pairs_per_s_join": P / (ms0 * 1e-3), "prefilter_pairs_per_s_pairwise": P / (ms1 * 1e-3),
                        "candidates_join": nc0, "candidates_pairwise": nc1, "modes_agree": nc0 == nc1,
                        "two_stage_s": t_full, "two_stage_pairs_per_s": P / t_full, "prefilter_hits": int(info["n_precluster_hits"]),
                        "phases_ms": {k: info[k] for k in ("sketch_ms", "index_ms", "prefilter_ms", "ani_ms", "engine_ms")},
                        "clusters": len(cl), "stage2": "in waves: only the (representative, genome) pairs the reference's two "
                        "passes evaluate (galah_b200_cluster_lazy, the default)",
                        "ani_pairs_evaluated": int(info["n_ani_pairs"]), "ani_waves": int(info["ani_waves"]),
                        "eager": {"two_stage_s": t_eager, "two_stage_pairs_per_s": P / t_eager,
                                  "ani_pairs_evaluated": int(info_e["n_ani_pairs"]),
                                  "phases_ms": {k: info_e[k] for k in ("sketch_ms", "index_ms", "prefilter_ms", "ani_ms", "engine_ms")},
                                  "clusters_identical": bool(cl == cl_e)}}
        del s2, v2, o2, t2, c2

    if not args.no_ingest:
        import tempfile
        rng = np.random.default_rng(SEED)
        n_files, glen, width = 48, 2_000_000, 80
        acgt = np.frombuffer(b"ACGT", np.uint8)
        files = []
        fam_base = None
        for g in range(n_files):
            if g % 4 == 0:
                fam_base = acgt[rng.integers(0, 4, size=glen)]
                seq = fam_base
            else:
                seq = fam_base.copy()
                hit = rng.uniform(size=glen) < 0.01 * (g % 4)
                seq[hit] = acgt[rng.integers(0, 4, size=int(hit.sum()))]
            seq = seq.reshape(-1, width)
            body = np.concatenate([seq, np.full((seq.shape[0], 1), 10, np.uint8)], axis=1).tobytes()
            files.append(b">genome_%d synthetic\n" % g + body)
        raw_bytes = sum(len(f) for f in files)
        gb.decode_fasta_device(files, unpack=False)
        meta, dec_ms = gb.decode_fasta_device(files, unpack=False)
        with tempfile.TemporaryDirectory(dir="/dev/shm" if os.path.isdir("/dev/shm") else None) as td:
            paths = []
            for g, data in enumerate(files):
                paths.append(os.path.join(td, f"g{g}.fna"))
                with open(paths[-1], "wb") as f:
                    f.write(data)
            gb.cluster(paths, precluster_ani=MIN_ANI, ani=ANI_PCT, min_aligned_fraction=MIN_AF)
            t0 = time.perf_counter()
            cl, info = gb.cluster(paths, precluster_ani=MIN_ANI, ani=ANI_PCT, min_aligned_fraction=MIN_AF)
            t_files = time.perf_counter() - t0
        out["ingest"] = {"workload": f"{n_files} synthetic FASTA files x {glen} bp, {width}-column lines ({raw_bytes} bytes), families of 4",
                         "k0_decode_ms": dec_ms, "k0_decode_gbytes_per_s": raw_bytes / (dec_ms * 1e-3) / 1e9,
                         "cluster_files_s": t_files, "cluster_files_fasta_gbytes_per_s": raw_bytes / t_files / 1e9,
                         "cluster_files_phases_ms": {k: info[k] for k in ("ingest_ms", "prefilter_ms", "ani_ms", "engine_ms", "total_ms")},
                         "clusters": len(cl), "prefilter_hits": int(info["n_precluster_hits"]),
                         "all_bases_ok": all(m["n_bases"] == glen and m["n_ambiguous"] == 0 for m in meta)}
    return out


if __name__ == "__main__":
    main()
