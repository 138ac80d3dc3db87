#!/usr/bin/env python
"""Summarises an .ncu-rep (ncu --set full) into the text committed under profiles/ and, with
--traffic-json, updates profiles/ncu_traffic.json (dram bytes per launch, read by bench.py).

    python tools/ncu_summary.py gpurun_out/x.ncu-rep "workload text" > profiles/r2_ncu_full_x.txt
"""
import csv
import io
import json
import os
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct", "sm__inst_executed.sum",
    "sm__inst_executed.sum.per_cycle_elapsed", "smsp__thread_inst_executed.sum",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes_read.sum.per_second", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "l1tex__data_bank_conflicts_pipe_lsu.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smsp__average_warp_latency_issue_stalled",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
]


def main():
    rep, workload = sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else ""
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    print(f"# {os.path.basename(rep)}: ncu --set full --clock-control none, one launch per row")
    print(f"# workload: {workload}")
    traffic = {}
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        name = d.get("Kernel Name", "?")
        print(f"\n== {name}")
        for h, u, v in zip(hdr, units, r):
            if any(h == k or h.startswith(k) for k in KEYS) and v not in ("", "n/a"):
                print(f"  {h} [{u}] = {v}")
        try:
            scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
            ur = units[hdr.index("dram__bytes_read.sum")]
            uw = units[hdr.index("dram__bytes_write.sum")]
            traffic[name.split("(")[0].split("<")[0]] = {
                "dram_bytes_read": int(float(d["dram__bytes_read.sum"]) * scale.get(ur, 1)),
                "dram_bytes_write": int(float(d["dram__bytes_write.sum"]) * scale.get(uw, 1)), "workload": workload}
        except (ValueError, KeyError):
            pass
    if "--traffic-json" in sys.argv:
        p = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "ncu_traffic.json")
        cur = json.load(open(p)) if os.path.exists(p) else {}
        cur.update(traffic)
        json.dump(cur, open(p, "w"), indent=1)


if __name__ == "__main__":
    main()
