#!/usr/bin/env python
"""Summarises an .ncu-rep (one `ncu --set full` capture) into the handful of counters DESIGN.md argues from.
usage: python scripts/ncu_summary.py gpurun_out/x.ncu-rep [--facts WORKLOAD RATINGS_PER_LAUNCH SOURCE.md] > profiles/x.md
--facts also records the per-launch counters bench.py's roofline object quotes (DRAM bytes, warp instructions, L2 reduction
sectors, lts throughput) of the FIRST kernel of the report under profiles/traffic_r2.json[WORKLOAD]."""
import csv
import json
import os
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sectors_srcunit_tex_op_read.sum", "lts__t_sectors_srcunit_tex_op_write.sum", "lts__t_sectors_srcunit_tex_op_red.sum",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tensor.sum",
    "sm__pipe_tensor_op_hmma_cycles_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "sm__inst_executed_pipe_fp64.sum", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_membar_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_sleeping_per_issue_active.ratio",
]


def num(x):
    return float(str(x).replace(",", ""))


def main():
    rep = sys.argv[1]
    facts = sys.argv[3:6] if len(sys.argv) >= 6 and sys.argv[2] == "--facts" else None
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(l for l in out.splitlines() if not l.startswith("==")))
    hdr, units = rows[0], rows[1]
    print("# ncu --set full summary of `%s`\n" % rep.split("/")[-1])
    for vals in rows[2:]:
        d = dict(zip(hdr, zip(vals, units)))
        print("## %s  (launch id %s)\n" % (d["Kernel Name"][0], d.get("ID", ("?",))[0]))
        print("| metric | value | unit |\n|---|---:|---|")
        for k in KEYS:
            if k in d and d[k][0] not in ("", "n/a"):
                print("| `%s` | %s | %s |" % (k, d[k][0], d[k][1]))
        print()
        if facts is not None:
            workload, n_ratings, source = facts[0], int(facts[1]), facts[2]
            path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "traffic_r2.json")
            allf = json.load(open(path)) if os.path.exists(path) else {}
            allf[workload] = {
                "kernel": d["Kernel Name"][0], "grid": d["launch__grid_size"][0], "registers": d["launch__registers_per_thread"][0],
                "duration_ms_under_ncu": num(d["gpu__time_duration.sum"][0]) * ({"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(d["gpu__time_duration.sum"][1], 1e-6)),
                "dram_bytes_per_launch": num(d["dram__bytes_read.sum"][0]) * _unit(d["dram__bytes_read.sum"][1]) + num(d["dram__bytes_write.sum"][0]) * _unit(d["dram__bytes_write.sum"][1]),
                "warp_inst_per_launch": num(d["smsp__inst_executed.sum"][0]),
                "l2_red_sectors_per_launch": num(d["lts__t_sectors_srcunit_tex_op_red.sum"][0]),
                "lts_throughput_pct": num(d["lts__throughput.avg.pct_of_peak_sustained_elapsed"][0]),
                "sm_throughput_pct": num(d["sm__throughput.avg.pct_of_peak_sustained_elapsed"][0]),
                "ratings_per_launch": n_ratings, "source": source}
            json.dump(allf, open(path, "w"), indent=1)
            facts = None


def _unit(u):
    return {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}.get(u, 1.0)


if __name__ == "__main__":
    main()
