#!/usr/bin/env python3
"""Key metrics of .ncu-rep files (development aid): python tools/ncu_summary.py [--json out.json] file.ncu-rep ..."""
import csv
import json
import subprocess
import sys

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sectors.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__m_xbar2l1tex_read_bytes.sum",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem"]
UNIT_SCALE = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}
args = sys.argv[1:]
jout = None
if args and args[0] == "--json":
    jout, args = args[1], args[2:]
summary = {}
for path in args:
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        print("--", path.split("/")[-1], "|", d.get("Kernel Name", "")[:90])
        for w in WANT:
            if w in d:
                print(f"  {w:84s} {d[w]:>16s} {units[hdr.index(w)]}")
        def val(name):
            u = units[hdr.index(name)]
            return float(d[name].replace(",", "")) * UNIT_SCALE.get(u, 1.0)
        summary[path.split("/")[-1].replace(".ncu-rep", "")] = {
            "kernel": d.get("Kernel Name", "")[:120],
            "duration_us": float(d["gpu__time_duration.sum"].replace(",", "")) * {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}[units[hdr.index("gpu__time_duration.sum")]],
            "dram_bytes": val("dram__bytes_read.sum") + val("dram__bytes_write.sum"),
        }
if jout:
    json.dump(summary, open(jout, "w"), indent=1)
