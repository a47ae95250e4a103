#!/usr/bin/env python3
"""Compact JSON summary of `ncu --set full` reports (one kernel launch each) for profiles/.
usage: tools/ncu_summary.py out.json name=report.ncu-rep [name=report.ncu-rep ...]
Byte-valued metrics are converted to bytes using the unit row of ncu's raw page."""
import csv, json, subprocess, sys
KEEP = {
    "gpu__time_duration.sum": "duration",
    "launch__registers_per_thread": "registers_per_thread",
    "launch__grid_size": "grid", "launch__block_size": "block",
    "launch__occupancy_limit_registers": "blocks_per_sm_by_registers",
    "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed": "fmaheavy_pipe_busy_pct",
    "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_elapsed": "alu_pipe_busy_pct",
    "sm__issue_active.avg.pct_of_peak_sustained_elapsed": "issue_slots_used_pct",
    "sm__inst_executed.sum": "warp_instructions",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio": "stall_wait",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio": "stall_math_pipe_throttle",
    "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio": "stall_dispatch",
    "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio": "stall_no_instruction",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio": "stall_long_scoreboard",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio": "stall_not_selected",
    "dram__bytes_read.sum": "dram_read_bytes", "dram__bytes_write.sum": "dram_write_bytes",
    "lts__t_sector_hit_rate.pct": "l2_hit_pct", "l1tex__t_sector_hit_rate.pct": "l1_hit_pct",
}
SCALE = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12, "ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}
out = {}
for arg in sys.argv[2:]:
    name, rep = arg.split("=", 1)
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    hdr, units, val = rows[0], rows[1], rows[-1]
    d = {"kernel": val[hdr.index("Kernel Name")].split("(")[0].replace("<unnamed>::", "")}
    for h, u, v in zip(hdr, units, val):
        if h in KEEP and v != "":
            x = float(v.replace(",", ""))
            if u in SCALE: x *= SCALE[u]
            key = KEEP[h] + ("_ms" if KEEP[h] == "duration" else "")
            d[key] = round(x, 4) if x < 1e6 else int(x)
    out[name] = d
json.dump(out, open(sys.argv[1], "w"), indent=1)
print(json.dumps(out, indent=1))
