#!/usr/bin/env python3
"""Where a kernel's time goes by code region, from the source page of an ncu report: splits the SASS into
regions at the given hex addresses and prints, per region, executed warp-instructions, stall samples and the
share of no-instruction (I-cache) / long-scoreboard samples.
usage: tools/ncu_regions.py <report.ncu-rep> [addr1 addr2 ...]   (no addresses: 16 equal slices)"""
import csv, io, subprocess, sys
rep = sys.argv[1]
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
lines = txt.splitlines()
start = next(i for i, l in enumerate(lines) if l.startswith('"Address"'))
rows = list(csv.DictReader(io.StringIO("\n".join(lines[start:]))))
base = int(rows[0]["Address"], 16)
addr = [int(r["Address"], 16) - base for r in rows]
cuts = [int(a, 16) for a in sys.argv[2:]] or [addr[-1] * k // 16 for k in range(1, 16)]
cuts = [0] + sorted(cuts) + [addr[-1] + 16]
tot_s = sum(int(r["# Samples"] or 0) for r in rows)
print(f"{'region':>18s} {'instrs':>7s} {'executed':>11s} {'samples':>8s} {'share':>6s} {'no_inst':>7s} {'long_sb':>7s} {'math':>6s} {'wait':>6s}")
for lo, hi in zip(cuts, cuts[1:]):
    sel = [r for a, r in zip(addr, rows) if lo <= a < hi]
    if not sel: continue
    ex = sum(int(r["Instructions Executed"] or 0) for r in sel)
    sm = sum(int(r["# Samples"] or 0) for r in sel)
    f = lambda k: sum(int(r[k] or 0) for r in sel)
    print(f"{lo:#8x}-{hi:#8x} {len(sel):7d} {ex:11.3e} {sm:8d} {100*sm/max(tot_s,1):5.1f}% {100*f('stall_no_inst')/max(sm,1):6.1f}% {100*f('stall_long_sb')/max(sm,1):6.1f}% {100*f('stall_math')/max(sm,1):5.1f}% {100*f('stall_wait')/max(sm,1):5.1f}%")
