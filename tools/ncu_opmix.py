#!/usr/bin/env python3
"""Dynamic opcode mix + stall samples of a kernel from an ncu report's source page.
usage: tools/ncu_opmix.py <report.ncu-rep> [top]"""
import csv, subprocess, sys, collections, io
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
lines = txt.splitlines()
start = next(i for i, l in enumerate(lines) if l.startswith('"Address"'))
rd = csv.DictReader(io.StringIO("\n".join(lines[start:])))
ex = collections.Counter(); samp = collections.Counter(); tot = 0; tots = 0
HEAVY2 = ("IMAD.WIDE", "IMAD.HI")
for r in rd:
    src = r["Source"].strip()
    toks = src.split()
    if not toks: continue
    op = toks[1] if toks[0].startswith("@") else toks[0]
    n = int(r["Instructions Executed"] or 0); s = int(r["# Samples"] or 0)
    ex[op] += n; samp[op] += s; tot += n; tots += s
print(f"total warp-instructions executed: {tot:.4e}   stall samples: {tots}")
def pipe(op):
    if op.startswith(HEAVY2): return "heavy2"
    if op.startswith(("IMAD", "FFMA", "FMUL", "FADD")): return "heavy1"
    if op.startswith(("IADD3", "LOP3", "SHF", "PRMT", "SEL", "ISETP", "LEA", "VIADD", "MOV", "IABS", "FLO", "POPC", "PLOP3", "ICMP", "IMNMX", "VIMNMX", "BMSK", "SGXT")): return "alu"
    if op.startswith(("LD", "ST", "RED", "ATOM")): return "lsu"
    return "other"
bypipe = collections.Counter()
for op, n in ex.items(): bypipe[pipe(op)] += n
heavy_slots = 2 * bypipe["heavy2"] + bypipe["heavy1"]
print("by pipe (warp-instr):", {k: f"{v:.3e} ({100*v/tot:.1f}%)" for k, v in bypipe.items()})
print(f"fmaheavy issue slots: {heavy_slots:.4e} = wide x2 {2*bypipe['heavy2']:.3e} + single {bypipe['heavy1']:.3e}  -> non-multiply share {100*bypipe['heavy1']/heavy_slots:.1f}%")
print(f"alu slots: {bypipe['alu']:.4e}  (alu/heavy slot ratio {bypipe['alu']/heavy_slots:.2f})")
for op, n in ex.most_common(top):
    print(f"{op:22s} {n:12.4e} {100*n/tot:6.2f}%   samples {100*samp[op]/max(tots,1):6.2f}%")
