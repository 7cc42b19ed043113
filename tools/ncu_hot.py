#!/usr/bin/env python
"""Where a kernel spends its time, from an .ncu-rep captured with --import-source on:
warp-stall samples and executed instructions per SASS opcode class and per region between barriers.
  python tools/ncu_hot.py file.ncu-rep"""
import csv, io, subprocess, sys, collections, re
rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = rows[1]
iS, iN, iE = hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
tot_s = tot_e = 0
by_op = collections.defaultdict(lambda: [0, 0])
regions, cur = [], [0, 0, 0, ""]
for r in rows[2:]:
    if len(r) <= iE:
        continue
    src = r[iS].strip()
    m = re.match(r"(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", src)
    op = m.group(1) if m else "?"
    s, e = int(r[iN] or 0), int(r[iE] or 0)
    tot_s += s; tot_e += e
    by_op[op][0] += s; by_op[op][1] += e
    cur[0] += s; cur[1] += e; cur[2] += 1
    if op == "BAR":
        cur[3] = src
        regions.append(cur); cur = [0, 0, 0, ""]
regions.append(cur)
print(f"total samples {tot_s}, warp instructions executed {tot_e}")
print("-- by opcode (samples %, executed %)")
for op, (s, e) in sorted(by_op.items(), key=lambda kv: -kv[1][0])[:18]:
    print(f"  {op:10s} {100*s/tot_s:6.2f} %   {100*e/tot_e:6.2f} %")
print("-- regions between barriers (static instr, samples %, executed %)")
for s, e, n, b in regions:
    print(f"  {n:6d} instr  {100*s/tot_s:6.2f} %  {100*e/tot_e:6.2f} %")
