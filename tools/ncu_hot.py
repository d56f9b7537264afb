#!/usr/bin/env python3
"""Per-instruction stall view of one kernel from an ncu report (source page):
    tools/ncu_hot.py <report.ncu-rep> <kernel regex> [top N]
Prints the stall-sample share by opcode class and the N instructions with the most
stall samples, with their dominant stall reason (needs --import-source / -lineinfo)."""
import collections
import csv
import subprocess
import sys

rep, pat = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{pat}"],
                     capture_output=True, text=True).stdout.splitlines()
# first kernel only
rows = list(csv.reader(out))
start = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[start]
body = []
for r in rows[start + 1:]:
    if not r or r[0] == "Kernel Name":
        break
    body.append(r)
si, ii, ei = hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
tot = sum(int(r[ii]) for r in body)
inst = sum(int(r[ei]) for r in body)
print(f"{len(body)} SASS instructions, {tot} samples, {inst} warp instructions executed")
byop = collections.Counter()
byreason = collections.Counter()
for r in body:
    op = r[si].split()
    op = op[1] if op[0].startswith("@") else op[0]
    byop[op.split(".")[0]] += int(r[ii])
    for c in stall_cols:
        byreason[hdr[c]] += int(r[c])
print("samples by opcode:", ", ".join(f"{k} {100 * v / tot:.1f}%" for k, v in byop.most_common(12)))
print("samples by reason:", ", ".join(f"{k} {100 * v / tot:.1f}%" for k, v in byreason.most_common(8)))
order = sorted(range(len(body)), key=lambda i: -int(body[i][ii]))[:top]
for i in sorted(order):
    r = body[i]
    reasons = sorted(((int(r[c]), hdr[c]) for c in stall_cols), reverse=True)[:2]
    prev = body[i - 1][si].strip() if i else ""
    print(f"{int(r[ii]):6d} {100 * int(r[ii]) / tot:5.2f}%  {r[si].strip():60s} <- {prev[:50]:50s} {reasons}")
