#!/usr/bin/env python
"""Top stalled SASS instructions of an .ncu-rep (source page). Usage: ncu_hot.py file.ncu-rep [N]"""
import csv, io, subprocess, sys
path = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
out = subprocess.check_output(["ncu", "-i", path, "--page", "source", "--csv"], text=True)
rows = list(csv.reader(io.StringIO(out)))
kern = None
for i, r in enumerate(rows):
    if r and r[0] == "Kernel Name":
        if kern is not None: break
        kern = r[1]; hdr = rows[i + 1]; start = i + 2
si = hdr.index("# Samples"); src = hdr.index("Source"); ex = hdr.index("Instructions Executed")
stalls = [j for j, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
data = []
for idx, r in enumerate(rows[start:]):
    if len(r) != len(hdr) or r[0] == "Kernel Name": break
    data.append((int(r[si]), idx, r))
tot = sum(d[0] for d in data)
print(kern, "total samples", tot, "instructions", len(data))
for n, idx, r in sorted(data, key=lambda t: -t[0])[:top]:
    why = sorted(((int(r[j]), hdr[j]) for j in stalls), reverse=True)[:2]
    print(f"{n:7d} {100*n/tot:5.1f}%  #{idx:4d} ex={r[ex]:>8s} {r[src].strip()[:80]:80s} {why}")
