#!/usr/bin/env python
"""Per-source-line shared-memory wavefronts / instructions of an ncu report, sorted by wavefronts.
usage: ncu_lines_smem.py report.ncu-rep [ncells] [top]"""
import csv, subprocess, sys, io
rep = sys.argv[1]; ncells = float(sys.argv[2]) if len(sys.argv) > 2 else 12288.0; top = int(sys.argv[3]) if len(sys.argv) > 3 else 45
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass,cuda"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Line No")
hdr = rows[hi]; ix = {}
for i, h in enumerate(hdr): ix.setdefault(h, i)
lines = [r for r in rows[hi + 1:] if r and r[0] != "" and len(r) >= len(hdr) - 2]
def g(r, k):
    try: return float(r[ix[k]] or 0)
    except Exception: return 0.0
lines.sort(key=lambda r: -g(r, "L1 Wavefronts Shared"))
print("line smemwf/cell ideal/cell inst/cell | src")
for r in lines[:top]:
    print(r[0], int(g(r, "L1 Wavefronts Shared") / ncells), int(g(r, "L1 Wavefronts Shared Ideal") / ncells) if "L1 Wavefronts Shared Ideal" in ix else -1,
          int(g(r, "Instructions Executed") / ncells), "|", r[1][:110])
