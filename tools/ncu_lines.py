#!/usr/bin/env python
"""Per-source-line summary of an ncu report (source page): stall samples, instructions, shared-memory wavefronts,
global tag requests.  usage: tools_ncu_lines.py report.ncu-rep [ncells] [top]"""
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
tot = sum(g(r, "# Samples") for r in lines); toti = sum(g(r, "Instructions Executed") for r in lines)
print("samples", tot, "warp-instr/cell", toti / ncells, "smem wf/cell", sum(g(r, "L1 Wavefronts Shared") for r in lines) / ncells,
      "global tags/cell", sum(g(r, "L1 Tag Requests Global") for r in lines) / ncells, "L2 sectors/cell", sum(g(r, "L2 Theoretical Sectors Global") for r in lines) / ncells)
st = [k for k in hdr if k.startswith("stall_") and "Not Issued" not in k]
print("stalls %:", {k[6:]: round(100 * sum(g(r, k) for r in lines) / tot, 1) for k in st if sum(g(r, k) for r in lines) / tot > 0.005})
lines.sort(key=lambda r: -g(r, "# Samples"))
print("line samp% inst/cell smemwf tags | barrier long_sb short_sb wait math mio lg | src")
for r in lines[:top]:
    print(r[0], round(100 * g(r, "# Samples") / tot, 1), int(g(r, "Instructions Executed") / ncells), int(g(r, "L1 Wavefronts Shared") / ncells), int(g(r, "L1 Tag Requests Global") / ncells), "|",
          *(int(g(r, "stall_" + k)) for k in ("barrier", "long_sb", "short_sb", "wait", "math", "mio", "lg")), "|", r[1][:95])
