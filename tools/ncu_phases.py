#!/usr/bin/env python
"""Per-function (phase) summary of an ncu report of the v7 kernel: executed warp instructions, shared wavefronts, samples per cell.
usage: ncu_phases.py report.ncu-rep [ncells]"""
import csv, io, re, subprocess, sys, os
rep = sys.argv[1]; ncells = float(sys.argv[2]) if len(sys.argv) > 2 else 12288.0
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "gridapmhd.jl_b200", "csrc")
files = {f: open(os.path.join(ROOT, f)).read().split("\n") for f in ("hdiv_v7.cu", "hdiv7_cell.h")}
def func_of(fname, lineno):
    src = files[fname]
    for i in range(min(lineno, len(src)) - 1, -1, -1):
        m = re.match(r"^(?:template.*\n)?(?:MHD_7HD|V7_NI|__device__ __forceinline__|__global__|static|inline|int|void)[^;(]*?\b(\w+)\s*\(", src[i])
        if m and not src[i].startswith(" "): return m.group(1)
        if src[i].startswith("hdiv_v7_jacobian_kernel("): return "kernel_body"
    return "?"
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass,cuda"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
agg = {}
cur = None; hdr = None; ix = {}
def g(r, k):
    try: return float(r[ix[k]] or 0)
    except Exception: return 0.0
for i, r in enumerate(rows):
    if r and r[0] == "File Path":
        cur = os.path.basename(r[1]); continue
    if r and r[0] == "Line No":
        hdr = r; ix = {}
        for j, h in enumerate(hdr): ix.setdefault(h, j)
        continue
    if not r or hdr is None or len(r) < len(hdr) - 2 or not r[0].isdigit(): continue
    ln = int(r[0])
    fn = func_of(cur, ln) if cur in files else "(" + str(cur) + ")"
    a = agg.setdefault(fn, [0, 0, 0])
    a[0] += g(r, "Instructions Executed"); a[1] += g(r, "L1 Wavefronts Shared"); a[2] += g(r, "# Samples")
ti = sum(a[0] for a in agg.values()); ts = sum(a[2] for a in agg.values())
print(f"{'function':24s} {'inst/cell':>10s} {'%':>6s} {'smemwf/cell':>12s} {'samples%':>9s}")
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    print(f"{k:24s} {a[0]/ncells:10.0f} {100*a[0]/ti:6.1f} {a[1]/ncells:12.0f} {100*a[2]/ts:9.1f}")
print(f"{'total':24s} {ti/ncells:10.0f} {'':6s} {sum(a[1] for a in agg.values())/ncells:12.0f}")
