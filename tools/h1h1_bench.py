"""Kernel times of the H1-H1 path at the cfg2 mesh (Hunt nc=(64,64), Ha=1000): Jacobian, residual, SpMV.
Usage (GPU box): python tools_h1h1_bench.py [nc] > gpurun_out/h1h1_bench.json"""
import json
import sys
import time

import numpy as np

import gridapmhd_jl_b200  # noqa: F401
from gridapmhd_jl_b200 import lib as L
from gridapmhd_jl_b200.applications import hunt_params, make_operator, setup_spaces

nc = int(sys.argv[1]) if len(sys.argv) > 1 else 64
L.init(0)
t0 = time.perf_counter()
p = hunt_params(nc=(nc, nc), B=(0.0, 1000.0, 0.0), current_disc="H1")
fes = setup_spaces(p)
t_host = time.perf_counter() - t0
op = make_operator(fes, p["fluid"])
t0 = time.perf_counter()
A = op.allocate_jacobian()
L.check(L.load().mhd_device_synchronize())
t_sym = time.perf_counter() - t0
import torch

x = torch.from_numpy(np.random.default_rng(1234).random(fes.ndofs)).cuda()
y = torch.empty_like(x)
r = torch.empty_like(x)
for _ in range(3):
    op.jacobian(x); op.residual_b(r, x); op.spmv(x, y)
L.check(L.load().mhd_profile_enable(1))
L.check(L.load().mhd_profile_reset())
n = 10
for _ in range(n):
    op.jacobian(x); op.residual_b(r, x); op.spmv(x, y)
jac_ms, nj = L.profile_get("jacobian")
res_ms, nr = L.profile_get("residual")
spmv_ms, ns = L.profile_get("spmv")
nent, nexcl = op.scatter_stats()
ncells = fes.mesh.ncells
alg = 8.0 * op.nnz + 2.0 * nent + ncells * (8 * 24 + 149 * 4 + 149 * 8 + 149 * 8)
out = {"workload": f"Hunt nc=({nc},{nc}) Ha=1000 H1-H1 (u Q2, p P1disc, phi Q3), newton convection", "ncells": ncells,
       "ndofs": fes.ndofs, "nnz": op.nnz, "entries": nent, "exclusive_entries": nexcl, "host_setup_s": t_host, "symbolic_s": t_sym,
       "jacobian_ms": jac_ms / nj, "residual_ms": res_ms / nr, "spmv_ms": spmv_ms / ns,
       "jacobian_Mcells_s": ncells / (jac_ms / nj) / 1e3, "jacobian_alg_GB": alg / 1e9, "jacobian_GBs": alg / (jac_ms / nj) / 1e6,
       "spmv_GBs": (12.0 * op.nnz + 20.0 * op.nrows) / (spmv_ms / ns) / 1e6}
print(json.dumps(out))
