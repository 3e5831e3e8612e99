"""Patch smoother at the cfg2 mesh (Hunt nc=(64,64), Ha=1000, zeta=10): setup (gather + invert) and apply times, memory,
and the residual history of one FGMRES(15) solve with the block-triangular preconditioner + inner patch-GMRES.
Usage (GPU box): python tools_patch_bench.py [nc] > gpurun_out/patch_bench.json"""
import json
import sys
import time

import numpy as np

import gridapmhd_jl_b200  # noqa: F401
from gridapmhd_jl_b200 import lib as L
from gridapmhd_jl_b200.applications import hunt_params, setup_spaces
from gridapmhd_jl_b200.feoperator import B200FEOperator, B200LinearSolver, B200SolverOptions

nc = int(sys.argv[1]) if len(sys.argv) > 1 else 64
Ha = float(sys.argv[2]) if len(sys.argv) > 2 else 1000.0
L.init(0)
p = hunt_params(nc=(nc, nc), B=(0.0, Ha, 0.0), solver="badia2024", zeta_u=10.0, zeta_j=10.0)
fes = setup_spaces(p)
op = B200FEOperator(fes, p["fluid"])
A = op.allocate_jacobian()
b = np.empty(op.nrows)
op.residual_and_jacobian_b(b, A, np.zeros(fes.ndofs))
L.check(L.load().mhd_profile_enable(1))
L.check(L.load().mhd_profile_reset())
opts = B200SolverOptions(m=15, maxiter=15, rtol=1e-7, atol=0.0, precond="block_tri", uj_solver="gmres_patch", uj_inner_its=20,
                         uj_inner_restart=20)
t0 = time.perf_counter()
ns = B200LinearSolver(opts).symbolic_setup(A).numerical_setup()
L.check(L.load().mhd_device_synchronize())
t_setup = time.perf_counter() - t0
for _ in range(2):
    ns.numerical_setup_b(A)
setup_ms, nset = L.profile_get("patch_setup")
nuj = fes.nfree["u"] + fes.nfree["j"]
r = np.random.default_rng(0).standard_normal(nuj)
import torch

rd = torch.from_numpy(r).cuda()
for _ in range(3):
    ns.patch_apply(rd)
L.check(L.load().mhd_profile_reset())
for _ in range(10):
    ns.patch_apply(rd)
apply_ms, napp = L.profile_get("patch_apply")
L.check(L.load().mhd_profile_reset())
dx = np.zeros(op.nrows)
t0 = time.perf_counter()
ns.solve_b(dx, -b)
t_solve = time.perf_counter() - t0
solve_apply_ms, solve_napp = L.profile_get("patch_apply")
spmv_ms, nspmv = L.profile_get("spmv")
out = {"workload": f"Hunt nc=({nc},{nc}) Ha={Ha:g} zeta=10, vertex-patch block-Jacobi smoother of the (u,j) block", "ncells": fes.mesh.ncells,
       "n_uj": nuj, "npatches": ns.npatches, "inverse_GB": ns.patch_entries * 8 / 1e9, "first_setup_wall_s": t_setup,
       "patch_setup_ms": setup_ms / max(nset, 1), "patch_apply_ms": apply_ms / napp,
       "patch_apply_GBs": ns.patch_entries * 8 / (apply_ms / napp) / 1e6,
       "solve": {"outer_its": ns.iters, "wall_s": t_solve, "history": [float(h) for h in ns.history],
                 "patch_applies": solve_napp, "patch_apply_ms": solve_apply_ms / max(solve_napp, 1), "spmvs": nspmv,
                 "spmv_ms": spmv_ms / max(nspmv, 1)}}
print(json.dumps(out))
