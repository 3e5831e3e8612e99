#!/usr/bin/env python
"""Convergence study of the device-resident linear solve at the benchmark size: Hunt nc=(64,64), Ha=500, unstretched mesh (the
published run hconv_ha00500ns100/summary.csv:2), FGMRES + block-triangular preconditioner, (u,j) block = inner GMRES
preconditioned by the vertex-patch smoother.  usage: solve_cfg2.py [nc] [Ha] [BL_adapted 0|1] zeta:inner:patch_its:omega[:m[:cycles]] ..."""
import json
import sys
import time

import numpy as np

sys.path.insert(0, __file__.rsplit("/", 2)[0])
import gridapmhd_jl_b200  # noqa: F401,E402
from gridapmhd_jl_b200 import lib as L  # noqa: E402
from gridapmhd_jl_b200.applications import hunt_params, setup_spaces  # noqa: E402
from gridapmhd_jl_b200.feoperator import B200FEOperator, B200LinearSolver, B200SolverOptions  # noqa: E402

nc = int(sys.argv[1]) if len(sys.argv) > 1 else 64
Ha = float(sys.argv[2]) if len(sys.argv) > 2 else 500.0
bl = bool(int(sys.argv[3])) if len(sys.argv) > 3 else False
cases = sys.argv[4:] or ["10:30:1:1.0"]
L.init(0)
for spec in cases:
    f = spec.split(":")
    zeta, inner, pits, omega = float(f[0]), int(f[1]), int(f[2]), float(f[3])
    m = int(f[4]) if len(f) > 4 else 30
    cycles = int(f[5]) if len(f) > 5 else 4
    p = hunt_params(nc=(nc, nc), B=(0.0, Ha, 0.0), BL_adapted=bl, zeta_u=zeta, zeta_j=zeta, solver="badia2024")
    fes = setup_spaces(p)
    op = B200FEOperator(fes, p["fluid"])
    A = op.allocate_jacobian()
    b = np.empty(op.nrows)
    x = np.zeros(fes.ndofs)
    op.residual_and_jacobian_b(b, A, x)
    t0 = time.perf_counter()
    opts = B200SolverOptions(m=m, maxiter=m * cycles, rtol=1e-10, atol=0.0, precond="block_tri", uj_solver="gmres_patch", uj_inner_its=inner,
                             uj_inner_restart=inner, patch_its=pits, patch_omega=omega)
    ns = B200LinearSolver(opts).symbolic_setup(A).numerical_setup()
    L.check(L.load().mhd_device_synchronize())
    t_setup = time.perf_counter() - t0
    dx = np.zeros(op.nrows)
    t0 = time.perf_counter()
    ns.solve_b(dx, -b)
    t_solve = time.perf_counter() - t0
    h = ns.history
    # true residual on the device
    y = op.spmv(dx)
    true = float(np.linalg.norm(y + b) / np.linalg.norm(b))
    print(json.dumps({"case": spec, "nc": nc, "Ha": Ha, "BL_adapted": bl, "ndofs": int(fes.ndofs), "iters": int(ns.iters), "setup_s": t_setup, "solve_s": t_solve,
                      "rel_residual_estimate": float(h[-1] / h[0]) if len(h) else None, "true_rel_residual": true,
                      "history_rel": [float(v / h[0]) for v in h[:: max(1, len(h) // 12)]]}), flush=True)
    ns.destroy()
    op.destroy()
L.finalize()
