"""Small cases for compute-sanitizer (memcheck / racecheck / synccheck / initcheck), one GPU:
    compute-sanitizer --tool racecheck python tools_sanitize_case.py [hdiv] [h1h1] [patch]
hdiv  : H1-HDiv fused + separate kernels, Newton and Picard+zeta variants (clean in round 1)
h1h1  : H1-H1 Jacobian / residual kernels, all convection x zeta_u variants        (NOT yet run under the sanitizer)
patch : patch gather + blocked Gauss-Jordan inversion + apply, and the block preconditioners that use them  (NOT yet run)
The H1-H1 and patch device code has only been through the CPU emulation (tests/emul, incl. AddressSanitizer) so far."""
import sys

import numpy as np

import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gridapmhd_jl_b200  # noqa: E402,F401
from gridapmhd_jl_b200 import lib as L  # noqa: E402
from gridapmhd_jl_b200.applications import hunt_params, make_operator, setup_spaces  # noqa: E402
from gridapmhd_jl_b200.feoperator import B200FEOperator, B200LinearSolver, B200SolverOptions, FluidParams  # noqa: E402

what = set(sys.argv[1:]) or {"hdiv"}
L.init(0)
if "hdiv" in what:
    for kw in (dict(zeta_u=0.0, zeta_j=0.0, convection="newton"), dict(zeta_u=2.0, zeta_j=3.0, convection="picard")):
        params = hunt_params(nc=(3, 2), B=(0.0, 10.0, 0.0), **kw)
        fes = setup_spaces(params)
        op = B200FEOperator(fes, params["fluid"])
        x = np.random.default_rng(1).random(fes.ndofs)
        A = op.allocate_jacobian()
        b = np.empty(op.nrows)
        op.residual_and_jacobian_b(b, A, x)
        op.jacobian(x)
        r = op.residual(x)
        op.spmv(x)
        print("ok hdiv", kw, float(np.abs(r - b).max()))
        op.destroy()
if "h1h1" in what:
    params = hunt_params(nc=(3, 2), B=(0.0, 10.0, 0.0), current_disc="H1")
    fes = setup_spaces(params)
    for conv in ("none", "picard", "newton"):
        for zu in (0.0, 2.0):
            fl = FluidParams(alpha=1.0, beta=1.0, gamma=100.0, zeta_u=zu, B=(0.1, 1.0, -0.2), f=(0.0, 0.0, 1.0), convection=conv)
            op = make_operator(fes, fl)
            x = np.random.default_rng(1).random(fes.ndofs)
            A = op.allocate_jacobian()
            b = np.empty(op.nrows)
            op.residual_and_jacobian_b(b, A, x)
            r = op.residual(x)
            op.spmv(x)
            print("ok h1h1", conv, zu, float(np.abs(r - b).max()))
            op.destroy()
if "patch" in what:
    params = hunt_params(nc=(3, 2), B=(0.0, 10.0, 0.0), solver="badia2024", zeta_u=5.0, zeta_j=5.0)
    fes = setup_spaces(params)
    op = B200FEOperator(fes, params["fluid"])
    A = op.jacobian(0.1 * np.random.default_rng(2).random(fes.ndofs))
    ns = B200LinearSolver(B200SolverOptions(m=10, maxiter=10, precond="block_tri", uj_solver="gmres_patch", uj_inner_its=5,
                                            uj_inner_restart=5, patch_its=2, patch_omega=0.5)).symbolic_setup(A).numerical_setup()
    dx = np.zeros(op.nrows)
    ns.solve_b(dx, np.random.default_rng(3).standard_normal(op.nrows))
    print("ok patch (u,j)", ns.iters, float(ns.history[-1] / ns.history[0]))
    ns.destroy()
    op.destroy()
    params = hunt_params(nc=(3, 2), B=(0.0, 10.0, 0.0), zeta_u=5.0, current_disc="H1")
    fes = setup_spaces(params)
    op = make_operator(fes, params["fluid"])
    A = op.jacobian(np.zeros(fes.ndofs))
    ns = B200LinearSolver(B200SolverOptions(m=10, maxiter=10, precond="h1h1_blocks", uj_solver="gmres_patch", uj_inner_its=5,
                                            uj_inner_restart=5)).symbolic_setup(A).numerical_setup()
    dx = np.zeros(op.nrows)
    ns.solve_b(dx, np.random.default_rng(3).standard_normal(op.nrows))
    print("ok patch h1h1 blocks", ns.iters, float(ns.history[-1] / ns.history[0]))
    ns.destroy()
    op.destroy()
L.finalize()
