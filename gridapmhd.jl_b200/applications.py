"""Drivers mirroring `GridapMHD.hunt(;kwargs...)` (src/Applications/hunt.jl:1-304) and `main(params)`
(src/main.jl:122-169) for the part of the call stack that reaches the hot path.

Only what configures the path is restated: reduced quantities (hunt.jl:115-136), mesh (hunt.jl:140-143),
params[:fluid]/[:bcs] (hunt.jl:149-193), the timed sections of `main` (`solve`, `residual`, `jacobian`;
main.jl:144-165).  VTK/BSON output and the analytical error norms stay with the Julia host.
"""
from __future__ import annotations

import math
import time

import numpy as np

from .feoperator import B200FEOperator, B200H1H1FEOperator, B200LinearSolver, B200SolverOptions, FluidParams, NewtonSolver
from .host.fespaces import FESpaces, setup_fe_spaces
from .host.fespaces_h1h1 import H1H1Spaces, setup_fe_spaces_h1h1
from .host.mesh import HexMesh, expansion_generate_mesh, hunt_generate_base_mesh


def hunt_reduced_quantities(nu=1.0, rho=1.0, sigma=1.0, B=(0.0, 10.0, 0.0), f=(0.0, 0.0, 1.0), L=1.0, u0=1.0,
                            formulation="cfd"):
    """hunt.jl:115-136.  Returns (alpha, beta, gamma, fbar, Bbar, Re, Ha, N)."""
    B0 = math.sqrt(sum(b * b for b in B))
    Re = u0 * L / nu
    Ha = B0 * L * math.sqrt(sigma / (rho * nu))
    N = Ha**2 / Re
    fbar = tuple((L / (rho * u0**2)) * x for x in f)
    Bbar = tuple(b / B0 for b in B)
    if formulation == "cfd":
        alpha, beta, gamma = 1.0, 1.0 / Re, N
    elif formulation == "mhd":
        alpha, beta, gamma = 1.0 / N, 1.0 / Ha**2, 1.0
        fbar = tuple(x / N for x in fbar)
    else:
        raise ValueError("Unknown formulation")
    return alpha, beta, gamma, fbar, Bbar, Re, Ha, N


def hunt_params(nc=(4, 4), nu=1.0, rho=1.0, sigma=1.0, B=(0.0, 10.0, 0.0), f=(0.0, 0.0, 1.0), zeta_u=0.0, zeta_j=0.0,
                L=1.0, u0=1.0, formulation="cfd", convection="newton", BL_adapted=True, kmap_x=1, kmap_y=1,
                solver="julia", nz=3, periodic_z=True, z_extent=(0.0, 0.1), tw=0.0, sigma_w1=0.1, sigma_w2=10.0,
                current_disc="RT"):
    """params Dict of `_hunt` (hunt.jl:88-193).  NOTE: `_hunt` never forwards its `convection` kwarg into
    params[:fluid] (hunt.jl:149-158), so the reference effectively runs Hunt with the `params_fluid` default
    `:newton` (parameters.jl:717); that is the default here too."""
    alpha, beta, gamma, fbar, Bbar, Re, Ha, N = hunt_reduced_quantities(nu, rho, sigma, B, f, L, u0, formulation)
    mesh = hunt_generate_base_mesh(nc, L=L, tw=tw, Ha=Ha, kmap_x=kmap_x, kmap_y=kmap_y, BL_adapted=BL_adapted, nz=nz,
                                   periodic_z=periodic_z, z_extent=z_extent)
    bcs = {"u": {"tags": ("noslip",) + (("zwalls",) if not periodic_z else ()), "values": None},
           "j": {"tags": ("insulating",)}}
    if current_disc == "H1":  # hunt.jl:185-187
        bcs["phi"] = {"tags": ("conducting",), "values": None}
    elif current_disc != "RT":
        raise ValueError("current_disc must be 'RT' (H1-HDiv) or 'H1' (H1-H1)")
    return {
        "model": mesh,
        "fluid": FluidParams(alpha=alpha, beta=beta, gamma=gamma, sigma=1.0, zeta_u=zeta_u, zeta_j=zeta_j, B=Bbar,
                             f=fbar, convection=convection),
        "bcs": bcs,
        # params[:fespaces][:current_disc] (hunt.jl:165): :RT -> H1-HDiv, :H1 -> H1-H1 (select_formulation,
        # parameters.jl:570-594)
        "current_disc": current_disc,
        "solver": solver,
        # params[:solid] (hunt.jl:168-176): sigma per cell = sigma_w1/sigma on solid_1, sigma_w2/sigma on solid_2; zeta = zeta_j
        "solid": None if tw <= 0.0 else {"cells": mesh.cell_tags["solid"],
                                         "sigma": np.where(mesh.cell_tags["solid_1"], sigma_w1 / sigma,
                                                           np.where(mesh.cell_tags["solid_2"], sigma_w2 / sigma, 1.0))},
        "info": {"Re": Re, "Ha": Ha, "N": N, "ncells": mesh.ncells},
    }


def u_inlet_parabolic(Z=4.0, b=0.2):
    """`u_inlet(:parabolic,...)` (expansion.jl:385): 36 Z (y-1/Z)(y+1/Z)(z-bZ)(z+bZ) e_x."""

    def fn(X):
        y, z = X[:, 1], X[:, 2]
        u = np.zeros_like(X)
        u[:, 0] = 36.0 * Z * (y - 1.0 / Z) * (y + 1.0 / Z) * (z - b * Z) * (z + b * Z)
        return u

    return fn


def expansion_params(level=1, Ha=1.0, N=1.0, zeta_u=0.0, zeta_j=0.0, formulation="mhd", convection="newton", Z=4.0, b=0.2,
                     solver="julia", perturb=0.0, mesh=None):
    """params Dict of `_expansion` (src/Applications/expansion.jl:40-181) for `solid_coupling=:none`,
    `inlet=:parabolic`: :mhd scaling alpha=1/N, beta=1/Ha^2, gamma=1 (:112-123), B=(0,1,0), f=0 (:133-143), u Dirichlet on
    ["inlet","wall"] = [u_in, 0] (:150-153), j.n = 0 on ["wall","inlet","outlet"] (:156-159).  The mesh is the p4est
    base mesh of expansion_mesher.jl refined `level` times (stand-in for the missing Expansion_68k/749k.msh) unless a
    `HexMesh` (e.g. from `read_gmsh41`) is passed."""
    Re = Ha**2 / N
    if formulation == "cfd":
        alpha, beta, gamma = 1.0, 1.0 / Re, N
    elif formulation == "mhd":
        alpha, beta, gamma = 1.0 / N, 1.0 / Ha**2, 1.0
    else:
        raise ValueError("Unknown formulation")
    mesh = mesh if mesh is not None else expansion_generate_mesh(level, perturb=perturb)
    return {
        "model": mesh,
        "fluid": FluidParams(alpha=alpha, beta=beta, gamma=gamma, sigma=1.0, zeta_u=zeta_u, zeta_j=zeta_j, B=(0.0, 1.0, 0.0),
                             f=(0.0, 0.0, 0.0), convection=convection),
        "bcs": {"u": {"tags": ("inlet", "wall"), "values": (u_inlet_parabolic(Z, b), None)},
                "j": {"tags": ("wall", "inlet", "outlet")}},
        "solver": solver,
        "info": {"Re": Re, "Ha": Ha, "N": N, "ncells": mesh.ncells},
    }


def setup_spaces(params) -> FESpaces | H1H1Spaces:
    """`setup_fe_spaces(params)` (src/fespaces.jl:13-46) on the host."""
    bcs = params["bcs"]
    utags = tuple(bcs["u"]["tags"])
    uvals = bcs["u"].get("values")
    if uvals is None:
        uvals = (None,) * len(utags)
    solid = params.get("solid")
    if params.get("current_disc", "RT") == "H1":
        ftags = tuple(bcs["phi"]["tags"])
        fvals = bcs["phi"].get("values") or (None,) * len(ftags)
        return setup_fe_spaces_h1h1(params["model"], u_tags=utags, u_values=tuple(uvals), phi_tags=ftags,
                                    phi_values=tuple(fvals), solid_cells=None if solid is None else solid["cells"])
    return setup_fe_spaces(params["model"], u_tags=utags, u_values=tuple(uvals), j_tags=tuple(bcs["j"]["tags"]),
                           solver=params.get("solver", "julia"), solid_cells=None if solid is None else solid["cells"],
                           cell_sigma=None if solid is None else solid["sigma"])


def make_operator(fes, fluid: FluidParams, nowned=None):
    """`_fe_operator(U,V,params)` (src/main.jl:207-233): the weak form follows the spaces (`weak_form`, weakforms.jl:2-31)."""
    cls = B200H1H1FEOperator if isinstance(fes, H1H1Spaces) else B200FEOperator
    return cls(fes, fluid, nowned)


def main(params, solve=True, res_assemble=False, jac_assemble=False, solver_opts: B200SolverOptions | None = None,
         newton_maxiter=10, newton_rtol=1e-6, verbose=False):
    """`main(params;output)` (src/main.jl:122-169): FE spaces -> FE operator -> [solve] -> [residual] -> [jacobian],
    with the PTimer sections of the reference (`fe_spaces`, `solve`, `residual`, `jacobian`) as wall-clock seconds."""
    from . import lib as L

    times = {}
    t0 = time.perf_counter()
    fes = setup_spaces(params)
    times["fe_spaces"] = time.perf_counter() - t0
    op = make_operator(fes, params["fluid"])
    x = np.zeros(fes.ndofs)  # initial_guess(::Val{:zero}) main.jl:302
    out = {"fes": fes, "op": op}
    if solve:
        t0 = time.perf_counter()
        if solver_opts is None and isinstance(fes, H1H1Spaces):
            # default_solver_params(Val(:h1h1blocks)) (src/parameters.jl:274-287) with the patch-smoothed block solvers
            solver_opts = B200SolverOptions(m=30, maxiter=60, rtol=newton_rtol / 10.0, atol=1e-14, precond="h1h1_blocks",
                                            uj_solver="gmres_patch", uj_inner_its=30, uj_inner_restart=30)
        nls = NewtonSolver(B200LinearSolver(solver_opts), maxiter=newton_maxiter, rtol=newton_rtol, verbose=verbose)
        x = nls.solve_b(x, op)
        L.check(L.load().mhd_device_synchronize())
        times["solve"] = time.perf_counter() - t0
        out["newton_log"] = nls.log
    if res_assemble:
        t0 = time.perf_counter()
        out["residual"] = op.residual(x)
        times["residual"] = time.perf_counter() - t0
    if jac_assemble:
        t0 = time.perf_counter()
        out["jacobian"] = op.jacobian(x)
        L.check(L.load().mhd_device_synchronize())
        times["jacobian"] = time.perf_counter() - t0
    out["x"] = x
    out["times"] = times
    return out


def hunt(nsums=10, post_process=True, **kwargs):
    """`hunt(;kwargs...)`: build params, call `main`, post-process, return the info dict (hunt.jl:212-303 subset:
    `time_*`, dof counts, and -- when `post_process` -- the norms `eu_l2 … jh_l2` of hunt.jl:247-260, integrated on the
    device against the analytical series with `nsums` terms, default as in the reference, hunt.jl:70)."""
    from .host.reffe import make_tables

    main_keys = ("solve", "res_assemble", "jac_assemble", "solver_opts", "newton_maxiter", "newton_rtol", "verbose")
    mk = {k: kwargs.pop(k) for k in main_keys if k in kwargs}
    phys = {k: kwargs.get(k, d) for k, d in (("nu", 1.0), ("rho", 1.0), ("sigma", 1.0), ("f", (0.0, 0.0, 1.0)),
                                              ("L", 1.0), ("u0", 1.0), ("B", (0.0, 10.0, 0.0)))}
    params = hunt_params(**kwargs)
    out = main(params, **mk)
    fes = out["fes"]
    info = dict(params["info"])
    info.update({f"ndofs_{f}": n for f, n in fes.nfree.items()})  # (u, p, j, phi) or, for current_disc = H1, (u, p, phi)
    info["ndofs"] = fes.ndofs
    h1h1 = "j" not in fes.nfree
    if post_process and h1h1:
        # hunt.jl:247-260 needs j_h; in the H1-H1 formulation it is the derived field sigma (-grad phi_h + u_h x B)
        # (src/weakforms.jl:121-135), which the device post-processing kernel does not evaluate: report that instead of failing
        info["post_process"] = "skipped: H1-H1 error norms are not built on the device (j_h is a derived field)"
    if post_process and not h1h1 and phys["L"] == 1.0:
        t0 = time.perf_counter()
        B0 = math.sqrt(sum(b * b for b in phys["B"]))
        norms = out["op"].hunt_error_norms(
            out["x"], make_tables(6), info["Ha"], nsums, u0=phys["u0"], jscale=phys["sigma"] * phys["u0"] * B0,  # hunt.jl:212-217
            a=phys["L"], mu=phys["rho"] * phys["nu"], sigma=phys["sigma"], grad_pz=-phys["f"][2] / phys["rho"])
        out["times"]["post_process"] = time.perf_counter() - t0
        info.update(norms)
    for k, v in out["times"].items():
        info[f"time_{k}"] = v
    return info, out
