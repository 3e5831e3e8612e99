"""GPU parity tests: the CUDA path, called through the C ABI, against the CPU oracle on the same seeded inputs.

Tolerances (BASELINE.json north_star): CSR structure bit-exact; matrix/residual values <= 1e-12 relative
(to the largest entry of the block); solutions <= 1e-10 relative.
"""
import numpy as np
import pytest

from gridapmhd_jl_b200.applications import hunt_params, setup_spaces
from gridapmhd_jl_b200.feoperator import B200FEOperator, FluidParams

pytestmark = pytest.mark.gpu

VAL_TOL = 1e-12
SOL_TOL = 1e-10


def oracle_params(fl: FluidParams):
    from oracle import mhd_oracle as O

    return O.FluidParams(fl.alpha, fl.beta, fl.gamma, fl.sigma, fl.zeta_u, fl.zeta_j, fl.B, fl.f, fl.g, fl.convection)


def make_case(nc=(4, 4), B=(0.0, 10.0, 0.0), solver="julia", **kw):
    params = hunt_params(nc=nc, B=B, solver=solver, **kw)
    fes = setup_spaces(params)
    return params, fes


def relerr(a, b):
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-300)


@pytest.fixture(scope="module")
def cfg1(mhdlib):
    """BASELINE config 1: Hunt nc=(4,4), Ha=10."""
    params, fes = make_case()
    op = B200FEOperator(fes, params["fluid"])
    yield params, fes, op
    op.destroy()


def test_symbolic_structure_bit_exact(cfg1):
    from oracle import mhd_oracle as O

    params, fes, op = cfg1
    A = op.allocate_jacobian()
    rowptr, colval = A.pattern()
    rp, cv = O.symbolic_csr(fes.cell_global_ids(), fes.ndofs)
    assert A.nnz == 381336  # cfg1 nnz
    assert op.nrows == 2610
    assert np.array_equal(rowptr, rp)
    assert np.array_equal(colval, cv)
    # 1-based Int64 (SparseMatrixCSC-style indices) and 0-based Int32 (PetscInt) views agree
    rp32 = np.empty(op.nrows + 1, dtype=np.int32)
    cv32 = np.empty(A.nnz, dtype=np.int32)
    from gridapmhd_jl_b200 import lib as L

    L.check(L.load().mhd_operator_get_csr(op.handle, L.ptr(rp32), L.ptr(cv32), 4, 1))
    assert np.array_equal(rp32 - 1, rp) and np.array_equal(cv32 - 1, cv)
    nent, nex = op.scatter_stats()
    assert nent == 453420  # scattered entries, SURVEY.md section 8
    assert 0 < nex < nent


@pytest.mark.parametrize(
    "conv,zu,zj",
    [("none", 0.0, 0.0), ("picard", 0.0, 0.0), ("newton", 0.0, 0.0), ("none", 10.0, 10.0), ("newton", 7.0, 3.0), ("picard", 2.0, 0.0)],
)
def test_jacobian_residual_values(cfg1, conv, zu, zj):
    from oracle import mhd_oracle as O

    params, fes, op = cfg1
    fl = FluidParams(alpha=0.7, beta=0.9, gamma=100.0, sigma=1.3, zeta_u=zu, zeta_j=zj, B=(0.1, 1.0, 0.2), f=(0.3, 0.1, 1.0),
                     g=(0.1, 0.2, 0.3), convection=conv)
    op.set_fluid(fl)
    x = np.random.default_rng(1234).random(fes.ndofs)
    A = op.jacobian(x)
    Ao = O.jacobian(fes, x, oracle_params(fl))
    rowptr, colval = A.pattern()
    assert np.array_equal(rowptr, Ao.indptr) and np.array_equal(colval, Ao.indices)
    assert relerr(A.nzval(), Ao.data) < VAL_TOL
    r = op.residual(x)
    ro = O.residual(fes, x, oracle_params(fl))
    assert relerr(r, ro) < VAL_TOL
    op.set_fluid(params["fluid"])


def test_jacobian_blockwise_relative(cfg1):
    """Every touched block separately within 1e-12 of its own scale (a tiny block must not hide behind gamma)."""
    from oracle import mhd_oracle as O

    params, fes, op = cfg1
    fl = FluidParams(alpha=1.0, beta=1.0, gamma=1e6, sigma=1.0, zeta_u=0.0, zeta_j=0.0, B=(0.0, 1.0, 0.0), f=(0, 0, 1), convection="newton")
    op.set_fluid(fl)
    x = np.random.default_rng(7).random(fes.ndofs)
    A = op.jacobian(x).to_scipy()
    Ao = O.jacobian(fes, x, oracle_params(fl))
    off = fes.offsets
    rng = {f: (off[f], off[f] + fes.nfree[f]) for f in off}
    for fr, fc in (("u", "u"), ("u", "p"), ("u", "j"), ("p", "u"), ("j", "j"), ("j", "u"), ("j", "phi"), ("phi", "j")):
        a = A[rng[fr][0] : rng[fr][1], rng[fc][0] : rng[fc][1]].toarray()
        b = Ao[rng[fr][0] : rng[fr][1], rng[fc][0] : rng[fc][1]].toarray()
        assert relerr(a, b) < VAL_TOL, (fr, fc)
    op.set_fluid(params["fluid"])


def test_device_pointer_path_matches_host_path(cfg1):
    import torch

    params, fes, op = cfg1
    x = np.random.default_rng(3).random(fes.ndofs)
    A = op.jacobian(x)
    v_host = A.nzval()
    xd = torch.from_numpy(x).cuda()
    op.jacobian(xd)
    torch.cuda.synchronize()
    assert np.array_equal(np.isfinite(A.nzval()), np.ones(A.nnz, dtype=bool))
    assert relerr(A.nzval(), v_host) < 1e-14  # atomics reorder sums only
    rd = op.residual(xd)
    assert relerr(rd.cpu().numpy(), op.residual(x)) < 1e-14


def test_fused_residual_and_jacobian(cfg1):
    """residual_and_jacobian!: the fused kernel gives the same matrix and residual as the two separate kernels."""
    from oracle import mhd_oracle as O

    params, fes, op = cfg1
    for conv, zu in (("newton", 0.0), ("none", 5.0), ("picard", 0.0)):
        fl = FluidParams(alpha=0.7, beta=0.9, gamma=100.0, sigma=1.3, zeta_u=zu, zeta_j=2.0, B=(0.1, 1.0, 0.2), f=(0.3, 0.1, 1.0),
                         g=(0.1, 0.2, 0.3), convection=conv)
        op.set_fluid(fl)
        x = np.random.default_rng(99).random(fes.ndofs)
        A = op.allocate_jacobian()
        b = np.empty(op.nrows)
        op.residual_and_jacobian_b(b, A, x)
        Ao = O.jacobian(fes, x, oracle_params(fl))
        assert relerr(A.nzval(), Ao.data) < VAL_TOL
        assert relerr(b, O.residual(fes, x, oracle_params(fl))) < VAL_TOL
    op.set_fluid(params["fluid"])


def test_spmv_dot_axpy(cfg1):
    import torch

    params, fes, op = cfg1
    rng = np.random.default_rng(5)
    x = rng.random(fes.ndofs)
    A = op.jacobian(x)
    As = A.to_scipy()
    v = rng.standard_normal(fes.ndofs)
    y = op.spmv(v)
    assert relerr(y, As @ v) < VAL_TOL
    w = rng.standard_normal(fes.ndofs)
    assert abs(op.dot(v, w) - v @ w) <= 1e-13 * np.abs(v * w).sum()
    y2 = w.copy()
    op.axpy(-0.37, v, y2)
    assert relerr(y2, w - 0.37 * v) < 1e-15
    # device-resident variants
    vd, wd = torch.from_numpy(v).cuda(), torch.from_numpy(w).cuda()
    yd = op.spmv(vd)
    assert relerr(yd.cpu().numpy(), As @ v) < VAL_TOL
    op.axpy(2.0, vd, wd)
    assert relerr(wd.cpu().numpy(), w + 2.0 * v) < 1e-15
    # fused Gram-Schmidt step
    k = 5
    V = np.linalg.qr(rng.standard_normal((fes.ndofs, k)))[0].T.copy()
    ww = rng.standard_normal(fes.ndofs)
    w0 = ww.copy()
    h = op.multi_dot_axpy(V, ww)
    assert relerr(h, V @ w0) < 1e-13
    assert relerr(ww, w0 - V.T @ (V @ w0)) < 1e-13


def test_edge_cases_and_errors(mhdlib):
    """Ragged / degenerate inputs and the error convention (no abort, message available)."""
    from gridapmhd_jl_b200 import lib as L

    params, fes = make_case(nc=(1, 1))  # single column of 3 periodic cells: only the 6 axis nodes of u are free
    assert fes.nfree["u"] == 18
    op = B200FEOperator(fes, params["fluid"])
    x = np.random.default_rng(0).random(fes.ndofs)
    from oracle import mhd_oracle as O

    A = op.jacobian(x)
    Ao = O.jacobian(fes, x, oracle_params(params["fluid"]))
    rowptr, colval = A.pattern()
    assert np.array_equal(rowptr, Ao.indptr) and np.array_equal(colval, Ao.indices)
    assert relerr(A.nzval(), Ao.data) < VAL_TOL
    assert relerr(op.residual(x), O.residual(fes, x, oracle_params(params["fluid"]))) < VAL_TOL
    # call-order violation reports MHD_E_STATE, bad arguments MHD_E_INVALID
    op2 = B200FEOperator(fes, params["fluid"])
    rc = L.load().mhd_jacobian(op2.handle, L.ptr(x), None)
    assert rc == -3 and b"symbolic" in L.load().mhd_last_error_string()
    rc = L.load().mhd_operator_get_csr(op2.handle, None, None, 8, 0)
    assert rc == -3
    op2.allocate_jacobian()
    rc = L.load().mhd_operator_get_csr(op2.handle, None, None, 3, 0)
    assert rc == -1
    bad = FluidParams(convection="newton").to_c()
    bad.convection = 9
    import ctypes as C

    assert L.load().mhd_operator_set_params(op2.handle, C.byref(bad)) == -1
    op.destroy()
    op2.destroy()


@pytest.mark.parametrize("solver,order", [("badia2024", ("u", "j", "p", "phi")), ("li2019", ("j", "u", "p", "phi"))])
def test_nonuniform_and_block_layout(mhdlib, solver, order):
    """Rectangular cell counts and the block layouts of `_multi_field_style` (src/fespaces.jl:4-9): badia2024 = ([u,j],p,phi),
    li2019 = (j,u,p,phi)."""
    from oracle import mhd_oracle as O

    params, fes = make_case(nc=(5, 3), B=(0.0, 50.0, 0.0), solver=solver)
    assert fes.field_order == order
    op = B200FEOperator(fes, params["fluid"])
    x = np.random.default_rng(11).random(fes.ndofs)
    A = op.jacobian(x)
    Ao = O.jacobian(fes, x, oracle_params(params["fluid"]))
    rowptr, colval = A.pattern()
    assert np.array_equal(rowptr, Ao.indptr) and np.array_equal(colval, Ao.indices)
    assert relerr(A.nzval(), Ao.data) < VAL_TOL
    assert relerr(op.residual(x), O.residual(fes, x, oracle_params(params["fluid"]))) < VAL_TOL
    v = np.random.default_rng(12).standard_normal(fes.ndofs)
    assert relerr(op.spmv(v), Ao @ v) < VAL_TOL
    op.destroy()


def test_golden_fixture(mhdlib):
    """Committed golden vectors (tests/golden/make_golden.py, generated with the oracle)."""
    import os

    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "hunt_nc3_ha20.npz"))
    params, fes = make_case(nc=(3, 3), B=(0.0, 20.0, 0.0))
    op = B200FEOperator(fes, params["fluid"])
    A = op.jacobian(g["x"])
    rowptr, colval = A.pattern()
    assert np.array_equal(rowptr, g["rowptr"]) and np.array_equal(colval, g["colval"])
    assert relerr(A.nzval(), g["nzval"]) < VAL_TOL
    assert relerr(op.residual(g["x"]), g["residual"]) < VAL_TOL
    assert relerr(op.spmv(g["v"]), g["Av"]) < VAL_TOL
    op.destroy()


def test_hunt_solve_matches_oracle_and_published_norms(mhdlib):
    """Hunt Ha=50, nc=(10,10) on the kmap=1 mesh of the published Gadi runs (hconv_ha00050ns500/summary.csv:7):
    Newton + device FGMRES with the Badia2024 block-triangular preconditioner (augmented Lagrangian zeta=20, exact
    device LU of the (u,j) block, cell-block mass inverses for p and phi) against
      * the oracle's sparse-LU solution: u and j within 1e-10 relative (p only up to its constant null mode), and
      * the reference's published 16-digit norms uh_l2, uh_h1, jh_l2.
    The AL terms vanish at the discrete solution (Pi_p div u_h = 0, div j_h = 0), so zeta does not change it."""
    from gridapmhd_jl_b200.feoperator import B200LinearSolver, B200SolverOptions, NewtonSolver
    from gridapmhd_jl_b200.host.reffe import make_tables
    from oracle import mhd_oracle as O

    Ha = 50.0
    params, fes = make_case(nc=(10, 10), B=(0.0, Ha, 0.0), BL_adapted=False, solver="badia2024", zeta_u=20.0, zeta_j=20.0)
    op = B200FEOperator(fes, params["fluid"])
    opts = B200SolverOptions(m=30, maxiter=30, rtol=1e-13, atol=1e-30, precond="block_tri", uj_solver="dense_lu")
    # two extra Newton steps act as iterative refinement of the (ill-conditioned, zeta-augmented) linear solve
    # (stop once below 1e-12 |r0|: further steps only move the iterate around inside the rounding ball of the residual)
    nls = NewtonSolver(B200LinearSolver(opts), maxiter=3, rtol=1e-12)
    x = nls.solve_b(np.zeros(fes.ndofs), op)
    assert nls.log[-1] < 1e-9 * nls.log[0], nls.log
    params0, fes0 = make_case(nc=(10, 10), B=(0.0, Ha, 0.0), BL_adapted=False, solver="badia2024")
    xo, _ = O.newton_lu(fes0, oracle_params(params0["fluid"]), min_iters=3)  # 2 refinement steps
    s, so = fes.split(x), fes0.split(xo)
    assert relerr(s["u"], so["u"]) < SOL_TOL
    assert relerr(s["j"], so["j"]) < SOL_TOL
    # p is defined up to a constant on Hunt (no pressure constraint)
    dp = s["p"] - so["p"]
    assert np.abs(dp - dp.mean()).max() < 1e-6 * max(1.0, np.abs(so["p"]).max())
    nr = O.solution_norms(fes, x, make_tables(6), u0=1.0, jscale=Ha)
    pins = dict(uh_l2=0.001125968494949451, uh_h1=0.009386206346670825, jh_l2=0.019669872964491745)
    for k, v in pins.items():
        assert abs(nr[k] - v) / v < 1e-9, (k, nr[k], v)
    op.destroy()


def test_fgmres_jacobi_inner_solver_reduces_residual(mhdlib):
    """The scalable configuration (inner Jacobi-GMRES on the (u,j) block, no dense factorisation): the device
    FGMRES must reduce the residual monotonically and report non-convergence through MHD_E_NOTCONV."""
    from gridapmhd_jl_b200.feoperator import B200LinearSolver, B200SolverOptions

    params, fes = make_case(nc=(4, 4), B=(0.0, 10.0, 0.0), solver="badia2024", zeta_u=1.0, zeta_j=1.0)
    op = B200FEOperator(fes, params["fluid"])
    x0 = np.zeros(fes.ndofs)
    A = op.allocate_jacobian()
    b = np.empty(op.nrows)
    op.residual_and_jacobian_b(b, A, x0)
    ns = B200LinearSolver(B200SolverOptions(m=30, maxiter=60, rtol=1e-12, atol=0.0, precond="block_tri", uj_inner_its=30,
                                            uj_inner_restart=30)).symbolic_setup(A).numerical_setup()
    dx = np.zeros(op.nrows)
    ns.solve_b(dx, -b)
    h = ns.history
    assert ns.iters == 60 and len(h) == 61
    assert np.all(np.diff(h) <= 1e-12 * h[0]) and h[-1] < 1e-2 * h[0]
    As = A.to_scipy()
    assert abs(np.linalg.norm(As @ dx + b) - ns.resnorm) < 1e-6 * h[0]  # reported residual is the true residual
    with pytest.raises(Exception):
        ns.solve_b(np.zeros(op.nrows), -b, raise_on_maxiter=True)
    ns.destroy()
    op.destroy()


def test_multigpu_parity_when_two_gpus_are_visible(mhdlib):
    """Spawns tests/multigpu_check.py under torchrun with 2 ranks (skipped on a single-GPU box; `gpurun --gpus 2`)."""
    import os
    import subprocess
    import sys

    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    here = os.path.dirname(os.path.abspath(__file__))
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
                          "127.0.0.1", "--master-port", "29511", os.path.join(here, "multigpu_check.py")],
                         capture_output=True, text=True, timeout=600)
    assert "MULTIGPU_OK 2" in out.stdout, out.stdout[-3000:] + out.stderr[-3000:]


def test_expansion_nonaffine_mesh_values(mhdlib):
    """Expansion configuration (BASELINE configs 3/4 stand-in): p4est base mesh of expansion_mesher.jl refined once,
    interior vertices perturbed -> general trilinear (non-affine) hexes, inhomogeneous inlet Dirichlet data,
    j.n = 0 on the whole boundary, Newton convection."""
    from gridapmhd_jl_b200.applications import expansion_params
    from oracle import mhd_oracle as O

    params = expansion_params(level=1, Ha=10.0, N=5.0, perturb=0.2, zeta_u=3.0, zeta_j=2.0)
    fes = setup_spaces(params)
    assert fes.mesh.ncells == 96 and np.abs(fes.dirichlet_values["u"]).max() > 1.0
    op = B200FEOperator(fes, params["fluid"])
    x = np.random.default_rng(5).random(fes.ndofs)
    A = op.allocate_jacobian()
    b = np.empty(op.nrows)
    op.residual_and_jacobian_b(b, A, x)
    Ao = O.jacobian(fes, x, oracle_params(params["fluid"]))
    rowptr, colval = A.pattern()
    assert np.array_equal(rowptr, Ao.indptr) and np.array_equal(colval, Ao.indices)
    assert relerr(A.nzval(), Ao.data) < VAL_TOL
    assert relerr(b, O.residual(fes, x, oracle_params(params["fluid"]))) < VAL_TOL
    assert relerr(op.residual(x), b) < 1e-13  # two kernels, different summation orders
    op.destroy()


def test_expansion_newton_solve_matches_oracle(mhdlib):
    """Full Newton solve of the (nonlinear, convection :newton) Expansion problem on the device against the oracle's
    Newton + sparse LU on the same discrete problem: u, j within 1e-10; phi up to its constant null mode
    (j.n = 0 on all of the boundary, no phi constraint: SURVEY.md Appendix G)."""
    from gridapmhd_jl_b200.applications import expansion_params
    from gridapmhd_jl_b200.feoperator import B200LinearSolver, B200SolverOptions, NewtonSolver
    from oracle import mhd_oracle as O

    params = expansion_params(level=1, Ha=10.0, N=5.0, perturb=0.15, zeta_u=20.0, zeta_j=20.0, solver="badia2024")
    fes = setup_spaces(params)
    op = B200FEOperator(fes, params["fluid"])
    opts = B200SolverOptions(m=40, maxiter=40, rtol=1e-13, atol=1e-30, precond="block_tri", uj_solver="dense_lu")
    # rtol 1e-12: stop at the end of the quadratic phase.  Iterating on (rtol 1e-15 is never met) makes the iterate wander
    # inside the rounding ball of the residual (atomics order differs from run to run): 5e-12 typically, 2e-10 at worst.
    nls = NewtonSolver(B200LinearSolver(opts), maxiter=12, rtol=1e-12)
    x = nls.solve_b(np.zeros(fes.ndofs), op)
    assert len(nls.log) >= 5 and nls.log[-1] < 1e-11 * nls.log[0], nls.log  # genuinely nonlinear: several Newton steps
    xo, _ = O.newton_lu(fes, oracle_params(params["fluid"]), maxiter=12, rtol=1e-15, min_iters=8)
    s, so = fes.split(x), fes.split(xo)
    assert relerr(s["u"], so["u"]) < SOL_TOL
    assert relerr(s["j"], so["j"]) < SOL_TOL
    assert relerr(s["p"], so["p"]) < 1e-7
    df = s["phi"] - so["phi"]
    assert np.abs(df - df.mean()).max() < 1e-7 * max(1.0, np.abs(so["phi"]).max())
    op.destroy()


def test_hunt_solid_walls_values_and_solve(mhdlib):
    """Hunt with conducting solid walls (reference test/seq/hunt_tests.jl:74-88: nc=(12,12), tw=0.2, kmap=3): u, p live
    on the fluid cells only, solid cells use jac/res_solid_h1_hdiv (weakforms.jl:314-338) with per-cell sigma."""
    from gridapmhd_jl_b200.feoperator import B200LinearSolver, B200SolverOptions, NewtonSolver
    from oracle import mhd_oracle as O

    params, fes = make_case(nc=(12, 12), B=(0.0, 50.0, 0.0), tw=0.2, BL_adapted=False, kmap_x=3, kmap_y=3, solver="badia2024",
                            zeta_u=20.0, zeta_j=20.0)
    assert fes.cell_solid.sum() == 132 and fes.nfree["u"] == 6498 and fes.nfree["p"] == 1200
    op = B200FEOperator(fes, params["fluid"])
    x = np.random.default_rng(21).random(fes.ndofs)
    A = op.allocate_jacobian()
    b = np.empty(op.nrows)
    op.residual_and_jacobian_b(b, A, x)
    Ao = O.jacobian(fes, x, oracle_params(params["fluid"]))
    rowptr, colval = A.pattern()
    assert np.array_equal(rowptr, Ao.indptr) and np.array_equal(colval, Ao.indices)
    assert relerr(A.nzval(), Ao.data) < VAL_TOL
    assert relerr(b, O.residual(fes, x, oracle_params(params["fluid"]))) < VAL_TOL
    assert relerr(op.residual(x), b) < 1e-13  # two kernels, different summation orders
    # solve (linear: the Hunt flow has no convection contribution)
    opts = B200SolverOptions(m=30, maxiter=30, rtol=1e-13, atol=1e-30, precond="block_tri", uj_solver="dense_lu")
    nls = NewtonSolver(B200LinearSolver(opts), maxiter=6, rtol=1e-12)  # extra steps = iterative refinement, until 1e-12 |r0|
    xs = nls.solve_b(np.zeros(fes.ndofs), op)
    xo, _ = O.newton_lu(fes, oracle_params(params["fluid"]), min_iters=6)
    s, so = fes.split(xs), fes.split(xo)
    assert relerr(s["u"], so["u"]) < SOL_TOL
    # the current is 3 orders of magnitude smaller than u x B here (conducting walls short-circuit it): measure its
    # error against the scale of the terms it balances in Ohm's law, sigma (u x B)
    assert np.abs(s["j"] - so["j"]).max() < SOL_TOL * max(np.abs(so["j"]).max(), np.abs(so["u"]).max())
    print("solid: rel err j (own scale) =", relerr(s["j"], so["j"]))
    op.destroy()


@pytest.mark.gpu
def test_hunt_post_processing_norms_match_oracle(mhdlib):
    """Device post-processing of hunt (hunt.jl:239-260): discrete norms and errors against the analytical Hunt series
    (hunt.jl:372-457), degree-6 quadrature, against the oracle (itself pinned to the published error norms, K6)."""
    from gridapmhd_jl_b200.host import reffe
    from oracle import mhd_oracle as O

    Ha = 30.0
    params = hunt_params(nc=(5, 4), B=(0.0, Ha, 0.0))
    fes = setup_spaces(params)
    op = B200FEOperator(fes, params["fluid"])
    # a smooth-ish state of the right magnitude plus noise: every term of the integrands is exercised
    x = 1e-3 * np.random.default_rng(11).random(fes.ndofs)
    T6 = reffe.make_tables(6)
    for nsums in (0, 7, 200):
        got = op.hunt_error_norms(x, T6, Ha, nsums, u0=1.3, jscale=Ha)
        want = O.hunt_error_norms(fes, x, T6, Ha, nsums, u0=1.3, jscale=Ha)
        want.update(O.solution_norms(fes, x, T6, u0=1.3, jscale=Ha))
        for k, v in want.items():
            assert abs(got[k] - v) <= 1e-11 * abs(v), (nsums, k, got[k], v)
    op.destroy()


@pytest.mark.gpu
def test_hunt_driver_end_to_end_reproduces_published_row(mhdlib):
    """`hunt(nc=(10,10), B=(0,50,0), nsums=500)` through the driver mirror: device Newton/FGMRES solve, then the device
    post-processing; all six norms of the reference's published run ha00050cx010 (hconv_ha00050ns500/summary.csv:7,
    kmap=1 = unstretched mesh) are reproduced from end to end."""
    from gridapmhd_jl_b200.applications import hunt
    from gridapmhd_jl_b200.feoperator import B200SolverOptions

    opts = B200SolverOptions(m=30, maxiter=30, rtol=1e-13, atol=1e-30, precond="block_tri", uj_solver="dense_lu")
    info, out = hunt(nc=(10, 10), B=(0.0, 50.0, 0.0), BL_adapted=False, solver="badia2024", zeta_u=20.0, zeta_j=20.0,
                     nsums=500, solve=True, solver_opts=opts, newton_maxiter=3, newton_rtol=1e-12)
    assert info["ndofs"] == 17298 and info["Ha"] == 50.0
    pins = dict(eu_l2=6.274034420523594e-5, eu_h1=0.0021489743289400043, ej_l2=0.0011055397108926523,
                uh_l2=0.001125968494949451, uh_h1=0.009386206346670825, jh_l2=0.019669872964491745)
    for k, v in pins.items():
        assert abs(info[k] - v) / v < 1e-8, (k, info[k], v)
    assert info["time_post_process"] < 5.0
    out["op"].destroy()


def test_hunt_nc64_ha500_device_solve_reproduces_published_row(mhdlib):
    """BASELINE config 2 size, end to end on the device with the repo's OWN preconditioner (no vendor LU): `hunt(nc=(64,64),
    B=(0,500,0))` on the kmap=1 mesh of the published run ha00500cx064 (analysis/gadi/results/2023_04/eaab9d14.../
    hconv_ha00500ns100/summary.csv:2 -- 732 690 dofs, 12 288 cells): Newton + FGMRES(30) + Badia2024 block-triangular
    preconditioner whose (u,j) block is an inner GMRES(30) preconditioned by the vertex-patch block-Jacobi smoother
    (src/Solvers/badia2024.jl:2-48 with the smoother of gmg.jl:62-81 in place of the direct block solver), augmented
    Lagrangian zeta = 10 (it vanishes at the discrete solution, so the published zeta = 0 norms are the target).
    DOF counts exact; uh_l2, uh_h1, jh_l2 and the error norms against the analytical series to 1e-8 relative."""
    from gridapmhd_jl_b200.applications import hunt
    from gridapmhd_jl_b200.feoperator import B200SolverOptions

    opts = B200SolverOptions(m=30, maxiter=120, rtol=1e-11, atol=1e-30, precond="block_tri", uj_solver="gmres_patch", uj_inner_its=30,
                             uj_inner_restart=30, patch_its=1, patch_omega=1.0)
    info, out = hunt(nc=(64, 64), B=(0.0, 500.0, 0.0), BL_adapted=False, solver="badia2024", zeta_u=10.0, zeta_j=10.0, nsums=100,
                     solve=True, solver_opts=opts, newton_maxiter=3, newton_rtol=1e-10)
    assert info["ndofs"] == 732690 and info["ndofs_u"] == 290322 and info["ndofs_j"] == 294912 and info["Ha"] == 500.0
    log = out["newton_log"]
    assert log[-1] <= 1e-9 * log[0], log
    pins = dict(uh_l2=5.831116781906231e-5, uh_h1=0.0016385112887929432, jh_l2=0.0030907363432802868,
                eu_l2=5.200746404046955e-7, eu_h1=9.23675484735579e-5, ej_l2=5.536545781986553e-5)
    for k, v in pins.items():
        assert abs(info[k] - v) / v < (1e-8 if k.endswith("h_l2") or k == "uh_h1" else 1e-5), (k, info[k], v)
    out["op"].destroy()


def test_expansion_6k_reference_mesh_assembly_and_spmv(mhdlib):
    """BASELINE configs 3/4 on the reference's own mesh: `meshes/Expansion_6k.msh` (5 317 nodes, 4 320 NON-AFFINE hexes; shipped
    as the reader-independent fixture tests/golden/expansion_6k_mesh.npz, written by tests/golden/make_golden.py expansion6k)
    with the Expansion parameterisation of src/Applications/expansion.jl:40-181 (:mhd scaling, Ha = 100, N = 3740 as in
    Turgalium_CIEMAT/sendExpansion.sh:49-56, parabolic inlet profile as Dirichlet data, Newton convection): CSR structure
    bit-exact, Jacobian and residual <= 1e-12 against the C oracle on EVERY row, SpMV against a host product."""
    import os

    from gridapmhd_jl_b200.applications import expansion_params
    from gridapmhd_jl_b200.host import mesh as M
    from oracle.parity import assembly_parity, spmv_parity

    m = M.load_mesh_npz(os.path.join(os.path.dirname(__file__), "golden", "expansion_6k_mesh.npz"))
    assert m.ncells == 4320 and m.coords.shape[0] == 5317
    params = expansion_params(Ha=100.0, N=3740.0, zeta_u=10.0, zeta_j=10.0, mesh=m)
    fes = setup_spaces(params)
    assert np.abs(fes.dirichlet_values["u"]).max() > 1.0  # the inlet profile is in
    fl = params["fluid"]
    op = B200FEOperator(fes, fl)
    assert op.kernel_version == 7
    x = np.random.default_rng(6).random(fes.ndofs)
    A = op.allocate_jacobian()
    b = np.empty(op.nrows)
    op.residual_and_jacobian_b(b, A, x)
    rowptr, colval = A.pattern()
    nz = A.nzval()
    par = assembly_parity(fes, oracle_params(fl), x, rowptr, colval, nz, b, op.nrows, ncells=m.ncells)
    assert par["cells"] == 4320 and par["rows_checked"] == op.nrows, par
    assert par["csr_bitexact"] and par["jac_rel"] < VAL_TOL and par["res_rel"] < VAL_TOL, par
    v = np.random.default_rng(7).standard_normal(fes.ndofs)
    assert spmv_parity(rowptr, colval, nz, v, op.spmv(v)) < VAL_TOL
    assert relerr(op.residual(x), b) < 1e-13
    op.destroy()
