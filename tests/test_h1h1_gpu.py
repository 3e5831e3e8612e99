"""GPU parity tests of the H1-H1 formulation through the C ABI (mhd_h1h1_operator_create and the
shared mhd_operator_* / mhd_jacobian / mhd_residual / mhd_spmv entry points) against the NumPy oracle
(oracle/mhd_oracle_h1h1.py).  Tolerances as in BASELINE.json: CSR structure bit-exact, values 1e-12 relative, solutions
1e-10 relative."""
import numpy as np
import pytest
import scipy.sparse.linalg as spla

from gridapmhd_jl_b200.applications import hunt_params, make_operator, setup_spaces, u_inlet_parabolic
from gridapmhd_jl_b200.feoperator import B200H1H1FEOperator, B200LinearSolver, B200SolverOptions, FluidParams
from gridapmhd_jl_b200.host import mesh as M
from gridapmhd_jl_b200.host.fespaces_h1h1 import setup_fe_spaces_h1h1
from oracle import mhd_oracle as O
from oracle import mhd_oracle_h1h1 as H

pytestmark = pytest.mark.gpu
VAL_TOL = 1e-12


def relerr(a, b):
    return np.abs(a - b).max() / np.abs(b).max()


def oprm(fl: FluidParams):
    return O.FluidParams(fl.alpha, fl.beta, fl.gamma, fl.sigma, fl.zeta_u, fl.zeta_j, fl.B, fl.f, fl.g, fl.convection)


@pytest.fixture(scope="module")
def cfg1(mhdlib):
    p = hunt_params(nc=(4, 4), B=(0.0, 10.0, 0.0), current_disc="H1")
    fes = setup_spaces(p)
    return p, fes


def test_h1h1_symbolic_structure_bit_exact(cfg1):
    p, fes = cfg1
    op = make_operator(fes, p["fluid"])
    assert isinstance(op, B200H1H1FEOperator)
    A = op.allocate_jacobian()
    rowptr, colval = A.pattern()
    rp, cv = H.symbolic_csr(fes.cell_global_ids(), fes.ndofs)
    assert op.nrows == op.ncols == fes.ndofs == 2361
    assert np.array_equal(rowptr, rp) and np.array_equal(colval, cv)
    r32, c32 = A.pattern(index_bytes=4, base=1)
    assert np.array_equal(r32, rp + 1) and np.array_equal(c32, cv + 1)
    nent, nexcl = op.scatter_stats()
    gids = fes.cell_global_ids()
    li, lj = np.nonzero(H.touched_mask())
    assert nent == int(((gids[:, li] >= 0) & (gids[:, lj] >= 0)).sum()) and 0 < nexcl < nent
    op.destroy()


@pytest.mark.parametrize("conv,zu", [("none", 0.0), ("picard", 0.0), ("newton", 0.0), ("none", 10.0), ("newton", 10.0)])
def test_h1h1_jacobian_residual_values(cfg1, conv, zu):
    p, fes = cfg1
    fl = FluidParams(alpha=1.0, beta=1.0, gamma=100.0, zeta_u=zu, B=(0.1, 1.0, -0.2), f=(0.0, 0.3, 1.0), convection=conv)
    op = make_operator(fes, fl)
    x = np.random.default_rng(1234).random(fes.ndofs)
    A = op.jacobian(x)
    Ao = H.jacobian(fes, x, oprm(fl))
    rowptr, colval = A.pattern()
    assert np.array_equal(rowptr, Ao.indptr) and np.array_equal(colval, Ao.indices)
    assert relerr(A.nzval(), Ao.data) < VAL_TOL
    ro = H.residual(fes, x, oprm(fl))
    assert relerr(op.residual(x), ro) < VAL_TOL
    # block-wise relative error: the small blocks are not hidden behind gamma-scaled ones
    o = fes.offsets
    S = A.to_scipy()
    for f1 in ("u", "p", "phi"):
        for f2 in ("u", "p", "phi"):
            blk = Ao[o[f1] : o[f1] + fes.nfree[f1], o[f2] : o[f2] + fes.nfree[f2]]
            if blk.nnz and np.abs(blk.data).max() > 0:
                mine = S[o[f1] : o[f1] + fes.nfree[f1], o[f2] : o[f2] + fes.nfree[f2]]
                assert np.abs((mine - blk).data).max() <= VAL_TOL * np.abs(blk.data).max() if (mine - blk).nnz else True
    # fused entry point (two launches for this formulation) and a second assembly on the same handle (values are re-zeroed)
    b = np.empty(op.nrows)
    op.residual_and_jacobian_b(b, A, x)
    assert relerr(b, ro) < VAL_TOL and relerr(A.nzval(), Ao.data) < VAL_TOL
    # Krylov building blocks on the H1-H1 matrix
    v = np.random.default_rng(5).standard_normal(fes.ndofs)
    assert relerr(op.spmv(v), Ao @ v) < VAL_TOL
    assert abs(op.dot(v, x) - v @ x) < 1e-12 * abs(v @ x)
    op.destroy()


def test_h1h1_device_pointer_path(cfg1):
    import torch

    p, fes = cfg1
    op = make_operator(fes, p["fluid"])
    x = np.random.default_rng(3).random(fes.ndofs)
    xd = torch.from_numpy(x).cuda()
    A = op.jacobian(xd)
    rd = op.residual(xd)
    torch.cuda.synchronize()
    Ao = H.jacobian(fes, x, oprm(p["fluid"]))
    assert relerr(A.nzval(), Ao.data) < VAL_TOL
    assert relerr(rd.cpu().numpy(), H.residual(fes, x, oprm(p["fluid"]))) < VAL_TOL
    op.destroy()


def test_h1h1_nonaffine_mesh_with_dirichlet_data(mhdlib):
    m = M.expansion_generate_mesh(1, perturb=0.15, seed=2)
    fes = setup_fe_spaces_h1h1(m, u_tags=("inlet", "wall"), u_values=(u_inlet_parabolic(), None), phi_tags=("outlet",),
                               phi_values=(lambda X: 1.0 + X[:, 1],))
    fl = FluidParams(alpha=0.2, beta=0.01, gamma=1.0, zeta_u=5.0, B=(0.0, 1.0, 0.0), convection="newton")
    op = make_operator(fes, fl)
    x = np.random.default_rng(7).random(fes.ndofs)
    A = op.jacobian(x)
    Ao = H.jacobian(fes, x, oprm(fl))
    rowptr, colval = A.pattern()
    assert np.array_equal(rowptr, Ao.indptr) and np.array_equal(colval, Ao.indices)
    assert relerr(A.nzval(), Ao.data) < VAL_TOL
    assert relerr(op.residual(x), H.residual(fes, x, oprm(fl))) < VAL_TOL
    op.destroy()


def test_h1h1_hunt_newton_solve_matches_oracle(mhdlib):
    """`hunt(nc=(6,6), B=(0,10,0), current_disc=:H1, solver=:julia)` (test/seq/hunt_tests.jl:27-37 at a smaller mesh):
    Newton with the device residual/Jacobian; the linear solves are the host's sparse LU on the device-assembled matrix,
    as `_solver(::Val{:julia})` (src/main.jl:181-186) does in the reference.  Solution within 1e-10 of the oracle's."""
    p = hunt_params(nc=(6, 6), B=(0.0, 10.0, 0.0), current_disc="H1")
    fes = setup_spaces(p)
    op = make_operator(fes, p["fluid"])
    A = op.allocate_jacobian()
    x = np.zeros(fes.ndofs)
    b = np.empty(op.nrows)
    op.residual_and_jacobian_b(b, A, x)
    r0 = np.linalg.norm(b)
    for _ in range(10):
        x += spla.splu(A.to_scipy().tocsc()).solve(-b)
        op.residual_and_jacobian_b(b, A, x)
        if np.linalg.norm(b) <= 1e-6 * r0:
            break
    assert np.linalg.norm(b) <= 1e-6 * r0
    xo, _ = H.newton_lu(fes, oprm(p["fluid"]))
    o = fes.offsets
    for f in ("u", "phi"):  # p is defined up to a constant on the enclosed duct
        s = slice(o[f], o[f] + fes.nfree[f])
        assert relerr(x[s], xo[s]) < 1e-10
    op.destroy()


def test_h1h1_device_fgmres_reduces_the_residual(cfg1):
    """the device FGMRES (MHD_PC_JACOBI) runs on an H1-H1 handle; the block-triangular preconditioner is refused"""
    from gridapmhd_jl_b200.lib import MhdError

    p, fes = cfg1
    fl = FluidParams(alpha=1.0, beta=1.0, gamma=1.0, zeta_u=0.0, B=(0.0, 1.0, 0.0), f=(0.0, 0.0, 1.0), convection="none")
    op = make_operator(fes, fl)
    A = op.jacobian(np.zeros(fes.ndofs))
    b = np.random.default_rng(0).standard_normal(fes.ndofs)
    ns = B200LinearSolver(B200SolverOptions(m=30, maxiter=60, rtol=1e-3, precond="jacobi")).symbolic_setup(A).numerical_setup()
    x = np.zeros(fes.ndofs)
    ns.solve_b(x, b)
    hist = ns.history
    assert len(hist) > 2 and hist[-1] < hist[0]
    res = np.linalg.norm(b - A.to_scipy() @ x)
    assert abs(res - hist[-1]) < 1e-6 * hist[0]
    ns.destroy()
    with pytest.raises(MhdError):
        B200LinearSolver(B200SolverOptions(precond="block_tri")).symbolic_setup(A).numerical_setup()
    op.destroy()


def test_h1h1_block_preconditioner_with_patch_solvers_converges(mhdlib):
    """H1H1BlockSolver (src/Solvers/h1h1blocks.jl:2-43; hunt(..., solver=:h1h1blocks, zeta_u=10) of test/seq/hunt_tests.jl:90-103 at a
    smaller mesh): outer FGMRES with the upper block-triangular preconditioner over (u,p,phi) whose u and phi blocks are
    inner GMRES(30) with vertex-patch solvers.  The NumPy restatement of the same algorithm converges to 1e-8 in 19 outer
    iterations; solution against sparse LU of the device matrix."""
    p = hunt_params(nc=(6, 6), B=(0.0, 20.0, 0.0), zeta_u=10.0, current_disc="H1")
    fes = setup_spaces(p)
    op = make_operator(fes, p["fluid"])
    A = op.allocate_jacobian()
    b = np.empty(op.nrows)
    op.residual_and_jacobian_b(b, A, np.zeros(fes.ndofs))
    opts = B200SolverOptions(m=30, maxiter=60, rtol=1e-8, atol=0.0, precond="h1h1_blocks", uj_solver="gmres_patch", uj_inner_its=30,
                             uj_inner_restart=30)
    ns = B200LinearSolver(opts).symbolic_setup(A).numerical_setup()
    dx = np.zeros(op.nrows)
    ns.solve_b(dx, -b, raise_on_maxiter=True)
    h = ns.history
    assert ns.iters <= 40 and h[-1] <= 1e-8 * h[0]
    As = A.to_scipy()
    assert abs(np.linalg.norm(As @ dx + b) - ns.resnorm) < 1e-6 * h[0]
    xo = spla.splu(As.tocsc()).solve(-b)
    o = fes.offsets
    for f in ("u", "phi"):
        sl = slice(o[f], o[f] + fes.nfree[f])
        assert relerr(dx[sl], xo[sl]) < 1e-5
    ns.destroy()
    # the Jacobi variant of the block solvers runs too (and is much weaker)
    nsj = B200LinearSolver(B200SolverOptions(m=30, maxiter=30, rtol=1e-8, atol=0.0, precond="h1h1_blocks", uj_solver="gmres_jacobi",
                                             uj_inner_its=30, uj_inner_restart=30)).symbolic_setup(A).numerical_setup()
    dj = np.zeros(op.nrows)
    nsj.solve_b(dj, -b)
    assert nsj.history[-1] < nsj.history[0] and nsj.history[-1] > h[-1]
    nsj.destroy()
    op.destroy()


def test_h1h1_hunt_driver_solves_on_the_device(mhdlib):
    """`main(params)` for `current_disc = :H1`: Newton + device FGMRES with the H1-H1 block preconditioner (no host solve);
    u and phi within 1e-5 of the oracle's Newton/LU solution"""
    from gridapmhd_jl_b200.applications import main

    p = hunt_params(nc=(6, 6), B=(0.0, 20.0, 0.0), zeta_u=10.0, current_disc="H1")
    out = main(p, newton_rtol=1e-9)
    fes = out["fes"]
    assert out["newton_log"][-1] <= 1e-8 * out["newton_log"][0], out["newton_log"]
    xo, _ = H.newton_lu(fes, oprm(p["fluid"]), rtol=1e-12)
    o = fes.offsets
    for f in ("u", "phi"):
        sl = slice(o[f], o[f] + fes.nfree[f])
        assert relerr(out["x"][sl], xo[sl]) < 1e-5
    out["op"].destroy()


def test_hunt_driver_with_current_disc_h1_returns_its_info_dict(mhdlib):
    """`hunt(current_disc=:H1, ...)` end to end (hunt.jl:212-303 subset): dof counts of the three fields, timings, solve log"""
    from gridapmhd_jl_b200.applications import hunt

    info, out = hunt(nc=(4, 4), B=(0.0, 20.0, 0.0), zeta_u=10.0, current_disc="H1", newton_rtol=1e-8)
    assert info["ndofs_u"] + info["ndofs_p"] + info["ndofs_phi"] == info["ndofs"] and "ndofs_j" not in info
    assert "skipped" in info["post_process"] and info["time_solve"] > 0.0
    assert out["newton_log"][-1] <= 1e-7 * out["newton_log"][0]
    out["op"].destroy()


def test_h1h1_golden_fixture(mhdlib):
    """Committed golden vectors (tests/golden/make_golden.py::main_h1h1, generated with the oracle)."""
    import os

    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "hunt_h1h1_nc2_ha20.npz"))
    p = hunt_params(nc=(2, 2), B=(0.0, 20.0, 0.0), zeta_u=5.0, current_disc="H1")
    fes = setup_spaces(p)
    op = make_operator(fes, p["fluid"])
    A = op.jacobian(g["x"])
    rowptr, colval = A.pattern()
    assert np.array_equal(rowptr, g["rowptr"]) and np.array_equal(colval, g["colval"])
    assert relerr(A.nzval(), g["nzval"]) < VAL_TOL
    assert relerr(op.residual(g["x"]), g["residual"]) < VAL_TOL
    assert relerr(op.spmv(g["v"]), g["Av"]) < VAL_TOL
    op.destroy()
