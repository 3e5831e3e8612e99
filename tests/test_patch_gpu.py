"""GPU tests of the vertex-patch block-Jacobi smoother of the (u,j) block (gmg_block_jacobi_smoothers,
src/Solvers/gmg.jl:62-81) through the C ABI: mhd_solver_set_patches / mhd_solver_setup / mhd_solver_patch_apply / mhd_solve."""
import numpy as np
import pytest
import scipy.sparse.linalg as spla

from gridapmhd_jl_b200.applications import hunt_params, setup_spaces
from gridapmhd_jl_b200.feoperator import B200FEOperator, B200LinearSolver, B200SolverOptions
from gridapmhd_jl_b200.host.patches import vertex_patches
from oracle import mhd_oracle as O

pytestmark = pytest.mark.gpu


def relerr(a, b):
    return np.abs(a - b).max() / np.abs(b).max()


@pytest.fixture(scope="module")
def case(mhdlib):
    p = hunt_params(nc=(6, 6), B=(0.0, 20.0, 0.0), solver="badia2024", zeta_u=10.0, zeta_j=10.0)
    fes = setup_spaces(p)
    fl = p["fluid"]
    prm = O.FluidParams(fl.alpha, fl.beta, fl.gamma, fl.sigma, fl.zeta_u, fl.zeta_j, fl.B, fl.f, fl.g, fl.convection)
    return p, fes, prm


def test_patch_apply_matches_the_oracle_additive_solver(case):
    """device gather + blocked Gauss-Jordan inverses + batched apply against numpy on the oracle's matrix, for the Jacobian at
    a random state (Newton convection: non-symmetric patches)"""
    p, fes, prm = case
    op = B200FEOperator(fes, p["fluid"])
    x = 0.1 * np.random.default_rng(0).random(fes.ndofs)
    A = op.jacobian(x)
    ns = B200LinearSolver(B200SolverOptions(precond="block_tri", uj_solver="gmres_patch")).symbolic_setup(A).numerical_setup()
    assert ns.npatches == fes.mesh.nverts
    nuj = fes.nfree["u"] + fes.nfree["j"]
    Ao = O.jacobian(fes, x, prm).tocsr()
    Auj = Ao[:nuj, :nuj].tocsr()
    ptr, dofs = vertex_patches(fes)
    r = np.random.default_rng(1).standard_normal(nuj)
    z = np.zeros(nuj)
    for k in range(len(ptr) - 1):
        q = dofs[ptr[k] : ptr[k + 1]]
        z[q] += np.linalg.solve(Auj[q][:, q].toarray(), r[q])
    zd = ns.patch_apply(r, omega=0.5)
    assert relerr(zd, 0.5 * z) < 1e-9
    # numerical_setup! after a new Jacobian refreshes the inverses
    x2 = 0.2 * np.random.default_rng(2).random(fes.ndofs)
    op.jacobian_b(A, x2)
    ns.numerical_setup_b(A)
    Auj2 = O.jacobian(fes, x2, prm).tocsr()[:nuj, :nuj].tocsr()
    z2 = np.zeros(nuj)
    for k in range(len(ptr) - 1):
        q = dofs[ptr[k] : ptr[k + 1]]
        z2[q] += np.linalg.solve(Auj2[q][:, q].toarray(), r[q])
    assert relerr(ns.patch_apply(r), z2) < 1e-9
    ns.destroy()
    op.destroy()


@pytest.mark.parametrize("patch_its,omega", [(1, 1.0), (3, 0.2)])
def test_block_tri_fgmres_with_patch_smoother_converges(case, patch_its, omega):
    """the configuration without any dense / direct (u,j) solve: outer FGMRES + block-triangular preconditioner whose (u,j)
    block is an inner GMRES(30) preconditioned by the patch smoother.  Converges to 1e-8 (the NumPy restatement of the same
    algorithm needs 11 outer iterations; point-Jacobi stagnates, see test_fgmres_jacobi_inner_solver_reduces_residual)."""
    p, fes, prm = case
    op = B200FEOperator(fes, p["fluid"])
    x0 = np.zeros(fes.ndofs)
    A = op.allocate_jacobian()
    b = np.empty(op.nrows)
    op.residual_and_jacobian_b(b, A, x0)
    opts = B200SolverOptions(m=30, maxiter=30, rtol=1e-8, atol=0.0, precond="block_tri", uj_solver="gmres_patch", uj_inner_its=30,
                             uj_inner_restart=30, patch_its=patch_its, patch_omega=omega)
    ns = B200LinearSolver(opts).symbolic_setup(A).numerical_setup()
    dx = np.zeros(op.nrows)
    ns.solve_b(dx, -b, raise_on_maxiter=True)
    h = ns.history
    assert ns.iters <= 20 and h[-1] <= 1e-8 * h[0]
    As = A.to_scipy()
    assert abs(np.linalg.norm(As @ dx + b) - ns.resnorm) < 1e-6 * h[0]
    xo = spla.splu(As.tocsc()).solve(-b)
    nuj = fes.nfree["u"] + fes.nfree["j"]
    assert relerr(dx[:nuj], xo[:nuj]) < 1e-5
    ns.destroy()
    op.destroy()


def test_patch_solver_argument_checks(case):
    import ctypes as C

    from gridapmhd_jl_b200 import lib as L

    p, fes, prm = case
    op = B200FEOperator(fes, p["fluid"])
    A = op.jacobian(np.zeros(fes.ndofs))
    ns = B200LinearSolver(B200SolverOptions(precond="block_tri", uj_solver="gmres_patch")).symbolic_setup(A).numerical_setup()
    lib = L.load()
    ptr = np.array([0, 2], dtype=np.int64)
    bad = np.array([5, 5], dtype=np.int32)  # not strictly increasing
    assert lib.mhd_solver_set_patches(ns.handle, 1, L.ptr(ptr), L.ptr(bad)) == -1
    big = np.arange(300, dtype=np.int32)
    assert lib.mhd_solver_set_patches(ns.handle, 1, L.ptr(np.array([0, 300], dtype=np.int64)), L.ptr(big)) == -4
    ns.destroy()
    # a Jacobi-inner solver refuses patches
    ns2 = B200LinearSolver(B200SolverOptions(precond="block_tri")).symbolic_setup(A).numerical_setup()
    assert lib.mhd_solver_set_patches(ns2.handle, 1, L.ptr(ptr), L.ptr(np.array([1, 2], dtype=np.int32))) == -3
    ns2.destroy()
    op.destroy()
