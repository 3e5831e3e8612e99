"""GPU tests of the sum-factorised H1-HDiv Jacobian kernel (csrc/hdiv_v7.cu): kernel selection from the tables, parity with
the oracle at the BENCHMARK configuration (Hunt nc=(64,64), Ha=1000, Newton convection -- BASELINE.json configs[1]),
the generic kernel of assembly.cu kept alive behind MHD_JAC_V7=0, and the deterministic (coloured) assembly mode.
Tolerances (BASELINE.json north_star): CSR structure bit-exact, values <= 1e-12 relative."""
import os

import numpy as np
import pytest

from gridapmhd_jl_b200.applications import hunt_params, setup_spaces
from gridapmhd_jl_b200.feoperator import B200FEOperator, FluidParams
from oracle import mhd_oracle as O

pytestmark = pytest.mark.gpu
VAL_TOL = 1e-12


def relerr(a, b):
    return np.abs(a - b).max() / np.abs(b).max()


def oprm(fl):
    return O.FluidParams(fl.alpha, fl.beta, fl.gamma, fl.sigma, fl.zeta_u, fl.zeta_j, fl.B, fl.f, fl.g, fl.convection)


@pytest.fixture()
def generic_kernel_env():
    old = os.environ.get("MHD_JAC_V7")
    os.environ["MHD_JAC_V7"] = "0"
    yield
    if old is None:
        del os.environ["MHD_JAC_V7"]
    else:
        os.environ["MHD_JAC_V7"] = old


def small_case():
    p = hunt_params(nc=(4, 4), B=(0.0, 10.0, 0.0))
    fes = setup_spaces(p)
    fl = FluidParams(alpha=0.7, beta=0.9, gamma=100.0, sigma=1.3, zeta_u=3.0, zeta_j=2.0, B=(0.1, 1.0, 0.2), f=(0.3, 0.1, 1.0),
                     g=(0.1, 0.2, 0.3), convection="newton")
    return fes, fl


def test_the_reference_elements_select_the_sum_factorised_kernel(mhdlib):
    fes, fl = small_case()
    op = B200FEOperator(fes, fl)
    assert op.kernel_version == 7
    op.destroy()


def test_tables_without_tensor_structure_keep_the_generic_kernel(mhdlib):
    """A basis that is not a tensor product (here: u basis function 5 += 0.1 * function 7) must not be fed to the
    sum-factorised kernel; the generic kernel integrates whatever the tables hold."""
    import copy

    fes, fl = small_case()
    fes2 = copy.copy(fes)
    T = copy.copy(fes.tables)
    nu = np.array(T.nu, dtype=np.float64, copy=True)
    dnu = np.array(T.dnu, dtype=np.float64, copy=True)
    nu.reshape(27, 27)[:, 5] += 0.1 * nu.reshape(27, 27)[:, 7]
    dnu.reshape(27, 27, 3)[:, 5, :] += 0.1 * dnu.reshape(27, 27, 3)[:, 7, :]
    T.nu, T.dnu = nu, dnu
    fes2.tables = T
    op = B200FEOperator(fes2, fl)
    assert op.kernel_version == 5
    x = np.random.default_rng(5).random(fes.ndofs)
    A = op.jacobian(x)
    Ao = O.jacobian(fes2, x, oprm(fl))
    assert relerr(A.nzval(), Ao.data) < VAL_TOL
    op.destroy()


@pytest.mark.parametrize("conv,zu,zj", [("newton", 0.0, 0.0), ("none", 10.0, 10.0)])
def test_generic_kernel_still_matches_the_oracle(mhdlib, generic_kernel_env, conv, zu, zj):
    fes, fl = small_case()
    fl = FluidParams(alpha=fl.alpha, beta=fl.beta, gamma=fl.gamma, sigma=fl.sigma, zeta_u=zu, zeta_j=zj, B=fl.B, f=fl.f, g=fl.g,
                     convection=conv)
    op = B200FEOperator(fes, fl)
    assert op.kernel_version == 5
    x = np.random.default_rng(1234).random(fes.ndofs)
    A = op.allocate_jacobian()
    b = np.empty(op.nrows)
    op.residual_and_jacobian_b(b, A, x)
    Ao = O.jacobian(fes, x, oprm(fl))
    rowptr, colval = A.pattern()
    assert np.array_equal(rowptr, Ao.indptr) and np.array_equal(colval, Ao.indices)
    assert relerr(A.nzval(), Ao.data) < VAL_TOL
    assert relerr(b, O.residual(fes, x, oprm(fl))) < VAL_TOL
    op.destroy()


def test_both_kernels_agree_entry_by_entry(mhdlib):
    fes, fl = small_case()
    x = np.random.default_rng(77).random(fes.ndofs)
    op7 = B200FEOperator(fes, fl)
    v7 = op7.jacobian(x).nzval().copy()
    op7.destroy()
    os.environ["MHD_JAC_V7"] = "0"
    try:
        op5 = B200FEOperator(fes, fl)
    finally:
        del os.environ["MHD_JAC_V7"]
    assert op5.kernel_version == 5
    v5 = op5.jacobian(x).nzval().copy()
    op5.destroy()
    assert relerr(v7, v5) < VAL_TOL


def test_cell_traversal_order_does_not_change_the_result(mhdlib):
    """The kernel walks the cells breadth-first over face neighbours (L2 locality of shared sectors, hdiv_v7.cu
    v7_build_cell_order); MHD_V7_ORDER=0 keeps the mesh order.  Same matrix and residual either way (summation order of the
    shared entries aside), both within tolerance of the oracle -- on a mesh with more cells than one wave of the grid."""
    p = hunt_params(nc=(12, 10), B=(0.0, 30.0, 0.0))
    fes = setup_spaces(p)
    _, fl = small_case()
    x = np.random.default_rng(5).random(fes.ndofs)
    out = {}
    for order in ("1", "0"):
        os.environ["MHD_V7_ORDER"] = order
        try:
            op = B200FEOperator(fes, fl)
            A = op.allocate_jacobian()
            b = np.empty(op.nrows)
            op.residual_and_jacobian_b(b, A, x)
            out[order] = (A.nzval().copy(), b.copy())
            op.destroy()
        finally:
            del os.environ["MHD_V7_ORDER"]
    Ao = O.jacobian(fes, x, oprm(fl))
    ro = O.residual(fes, x, oprm(fl))
    for order in ("1", "0"):
        assert relerr(out[order][0], Ao.data) < VAL_TOL
        assert relerr(out[order][1], ro) < VAL_TOL
    assert relerr(out["1"][0], out["0"][0]) < 1e-14


def test_deterministic_mode_is_bit_reproducible_and_matches_the_oracle(mhdlib):
    fes, fl = small_case()
    op = B200FEOperator(fes, fl)
    x = np.random.default_rng(3).random(fes.ndofs)
    op.allocate_jacobian()
    ncol = op.set_deterministic(True)
    assert 8 <= ncol <= 64
    runs = []
    for _ in range(4):
        A = op.allocate_jacobian()
        b = np.empty(op.nrows)
        op.residual_and_jacobian_b(b, A, x)
        runs.append((A.nzval().copy(), b.copy()))
    for nz, b in runs[1:]:
        assert np.array_equal(nz, runs[0][0])
        assert np.array_equal(b, runs[0][1])  # the fused residual follows the colours too
    Ao = O.jacobian(fes, x, oprm(fl))
    assert relerr(runs[0][0], Ao.data) < VAL_TOL
    assert relerr(runs[0][1], O.residual(fes, x, oprm(fl))) < VAL_TOL
    op.set_deterministic(False)
    assert relerr(op.jacobian(x).nzval(), Ao.data) < VAL_TOL
    op.destroy()


def test_parity_at_the_benchmark_configuration(mhdlib):
    """BASELINE configs[1]: Hunt nc=(64,64), Ha=1000, 12 288 cells, 146 981 976 nnz -- the size every bench number is quoted
    on.  CSR structure of the complete rows of a 2 048-cell block bit-exact, their values and residual rows <= 1e-12 against
    the C oracle, and the SpMV against a host product on the device's own matrix."""
    from oracle.parity import assembly_parity, spmv_parity

    p = hunt_params(nc=(64, 64), B=(0.0, 1000.0, 0.0))
    fes = setup_spaces(p)
    fl = p["fluid"]
    assert fl.convection == "newton"
    op = B200FEOperator(fes, fl)
    assert op.kernel_version == 7
    x = np.random.default_rng(1234).random(fes.ndofs)
    A = op.allocate_jacobian()
    assert A.nnz == 146981976 and op.nrows == 732690
    b = np.empty(op.nrows)
    op.residual_and_jacobian_b(b, A, x)
    rowptr, colval = A.pattern()
    nz = A.nzval()
    for seed_cell in (0, 6000):
        par = assembly_parity(fes, oprm(fl), x, rowptr, colval, nz, b, op.nrows, ncells=2048, seed_cell=seed_cell)
        assert par["csr_bitexact"] and par["rows_checked"] > 50000, par
        assert par["jac_rel"] < VAL_TOL and par["res_rel"] < VAL_TOL, par
    y = op.spmv(x)
    assert spmv_parity(rowptr, colval, nz, x, y) < VAL_TOL
    op.destroy()
