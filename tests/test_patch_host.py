"""Vertex-patch block-Jacobi smoother of the (u,j) block, CPU side: the host's patch topology, the
DEVICE inversion code of gridapmhd.jl_b200/csrc/patch_cell.h executed on the CPU (tests/emul/emul_patch.cpp) against
numpy.linalg.inv, and the algorithm itself on the oracle's matrix (the smoother makes the inner GMRES converge where
point-Jacobi stagnates)."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import gridapmhd_jl_b200  # noqa: F401
from gridapmhd_jl_b200.applications import hunt_params, setup_spaces
from gridapmhd_jl_b200.host.patches import CORNER_DOFS, vertex_patches
from oracle import mhd_oracle as O

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def emul():
    so = os.path.join(HERE, "emul", "libemul_patch.so")
    src = os.path.join(HERE, "emul", "emul_patch.cpp")
    hdr = os.path.join(HERE, "..", "gridapmhd.jl_b200", "csrc", "patch_cell.h")
    if not os.path.exists(so) or max(os.path.getmtime(src), os.path.getmtime(hdr)) > os.path.getmtime(so):
        subprocess.check_call(["g++", "-O2", "-shared", "-fPIC", "-std=c++17", "-o", so, src])
    return C.CDLL(so)


@pytest.mark.parametrize("n", [1, 2, 15, 16, 17, 45, 144, 225, 256])
def test_device_inversion_code_matches_numpy(emul, n):
    """positive-real matrices with a dominant skew part (the structure of the (u,j) patches), rows shuffled so that partial
    pivoting has to act; forward and reverse thread order (a race inside a phase would show in one of them)"""
    rng = np.random.default_rng(n)
    S = rng.standard_normal((n, n))
    S = S @ S.T / n + np.eye(n)
    K = rng.standard_normal((n, n))
    A = S + 1e4 * (K - K.T)
    A = A[rng.permutation(n)]
    ref = np.linalg.inv(A)
    for rev in (0, 1):
        M = np.ascontiguousarray(A.copy())
        sing = emul.emul_patch_invert(M.ctypes.data_as(C.c_void_p), n, 256, rev)
        assert sing == 0
        assert np.abs(M - ref).max() <= 1e-9 * np.abs(ref).max()
        assert np.abs(M @ A - np.eye(n)).max() < 1e-8


def test_device_inversion_reports_singular_patches(emul):
    A = np.zeros((8, 8))
    A[:7, :7] = np.eye(7)
    assert emul.emul_patch_invert(A.ctypes.data_as(C.c_void_p), 8, 256, 0) == 1
    assert np.isfinite(A).all()


def test_vertex_patches_of_the_hunt_mesh():
    p = hunt_params(nc=(4, 4), B=(0.0, 10.0, 0.0), solver="badia2024")
    fes = setup_spaces(p)
    ptr, dofs = vertex_patches(fes)
    nuj = fes.nfree["u"] + fes.nfree["j"]
    sizes = np.diff(ptr)
    assert len(ptr) == fes.mesh.nverts + 1 and CORNER_DOFS.shape == (8, 48)
    assert sizes.max() == 225  # interior vertex: 27 Q2 nodes x 3 + 12 faces x 4 + 8 cells x 12
    assert dofs.min() >= 0 and dofs.max() < nuj
    for k in range(len(ptr) - 1):
        d = dofs[ptr[k] : ptr[k + 1]]
        assert np.all(np.diff(d) > 0)
    cover = np.bincount(dofs, minlength=nuj)
    assert cover.min() >= 1 and cover.max() == 8  # every free (u,j) dof is in a patch; cell-interior dofs in all 8 corners


def test_patch_smoother_makes_the_inner_gmres_converge_on_the_oracle_matrix():
    """Hunt nc=(6,6), Ha=20, augmented Lagrangian zeta=10: GMRES(30) on the (u,j) block stagnates with point-Jacobi and
    converges with the additive vertex-patch solver; Richardson(10, 0.2) sweeps (the reference's smoother setting,
    gmg.jl:62) converge faster per iteration."""
    p = hunt_params(nc=(6, 6), B=(0.0, 20.0, 0.0), solver="badia2024", zeta_u=10.0, zeta_j=10.0)
    fes = setup_spaces(p)
    fl = p["fluid"]
    prm = O.FluidParams(fl.alpha, fl.beta, fl.gamma, fl.sigma, fl.zeta_u, fl.zeta_j, fl.B, fl.f, fl.g, fl.convection)
    x0 = np.zeros(fes.ndofs)
    A = O.jacobian(fes, x0, prm).tocsr()
    nuj = fes.nfree["u"] + fes.nfree["j"]
    Auj = A[:nuj, :nuj].tocsr()
    rhs = -O.residual(fes, x0, prm)[:nuj]
    ptr, dofs = vertex_patches(fes)
    patches = [dofs[ptr[i] : ptr[i + 1]] for i in range(len(ptr) - 1)]
    inv = [np.linalg.inv(Auj[q][:, q].toarray()) for q in patches]

    def asm(r):
        z = np.zeros_like(r)
        for q, Ai in zip(patches, inv):
            z[q] += Ai @ r[q]
        return z

    def richardson(r, its=10, w=0.2):
        z = np.zeros_like(r)
        for _ in range(its):
            z += w * asm(r - Auj @ z)
        return z

    dj = 1.0 / Auj.diagonal()
    _, _, hj = O.fgmres(Auj, rhs, M=lambda r: dj * r, m=30, maxiter=60, rtol=1e-8, atol=0)
    _, _, ha = O.fgmres(Auj, rhs, M=asm, m=30, maxiter=60, rtol=1e-8, atol=0)
    _, _, hr = O.fgmres(Auj, rhs, M=richardson, m=30, maxiter=20, rtol=1e-8, atol=0)
    assert hj[-1] > 0.5 * hj[0]
    assert ha[-1] < 0.05 * ha[0]
    assert hr[-1] < 1e-2 * hr[0]


def test_vertex_patches_of_the_h1h1_spaces():
    from gridapmhd_jl_b200.host.patches import vertex_patches_h1h1

    p = hunt_params(nc=(4, 4), B=(0.0, 10.0, 0.0), current_disc="H1")
    fes = setup_spaces(p)
    (pu, du), (pf, df) = vertex_patches_h1h1(fes)
    nu, npp = fes.nfree["u"], fes.nfree["p"]
    assert np.diff(pu).max() == 81 and np.diff(pf).max() == 125  # 27 Q2 nodes x 3 ; 5^3 Q3 nodes
    assert du.min() >= 0 and du.max() < nu and df.min() >= nu + npp and df.max() < fes.ndofs
    assert np.bincount(du, minlength=nu).min() >= 1
    assert np.bincount(df - (nu + npp), minlength=fes.nfree["phi"]).min() >= 1
    for ptr, d in ((pu, du), (pf, df)):
        for k in range(len(ptr) - 1):
            assert np.all(np.diff(d[ptr[k] : ptr[k + 1]]) > 0)
