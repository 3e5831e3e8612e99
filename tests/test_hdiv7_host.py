"""The fully sum-factorised H1-HDiv Jacobian ("v7": gridapmhd.jl_b200/csrc/hdiv7_cell.h + hdiv7_tables.h) executed on the CPU
(tests/emul/emul_hdiv7.cpp, the kernel's phase sequence incl. the two staging buffers) against the oracle's dense cell matrices:
tensor structure discovered from the plain tables, all convection x zeta variants, non-affine cells, solid cells."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import gridapmhd_jl_b200  # noqa: F401
from gridapmhd_jl_b200.applications import hunt_params, setup_spaces
from gridapmhd_jl_b200.host import mesh as M
from gridapmhd_jl_b200.host.fespaces import setup_fe_spaces
from oracle import mhd_oracle as O

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "..", "gridapmhd.jl_b200", "csrc")


@pytest.fixture(scope="module")
def emul():
    so = os.path.join(HERE, "emul", "libemul_hdiv7.so")
    srcs = [os.path.join(HERE, "emul", "emul_hdiv7.cpp"), os.path.join(CSRC, "hdiv7_cell.h"), os.path.join(CSRC, "hdiv7_tables.h")]
    if not os.path.exists(so) or max(os.path.getmtime(f) for f in srcs) > os.path.getmtime(so):
        subprocess.check_call(["g++", "-O2", "-shared", "-fPIC", "-std=c++17", "-o", so, srcs[0]])
    lib = C.CDLL(so)
    lib.emul_hdiv7_cells.restype = C.c_longlong
    return lib


def device_gids(fes):
    off = fes.offsets
    doff, o = {}, 0
    for f in ("u", "p", "j", "phi"):
        doff[f] = o
        o += fes.ndir[f]
    cols = []
    for f in ("u", "p", "j", "phi"):
        ids = fes.cell_dofs[f]
        cols.append(np.where(ids > 0, ids - 1 + off[f], np.where(ids < 0, -(doff[f] + (-ids - 1)) - 1, -o - 1)))
    dirv = np.concatenate([fes.dirichlet_values[f] for f in ("u", "p", "j", "phi")] + [np.zeros(1)])
    return np.ascontiguousarray(np.concatenate(cols, axis=1), dtype=np.int32), dirv


def run(lib, fes, x, prm, nt=256, reverse=0, tables=None, residual=False):
    m, T = fes.mesh, tables or fes.tables
    gids, dirv = device_gids(fes)
    nc = m.ncells
    K = np.zeros((nc, 129, 129))
    a = lambda v, dt=np.float64: np.ascontiguousarray(v, dtype=dt)
    P = lambda v: None if v is None else v.ctypes.data_as(C.c_void_p)
    coords, cn, js = a(m.coords), a(m.cell_nodes, np.int32), a(fes.j_sign, np.int8)
    solid = None if fes.cell_solid is None else a(fes.cell_solid, np.uint8)
    sig = None if fes.cell_sigma is None else a(fes.cell_sigma)
    tabs = [a(t) for t in (T.w, T.geo_grad, T.nu, T.dnu, T.pp, T.psi, T.dpsi, T.chi)]
    pv = a([prm.alpha, prm.beta, prm.gamma, prm.sigma, prm.zeta_u, prm.zeta_j, *prm.B, *prm.f, *prm.g])
    R = np.zeros((nc, 129)) if residual else None
    xx = a(x)
    conv = {"none": 0, "picard": 1, "newton": 2}[prm.convection]
    nbad = lib.emul_hdiv7_cells(C.c_longlong(nc), P(coords), P(cn), P(gids), P(js), P(solid), P(sig), P(dirv), P(xx), *[P(t) for t in tabs],
                                P(pv), conv, nt, reverse, P(K), P(R))
    return (nbad, K, R) if residual else (nbad, K)


def compare(K, Ko, solid=None):
    mask = O.touched_mask()
    assert np.isfinite(K).all()
    assert np.abs(K[:, ~mask]).max() == 0.0
    if solid is not None:  # the u / p dofs of a solid cell are absent: the device drops those entries through the map
        K = K.copy()
        K[solid, :85, :] = 0.0
        K[solid, :, :85] = 0.0
    for r in ((0, 81), (81, 85), (85, 121), (121, 129)):
        for c in ((0, 81), (81, 85), (85, 121), (121, 129)):
            blk = Ko[:, r[0] : r[1], c[0] : c[1]]
            if np.abs(blk).max() > 0:
                assert np.abs(K[:, r[0] : r[1], c[0] : c[1]] - blk).max() <= 1e-12 * np.abs(blk).max(), (r, c)


@pytest.mark.parametrize("conv,zu,zj", [("none", 0.0, 0.0), ("picard", 0.0, 0.0), ("newton", 0.0, 0.0), ("none", 7.0, 3.0),
                                         ("newton", 7.0, 3.0)])
def test_v7_cell_code_matches_the_oracle(emul, conv, zu, zj):
    p = hunt_params(nc=(3, 3), B=(0.0, 20.0, 0.0))
    fes = setup_spaces(p)
    fl = p["fluid"]
    prm = O.FluidParams(alpha=fl.alpha, beta=fl.beta, gamma=fl.gamma, sigma=0.7, zeta_u=zu, zeta_j=zj, B=(0.2, 1.0, -0.3), convection=conv)
    x = np.random.default_rng(1).random(fes.ndofs)
    Ko = O.cell_jacobians(fes.tables, fes.mesh.cell_coords(), fes.cell_state(x), fes.j_sign, prm)
    for nt, rev in ((256, 0), (256, 1), (96, 0)):
        nbad, K = run(emul, fes, x, prm, nt=nt, reverse=rev)
        assert nbad == 0  # >= 0: the tables have the tensor structure; every map entry stored exactly once at its (row, col)
        compare(K, Ko)


def test_v7_cell_code_on_nonaffine_cells(emul):
    m = M.expansion_generate_mesh(0, perturb=0.2, seed=1)
    from gridapmhd_jl_b200.applications import u_inlet_parabolic

    fes = setup_fe_spaces(m, u_tags=("inlet", "wall"), u_values=(u_inlet_parabolic(), None), j_tags=("wall", "inlet", "outlet"))
    prm = O.FluidParams(alpha=0.5, beta=0.01, gamma=1.0, sigma=1.0, zeta_u=2.0, zeta_j=2.0, B=(0.3, 1.0, 0.1), convection="newton")
    x = np.random.default_rng(2).random(fes.ndofs)
    Ko = O.cell_jacobians(fes.tables, m.cell_coords(), fes.cell_state(x), fes.j_sign, prm)
    nbad, K = run(emul, fes, x, prm)
    assert nbad == 0
    compare(K, Ko)


def test_v7_cell_code_with_solid_walls(emul):
    p = hunt_params(nc=(12, 12), B=(0.0, 50.0, 0.0), tw=0.2, BL_adapted=False, kmap_x=3, kmap_y=3, zeta_j=2.0)  # hunt_tests.jl:74-88
    fes = setup_spaces(p)
    fl = p["fluid"]
    prm = O.FluidParams(fl.alpha, fl.beta, fl.gamma, fl.sigma, fl.zeta_u, fl.zeta_j, fl.B, fl.f, fl.g, fl.convection)
    x = np.random.default_rng(3).random(fes.ndofs)
    assert fes.cell_solid is not None and fes.cell_solid.any()
    Ko = O.cell_jacobians(fes.tables, fes.mesh.cell_coords(), fes.cell_state(x), fes.j_sign, prm, fes.cell_solid, fes.cell_sigma)
    nbad, K = run(emul, fes, x, prm)
    assert nbad == 0
    compare(K, Ko, solid=fes.cell_solid)


def test_v7_structure_discovery_rejects_non_tensor_tables(emul):
    """a basis that is not a tensor product (here: one Q2 function perturbed at one point) must be refused -> generic kernel"""
    import copy

    p = hunt_params(nc=(2, 2), B=(0.0, 20.0, 0.0))
    fes = setup_spaces(p)
    T = copy.deepcopy(fes.tables)
    T.nu[5, 7] += 1e-6
    prm = O.FluidParams()
    nbad, _ = run(emul, fes, np.zeros(fes.ndofs), prm, tables=T)
    assert nbad == -1


def compare_residual(R, Ro, solid=None):
    assert np.isfinite(R).all()
    if solid is not None:
        R = R.copy()
        R[solid, :85] = 0.0
    for r in ((0, 81), (81, 85), (85, 121), (121, 129)):
        blk = Ro[:, r[0] : r[1]]
        assert np.abs(R[:, r[0] : r[1]] - blk).max() <= 1e-12 * max(np.abs(blk).max(), 1e-300), r


@pytest.mark.parametrize("conv,zu,zj", [("none", 0.0, 0.0), ("picard", 0.0, 0.0), ("newton", 0.0, 0.0), ("none", 7.0, 3.0),
                                         ("newton", 7.0, 3.0)])
def test_v7_fused_residual_matches_the_oracle(emul, conv, zu, zj):
    """res_fluid_h1_hdiv by sum factorisation, interleaved with the Jacobian phases (the Jacobian must come out unchanged)"""
    p = hunt_params(nc=(3, 3), B=(0.0, 20.0, 0.0))
    fes = setup_spaces(p)
    fl = p["fluid"]
    prm = O.FluidParams(alpha=fl.alpha, beta=fl.beta, gamma=fl.gamma, sigma=0.7, zeta_u=zu, zeta_j=zj, B=(0.2, 1.0, -0.3), f=(0.3, -0.1, 1.0),
                        g=(0.1, 0.2, -0.3), convection=conv)
    x = np.random.default_rng(11).random(fes.ndofs)
    st = fes.cell_state(x)
    Ko = O.cell_jacobians(fes.tables, fes.mesh.cell_coords(), st, fes.j_sign, prm)
    Ro = O.cell_residuals(fes.tables, fes.mesh.cell_coords(), st, fes.j_sign, prm)
    for nt, rev in ((256, 0), (256, 1), (96, 0)):
        nbad, K, R = run(emul, fes, x, prm, nt=nt, reverse=rev, residual=True)
        assert nbad == 0
        compare(K, Ko)
        compare_residual(R, Ro)


def test_v7_fused_residual_on_nonaffine_cells_with_dirichlet_data(emul):
    m = M.expansion_generate_mesh(0, perturb=0.2, seed=1)
    from gridapmhd_jl_b200.applications import u_inlet_parabolic

    fes = setup_fe_spaces(m, u_tags=("inlet", "wall"), u_values=(u_inlet_parabolic(), None), j_tags=("wall", "inlet", "outlet"))
    prm = O.FluidParams(alpha=0.5, beta=0.01, gamma=1.0, sigma=1.0, zeta_u=2.0, zeta_j=2.0, B=(0.3, 1.0, 0.1), f=(0.1, 0.2, 0.3), g=(0.3, 0.2, 0.1),
                        convection="newton")
    x = np.random.default_rng(2).random(fes.ndofs)
    st = fes.cell_state(x)
    Ro = O.cell_residuals(fes.tables, m.cell_coords(), st, fes.j_sign, prm)
    nbad, K, R = run(emul, fes, x, prm, residual=True)
    assert nbad == 0
    compare(K, O.cell_jacobians(fes.tables, m.cell_coords(), st, fes.j_sign, prm))
    compare_residual(R, Ro)


def test_v7_fused_residual_with_solid_walls(emul):
    p = hunt_params(nc=(12, 12), B=(0.0, 50.0, 0.0), tw=0.2, BL_adapted=False, kmap_x=3, kmap_y=3, zeta_j=2.0)
    fes = setup_spaces(p)
    fl = p["fluid"]
    prm = O.FluidParams(fl.alpha, fl.beta, fl.gamma, fl.sigma, fl.zeta_u, fl.zeta_j, fl.B, fl.f, (0.1, 0.2, 0.3), fl.convection)
    x = np.random.default_rng(3).random(fes.ndofs)
    st = fes.cell_state(x)
    Ro = O.cell_residuals(fes.tables, fes.mesh.cell_coords(), st, fes.j_sign, prm, fes.cell_solid, fes.cell_sigma)
    nbad, K, R = run(emul, fes, x, prm, residual=True)
    assert nbad == 0
    compare_residual(R, Ro, solid=fes.cell_solid)
