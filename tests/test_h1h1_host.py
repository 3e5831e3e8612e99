"""H1-H1 formulation, CPU side: the continuous Q3 space on the host, the NumPy oracle's pins
(FD-Jacobian, manufactured in-space solutions, block identities, Hunt solution against the analytical series), and the
DEVICE code of gridapmhd.jl_b200/csrc/h1h1_cell.h executed on the CPU (tests/emul/emul_h1h1.cpp: same phase
functions, the CTA replaced by a loop over thread ids) against the oracle, cell by cell."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import gridapmhd_jl_b200  # noqa: F401
from gridapmhd_jl_b200 import lib as L
from gridapmhd_jl_b200.applications import hunt_params, setup_spaces
from gridapmhd_jl_b200.host import mesh as M
from gridapmhd_jl_b200.host.fespaces_h1h1 import (Q3_NODE_XI, H1H1Spaces, make_tables_h1h1, q3_tabulate,
                                                   setup_fe_spaces_h1h1)
from gridapmhd_jl_b200.host.reffe import Q2_NODE_XI, make_tables, q1_tabulate
from oracle import mhd_oracle as O
from oracle import mhd_oracle_h1h1 as H

HERE = os.path.dirname(os.path.abspath(__file__))


def oracle_params(fl, **kw):
    d = dict(alpha=fl.alpha, beta=fl.beta, gamma=fl.gamma, sigma=fl.sigma, zeta_u=fl.zeta_u, zeta_j=fl.zeta_j, B=fl.B,
             f=fl.f, g=fl.g, convection=fl.convection)
    d.update(kw)
    return O.FluidParams(**d)


def test_q3_basis_is_nodal_and_reproduces_cubics():
    v, g = q3_tabulate(Q3_NODE_XI)
    assert np.allclose(v, np.eye(64), atol=1e-12)
    pts = np.random.default_rng(0).random((20, 3))
    v, g = q3_tabulate(pts)
    f = lambda x: x[:, 0] ** 3 - 2 * x[:, 0] * x[:, 1] ** 2 * x[:, 2] + x[:, 2] ** 3 + 1.0
    assert np.allclose(v @ f(Q3_NODE_XI), f(pts), atol=1e-12)
    dfdx = 3 * pts[:, 0] ** 2 - 2 * pts[:, 1] ** 2 * pts[:, 2]
    assert np.allclose(g[:, :, 0] @ f(Q3_NODE_XI), dfdx, atol=1e-11)


def _check_continuity(fes: H1H1Spaces, period_z=None):
    """every global phi label sits at ONE physical point, whichever cell looks at it"""
    lab = fes.extra["phi_labels"].ravel()
    xyz = fes.phi_node_coords.reshape(-1, 3).copy()
    if period_z is not None:
        z0, Lz = period_z
        xyz[:, 2] = z0 + np.mod(xyz[:, 2] - z0 + 1e-9, Lz) - 1e-9
    order = np.argsort(lab, kind="stable")
    lab_s, xyz_s = lab[order], xyz[order]
    first = np.r_[True, lab_s[1:] != lab_s[:-1]]
    ref = xyz_s[np.maximum.accumulate(np.where(first, np.arange(len(lab_s)), 0))]
    assert np.abs(xyz_s - ref).max() < 1e-9
    # and distinct labels sit at distinct points
    pts = np.round(xyz_s[first], 7)
    assert len(np.unique(pts, axis=0)) == first.sum()


def test_q3_space_on_the_hunt_mesh_counts_and_continuity():
    p = hunt_params(nc=(4, 4), B=(0.0, 10.0, 0.0), current_disc="H1")
    fes = setup_spaces(p)
    assert isinstance(fes, H1H1Spaces)
    # 13 x 13 nodes in the cross-section, 9 periodic layers; Dirichlet on the two conducting walls y = +-1
    assert fes.nfree["phi"] + fes.ndir["phi"] == 13 * 13 * 9 and fes.ndir["phi"] == 2 * 13 * 9
    assert fes.nfree["u"] == 882 and fes.nfree["p"] == 192  # same u, p spaces as H1-HDiv (cfg1 dof counts)
    assert fes.field_order == ("u", "p", "phi") and fes.offsets["phi"] == 882 + 192
    for f in ("u", "p", "phi"):
        ids = fes.cell_dofs[f]
        assert ids.max() == fes.nfree[f] and (ids != 0).all()
    _check_continuity(fes, period_z=(0.0, 0.1))


def test_q3_space_is_continuous_on_a_nonaffine_unstructured_mesh():
    m = M.expansion_generate_mesh(0, perturb=0.15, seed=3)
    fes = setup_fe_spaces_h1h1(m, u_tags=("inlet", "wall"), u_values=(None, None), phi_tags=("outlet",))
    _check_continuity(fes)
    # scramble the vertex numbering cell by cell orientation: reversing the global vertex ids must not break continuity
    m2 = M.expansion_generate_mesh(0, perturb=0.15, seed=3)
    perm = np.random.default_rng(5).permutation(m2.nverts)
    m2.cell_verts = perm[m2.cell_verts]
    for tags in (m2.vertex_tags,):
        for k in tags:
            t = np.zeros_like(tags[k])
            t[perm] = tags[k]
            tags[k] = t
    M.build_topology(m2)
    fv = {k: v for k, v in m.face_tags.items()}
    # re-tag through geometry-independent face vertex sets
    ev_old, fv_old = M.entity_vertices(m)
    ev_new, fv_new = M.entity_vertices(m2)
    key_old = {tuple(sorted(perm[r])): i for i, r in enumerate(fv_old)}
    idx = np.array([key_old[tuple(sorted(r))] for r in fv_new])
    for k in fv:
        M.tag_from_boundary_faces(m2, k, fv[k][idx])
    fes2 = setup_fe_spaces_h1h1(m2, u_tags=("inlet", "wall"), u_values=(None, None), phi_tags=("outlet",))
    _check_continuity(fes2)
    assert fes2.nfree == fes.nfree and fes2.ndir == fes.ndir


@pytest.fixture(scope="module")
def small_case():
    p = hunt_params(nc=(3, 3), B=(0.0, 20.0, 0.0), current_disc="H1")
    return p, setup_spaces(p)


def test_oracle_jacobian_is_the_derivative_of_the_residual(small_case):
    p, fes = small_case
    prm = oracle_params(p["fluid"], zeta_u=3.0, B=(0.3, 1.0, -0.2), convection="newton")
    rng = np.random.default_rng(0)
    x, d = rng.random(fes.ndofs), rng.standard_normal(fes.ndofs)
    A = H.jacobian(fes, x, prm)
    eps = 1e-6
    fd = (H.residual(fes, x + eps * d, prm) - H.residual(fes, x - eps * d, prm)) / (2 * eps)
    assert np.abs(fd - A @ d).max() / np.abs(A @ d).max() < 1e-8
    # picard = newton without the conv(du, grad u) term: both are exact for a state with u = 0
    x0 = x.copy()
    x0[: fes.nfree["u"]] = 0.0
    Ap = H.jacobian(fes, x0, oracle_params(p["fluid"], convection="picard"))
    An = H.jacobian(fes, x0, oracle_params(p["fluid"], convection="newton"))
    assert np.abs(Ap.data - An.data).max() < 1e-13 * np.abs(An.data).max()


def _interpolant_state(mesh, u_fn, p_fn, phi_fn):
    X = mesh.cell_coords()
    g2, _ = q1_tabulate(Q2_NODE_XI)
    g3, _ = q1_tabulate(Q3_NODE_XI)
    xu = np.einsum("av,cvi->cai", g2, X)
    xf = np.einsum("av,cvi->cai", g3, X)
    xp = X[:, [0, 1, 2, 4], :]  # P1disc is nodal at the reference simplex vertices
    nc = X.shape[0]
    u = u_fn(xu.reshape(-1, 3)).reshape(nc, 27, 3)
    st = np.concatenate([u.transpose(0, 2, 1).reshape(nc, 81), p_fn(xp.reshape(-1, 3)).reshape(nc, 4),
                         phi_fn(xf.reshape(-1, 3)).reshape(nc, 64)], axis=1)
    return st


@pytest.mark.parametrize("case", ["cancelling_current", "pure_potential"])
def test_oracle_residual_vanishes_on_manufactured_in_space_solutions(case):
    """Strong form: -beta lap u + grad p - gamma j x B = f, div u = 0, div j = 0 with j = u x B - grad phi (Ohm's law
    eliminated).  (a) u = (y,x,0), B = e_z, phi = (x^2-y^2)/2 => u x B = grad phi, j = 0, f = grad p = (1,0,0):
    the Lorentz uu term cancels against the u-phi term and the phi-u term against the Laplacian.
    (b) u = 0, phi = z, B = e_y => j = -e_z, j x B = e_x, f = -gamma e_x."""
    m = M.cartesian_hex_mesh((0.0, 1.0, 0.0, 1.5, 0.0, 0.5), (3, 2, 2))
    allb = m.face_ncells == 1
    M.tag_from_boundary_faces(m, "all", allb)
    if case == "cancelling_current":
        u_fn = lambda x: np.stack([x[:, 1], x[:, 0], 0 * x[:, 0]], axis=1)
        p_fn = lambda x: x[:, 0]
        phi_fn = lambda x: 0.5 * (x[:, 0] ** 2 - x[:, 1] ** 2)
        prm = O.FluidParams(alpha=1.0, beta=0.7, gamma=13.0, B=(0.0, 0.0, 1.0), f=(1.0, 0.0, 0.0), convection="none")
    else:
        u_fn = lambda x: np.zeros_like(x)
        p_fn = lambda x: 0 * x[:, 0]
        phi_fn = lambda x: x[:, 2]
        prm = O.FluidParams(alpha=1.0, beta=0.7, gamma=13.0, B=(0.0, 1.0, 0.0), f=(-13.0, 0.0, 0.0), convection="newton")
    fes = setup_fe_spaces_h1h1(m, u_tags=("all",), u_values=(u_fn,), phi_tags=("all",), phi_values=(phi_fn,))
    st = _interpolant_state(m, u_fn, p_fn, phi_fn)
    R = H.cell_residuals(fes.tables, m.cell_coords(), st, prm)
    gids = fes.cell_global_ids()
    r = np.zeros(fes.ndofs)
    np.add.at(r, gids[gids >= 0], R[gids >= 0])
    assert np.abs(r).max() < 1e-12
    # the Dirichlet data of the spaces is the same interpolant: cell_state of the free part reproduces it
    x = np.zeros(fes.ndofs)
    x[gids[gids >= 0]] = st[gids >= 0]
    assert np.abs(fes.cell_state(x) - st).max() < 1e-13
    assert np.abs(H.residual(fes, x, prm)).max() < 1e-12


def test_oracle_block_identities(small_case):
    p, fes = small_case
    prm = oracle_params(p["fluid"], convection="none", gamma=400.0)
    A = H.jacobian(fes, np.zeros(fes.ndofs), prm).tocsr()
    o = fes.offsets
    nu, npp, nf = fes.nfree["u"], fes.nfree["p"], fes.nfree["phi"]
    su, spp, sf = slice(o["u"], o["u"] + nu), slice(o["p"], o["p"] + npp), slice(o["phi"], o["phi"] + nf)
    Kup, Kpu = A[su, spp].toarray(), A[spp, su].toarray()
    Kuf, Kfu, Kff = A[su, sf].toarray(), A[sf, su].toarray(), A[sf, sf].toarray()
    assert np.abs(Kpu - Kup.T).max() < 1e-14
    assert np.abs(Kfu - Kuf.T / prm.gamma).max() < 1e-13 * np.abs(Kfu).max()
    assert np.abs(Kff - Kff.T).max() < 1e-13 and np.linalg.eigvalsh(Kff).min() > 0  # Dirichlet walls: SPD
    assert A[spp, spp].nnz == 0 and A[spp, sf].nnz == 0 and A[sf, spp].nnz == 0  # never inserted
    assert np.abs(Kup @ np.ones(npp)).max() < 1e-13  # constant-pressure null mode of the enclosed periodic duct
    Kuu = A[su, su].toarray()
    assert np.abs(Kuu - Kuu.T).max() < 1e-12 * np.abs(Kuu).max()  # no convection: symmetric


@pytest.mark.slow
def test_oracle_hunt_solution_matches_the_analytical_series_like_h1hdiv():
    """Solution-level pin (the reference publishes no H1-H1 numbers): at nc=(10,10), Ha=50 the H1-H1 velocity error against
    the analytical Hunt series is as small as that of the H1-HDiv solution, whose norms are pinned to the reference's
    published 16-digit values in test_oracle_pins.py."""
    p = hunt_params(nc=(10, 10), B=(0.0, 50.0, 0.0), current_disc="H1")
    fes = setup_spaces(p)
    prm = oracle_params(p["fluid"])
    x, hist = H.newton_lu(fes, prm)
    assert hist[-1] < 1e-10 * hist[0]
    n = H.hunt_norms(fes, x, make_tables_h1h1(6), p["fluid"].B, 50.0, 500, jscale=50.0)
    # H1-HDiv at the same mesh (hconv_ha00050ns500/summary.csv row nc=10): uh_l2 = 1.1222e-3, eu_l2 = 1.46e-5, eu_h1 = 6.16e-4
    assert abs(n["uh_l2"] - 1.1222118875706497e-3) < 3e-6
    assert n["eu_l2"] < 1.6e-5 and n["eu_h1"] < 6.3e-4 and n["ej_l2"] < 5.1e-4


def test_h1h1_operator_matches_hunt_solution_coarse():
    """fast version of the pin above: nc=(6,6), Ha=10; H1-H1 and H1-HDiv discrete velocities agree to discretisation error"""
    from gridapmhd_jl_b200.host.fespaces import setup_fe_spaces

    p = hunt_params(nc=(6, 6), B=(0.0, 10.0, 0.0), current_disc="H1")
    fes = setup_spaces(p)
    prm = oracle_params(p["fluid"])
    x, hist = H.newton_lu(fes, prm)
    assert hist[-1] < 1e-10 * hist[0]
    n = H.hunt_norms(fes, x, make_tables_h1h1(6), p["fluid"].B, 10.0, 200, jscale=10.0)
    f2 = setup_fe_spaces(p["model"])
    x2, _ = O.newton_lu(f2, prm)
    n2 = O.solution_norms(f2, x2, make_tables(6), jscale=10.0)
    e2 = O.hunt_error_norms(f2, x2, make_tables(6), 10.0, 200, jscale=10.0)
    assert abs(n["uh_l2"] - n2["uh_l2"]) < 2e-3 * n2["uh_l2"]
    assert abs(n["jh_l2"] - n2["jh_l2"]) < 1e-2 * n2["jh_l2"]
    assert n["eu_l2"] < 1.5 * e2["eu_l2"] and n["eu_h1"] < 1.5 * e2["eu_h1"]


# ---------------------------------------------------------------------------------------------------------------
# the device code on the CPU


@pytest.fixture(scope="module")
def emul():
    so = os.path.join(HERE, "emul", "libemul_h1h1.so")
    src = os.path.join(HERE, "emul", "emul_h1h1.cpp")
    hdr = os.path.join(HERE, "..", "gridapmhd.jl_b200", "csrc", "h1h1_cell.h")
    if not os.path.exists(so) or max(os.path.getmtime(src), os.path.getmtime(hdr)) > os.path.getmtime(so):
        subprocess.check_call(["g++", "-O2", "-shared", "-fPIC", "-std=c++17", "-o", so, src])
    lib = C.CDLL(so)
    lib.emul_h1h1_cells.restype = C.c_longlong
    return lib


def device_gids(fes: H1H1Spaces):
    """the int32 table mhd_h1h1_operator_create builds: >= 0 free id, < 0: -(index into the concatenated Dirichlet values)-1"""
    off = fes.offsets
    doff, o = {}, 0
    for f in ("u", "p", "phi"):
        doff[f] = o
        o += fes.ndir[f]
    cols = []
    for f in ("u", "p", "phi"):
        ids = fes.cell_dofs[f]
        cols.append(np.where(ids > 0, ids - 1 + off[f], np.where(ids < 0, -(doff[f] + (-ids - 1)) - 1, -o - 1)))
    dirv = np.concatenate([fes.dirichlet_values[f] for f in ("u", "p", "phi")] + [np.zeros(1)])
    return np.ascontiguousarray(np.concatenate(cols, axis=1), dtype=np.int32), dirv


def run_emul(lib, fes, x, prm: O.FluidParams, nt=256, reverse=0, jac=True, res=True):
    m, T = fes.mesh, fes.tables
    gids, dirv = device_gids(fes)
    nc = m.ncells
    K = np.zeros((nc, 149, 149)) if jac else None
    R = np.zeros((nc, 149)) if res else None
    a = lambda v, dt=np.float64: np.ascontiguousarray(v, dtype=dt)
    coords, cn = a(m.coords), a(m.cell_nodes, np.int32)
    tabs = [a(T.w), a(T.geo_grad), a(T.nu), a(T.dnu), a(T.pp), a(T.dphi3)]
    pv = a([prm.alpha, prm.beta, prm.gamma, prm.zeta_u, *prm.B, *prm.f])
    xx = a(x)
    conv = {"none": 0, "picard": 1, "newton": 2}[prm.convection]
    P = lambda v: None if v is None else v.ctypes.data_as(C.c_void_p)
    nbad = lib.emul_h1h1_cells(C.c_longlong(nc), P(coords), P(cn), P(gids), P(dirv), P(xx), *[P(t) for t in tabs], P(pv),
                               conv, nt, reverse, P(K), P(R))
    return nbad, K, R


@pytest.mark.parametrize("conv,zu", [("none", 0.0), ("picard", 0.0), ("newton", 0.0), ("none", 7.0), ("newton", 7.0)])
def test_device_cell_code_matches_the_oracle(emul, small_case, conv, zu):
    p, fes = small_case
    prm = oracle_params(p["fluid"], convection=conv, zeta_u=zu, B=(0.2, 1.0, -0.3), f=(0.1, -0.2, 1.0))
    x = np.random.default_rng(1).random(fes.ndofs)
    Ko = H.cell_jacobians(fes.tables, fes.mesh.cell_coords(), fes.cell_state(x), prm)
    Ro = H.cell_residuals(fes.tables, fes.mesh.cell_coords(), fes.cell_state(x), prm)
    mask = H.touched_mask()
    for nt, rev in ((256, 0), (256, 1), (96, 0)):
        nbad, K, R = run_emul(emul, fes, x, prm, nt=nt, reverse=rev)
        assert nbad == 0  # every entry of the enumeration stored exactly once at the (row, col) the symbolic phase assumes
        assert np.isfinite(K).all() and np.isfinite(R).all()  # no phase read data of a later (or the same) phase
        assert np.abs(K[:, ~mask]).max() == 0.0
        for blk in ((slice(0, 81), slice(0, 81)), (slice(0, 81), slice(81, 85)), (slice(81, 85), slice(0, 81)),
                    (slice(0, 81), slice(85, 149)), (slice(85, 149), slice(0, 81)), (slice(85, 149), slice(85, 149))):
            assert np.abs(K[:, blk[0], blk[1]] - Ko[:, blk[0], blk[1]]).max() <= 1e-13 * np.abs(Ko[:, blk[0], blk[1]]).max()
        assert np.abs(R - Ro).max() <= 1e-13 * np.abs(Ro).max()
    # the fused residual + Jacobian phase sequence (h1h1_jacobian_kernel<.,.,RES=true>, MHD_H1H1_FUSED=1): same results
    for rev in (2, 3):
        nbad, K, R = run_emul(emul, fes, x, prm, nt=256, reverse=rev)
        assert nbad == 0 and np.isfinite(K).all() and np.isfinite(R).all()
        assert np.abs(K - Ko).max() <= 1e-13 * np.abs(Ko).max() and np.abs(R - Ro).max() <= 1e-13 * np.abs(Ro).max()


def test_device_cell_code_on_nonaffine_cells_with_dirichlet_data(emul):
    m = M.expansion_generate_mesh(0, perturb=0.2, seed=1)
    from gridapmhd_jl_b200.applications import u_inlet_parabolic

    fes = setup_fe_spaces_h1h1(m, u_tags=("inlet", "wall"), u_values=(u_inlet_parabolic(), None), phi_tags=("outlet",),
                               phi_values=(lambda X: 1.0 + X[:, 1],))
    prm = O.FluidParams(alpha=0.5, beta=0.01, gamma=1.0, zeta_u=2.0, B=(0.0, 1.0, 0.0), convection="newton")
    x = np.random.default_rng(2).random(fes.ndofs)
    Ko = H.cell_jacobians(fes.tables, m.cell_coords(), fes.cell_state(x), prm)
    Ro = H.cell_residuals(fes.tables, m.cell_coords(), fes.cell_state(x), prm)
    nbad, K, R = run_emul(emul, fes, x, prm)
    assert nbad == 0
    assert np.abs(K - Ko).max() <= 1e-13 * np.abs(Ko).max() and np.abs(R - Ro).max() <= 1e-13 * np.abs(Ro).max()


def test_entry_enumeration_exported_by_the_library_covers_the_touched_blocks_once():
    lib = L.load()
    n = C.c_int64()
    L.check(lib.mhd_h1h1_entry_order(None, C.byref(n)))
    assert n.value == int(H.touched_mask().sum()) == 21673
    order = np.zeros(n.value, dtype=np.uint16)
    L.check(lib.mhd_h1h1_entry_order(order.ctypes.data_as(C.POINTER(C.c_uint16)), C.byref(n)))
    hit = np.zeros((149, 149), dtype=np.int64)
    np.add.at(hit, (order >> 8, order & 0xFF), 1)
    assert np.array_equal(hit, H.touched_mask().astype(np.int64))


def test_golden_fixture_h1h1_oracle_and_emulated_device_code(emul):
    """Committed golden vectors of the H1-H1 path (tests/golden/make_golden.py::main_h1h1): the oracle reproduces them, and
    so does the device cell code run on the CPU, assembled with the oracle's scatter."""
    g = np.load(os.path.join(HERE, "golden", "hunt_h1h1_nc2_ha20.npz"))
    p = hunt_params(nc=(2, 2), B=(0.0, 20.0, 0.0), zeta_u=5.0, current_disc="H1")
    fes = setup_spaces(p)
    assert [fes.nfree[f] for f in ("u", "p", "phi")] == g["ndofs"].tolist()
    prm = oracle_params(p["fluid"])
    A = H.jacobian(fes, g["x"], prm)
    assert np.array_equal(A.indptr, g["rowptr"]) and np.array_equal(A.indices, g["colval"])
    assert np.abs(A.data - g["nzval"]).max() <= 1e-13 * np.abs(g["nzval"]).max()
    assert np.abs(H.residual(fes, g["x"], prm) - g["residual"]).max() <= 1e-13 * np.abs(g["residual"]).max()
    nbad, K, R = run_emul(emul, fes, g["x"], prm)
    assert nbad == 0
    gids = fes.cell_global_ids()
    data = H._assemble(K, gids, fes.ndofs, (g["rowptr"], g["colval"].astype(np.int64)))
    assert np.abs(data - g["nzval"]).max() <= 1e-12 * np.abs(g["nzval"]).max()
    r = np.zeros(fes.ndofs)
    np.add.at(r, gids[gids >= 0], R[gids >= 0])
    assert np.abs(r - g["residual"]).max() <= 1e-12 * np.abs(g["residual"]).max()


def test_device_code_is_clean_under_address_sanitizer(tmp_path):
    """h1h1_cell.h, patch_cell.h and hdiv7_cell.h under -fsanitize=address,undefined (tests/emul/sanitize_main.cpp): no out-of-range index
    into the cell's shared data for any template variant / thread count / patch size"""
    exe = tmp_path / "sanitize_emul"
    srcs = [os.path.join(HERE, "emul", f) for f in ("sanitize_main.cpp", "emul_h1h1.cpp", "emul_patch.cpp", "emul_hdiv7.cpp")]
    build = subprocess.run(["g++", "-O1", "-g", "-fsanitize=address,undefined", "-fno-omit-frame-pointer", "-std=c++17", "-o", str(exe)] + srcs,
                           capture_output=True, text=True)
    if build.returncode != 0 and ("asan" in build.stderr or "ubsan" in build.stderr or "sanitize" in build.stderr):
        pytest.skip("sanitizer runtime not installed: " + build.stderr[-200:])
    assert build.returncode == 0, build.stderr
    run = subprocess.run([str(exe)], capture_output=True, text=True)
    assert run.returncode == 0 and "bad=0" in run.stdout, run.stdout + run.stderr


REF_MESH = "/root/reference/meshes/Expansion_710.msh"


@pytest.mark.skipif(not os.path.exists(REF_MESH), reason="reference checkout not present (GPU box)")
def test_h1h1_spaces_on_the_reference_gmsh_expansion_mesh(emul):
    """the reference's own Gmsh fixture (240 graded hexes, Gmsh vertex order permuted by the reader): the Q3 space is
    continuous on it, the oracle Jacobian is the derivative of the residual, and the device cell code agrees with the oracle"""
    from gridapmhd_jl_b200.applications import u_inlet_parabolic

    m = M.read_gmsh41(REF_MESH)
    fes = setup_fe_spaces_h1h1(m, u_tags=("inlet", "wall"), u_values=(u_inlet_parabolic(), None), phi_tags=("outlet",))
    _check_continuity(fes)
    assert fes.ndir["phi"] > 0 and fes.ndir["u"] > 0
    prm = O.FluidParams(alpha=1.0 / 5.0, beta=1.0 / 100.0, gamma=1.0, B=(0.0, 1.0, 0.0), convection="newton")
    rng = np.random.default_rng(0)
    x, d = rng.random(fes.ndofs), rng.standard_normal(fes.ndofs)
    A = H.jacobian(fes, x, prm)
    eps = 1e-6
    fd = (H.residual(fes, x + eps * d, prm) - H.residual(fes, x - eps * d, prm)) / (2 * eps)
    assert np.abs(fd - A @ d).max() / np.abs(A @ d).max() < 1e-7
    Ko = H.cell_jacobians(fes.tables, m.cell_coords(), fes.cell_state(x), prm)
    nbad, K, R = run_emul(emul, fes, x, prm, res=False)
    assert nbad == 0 and np.abs(K - Ko).max() <= 1e-12 * np.abs(Ko).max()
