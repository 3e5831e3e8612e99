"""CPU-only tests of the host logic and of the C-ABI library surface (no compute calls without a GPU)."""
import ctypes
import os
import re

import numpy as np
import pytest

from gridapmhd_jl_b200.host import fespaces as F
from gridapmhd_jl_b200.host import mesh as M
from gridapmhd_jl_b200.host import reffe

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


# ---- C ABI surface ---------------------------------------------------------------------------
def _declared_symbols():
    txt = open(os.path.join(ROOT, "include", "mhdb200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(mhd_[a-z0-9_]+)\s*\(", txt)))


def test_library_loads_and_exports_every_declared_symbol():
    from gridapmhd_jl_b200 import lib as L

    lib = L.load()
    declared = _declared_symbols()
    assert len(declared) >= 30
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/mhdb200.h but not exported"
    # and the ctypes table covers exactly the header
    assert sorted(L.SIGNATURES) == declared


def test_no_cpu_fallback_fails_loudly():
    """Without a CUDA device the product path must refuse to run (no silent CPU fallback)."""
    from gridapmhd_jl_b200 import lib as L

    try:
        import torch

        if torch.cuda.is_available():
            pytest.skip("GPU present")
    except ImportError:
        pass
    rc = L.load().mhd_init(0)
    assert rc == -2
    assert b"no CPU fallback" in L.load().mhd_last_error_string()
    with pytest.raises(L.MhdError):
        L.init(0)


def test_c_struct_layouts_match_header_sizes(tmp_path):
    """ctypes mirrors of the ABI structs have the sizes the C compiler gives the header's structs."""
    import subprocess

    from gridapmhd_jl_b200 import lib as L

    src = tmp_path / "sizes.c"
    src.write_text('#include <stdio.h>\n#include "mhdb200.h"\nint main(void){printf("%zu %zu %zu %zu %zu %zu %zu %zu\\n", sizeof(mhd_mesh_t),'
                   " sizeof(mhd_tables_t), sizeof(mhd_layout_t), sizeof(mhd_params_t), sizeof(mhd_solver_opts_t), sizeof(mhd_hunt_post_t),"
                   " sizeof(mhd_tables_h1h1_t), sizeof(mhd_layout_h1h1_t)); return 0;}\n")
    exe = tmp_path / "sizes"
    subprocess.check_call(["/usr/bin/gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    sizes = [int(v) for v in subprocess.check_output([str(exe)]).split()]
    mine = [ctypes.sizeof(t) for t in (L.mhd_mesh_t, L.mhd_tables_t, L.mhd_layout_t, L.mhd_params_t, L.mhd_solver_opts_t, L.mhd_hunt_post_t,
                                       L.mhd_tables_h1h1_t, L.mhd_layout_h1h1_t)]
    assert mine == sizes, (mine, sizes)


def test_scatter_map_entry_order_covers_every_touched_pair_once():
    """The enumeration shared by the symbolic phase (map construction) and the Jacobian kernel (map consumption):
    every touched (row, col) pair of the 8 blocks of jac_fluid_h1_hdiv (weakforms.jl:311) exactly once, nothing else,
    job ranges padded with 0xFFFF to multiples of 32."""
    from gridapmhd_jl_b200 import lib as L
    from oracle import mhd_oracle as O

    lib = L.load()
    n = ctypes.c_int64()
    assert lib.mhd_map_entry_order(None, ctypes.byref(n)) == 0
    assert n.value % 32 == 0 and 14913 <= n.value <= 16384
    order = np.zeros(n.value, dtype=np.uint16)
    assert lib.mhd_map_entry_order(order.ctypes.data_as(ctypes.POINTER(ctypes.c_uint16)), ctypes.byref(n)) == 0
    real = order[order != 0xFFFF]
    assert len(real) == 14913 and len(np.unique(real)) == 14913
    seen = np.zeros((129, 129), dtype=bool)
    seen[real >> 8, real & 0xFF] = True
    assert np.array_equal(seen, O.touched_mask())  # the permutation acts inside each field: same block structure
    # padding only at the end of 32-aligned job ranges: a pad code is never followed by a real entry inside a 32-block
    blocks = (order != 0xFFFF).reshape(-1, 32)
    assert np.all(np.diff(blocks.astype(np.int8), axis=1) <= 0)


# ---- reference elements --------------------------------------------------------------------------
def test_quadrature_degree5_is_27_point_gauss():
    xi, w = reffe.quadrature_for_degree(5)
    assert len(w) == 27 and abs(w.sum() - 1.0) < 1e-15
    # exact for x^5 y^4 z^3 on [0,1]^3
    assert abs((w * xi[:, 0] ** 5 * xi[:, 1] ** 4 * xi[:, 2] ** 3).sum() - 1 / 6 / 5 / 4) < 1e-15
    assert len(reffe.quadrature_for_degree(6)[1]) == 64


def test_q2_and_q1_bases_are_nodal_and_partition_unity():
    val, grad = reffe.q2_tabulate(reffe.Q2_NODE_XI)
    assert np.allclose(val, np.eye(27), atol=1e-14)
    pts = np.random.default_rng(0).random((10, 3))
    v, g = reffe.q2_tabulate(pts)
    assert np.allclose(v.sum(axis=1), 1.0) and np.allclose(g.sum(axis=1), 0.0, atol=1e-13)
    v1, g1 = reffe.q1_tabulate(pts)
    assert np.allclose(v1.sum(axis=1), 1.0) and np.allclose(g1.sum(axis=1), 0.0, atol=1e-13)
    assert np.allclose(reffe.p1_tabulate(pts).sum(axis=1), 1.0)


def test_rt1_basis_is_dual_to_its_moments_and_div_is_consistent():
    C = reffe.rt1_coefficients()
    assert C.shape == (36, 36) and np.linalg.matrix_rank(C) == 36
    # normal flux of every basis function through face f vanishes unless the function belongs to f
    g, gw = reffe.gauss_legendre_01(4)
    S, T = np.meshgrid(g, g, indexing="ij")
    for f in range(6):
        ax, side = reffe.HEX_FACE_AXIS[f], reffe.HEX_FACE_SIDE[f]
        others = [d for d in range(3) if d != ax]
        pts = np.zeros((16, 3))
        pts[:, ax] = side
        pts[:, others[0]] = S.ravel()
        pts[:, others[1]] = T.ravel()
        val, _ = reffe.rt1_tabulate(pts)
        vn = val[:, :, ax]
        mask = np.ones(36, dtype=bool)
        mask[4 * f : 4 * f + 4] = False
        assert np.abs(vn[:, mask]).max() < 1e-12
    # divergence theorem per basis function: int div psi = sum of face fluxes = sum over the 4 face moments (q_k sum to 1)
    pts3, w3 = reffe.tensor_quadrature(4)
    _, div = reffe.rt1_tabulate(pts3)
    total_div = (w3[:, None] * div).sum(axis=0)
    expected = np.zeros(36)
    expected[:24] = 1.0  # each face function has unit moment against one bilinear q_k, and sum_k q_k = 1
    assert np.allclose(total_div, expected, atol=1e-12)
    # tabulated divergence equals the finite-difference divergence of the tabulated values
    p0 = np.array([[0.3, 0.6, 0.2]])
    h = 1e-6
    fd = np.zeros(36)
    for d in range(3):
        e = np.zeros(3)
        e[d] = h
        fd += (reffe.rt1_tabulate(p0 + e)[0][0, :, d] - reffe.rt1_tabulate(p0 - e)[0][0, :, d]) / (2 * h)
    assert np.allclose(fd, reffe.rt1_tabulate(p0)[1][0], atol=1e-8)


# ---- meshes ----------------------------------------------------------------------------------------
def test_hunt_stretch_map_restates_the_reference_formula():
    """strechMHD (hunt_mesher.jl:5-28): end points fixed, monotone, clustered towards the walls."""
    Ha = 10.0
    cmap = M.hunt_stretch_map(1.0, Ha)
    x = np.zeros((5, 3))
    x[:, 0] = [-1, -0.5, 0, 0.5, 1]
    x[:, 1] = [-1, -0.5, 0, 0.5, 1]
    y = cmap(x)
    assert np.allclose(y[[0, 2, 4], 0], [-1, 0, 1]) and np.allclose(y[[0, 2, 4], 1], [-1, 0, 1])
    assert np.all(np.diff(y[:, 0]) > 0) and y[3, 0] > 0.5 and y[1, 0] < -0.5  # clustered towards +-1
    f = np.sqrt(Ha / (Ha - 1))
    c = (f + 1) / (f - 1)
    assert abs(y[3, 1] - f * (c**0.5 - 1) / (1 + c**0.5)) < 1e-14


def test_hunt_mesh_topology_and_tags():
    m = M.hunt_generate_base_mesh((4, 3), Ha=10.0)
    assert m.ncells == 36 and m.nverts == 5 * 4 * 3  # periodic z: 3 vertex levels
    assert m.nfaces == (5 * 3 + 4 * 4) * 3 + 4 * 3 * 3
    assert m.face_tags["insulating"].sum() == 2 * 3 * 3 and m.face_tags["noslip"].sum() == 2 * 3 * 3 + 2 * 4 * 3
    assert np.all(m.face_ncells[m.face_tags["noslip"]] == 1)
    X = m.cell_coords()
    vol = np.abs(np.einsum("ci,ci->c", np.cross(X[:, 1] - X[:, 0], X[:, 2] - X[:, 0]), X[:, 4] - X[:, 0])).sum()
    assert abs(vol - 4 * 0.1) < 1e-12


def test_gmsh_reader_on_a_generated_file(tmp_path):
    """Gmsh 4.1 ASCII subset: 2 hexes in gmsh node order + physical names."""
    msh = tmp_path / "two_hex.msh"
    nodes = [(x, y, z) for z in (0, 1) for y in (0, 1) for x in (0, 1, 2)]
    nid = lambda x, y, z: 1 + x + 3 * (y + 2 * z)
    hexes = [[nid(i, 0, 0), nid(i + 1, 0, 0), nid(i + 1, 1, 0), nid(i, 1, 0), nid(i, 0, 1), nid(i + 1, 0, 1), nid(i + 1, 1, 1), nid(i, 1, 1)] for i in (0, 1)]
    inlet = [nid(0, 0, 0), nid(0, 1, 0), nid(0, 1, 1), nid(0, 0, 1)]
    lines = ["$MeshFormat", "4.1 0 8", "$EndMeshFormat", "$PhysicalNames", "2", '2 1 "inlet"', '3 2 "fluid"', "$EndPhysicalNames",
             "$Entities", "0 0 1 1", "1 0 0 0 0 1 1 1 1 0", "1 0 0 0 2 1 1 1 2 0", "$EndEntities",
             "$Nodes", f"1 {len(nodes)} 1 {len(nodes)}", f"3 1 0 {len(nodes)}"] + [str(i + 1) for i in range(len(nodes))] + \
            [f"{x} {y} {z}" for x, y, z in nodes] + ["$EndNodes", "$Elements", "2 3 1 3", "2 1 3 1", "1 " + " ".join(map(str, inlet)),
                                                     "3 1 5 2", "2 " + " ".join(map(str, hexes[0])), "3 " + " ".join(map(str, hexes[1])), "$EndElements"]
    msh.write_text("\n".join(lines) + "\n")
    m = M.read_gmsh41(str(msh))
    assert m.ncells == 2 and m.coords.shape == (12, 3)
    assert m.face_tags["inlet"].sum() == 1 and m.face_tags["boundary"].sum() == 10
    assert m.cell_tags["fluid"].all()
    X = m.cell_coords()
    # lexicographic local order: vertex 1 is +x of vertex 0, vertex 2 is +y, vertex 4 is +z
    assert np.allclose(X[:, 1] - X[:, 0], [1, 0, 0]) and np.allclose(X[:, 2] - X[:, 0], [0, 1, 0]) and np.allclose(X[:, 4] - X[:, 0], [0, 0, 1])
    fine = M.refine_uniform(m)
    assert fine.ncells == 16 and fine.face_tags["inlet"].sum() == 4 and fine.face_tags["boundary"].sum() == 40


def test_fe_spaces_dirichlet_and_sign_conventions():
    m = M.hunt_generate_base_mesh((3, 3), Ha=10.0)
    fes = F.setup_fe_spaces(m)
    # every interior face has exactly one +1 and one -1 cell; boundary faces only +1
    s = fes.j_sign[:, :24].reshape(m.ncells, 6, 4)[:, :, 0]
    tot = np.zeros(m.nfaces)
    np.add.at(tot, m.cell_faces.ravel(), s.ravel())
    assert np.all(tot[m.face_ncells == 2] == 0) and np.all(tot[m.face_ncells == 1] == 1)
    # ids: 1-based signed, contiguous
    for f in ("u", "p", "j", "phi"):
        ids = fes.cell_dofs[f]
        assert ids.max() == fes.nfree[f] and (ids != 0).all()
        if fes.ndir[f]:
            assert ids.min() == -fes.ndir[f]
    assert fes.offsets == {"u": 0, "p": fes.nfree["u"], "j": fes.nfree["u"] + fes.nfree["p"],
                           "phi": fes.nfree["u"] + fes.nfree["p"] + fes.nfree["j"]}
    fb = F.setup_fe_spaces(m, solver="badia2024")
    assert fb.field_order == ("u", "j", "p", "phi") and fb.offsets["p"] == fb.nfree["u"] + fb.nfree["j"]


def test_cartesian_and_rcb_partitions():
    part = M.cartesian_partition((8, 6, 3), (2, 2, 1))
    assert part.shape == (144,) and set(part) == {0, 1, 2, 3}
    assert np.bincount(part).tolist() == [36, 36, 36, 36]
    c = np.random.default_rng(0).random((1000, 3))
    p = M.rcb_partition(c, 8)
    assert np.bincount(p).min() >= 120 and np.bincount(p).max() <= 130


REF_MESH = "/root/reference/meshes/Expansion_710.msh"


@pytest.mark.skipif(not os.path.exists(REF_MESH), reason="reference checkout not present (GPU box)")
def test_gmsh_reader_on_the_reference_expansion_mesh():
    """The reference's own fixture (402 nodes / 240 hexes, graded cells, physical names
    inlet/outlet/wall/fluid): read it, build the Expansion spaces on it and check the oracle Jacobian against its FD."""
    from gridapmhd_jl_b200.applications import expansion_params, setup_spaces
    from oracle import mhd_oracle as O

    m = M.read_gmsh41(REF_MESH)
    assert m.ncells == 240 and m.coords.shape[0] == 402
    for tag in ("inlet", "outlet", "wall"):
        assert m.face_tags[tag].sum() > 0
    assert (m.face_tags["inlet"] | m.face_tags["outlet"] | m.face_tags["wall"]).sum() == (m.face_ncells == 1).sum()
    assert m.cell_tags["fluid"].all()
    X = m.cell_coords()
    assert abs(X[..., 0].min() + 8) < 1e-9 and abs(X[..., 0].max() - 8) < 1e-9  # x in [-8, 8]
    params = expansion_params(Ha=10.0, N=5.0, mesh=m)
    fes = setup_spaces(params)
    _, det, _ = O.cell_geometry(fes.tables, X)
    assert det.min() > 0  # positively oriented after the Gmsh -> lexicographic vertex permutation
    fl = params["fluid"]
    prm = O.FluidParams(fl.alpha, fl.beta, fl.gamma, fl.sigma, fl.zeta_u, fl.zeta_j, fl.B, fl.f, fl.g, fl.convection)
    rng = np.random.default_rng(0)
    x, d = rng.random(fes.ndofs), rng.standard_normal(fes.ndofs)
    A = O.jacobian(fes, x, prm)
    eps = 1e-6
    fd = (O.residual(fes, x + eps * d, prm) - O.residual(fes, x - eps * d, prm)) / (2 * eps)
    assert np.abs(fd - A @ d).max() / np.abs(A @ d).max() < 1e-6


def test_sum_factorised_uu_block_matches_the_oracle():
    """design study for the next Jacobian kernel: the uu block by sum factorisation = the oracle's dense block"""
    import importlib.util

    spec = importlib.util.spec_from_file_location("sumfac", os.path.join(ROOT, "tools", "sumfac_study.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    mod.main()  # asserts < 1e-12


