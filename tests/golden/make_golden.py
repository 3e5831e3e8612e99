"""Generates the committed golden fixtures with the CPU oracle (run from the repo root:
`python tests/golden/make_golden.py`).  The reference (Julia/Gridap) cannot run in the build container, so the
vectors come from oracle/mhd_oracle.py, which is itself pinned to the reference's published Hunt norms
(tests/test_oracle_pins.py)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import gridapmhd_jl_b200  # noqa: E402,F401
from gridapmhd_jl_b200.applications import hunt_params, setup_spaces  # noqa: E402
from oracle import mhd_oracle as O  # noqa: E402


def main():
    params = hunt_params(nc=(3, 3), B=(0.0, 20.0, 0.0))
    fes = setup_spaces(params)
    fl = params["fluid"]
    prm = O.FluidParams(fl.alpha, fl.beta, fl.gamma, fl.sigma, fl.zeta_u, fl.zeta_j, fl.B, fl.f, fl.g, fl.convection)
    rng = np.random.default_rng(20261017)
    x = rng.random(fes.ndofs)
    v = rng.standard_normal(fes.ndofs)
    A = O.jacobian(fes, x, prm)
    r = O.residual(fes, x, prm)
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "hunt_nc3_ha20.npz")
    np.savez_compressed(out, x=x, v=v, rowptr=A.indptr.astype(np.int64), colval=A.indices.astype(np.int64), nzval=A.data,
                        residual=r, Av=A @ v, ndofs=np.array([fes.nfree[f] for f in ("u", "p", "j", "phi")]))
    print("wrote", out, os.path.getsize(out), "bytes; nnz", A.nnz)


def main_expansion6k():
    """BASELINE configs 3/4 stand-in: the reference's own `meshes/Expansion_6k.msh` (5 317 nodes, 4 320 non-affine
    hexes, physical names inlet / outlet / wall / fluid) converted by host/mesh.py::read_gmsh41 into the reader-independent
    fixture tests/golden/expansion_6k_mesh.npz -- the GPU box has no /root/reference, the mesh has to travel with the repo."""
    from gridapmhd_jl_b200.host import mesh as M

    m = M.read_gmsh41("/root/reference/meshes/Expansion_6k.msh")
    assert m.ncells == 4320 and m.coords.shape[0] == 5317
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "expansion_6k_mesh.npz")
    M.save_mesh_npz(m, out)
    m2 = M.load_mesh_npz(out)
    assert np.array_equal(m2.cell_nodes, m.cell_nodes) and np.array_equal(m2.coords, m.coords)
    for k in m.face_tags:
        assert np.array_equal(m2.face_tags[k], m.face_tags[k]), k
    print("wrote", out, os.path.getsize(out), "bytes")


def main_h1h1():
    """H1-H1 formulation (oracle/mhd_oracle_h1h1.py): Hunt nc=(2,2), Ha=20, zeta_u=5, Newton convection."""
    from oracle import mhd_oracle_h1h1 as H

    params = hunt_params(nc=(2, 2), B=(0.0, 20.0, 0.0), zeta_u=5.0, current_disc="H1")
    fes = setup_spaces(params)
    fl = params["fluid"]
    prm = O.FluidParams(fl.alpha, fl.beta, fl.gamma, fl.sigma, fl.zeta_u, fl.zeta_j, fl.B, fl.f, fl.g, fl.convection)
    rng = np.random.default_rng(20261018)
    x = rng.random(fes.ndofs)
    v = rng.standard_normal(fes.ndofs)
    A = H.jacobian(fes, x, prm)
    r = H.residual(fes, x, prm)
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "hunt_h1h1_nc2_ha20.npz")
    np.savez_compressed(out, x=x, v=v, rowptr=A.indptr.astype(np.int64), colval=A.indices.astype(np.int32), nzval=A.data,
                        residual=r, Av=A @ v, ndofs=np.array([fes.nfree[f] for f in ("u", "p", "phi")]))
    print("wrote", out, os.path.getsize(out), "bytes; nnz", A.nnz)


if __name__ == "__main__" and len(sys.argv) > 1 and sys.argv[1] == "expansion6k":
    main_expansion6k()
    sys.exit(0)
if __name__ == "__main__":
    main()
    main_h1h1()
