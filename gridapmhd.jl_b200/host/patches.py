"""Vertex-star patches of the (u,j) block for the patch-based block-Jacobi smoother.

Mirrors what the reference hands to `PatchBasedSmoothers.BlockJacobiSolver(space, ptopo; assembly = :star)` with
`ptopo = Geometry.PatchTopology(ReferenceFE{0}, model)` in `gmg_block_jacobi_smoothers` (src/Solvers/gmg.jl:62-81):
one patch per mesh vertex v, holding the free dofs of every entity (vertex, edge, face, cell) whose closure contains v
-- the dofs interior to the star of v.  For Q2 velocity + RT1 current on an interior vertex of a hex mesh that is
27 nodes x 3 + 12 faces x 4 + 8 cells x 12 = 225 dofs.
"""
from __future__ import annotations

import numpy as np

from .fespaces import FESpaces
from .reffe import HEX_FACES, HEX_VERTS, Q2_NODE_IJK


def _corner_local_dofs():
    """[8, 48] local (u,j) dof slots of a cell (u: a + 27 c in 0..80, j: 85 + m) that belong to the patch of corner k."""
    out = []
    for k in range(8):
        vk = HEX_VERTS[k]
        nodes = [a for a in range(27) if all(Q2_NODE_IJK[a][d] != 2 * (1 - vk[d]) for d in range(3))]
        assert len(nodes) == 8
        u = [a + 27 * c for c in range(3) for a in nodes]
        faces = [f for f in range(6) if k in HEX_FACES[f]]
        assert len(faces) == 3
        j = [85 + 4 * f + i for f in faces for i in range(4)] + [85 + 24 + i for i in range(12)]
        out.append(u + j)
    return np.array(out, dtype=np.int64)


CORNER_DOFS = _corner_local_dofs()


def vertex_patches(fes: FESpaces):
    """(patch_ptr int64 [nverts+1], patch_dofs int32) -- sorted global ids (0-based, layout of `fes.field_order`) of the
    free u and j dofs in the star of every vertex.  Empty patches (all dofs Dirichlet) are kept with zero length."""
    gids = fes.cell_global_ids()  # [nc,129], -1 = Dirichlet / absent
    nc = gids.shape[0]
    verts = np.repeat(fes.mesh.cell_verts.reshape(nc, 8, 1), CORNER_DOFS.shape[1], axis=2)  # [nc,8,48]
    dofs = gids[:, CORNER_DOFS]  # [nc,8,48]
    ok = dofs >= 0
    key = verts[ok].astype(np.int64) * (gids.max() + 2) + dofs[ok]
    key = np.unique(key)
    v = key // (gids.max() + 2)
    d = key % (gids.max() + 2)
    ptr = np.zeros(fes.mesh.nverts + 1, dtype=np.int64)
    np.add.at(ptr, v + 1, 1)
    return np.cumsum(ptr), d.astype(np.int32)
