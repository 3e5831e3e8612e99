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


def _patches_from_corner_slots(fes, corner_slots: np.ndarray):
    """CSR (ptr, dofs) of the vertex patches given, per cell corner, the local dof slots that belong to that corner's patch."""
    gids = fes.cell_global_ids()
    nc = gids.shape[0]
    stride = int(gids.max()) + 2
    verts = np.repeat(fes.mesh.cell_verts.reshape(nc, 8, 1), corner_slots.shape[1], axis=2)
    dofs = gids[:, corner_slots]
    ok = dofs >= 0
    key = np.unique(verts[ok].astype(np.int64) * stride + dofs[ok])
    v, d = key // stride, key % stride
    ptr = np.zeros(fes.mesh.nverts + 1, dtype=np.int64)
    np.add.at(ptr, v + 1, 1)
    return np.cumsum(ptr), d.astype(np.int32)


def vertex_patches_h1h1(fes):
    """H1-H1 spaces (u,p,phi): ((ptr_u, dofs_u), (ptr_phi, dofs_phi)) -- the vertex-star patches of the velocity block (27
    Q2 nodes x 3 = 81 dofs on an interior vertex) and of the potential block (5^3 = 125 Q3 nodes), global row ids of the
    layout `fes.field_order`; what `gmg_block_jacobi_smoothers` builds for the u-block GMG of `gmg_solver(::Val{:H1H1},
    ::Val{:h1h1blocks}, params)` (src/Solvers/gmg.jl:107-140)."""
    from .fespaces_h1h1 import Q3_NODE_IJK

    u_slots, f_slots = [], []
    for k in range(8):
        vk = HEX_VERTS[k]
        nodes = [a for a in range(27) if all(Q2_NODE_IJK[a][d] != 2 * (1 - vk[d]) for d in range(3))]
        u_slots.append([a + 27 * c for c in range(3) for a in nodes])
        f_slots.append([85 + l for l in range(64) if all(Q3_NODE_IJK[l][d] != 3 * (1 - vk[d]) for d in range(3))])
    return (_patches_from_corner_slots(fes, np.array(u_slots, dtype=np.int64)),
            _patches_from_corner_slots(fes, np.array(f_slots, dtype=np.int64)))


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
