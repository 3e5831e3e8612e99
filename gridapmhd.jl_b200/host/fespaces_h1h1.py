"""FE-space setup on the host for the H1-H1 formulation (u, p, phi; the current is eliminated).

Mirrors the `formulation in (:H1H1,:HDivH1)` branch of `setup_fe_spaces` (`src/fespaces.jl:32-41`) with the element
choice of `params_current_discretization(:H1, HEX, ...)` (`src/parameters.jl:528-532`):
u   = Q2 vector Lagrangian, H1-conforming, Dirichlet on `bcs[:u][:tags]`            (fespaces.jl:48-65)
p   = P1 discontinuous                                                              (fespaces.jl:67-77)
phi = Q3 scalar Lagrangian (order k+1 = 3), H1-conforming, Dirichlet on `bcs[:phi][:tags]`
      (fespaces.jl:97-114, `conformity == :H1` branch; Hunt: "conducting", hunt.jl:185-187)
and the layout `_multi_field_style(::Val{:h1h1blocks})` = (u,p,phi) (fespaces.jl:9).

The Q3 space is continuous across cells of an arbitrary conforming hex mesh: the two interior nodes of an edge and
the four interior nodes of a face are identified through the vertex they are nearest to (global vertex id), which is
invariant under the rotations/reflections with which neighbouring cells see the shared entity.

The quadrature is the same `Quadrature(HEX,5)` as for H1-HDiv: q = max(2, 5, 4, 4, 2*(3-1)) = 5 (parameters.jl:382-388).
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

from .fespaces import _first_touch_numbering
from .mesh import HexMesh
from .reffe import HEX_EDGES, HEX_FACES, Q2_NODE_XI, Tables, _lagrange_1d, make_tables, q1_tabulate

FIELDS_H1H1 = ("u", "p", "phi")
NDOFS_H1H1 = {"u": 81, "p": 4, "phi": 64}
NLOC_H1H1 = 149

# local Q3 dof l = i + 4 j + 16 k, node (i,j,k)/3
Q3_NODE_IJK = np.array([[i, j, k] for k in range(4) for j in range(4) for i in range(4)], dtype=np.int64)
Q3_NODE_XI = Q3_NODE_IJK / 3.0


def q3_tabulate(pts: np.ndarray):
    """Scalar Q3 Lagrange basis on equispaced nodes: values [nq,64], reference gradients [nq,64,3]."""
    nodes = np.array([0.0, 1.0 / 3.0, 2.0 / 3.0, 1.0])
    v = [_lagrange_1d(nodes, pts[:, d]) for d in range(3)]
    val = np.empty((len(pts), 64))
    grad = np.empty((len(pts), 64, 3))
    for a, (i, j, k) in enumerate(Q3_NODE_IJK):
        val[:, a] = v[0][0][i] * v[1][0][j] * v[2][0][k]
        grad[:, a, 0] = v[0][1][i] * v[1][0][j] * v[2][0][k]
        grad[:, a, 1] = v[0][0][i] * v[1][1][j] * v[2][0][k]
        grad[:, a, 2] = v[0][0][i] * v[1][0][j] * v[2][1][k]
    return val, grad


@dataclass
class TablesH1H1:
    """What `mhd_tables_h1h1_t` carries: the H1-HDiv tables that are reused (w, geometry, Q2, P1) + Q3."""

    base: Tables
    phi3: np.ndarray  # [nq,64]
    dphi3: np.ndarray  # [nq,64,3]

    def __getattr__(self, name):  # w, geo_grad, nu, dnu, pp, ... of the shared tables
        return getattr(self.base, name)


def make_tables_h1h1(qdegree: int = 5) -> TablesH1H1:
    base = make_tables(qdegree)
    v, g = q3_tabulate(base.xi)
    return TablesH1H1(base=base, phi3=v, dphi3=g)


def _q3_local_entities():
    """Per local Q3 node: (dimension of its entity, local entity index, nearest local vertex)."""
    dims, ents, near = [], [], []
    edge_lookup = {tuple(sorted(e)): n for n, e in enumerate(HEX_EDGES.tolist())}
    face_lookup = {tuple(sorted(f)): n for n, f in enumerate(HEX_FACES.tolist())}
    for i, j, k in Q3_NODE_IJK.tolist():
        idx = (i, j, k)
        free = [d for d in range(3) if idx[d] in (1, 2)]
        nv = [1 if idx[d] >= 2 else 0 for d in range(3)]
        vnear = nv[0] + 2 * nv[1] + 4 * nv[2]

        def corners(free_axes):
            out = []
            for m in range(1 << len(free_axes)):
                c = list(nv)
                for b, d in enumerate(free_axes):
                    c[d] = (m >> b) & 1
                out.append(c[0] + 2 * c[1] + 4 * c[2])
            return tuple(sorted(out))

        dims.append(len(free))
        near.append(vnear)
        if len(free) == 0:
            ents.append(vnear)
        elif len(free) == 1:
            ents.append(edge_lookup[corners(free)])
        elif len(free) == 2:
            ents.append(face_lookup[corners(free)])
        else:
            ents.append((i - 1) + 2 * (j - 1) + 4 * (k - 1))
    return np.array(dims), np.array(ents), np.array(near)


Q3_DIM, Q3_ENT, Q3_NEAR = _q3_local_entities()


def q3_global_labels(mesh: HexMesh):
    """[ncells,64] global node labels of the continuous Q3 space and the per-label entity (dim, id) arrays."""
    nc = mesh.ncells
    cv = mesh.cell_verts
    o_e = mesh.nverts
    o_f = o_e + 2 * mesh.nedges
    o_c = o_f + 4 * mesh.nfaces
    lab = np.empty((nc, 64), dtype=np.int64)
    for l in range(64):
        d, e, vn = int(Q3_DIM[l]), int(Q3_ENT[l]), int(Q3_NEAR[l])
        g_near = cv[:, vn]
        if d == 0:
            lab[:, l] = g_near
        elif d == 1:
            gv = cv[:, HEX_EDGES[e]]  # [nc,2]
            rank = (gv < g_near[:, None]).sum(axis=1)
            lab[:, l] = o_e + 2 * mesh.cell_edges[:, e] + rank
        elif d == 2:
            gv = cv[:, HEX_FACES[e]]  # [nc,4]
            rank = (gv < g_near[:, None]).sum(axis=1)
            lab[:, l] = o_f + 4 * mesh.cell_faces[:, e] + rank
        else:
            lab[:, l] = o_c + 8 * np.arange(nc) + e
    return lab, (o_e, o_f, o_c, o_c + 8 * nc)


@dataclass
class H1H1Spaces:
    mesh: HexMesh
    tables: TablesH1H1
    cell_dofs: dict  # field -> [ncells, ndofs] signed 1-based per-field ids (0 = absent)
    nfree: dict
    ndir: dict
    dirichlet_values: dict
    field_order: tuple = FIELDS_H1H1
    cell_solid: np.ndarray | None = None
    phi_node_coords: np.ndarray | None = None  # [ncells,64,3]
    extra: dict = field(default_factory=dict)

    @property
    def offsets(self) -> dict:
        off, o = {}, 0
        for f in self.field_order:
            off[f] = o
            o += self.nfree[f]
        return off

    @property
    def ndofs(self) -> int:
        return sum(self.nfree.values())

    def cell_global_ids(self) -> np.ndarray:
        """[ncells,149] 0-based global free ids in local order (u,p,phi); -1 where Dirichlet or absent."""
        off = self.offsets
        return np.concatenate([np.where(self.cell_dofs[f] > 0, self.cell_dofs[f] - 1 + off[f], -1) for f in FIELDS_H1H1],
                              axis=1)

    def cell_state(self, x: np.ndarray) -> np.ndarray:
        """[ncells,149] local values (free from x, Dirichlet from the stored values, absent = 0)."""
        off = self.offsets
        out = []
        for f in FIELDS_H1H1:
            ids = self.cell_dofs[f]
            dv = self.dirichlet_values[f]
            free = x[np.where(ids > 0, ids - 1 + off[f], 0)]
            dirv = dv[np.where(ids < 0, -ids - 1, 0)] if len(dv) else np.zeros_like(free)
            out.append(np.where(ids > 0, free, np.where(ids < 0, dirv, 0.0)))
        return np.concatenate(out, axis=1)

    def split(self, x: np.ndarray) -> dict:
        off = self.offsets
        return {f: x[off[f] : off[f] + self.nfree[f]] for f in FIELDS_H1H1}


def setup_fe_spaces_h1h1(mesh: HexMesh, u_tags=("noslip",), u_values=(None,), phi_tags=("conducting",),
                         phi_values=(None,), tables: TablesH1H1 | None = None,
                         solid_cells: np.ndarray | None = None) -> H1H1Spaces:
    """Build (u, p, phi).  `u_values[i]` / `phi_values[i]` is None (zero) or a callable x[n,3] -> values for tag i;
    where several tags meet on a node the later tag wins.  u and p live on the fluid cells only; phi on all cells."""
    tables = tables or make_tables_h1h1(5)
    nc = mesh.ncells
    X = mesh.cell_coords()
    fluid = np.ones(nc, dtype=bool) if solid_cells is None else ~np.asarray(solid_cells, dtype=bool)

    # ---- u: as in the H1-HDiv spaces (fespaces.py)
    o_e = mesh.nverts
    o_f = o_e + mesh.nedges
    o_c = o_f + mesh.nfaces
    unodes = np.concatenate([mesh.cell_verts, o_e + mesh.cell_edges, o_f + mesh.cell_faces, o_c + np.arange(nc)[:, None]],
                            axis=1)
    nnodes = o_c + nc
    node_dir = np.zeros(nnodes, dtype=bool)
    node_tagidx = -np.ones(nnodes, dtype=np.int64)
    for ti, tag in enumerate(u_tags):
        m = np.concatenate([mesh.vertex_tags[tag], mesh.edge_tags[tag], mesh.face_tags[tag], np.zeros(nc, dtype=bool)])
        node_dir |= m
        node_tagidx[m] = ti
    node_ids_f, nfree_n, ndir_n, _, dir_nodes = _first_touch_numbering(unodes[fluid], node_dir)
    node_ids = np.zeros(unodes.shape, dtype=np.int64)
    node_ids[fluid] = node_ids_f
    sgn = np.sign(node_ids)
    base = 3 * (np.abs(node_ids) - 1)
    cd_u = np.concatenate([np.where(sgn != 0, sgn * (base + c + 1), 0) for c in range(3)], axis=1)
    gv, _ = q1_tabulate(Q2_NODE_XI)
    node_xyz = np.einsum("av,cvi->cai", gv, X)
    dir_u = np.zeros(3 * ndir_n)
    if ndir_n:
        flat = unodes[fluid].ravel()
        uniq, first = np.unique(flat, return_index=True)
        label_first = np.zeros(nnodes, dtype=np.int64)
        label_first[uniq] = first
        xyz = node_xyz[fluid].reshape(-1, 3)[label_first[dir_nodes]]
        tix = node_tagidx[dir_nodes]
        vals = np.zeros((ndir_n, 3))
        for ti, fn in enumerate(u_values):
            if fn is not None and np.any(tix == ti):
                vals[tix == ti] = fn(xyz[tix == ti])
        dir_u = vals.reshape(-1)

    # ---- p: cell-local on the fluid cells
    fnum = np.cumsum(fluid) - 1
    cd_p = np.where(fluid[:, None], 1 + 4 * fnum[:, None] + np.arange(4)[None, :], 0)

    # ---- phi: continuous Q3 on the whole model
    lab, (q_e, q_f, q_c, nlab) = q3_global_labels(mesh)
    lab_dir = np.zeros(nlab, dtype=bool)
    lab_tag = -np.ones(nlab, dtype=np.int64)
    for ti, tag in enumerate(phi_tags):
        m = np.concatenate([mesh.vertex_tags[tag], np.repeat(mesh.edge_tags[tag], 2), np.repeat(mesh.face_tags[tag], 4),
                            np.zeros(8 * nc, dtype=bool)])
        lab_dir |= m
        lab_tag[m] = ti
    cd_phi, nfree_phi, ndir_phi, _, dir_labs = _first_touch_numbering(lab, lab_dir)
    gq, _ = q1_tabulate(Q3_NODE_XI)
    phi_xyz = np.einsum("av,cvi->cai", gq, X)
    dir_phi = np.zeros(ndir_phi)
    if ndir_phi and any(fn is not None for fn in phi_values):
        flat = lab.ravel()
        uniq, first = np.unique(flat, return_index=True)
        label_first = np.zeros(nlab, dtype=np.int64)
        label_first[uniq] = first
        xyz = phi_xyz.reshape(-1, 3)[label_first[dir_labs]]
        tix = lab_tag[dir_labs]
        for ti, fn in enumerate(phi_values):
            if fn is not None and np.any(tix == ti):
                dir_phi[tix == ti] = fn(xyz[tix == ti])

    return H1H1Spaces(
        mesh=mesh,
        tables=tables,
        cell_dofs={"u": cd_u, "p": cd_p, "phi": cd_phi},
        nfree={"u": 3 * nfree_n, "p": 4 * int(fluid.sum()), "phi": nfree_phi},
        ndir={"u": 3 * ndir_n, "p": 0, "phi": ndir_phi},
        dirichlet_values={"u": dir_u, "p": np.zeros(0), "phi": dir_phi},
        cell_solid=None if solid_cells is None else ~fluid,
        phi_node_coords=phi_xyz,
        extra={"phi_labels": lab},
    )
