"""FE-space setup on the host: cell->dof tables, Dirichlet data and the multi-field layout.

Mirrors `setup_fe_spaces` (`src/fespaces.jl:13-46`) for the H1-HDiv formulation:
u  = Q2 vector Lagrangian, H1-conforming, Dirichlet on `bcs[:u][:tags]`       (fespaces.jl:48-65)
p  = P1 discontinuous (L2)                                                      (fespaces.jl:67-77)
j  = RT1, HDiv-conforming, Dirichlet (normal flux) on `bcs[:j][:tags]`          (fespaces.jl:79-95)
phi= Q1 discontinuous (L2)                                                      (fespaces.jl:97-114)
and the layouts of `_multi_field_style` (fespaces.jl:4-9).

Index convention handed to the C ABI (Gridap's): per-field, 1-based, signed; id<0 is a Dirichlet dof
and -id indexes the field's Dirichlet-value array (1-based).
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

from .mesh import HexMesh
from .reffe import HEX_FACES, Q2_NODE_XI, Tables, make_tables, q1_tabulate

FIELDS = ("u", "p", "j", "phi")
NDOFS = {"u": 81, "p": 4, "j": 36, "phi": 8}

# `_multi_field_style` (src/fespaces.jl:4-9): order of the fields in the global vector
FIELD_ORDER = {
    "julia": ("u", "p", "j", "phi"),  # ConsecutiveMultiFieldStyle
    "petsc": ("u", "p", "j", "phi"),
    "badia2024": ("u", "j", "p", "phi"),  # BlockMultiFieldStyle(3,(2,1,1),(1,3,2,4)): ([u,j],p,phi)
    "li2019": ("j", "u", "p", "phi"),  # BlockMultiFieldStyle(4,(1,1,1,1),(3,1,2,4))
    "b200": ("u", "j", "p", "phi"),  # device FGMRES + block-triangular preconditioner
}


def _first_touch_numbering(labels: np.ndarray, is_dir: np.ndarray):
    """Number labels in order of first appearance in `labels.ravel()`; Dirichlet labels are numbered
    separately. Returns (ids signed 1-based with the shape of labels, nfree, ndir)."""
    flat = labels.ravel()
    uniq, first = np.unique(flat, return_index=True)
    order = np.argsort(first, kind="stable")
    uniq = uniq[order]
    dmask = is_dir[uniq]
    ids_of = np.zeros(int(flat.max()) + 1, dtype=np.int64)
    free = uniq[~dmask]
    dirs = uniq[dmask]
    ids_of[free] = np.arange(1, len(free) + 1)
    ids_of[dirs] = -np.arange(1, len(dirs) + 1)
    return ids_of[flat].reshape(labels.shape), len(free), len(dirs), free, dirs


@dataclass
class FESpaces:
    mesh: HexMesh
    tables: Tables
    cell_dofs: dict  # field -> [ncells, ndofs] signed 1-based per-field ids
    nfree: dict  # field -> int
    ndir: dict  # field -> int
    dirichlet_values: dict  # field -> [ndir] float64
    j_sign: np.ndarray  # [ncells,36] int8
    field_order: tuple = FIELD_ORDER["julia"]
    u_node_coords: np.ndarray | None = None  # [ncells,27,3] physical coordinates of the Q2 nodes
    cell_unodes: np.ndarray | None = None  # [ncells,27] global Q2 node labels
    cell_solid: np.ndarray | None = None  # [ncells] bool: solid cells carry only j and phi (u, p ids are 0 = absent)
    cell_sigma: np.ndarray | None = None  # [ncells] conductivity used on solid cells (params[:solid][:sigma])
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
        """[ncells,129] 0-based global free ids in local order (u,p,j,phi); -1 where Dirichlet."""
        off = self.offsets
        cols = []
        for f in FIELDS:
            ids = self.cell_dofs[f]
            cols.append(np.where(ids > 0, ids - 1 + off[f], -1))
        return np.concatenate(cols, axis=1)

    def cell_state(self, x: np.ndarray) -> np.ndarray:
        """[ncells,129] local values (free from x, Dirichlet from the stored values)."""
        off = self.offsets
        out = []
        for f in FIELDS:
            ids = self.cell_dofs[f]
            dv = self.dirichlet_values[f]
            free = x[np.where(ids > 0, ids - 1 + off[f], 0)]
            dirv = dv[np.where(ids < 0, -ids - 1, 0)] if len(dv) else np.zeros_like(free)
            out.append(np.where(ids > 0, free, np.where(ids < 0, dirv, 0.0)))  # id 0 = absent dof (value 0)
        return np.concatenate(out, axis=1)

    def split(self, x: np.ndarray) -> dict:
        off = self.offsets
        return {f: x[off[f] : off[f] + self.nfree[f]] for f in FIELDS}


def setup_fe_spaces(mesh: HexMesh, u_tags=("noslip",), u_values=(None,), j_tags=("insulating",),
                    solver: str = "julia", tables: Tables | None = None, solid_cells: np.ndarray | None = None,
                    cell_sigma: np.ndarray | None = None) -> FESpaces:
    """Build the four spaces. `u_values[i]` is None (zero) or a callable x[n,3] -> u[n,3] for tag i;
    where several tags meet on a node the later tag in the list wins (expansion.jl:150-153).
    Normal-flux Dirichlet values for j are zero (hunt.jl:182, expansion.jl:156-159)."""
    tables = tables or make_tables(5)
    nc = mesh.ncells
    X = mesh.cell_coords()

    # ---- u: Q2 nodes = vertices | edges | faces | cell interiors
    o_e = mesh.nverts
    o_f = o_e + mesh.nedges
    o_c = o_f + mesh.nfaces
    unodes = np.concatenate(
        [mesh.cell_verts, o_e + mesh.cell_edges, o_f + mesh.cell_faces, o_c + np.arange(nc)[:, None]], axis=1
    )
    nnodes = o_c + nc
    node_dir = np.zeros(nnodes, dtype=bool)
    node_tagidx = -np.ones(nnodes, dtype=np.int64)
    for ti, tag in enumerate(u_tags):
        m = np.concatenate([mesh.vertex_tags[tag], mesh.edge_tags[tag], mesh.face_tags[tag], np.zeros(nc, dtype=bool)])
        node_dir |= m
        node_tagidx[m] = ti
    fluid = np.ones(nc, dtype=bool) if solid_cells is None else ~np.asarray(solid_cells, dtype=bool)
    # u and p live on the fluid triangulation only (fe_space_u/fe_space_p use params[:Omega_f], fespaces.jl:51,68)
    node_ids_f, nfree_n, ndir_n, free_nodes, dir_nodes = _first_touch_numbering(unodes[fluid], node_dir)
    node_ids = np.zeros(unodes.shape, dtype=np.int64)
    node_ids[fluid] = node_ids_f
    comp = np.arange(3)
    # local dof a + 27 c ; global id 3*(node-1)+c+1, sign preserved
    sgn = np.sign(node_ids)
    base = 3 * (np.abs(node_ids) - 1)
    cd_u = np.concatenate([np.where(sgn != 0, sgn * (base + c + 1), 0) for c in comp], axis=1)
    # node coordinates (trilinear map of the reference node positions)
    gv, _ = q1_tabulate(Q2_NODE_XI)
    node_xyz = np.einsum("av,cvi->cai", gv, X)
    dir_u = np.zeros(3 * ndir_n)
    if ndir_n:
        # coordinates and tag of each Dirichlet node (first cell occurrence)
        flat = unodes[fluid].ravel()
        _, first = np.unique(flat, return_index=True)
        label_first = np.zeros(nnodes, dtype=np.int64)
        label_first[np.unique(flat)] = first
        fidx = label_first[dir_nodes]
        xyz = node_xyz[fluid].reshape(-1, 3)[fidx]
        tix = node_tagidx[dir_nodes]
        vals = np.zeros((ndir_n, 3))
        for ti, fn in enumerate(u_values):
            if fn is None:
                continue
            m = tix == ti
            if np.any(m):
                vals[m] = fn(xyz[m])
        dir_u = vals.reshape(-1)

    # ---- p, phi: cell-local
    fnum = np.cumsum(fluid) - 1  # fluid cell counter
    cd_p = np.where(fluid[:, None], 1 + 4 * fnum[:, None] + np.arange(4)[None, :], 0)
    cd_phi = 1 + 8 * np.arange(nc)[:, None] + np.arange(8)[None, :]

    # ---- j: 4 dofs per face (one per face vertex, matched across cells by global vertex id) + 12 interior
    fv = mesh.cell_verts[:, HEX_FACES]  # [nc,6,4]
    slot = np.argsort(np.argsort(fv, axis=2), axis=2)  # rank of each local face vertex among the face's ids
    flabel = 4 * mesh.cell_faces[:, :, None] + slot  # [nc,6,4]
    ilabel = 4 * mesh.nfaces + 12 * np.arange(nc)[:, None] + np.arange(12)[None, :]
    jlabels = np.concatenate([flabel.reshape(nc, 24), ilabel], axis=1)
    jdir = np.zeros(4 * mesh.nfaces + 12 * nc, dtype=bool)
    for tag in j_tags:
        fm = mesh.face_tags[tag]
        jdir[: 4 * mesh.nfaces] |= np.repeat(fm, 4)
    cd_j, nfree_j, ndir_j, _, _ = _first_touch_numbering(jlabels, jdir)
    # sign flip: the second cell around a facet flips its face dofs
    is_first = mesh.face_first_cell[mesh.cell_faces] == np.arange(nc)[:, None]  # [nc,6]
    js = np.ones((nc, 36), dtype=np.int8)
    js[:, :24] = np.repeat(np.where(is_first, 1, -1), 4, axis=1)

    return FESpaces(
        mesh=mesh,
        tables=tables,
        cell_dofs={"u": cd_u, "p": cd_p, "j": cd_j, "phi": cd_phi},
        nfree={"u": 3 * nfree_n, "p": 4 * int(fluid.sum()), "j": nfree_j, "phi": 8 * nc},
        ndir={"u": 3 * ndir_n, "p": 0, "j": ndir_j, "phi": 0},
        dirichlet_values={"u": dir_u, "p": np.zeros(0), "j": np.zeros(ndir_j), "phi": np.zeros(0)},
        j_sign=js,
        field_order=FIELD_ORDER[solver],
        u_node_coords=node_xyz,
        cell_unodes=unodes,
        cell_solid=None if solid_cells is None else ~fluid,
        cell_sigma=None if cell_sigma is None else np.asarray(cell_sigma, dtype=np.float64),
    )
