"""Hexahedral meshes for the Hunt and Expansion configurations (host side).

Mirrors the inputs the reference builds in `src/Meshers/hunt_mesher.jl:5-46,94-121`
(stretched periodic Cartesian duct + tags) and reads in
`src/Applications/expansion.jl:267-282` (Gmsh 4.1 ASCII hex meshes).

A `HexMesh` separates geometry (node coordinates per cell) from topology
(vertex ids after periodic identification), which is what a periodic
`CartesianDiscreteModel` does in Gridap.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field

import numpy as np

from .reffe import HEX_EDGES, HEX_FACES


@dataclass
class HexMesh:
    coords: np.ndarray  # [nnodes,3] geometric node coordinates
    cell_nodes: np.ndarray  # [ncells,8] geometric node ids, lexicographic local order
    cell_verts: np.ndarray  # [ncells,8] topological vertex ids (periodic images identified)
    # boundary tags: name -> boolean masks over topological entities (filled by `build_topology`)
    vertex_tags: dict = field(default_factory=dict)
    edge_tags: dict = field(default_factory=dict)
    face_tags: dict = field(default_factory=dict)
    cell_tags: dict = field(default_factory=dict)  # name -> bool[ncells]
    # topology
    cell_edges: np.ndarray | None = None  # [ncells,12]
    cell_faces: np.ndarray | None = None  # [ncells,6]
    nverts: int = 0
    nedges: int = 0
    nfaces: int = 0
    face_first_cell: np.ndarray | None = None  # [nfaces] lowest cell id around the face
    face_ncells: np.ndarray | None = None  # [nfaces] 1 (boundary) or 2
    grid_shape: tuple | None = None  # (nx,ny,nz) for structured meshes

    @property
    def ncells(self) -> int:
        return self.cell_nodes.shape[0]

    def cell_coords(self) -> np.ndarray:
        return self.coords[self.cell_nodes]  # [ncells,8,3]


def _unique_rows(keys: np.ndarray):
    """ids of rows of an integer key matrix; returns (ids, nunique, first_index)."""
    uniq, first, inv = np.unique(keys, axis=0, return_index=True, return_inverse=True)
    return inv.reshape(-1), len(uniq), first


def build_topology(mesh: HexMesh) -> HexMesh:
    """Global edge and face ids from the topological vertex ids."""
    cv = mesh.cell_verts
    nc = cv.shape[0]
    mesh.nverts = int(cv.max()) + 1
    ek = np.sort(cv[:, HEX_EDGES], axis=2).reshape(nc * 12, 2)
    eid, mesh.nedges, _ = _unique_rows(ek)
    mesh.cell_edges = eid.reshape(nc, 12)
    fk = np.sort(cv[:, HEX_FACES], axis=2).reshape(nc * 6, 4)
    fid, mesh.nfaces, _ = _unique_rows(fk)
    mesh.cell_faces = fid.reshape(nc, 6)
    cells = np.repeat(np.arange(nc), 6)
    first = np.full(mesh.nfaces, nc, dtype=np.int64)
    np.minimum.at(first, fid, cells)
    mesh.face_first_cell = first
    mesh.face_ncells = np.bincount(fid, minlength=mesh.nfaces)
    return mesh


def entity_vertices(mesh: HexMesh):
    """Vertex lists of each global edge [nedges,2] and face [nfaces,4] (one representative cell's order)."""
    cv = mesh.cell_verts
    ev = np.zeros((mesh.nedges, 2), dtype=np.int64)
    ev[mesh.cell_edges.reshape(-1)] = cv[:, HEX_EDGES].reshape(-1, 2)
    fv = np.zeros((mesh.nfaces, 4), dtype=np.int64)
    fv[mesh.cell_faces.reshape(-1)] = cv[:, HEX_FACES].reshape(-1, 4)
    return ev, fv


def tag_from_boundary_faces(mesh: HexMesh, name: str, face_mask: np.ndarray):
    """Tag a set of faces and their closure (edges, vertices), like a Gridap face labeling built from
    boundary entities."""
    ev, fv = entity_vertices(mesh)
    vmask = np.zeros(mesh.nverts, dtype=bool)
    vmask[fv[face_mask].reshape(-1)] = True
    # an edge belongs to the closure iff it is an edge of a tagged face
    emask = np.zeros(mesh.nedges, dtype=bool)
    # edges of faces: use cell-local relation face -> edges through vertex pairs
    face_of_cell = mesh.cell_faces
    tagged_cf = face_mask[face_of_cell]  # [nc,6]
    for f in range(6):
        fvl = set(HEX_FACES[f].tolist())
        for e in range(12):
            if set(HEX_EDGES[e].tolist()) <= fvl:
                emask[mesh.cell_edges[tagged_cf[:, f], e]] = True
    mesh.face_tags[name] = face_mask.copy()
    mesh.edge_tags[name] = emask
    mesh.vertex_tags[name] = vmask
    return mesh


# ----------------------------------------------------------------------------
# Hunt duct (src/Meshers/hunt_mesher.jl)


def strech_mhd(x: np.ndarray, domain, factor, dirs=(0, 1)) -> np.ndarray:
    """Smolentsev stretching, restating `strechMHD` (hunt_mesher.jl:5-28). x: [n,3]."""
    y = x.copy()
    for i, d in enumerate(dirs):
        xi0, xi1 = domain[2 * i], domain[2 * i + 1]
        l = xi1 - xi0
        f = factor[i]
        c = (f + 1.0) / (f - 1.0)
        if l > 0:
            m = (x[:, d] >= xi0) & (x[:, d] <= xi1)
        else:
            m = (x[:, d] >= xi1) & (x[:, d] <= xi0)
        t = (x[m, d] - xi0) / l
        ts = f * (c**t - 1.0) / (1.0 + c**t)
        y[m, d] = ts * l + xi0
    return y


def hunt_stretch_map(L: float, Ha: float, kmap_x=1, kmap_y=1, BL_adapted=True):
    """`hunt_stretch_map` (hunt_mesher.jl:30-46)."""
    strech_Ha = math.sqrt(Ha / (Ha - 1.0))
    strech_side = math.sqrt(math.sqrt(Ha) / (math.sqrt(Ha) - 1.0))

    def map1(x):
        y = strech_mhd(x, (0.0, -L, 0.0, -L), (strech_side, strech_Ha))
        return strech_mhd(y, (0.0, L, 0.0, L), (strech_side, strech_Ha))

    def map2(x):
        y = x.copy()
        y[:, 0] = np.sign(x[:, 0]) * np.abs(L * x[:, 0]) ** (1.0 / kmap_x)
        y[:, 1] = np.sign(x[:, 1]) * np.abs(L * x[:, 1]) ** (1.0 / kmap_y)
        return y

    return map1 if BL_adapted else map2


def cartesian_hex_mesh(domain, nc, periodic=(False, False, False), coord_map=None) -> HexMesh:
    """Cartesian hex mesh, cells lexicographic with x fastest (Gridap `CartesianDiscreteModel`)."""
    nx, ny, nz = nc
    xs = np.linspace(domain[0], domain[1], nx + 1)
    ys = np.linspace(domain[2], domain[3], ny + 1)
    zs = np.linspace(domain[4], domain[5], nz + 1)
    Z, Y, X = np.meshgrid(zs, ys, xs, indexing="ij")
    coords = np.stack([X.ravel(), Y.ravel(), Z.ravel()], axis=1)
    ref_coords = coords.copy()
    if coord_map is not None:
        coords = coord_map(coords)

    def nid(i, j, k):
        return i + (nx + 1) * (j + (ny + 1) * k)

    tn = [nx if periodic[0] else nx + 1, ny if periodic[1] else ny + 1, nz if periodic[2] else nz + 1]

    def tid(i, j, k):
        return (i % tn[0]) + tn[0] * ((j % tn[1]) + tn[1] * (k % tn[2]))

    K, J, I = np.meshgrid(np.arange(nz), np.arange(ny), np.arange(nx), indexing="ij")
    I, J, K = I.ravel(), J.ravel(), K.ravel()
    cn = np.empty((nx * ny * nz, 8), dtype=np.int64)
    cvt = np.empty_like(cn)
    a = 0
    for dk in (0, 1):
        for dj in (0, 1):
            for di in (0, 1):
                cn[:, a] = nid(I + di, J + dj, K + dk)
                cvt[:, a] = tid(I + di, J + dj, K + dk)
                a += 1
    mesh = HexMesh(coords=coords, cell_nodes=cn, cell_verts=cvt, grid_shape=(nx, ny, nz))
    mesh._ref_coords = ref_coords  # unmapped coordinates (used for Cartesian boundary tags)
    build_topology(mesh)
    return mesh


def _cartesian_face_masks(mesh: HexMesh, domain):
    """Boolean masks of the boundary faces on each of the 6 sides: dict {(axis,side): mask}."""
    _, fv = entity_vertices(mesh)
    # face vertex (topological) -> a geometric coordinate: use the cell geometric nodes instead
    fcoord = np.zeros((mesh.nfaces, 4, 3))
    rc = mesh._ref_coords[mesh.cell_nodes]  # [nc,8,3]
    fcoord[mesh.cell_faces.reshape(-1)] = rc[:, HEX_FACES].reshape(-1, 4, 3)
    out = {}
    tol = 1e-12
    for ax in range(3):
        for side in (0, 1):
            val = domain[2 * ax + side]
            m = np.all(np.abs(fcoord[:, :, ax] - val) < tol, axis=1) & (mesh.face_ncells == 1)
            out[(ax, side)] = m
    return out


def hunt_generate_base_mesh(nc, L=1.0, tw=0.0, Ha=10.0, kmap_x=1, kmap_y=1, BL_adapted=True, nz=3,
                            periodic_z=True, z_extent=(0.0, 0.1)) -> HexMesh:
    """`hunt_generate_base_mesh` (hunt_mesher.jl:128-138) + `hunt_add_tags!` (:94-100, tw == 0 branch).

    Tags (CartesianDiscreteModel entity ids 1-26): `noslip` = every x/y wall (ids 1-20,23-26),
    `insulating` = x = -1 and x = +1 faces (25,26), `conducting` = y walls (23,24 + lower-dim)."""
    Lt = L + tw
    cmap = hunt_stretch_map(Lt, Ha, kmap_x, kmap_y, BL_adapted)
    domain = (-1.0, 1.0, -1.0, 1.0, z_extent[0], z_extent[1])
    mesh = cartesian_hex_mesh(domain, (nc[0], nc[1], nz), periodic=(False, False, periodic_z), coord_map=cmap)
    fm = _cartesian_face_masks(mesh, domain)
    xw = fm[(0, 0)] | fm[(0, 1)]
    yw = fm[(1, 0)] | fm[(1, 1)]
    if tw > 0.0:
        # hunt_add_tags!, tw > 0 branch (hunt_mesher.jl:60-93): cells whose vertices all lie beyond |x| > L are solid_1,
        # beyond |y| > L solid_2; the fluid/solid interface is `noslip`; the outer x/y boundary is `insulating`
        X = mesh.cell_coords()
        tol = 1.0e-9
        s1 = np.all((X[:, :, 0] > L - tol) | (X[:, :, 0] < -L + tol), axis=1)
        s2 = np.all((X[:, :, 1] > L - tol) | (X[:, :, 1] < -L + tol), axis=1) & ~s1
        solid = s1 | s2
        mesh.cell_tags["solid_1"], mesh.cell_tags["solid_2"] = s1, s2
        mesh.cell_tags["solid"], mesh.cell_tags["fluid"] = solid, ~solid
        nsolid = np.zeros(mesh.nfaces, dtype=np.int64)
        nfluid = np.zeros(mesh.nfaces, dtype=np.int64)
        np.add.at(nsolid, mesh.cell_faces[solid].ravel(), 1)
        np.add.at(nfluid, mesh.cell_faces[~solid].ravel(), 1)
        tag_from_boundary_faces(mesh, "noslip", (nsolid > 0) & (nfluid > 0))
        tag_from_boundary_faces(mesh, "insulating", xw | yw)
        tag_from_boundary_faces(mesh, "conducting", np.zeros(mesh.nfaces, dtype=bool))
        if not periodic_z:
            tag_from_boundary_faces(mesh, "zwalls", fm[(2, 0)] | fm[(2, 1)])
        return mesh
    tag_from_boundary_faces(mesh, "noslip", xw | yw)
    tag_from_boundary_faces(mesh, "insulating", xw)
    tag_from_boundary_faces(mesh, "conducting", yw)
    if not periodic_z:
        zw = fm[(2, 0)] | fm[(2, 1)]
        tag_from_boundary_faces(mesh, "zwalls", zw)
    mesh.cell_tags["fluid"] = np.ones(mesh.ncells, dtype=bool)
    return mesh


# ----------------------------------------------------------------------------
# Expansion (sudden expansion duct), src/Meshers/expansion_mesher.jl


def expansion_generate_mesh(level: int = 0, perturb: float = 0.0, seed: int = 0) -> HexMesh:
    """The 12-hex base mesh of `expansion_generate_base_mesh` (expansion_mesher.jl:92-118: 4 blocks refined (1,3,1),
    mapped by `coordinate_transformation` :3-13) after `level` uniform refinements (what p4est produces there).

    In physical coordinates the base mesh is the T-shaped union of boxes with X-grid {-8,-16/3,-8/3,0,8/3,16/3,8},
    Y-grid {-1,-1/4,1/4,1} (the cubic y_stretch maps 0,1,2,3 to these), Z in [-1,1]; for X<0 only the inlet channel
    |Y|<=1/4 exists.  Tags: `inlet` (X=-8), `outlet` (X=8), `wall` (rest of the boundary), cell tag `fluid`.
    `perturb` > 0 moves interior vertices randomly by that fraction of the local cell size: general (non-affine)
    trilinear hexes like the sheared cells of meshes/Expansion_710.msh."""
    k = 2**level
    nx, ny, nz = 6 * k, 3 * k, k
    xs = np.linspace(-8.0, 8.0, nx + 1)
    yb = np.array([-1.0, -0.25, 0.25, 1.0])
    ys = np.concatenate([np.linspace(yb[i], yb[i + 1], k + 1)[:-1] for i in range(3)] + [[1.0]])
    zs = np.linspace(-1.0, 1.0, nz + 1)
    Z, Y, X = np.meshgrid(zs, ys, xs, indexing="ij")
    coords = np.stack([X.ravel(), Y.ravel(), Z.ravel()], axis=1)

    def nid(i, j, l):
        return i + (nx + 1) * (j + (ny + 1) * l)

    K, J, I = np.meshgrid(np.arange(nz), np.arange(ny), np.arange(nx), indexing="ij")
    I, J, K = I.ravel(), J.ravel(), K.ravel()
    keep = (I >= nx // 2) | ((J >= k) & (J < 2 * k))
    I, J, K = I[keep], J[keep], K[keep]
    cn = np.stack([nid(I + di, J + dj, K + dk) for dk in (0, 1) for dj in (0, 1) for di in (0, 1)], axis=1)
    used, inv = np.unique(cn, return_inverse=True)
    cn = inv.reshape(cn.shape)
    coords = coords[used]
    mesh = HexMesh(coords=coords, cell_nodes=cn, cell_verts=cn.copy())
    build_topology(mesh)
    _, fv = entity_vertices(mesh)
    fx = coords[fv][:, :, 0]
    bnd = mesh.face_ncells == 1
    inlet = bnd & np.all(np.abs(fx + 8.0) < 1e-12, axis=1)
    outlet = bnd & np.all(np.abs(fx - 8.0) < 1e-12, axis=1)
    tag_from_boundary_faces(mesh, "inlet", inlet)
    tag_from_boundary_faces(mesh, "outlet", outlet)
    tag_from_boundary_faces(mesh, "wall", bnd & ~inlet & ~outlet)
    tag_from_boundary_faces(mesh, "boundary", bnd)
    mesh.cell_tags["fluid"] = np.ones(mesh.ncells, dtype=bool)
    if perturb > 0.0:
        interior = ~mesh.vertex_tags["boundary"]
        h = np.array([16.0 / nx, 0.5 / k, 2.0 / nz])
        rng = np.random.default_rng(seed)
        mesh.coords = coords + interior[:, None] * (rng.random(coords.shape) - 0.5) * 2.0 * perturb * h
    return mesh


# ----------------------------------------------------------------------------
# Gmsh 4.1 ASCII reader (hex8 + quad4 physical groups), Expansion meshes

# gmsh hex8 node order (0,0,0),(1,0,0),(1,1,0),(0,1,0),(0,0,1),(1,0,1),(1,1,1),(0,1,1) -> lexicographic
_GMSH_HEX_TO_LEX = np.array([0, 1, 3, 2, 4, 5, 7, 6])


def read_gmsh41(path: str) -> HexMesh:
    """Minimal Gmsh 4.1 ASCII reader: nodes, hex8 cells, quad4 boundary faces with physical names
    (the subset `GmshDiscreteModel` needs for `meshes/Expansion_*.msh`)."""
    with open(path, "r") as fh:
        lines = fh.read().split("\n")
    pos = 0

    def find(tag):
        nonlocal pos
        while lines[pos].strip() != tag:
            pos += 1
        pos += 1

    find("$PhysicalNames")
    nphys = int(lines[pos]); pos += 1
    phys = {}
    for _ in range(nphys):
        d, t, name = lines[pos].split(maxsplit=2); pos += 1
        phys[(int(d), int(t))] = name.strip().strip('"')
    find("$Entities")
    npnt, ncur, nsur, nvol = map(int, lines[pos].split()); pos += 1
    ent_phys = {}
    for _ in range(npnt):
        p = lines[pos].split(); pos += 1
        n = int(p[4]); ent_phys[(0, int(p[0]))] = [int(v) for v in p[5:5 + n]]
    for dim, cnt in ((1, ncur), (2, nsur), (3, nvol)):
        for _ in range(cnt):
            p = lines[pos].split(); pos += 1
            n = int(p[7]); ent_phys[(dim, int(p[0]))] = [int(v) for v in p[8:8 + n]]
    find("$Nodes")
    nblocks, nnodes, _, maxtag = map(int, lines[pos].split()); pos += 1
    coords = np.zeros((maxtag + 1, 3))
    present = np.zeros(maxtag + 1, dtype=bool)
    for _ in range(nblocks):
        _, _, _, nb = map(int, lines[pos].split()); pos += 1
        tags = [int(lines[pos + i]) for i in range(nb)]; pos += nb
        for i in range(nb):
            coords[tags[i]] = [float(v) for v in lines[pos + i].split()]
        present[tags] = True
        pos += nb
    find("$Elements")
    nblocks, _, _, _ = map(int, lines[pos].split()); pos += 1
    hexes, hex_phys, quads, quad_phys = [], [], [], []
    for _ in range(nblocks):
        edim, etag, etype, nb = map(int, lines[pos].split()); pos += 1
        names = [phys[(edim, t)] for t in ent_phys.get((edim, etag), []) if (edim, t) in phys]
        for i in range(nb):
            p = [int(v) for v in lines[pos + i].split()]
            if etype == 5:
                hexes.append(p[1:9]); hex_phys.append(names)
            elif etype == 3:
                quads.append(p[1:5]); quad_phys.append(names)
        pos += nb
    renum = -np.ones(maxtag + 1, dtype=np.int64)
    used = np.unique(np.array(hexes).reshape(-1))
    renum[used] = np.arange(len(used))
    cn = renum[np.array(hexes, dtype=np.int64)][:, _GMSH_HEX_TO_LEX]
    mesh = HexMesh(coords=coords[used], cell_nodes=cn, cell_verts=cn.copy())
    build_topology(mesh)
    # orientation check: positive Jacobian at the first vertex, otherwise mirror the cell
    X = mesh.cell_coords()
    det = np.einsum("ci,ci->c", np.cross(X[:, 1] - X[:, 0], X[:, 2] - X[:, 0]), X[:, 4] - X[:, 0])
    if np.any(det < 0):
        flip = det < 0
        cn[flip] = cn[flip][:, [1, 0, 3, 2, 5, 4, 7, 6]]
        mesh = HexMesh(coords=coords[used], cell_nodes=cn, cell_verts=cn.copy())
        build_topology(mesh)
    # physical names on boundary faces
    _, fv = entity_vertices(mesh)
    fkey = {tuple(sorted(r)): i for i, r in enumerate(fv.tolist())}
    names_all = sorted({n for ns in quad_phys for n in ns})
    masks = {n: np.zeros(mesh.nfaces, dtype=bool) for n in names_all}
    for q, ns in zip(quads, quad_phys):
        key = tuple(sorted(renum[q].tolist()))
        fi = fkey.get(key)
        if fi is None:
            continue
        for n in ns:
            masks[n][fi] = True
    for n, m in masks.items():
        tag_from_boundary_faces(mesh, n, m)
    for n in sorted({n for ns in hex_phys for n in ns}):
        mesh.cell_tags[n] = np.array([n in ns for ns in hex_phys])
    tag_from_boundary_faces(mesh, "boundary", mesh.face_ncells == 1)
    return mesh


def save_mesh_npz(mesh: HexMesh, path: str) -> None:
    """Compact, reader-independent dump of an unstructured hex mesh: node coordinates, cell -> node ids (lexicographic
    local order) and, per boundary tag, the tagged faces as sorted vertex quadruples; cell tags as masks."""
    _, fv = entity_vertices(mesh)
    out = {"coords": mesh.coords, "cell_nodes": mesh.cell_nodes.astype(np.int32)}
    for name, mask in mesh.face_tags.items():
        out["face_tag__" + name] = np.sort(fv[np.asarray(mask, dtype=bool)], axis=1).astype(np.int32)
    for name, mask in mesh.cell_tags.items():
        out["cell_tag__" + name] = np.asarray(mask, dtype=bool)
    np.savez_compressed(path, **out)


def load_mesh_npz(path: str) -> HexMesh:
    """Inverse of `save_mesh_npz`: rebuilds the topology and re-attaches the tags (vertex / edge tags follow the faces)."""
    d = np.load(path)
    cn = d["cell_nodes"].astype(np.int64)
    mesh = HexMesh(coords=d["coords"], cell_nodes=cn, cell_verts=cn.copy())
    build_topology(mesh)
    _, fv = entity_vertices(mesh)
    fkey = {tuple(r): i for i, r in enumerate(np.sort(fv, axis=1).tolist())}
    for key in d.files:
        if key.startswith("face_tag__"):
            mask = np.zeros(mesh.nfaces, dtype=bool)
            for quad in d[key].tolist():
                mask[fkey[tuple(quad)]] = True
            tag_from_boundary_faces(mesh, key[len("face_tag__"):], mask)
        elif key.startswith("cell_tag__"):
            mesh.cell_tags[key[len("cell_tag__"):]] = d[key].astype(bool)
    return mesh


def refine_uniform(mesh: HexMesh) -> HexMesh:
    """Uniform 1:8 refinement with trilinear interpolation of the geometry (stand-in generator for the
    missing `Expansion_68k/749k.msh`). Boundary face tags are inherited."""
    if np.any(mesh.cell_nodes != mesh.cell_verts):
        raise NotImplementedError("refinement of periodic meshes")
    nc = mesh.ncells
    nv = mesh.coords.shape[0]
    ev, fv = entity_vertices(mesh)
    X = mesh.coords
    new_coords = np.concatenate(
        [X, X[ev].mean(axis=1), X[fv].mean(axis=1), X[mesh.cell_nodes].mean(axis=1)], axis=0
    )
    from .reffe import Q2_NODE_IJK

    o_e, o_f, o_c = nv, nv + mesh.nedges, nv + mesh.nedges + mesh.nfaces
    n27 = np.concatenate(
        [mesh.cell_nodes, o_e + mesh.cell_edges, o_f + mesh.cell_faces, o_c + np.arange(nc)[:, None]], axis=1
    )
    lut = {tuple(ijk): a for a, ijk in enumerate(Q2_NODE_IJK.tolist())}
    children = []
    child_parent_face = []  # for each child: for each of 6 faces, the parent's local face if on it else -1
    for ck in (0, 1):
        for cj in (0, 1):
            for ci in (0, 1):
                loc = [lut[(ci + di, cj + dj, ck + dk)] for dk in (0, 1) for dj in (0, 1) for di in (0, 1)]
                children.append(n27[:, loc])
                pf = [-1] * 6
                if ck == 0: pf[0] = 0
                if ck == 1: pf[1] = 1
                if cj == 0: pf[2] = 2
                if cj == 1: pf[3] = 3
                if ci == 0: pf[4] = 4
                if ci == 1: pf[5] = 5
                child_parent_face.append(pf)
    cn = np.stack(children, axis=1).reshape(nc * 8, 8)
    fine = HexMesh(coords=new_coords, cell_nodes=cn, cell_verts=cn.copy())
    build_topology(fine)
    cpf = np.array(child_parent_face)  # [8,6]
    fcf = fine.cell_faces.reshape(nc, 8, 6)
    for name, pm in mesh.face_tags.items():
        fm = np.zeros(fine.nfaces, dtype=bool)
        for ch in range(8):
            for f in range(6):
                if cpf[ch, f] >= 0:
                    sel = pm[mesh.cell_faces[:, cpf[ch, f]]]
                    fm[fcf[sel, ch, f]] = True
        tag_from_boundary_faces(fine, name, fm)
    for name, cm in mesh.cell_tags.items():
        fine.cell_tags[name] = np.repeat(cm, 8)
    return fine


# ----------------------------------------------------------------------------
# cell partitions (multi-GPU)


def cartesian_partition(grid_shape, np_xyz) -> np.ndarray:
    """cell -> part for a block partition of a Cartesian cell grid, GridapDistributed-style
    (`CartesianDiscreteModel(ranks,(px,py,1),...)`, hunt_mesher.jl:116-118). Parts x fastest."""
    nx, ny, nz = grid_shape
    px, py, pz = np_xyz

    def split(n, p):
        # uniform block partition: first (n % p) parts get one extra cell
        base, rem = divmod(n, p)
        sizes = [base + (1 if i < rem else 0) for i in range(p)]
        return np.repeat(np.arange(p), sizes)

    ox, oy, oz = split(nx, px), split(ny, py), split(nz, pz)
    K, J, I = np.meshgrid(np.arange(nz), np.arange(ny), np.arange(nx), indexing="ij")
    return (ox[I] + px * (oy[J] + py * oz[K])).ravel()


def rcb_partition(centroids: np.ndarray, nparts: int) -> np.ndarray:
    """Recursive coordinate bisection (stand-in for the METIS partition GridapGmsh uses)."""
    part = np.zeros(len(centroids), dtype=np.int64)

    def rec(idx, p0, n):
        if n == 1:
            part[idx] = p0
            return
        c = centroids[idx]
        ax = int(np.argmax(c.max(axis=0) - c.min(axis=0)))
        order = idx[np.argsort(c[:, ax], kind="stable")]
        nl = n // 2
        cut = (len(order) * nl) // n
        rec(order[:cut], p0, nl)
        rec(order[cut:], p0 + nl, n - nl)

    rec(np.arange(len(centroids)), 0, nparts)
    return part
