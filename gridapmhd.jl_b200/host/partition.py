"""Domain decomposition of the FE spaces for one-process-per-GPU runs (host side).

Mirrors what GridapDistributed/PartitionedArrays give the reference:
  * cells are partitioned (`CartesianDiscreteModel(ranks,(px,py,1),...)`, hunt_mesher.jl:116-118; METIS for Gmsh models),
    each rank also holds one layer of ghost cells;
  * every free dof has exactly one owner rank; local numbering = owned dofs first, then ghosts (PVector own/ghost layout);
  * `consistent!` (owner -> ghost) is described by per-neighbour send/receive index lists.
Differences by design: ghost cells are integrated redundantly ("fully assembled rows") so no matrix entry ever
crosses NVLink; only rows of owned dofs are assembled on a rank.
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

from .fespaces import FESpaces
from .fespaces_h1h1 import H1H1Spaces
from .mesh import HexMesh, build_topology, cartesian_partition


def _fields(fes) -> tuple:
    """("u","p","j","phi") for the H1-HDiv spaces, ("u","p","phi") for the H1-H1 spaces"""
    return tuple(fes.cell_dofs.keys())


def dof_owners(fes: FESpaces, cell_part: np.ndarray) -> dict:
    """field -> owner rank of each free dof (0-based per-field id): the lowest part among the cells around it."""
    out = {}
    nparts = int(cell_part.max()) + 1
    for f in _fields(fes):
        ids = fes.cell_dofs[f]
        own = np.full(fes.nfree[f], nparts, dtype=np.int64)
        free = ids > 0
        parts = np.broadcast_to(cell_part[:, None], ids.shape)
        np.minimum.at(own, ids[free] - 1, parts[free])
        out[f] = own
    return out


@dataclass
class LocalSets:
    cells: np.ndarray  # global ids of the local cells, owned first
    nowned_cells: int
    owned: dict  # field -> sorted global free ids (0-based) owned by the rank
    ghost: dict  # field -> sorted global free ids of the ghosts


def local_sets(fes: FESpaces, cell_part: np.ndarray, owners: dict, rank: int) -> LocalSets:
    owned_cells = np.nonzero(cell_part == rank)[0]
    touch = np.zeros(fes.mesh.ncells, dtype=bool)
    for f in _fields(fes):
        ids = fes.cell_dofs[f]
        free = ids > 0
        mine = np.zeros(ids.shape, dtype=bool)
        mine[free] = owners[f][ids[free] - 1] == rank
        touch |= mine.any(axis=1)
    ghost_cells = np.nonzero(touch & (cell_part != rank))[0]
    cells = np.concatenate([owned_cells, ghost_cells])
    owned, ghost = {}, {}
    for f in _fields(fes):
        ids = fes.cell_dofs[f][cells]
        g = np.unique(ids[ids > 0] - 1)
        o = owners[f][g] == rank
        owned[f] = g[o]
        ghost[f] = g[~o]
    return LocalSets(cells=cells, nowned_cells=len(owned_cells), owned=owned, ghost=ghost)


@dataclass
class PartitionedSpaces:
    fes: FESpaces  # local spaces (local cells, local per-field numbering: owned first, then ghosts)
    rank: int
    nparts: int
    nowned: dict
    nowned_cells: int
    cells: np.ndarray
    own_global: dict  # field -> global ids of the owned dofs (local order)
    ghost_global: dict
    neigh: np.ndarray = field(default_factory=lambda: np.zeros(0, np.int32))
    send_ptr: np.ndarray = field(default_factory=lambda: np.zeros(1, np.int64))
    send_idx: np.ndarray = field(default_factory=lambda: np.zeros(0, np.int32))
    recv_ptr: np.ndarray = field(default_factory=lambda: np.zeros(1, np.int64))
    recv_idx: np.ndarray = field(default_factory=lambda: np.zeros(0, np.int32))

    @property
    def nrows(self):
        return sum(self.nowned.values())

    @property
    def ncols(self):
        return sum(self.fes.nfree.values())

    def offsets(self):
        """(own_off, ghost_off) of each field in the local vector [all owned | all ghosts] (library layout)."""
        own, gh, o = {}, {}, 0
        for f in self.fes.field_order:
            own[f] = o
            o += self.nowned[f]
        for f in self.fes.field_order:
            gh[f] = o
            o += self.fes.nfree[f] - self.nowned[f]
        return own, gh

    def local_vector_ids(self):
        """global vector index (global layout of the un-partitioned spaces) of each local vector entry."""
        goff = self._global_offsets
        own_off, gh_off = self.offsets()
        out = np.empty(self.ncols, dtype=np.int64)
        for f in self.fes.field_order:
            no = self.nowned[f]
            out[own_off[f] : own_off[f] + no] = goff[f] + self.own_global[f]
            ng = len(self.ghost_global[f])
            out[gh_off[f] : gh_off[f] + ng] = goff[f] + self.ghost_global[f]
        return out


def partition_fespaces(fes: FESpaces, cell_part: np.ndarray, rank: int) -> PartitionedSpaces:
    """Local spaces + halo plan of `rank`. Every rank can call this on the same global spaces; the send/receive
    lists of two neighbours are ordered identically (field order, then global id), so no handshake is needed."""
    nparts = int(cell_part.max()) + 1
    owners = dof_owners(fes, cell_part)
    sets = {rank: local_sets(fes, cell_part, owners, rank)}
    me = sets[rank]
    m = fes.mesh
    # ---- local mesh
    cn = m.cell_nodes[me.cells]
    used, inv = np.unique(cn, return_inverse=True)
    lmesh = HexMesh(coords=m.coords[used], cell_nodes=inv.reshape(cn.shape), cell_verts=m.cell_verts[me.cells])
    # ---- local dof tables
    cell_dofs, nfree, nowned = {}, {}, {}
    for f in _fields(fes):
        ids = fes.cell_dofs[f][me.cells]
        lut = np.zeros(fes.nfree[f] + 1, dtype=np.int64)
        no, ng = len(me.owned[f]), len(me.ghost[f])
        lut[me.owned[f] + 1] = np.arange(1, no + 1)
        lut[me.ghost[f] + 1] = np.arange(no + 1, no + ng + 1)
        cell_dofs[f] = np.where(ids > 0, lut[np.where(ids > 0, ids, 0)], ids)
        nfree[f] = no + ng
        nowned[f] = no
    if isinstance(fes, H1H1Spaces):
        lfes = H1H1Spaces(mesh=lmesh, tables=fes.tables, cell_dofs=cell_dofs, nfree=nfree, ndir=dict(fes.ndir),
                          dirichlet_values=fes.dirichlet_values, field_order=fes.field_order,
                          cell_solid=None if fes.cell_solid is None else fes.cell_solid[me.cells],
                          phi_node_coords=None if fes.phi_node_coords is None else fes.phi_node_coords[me.cells])
    else:
        lfes = FESpaces(mesh=lmesh, tables=fes.tables, cell_dofs=cell_dofs, nfree=nfree, ndir=dict(fes.ndir),
                        dirichlet_values=fes.dirichlet_values, j_sign=fes.j_sign[me.cells], field_order=fes.field_order,
                        u_node_coords=None if fes.u_node_coords is None else fes.u_node_coords[me.cells],
                        cell_solid=None if fes.cell_solid is None else fes.cell_solid[me.cells],
                        cell_sigma=None if fes.cell_sigma is None else fes.cell_sigma[me.cells])
    ps = PartitionedSpaces(fes=lfes, rank=rank, nparts=nparts, nowned=nowned, nowned_cells=me.nowned_cells, cells=me.cells,
                           own_global=me.owned, ghost_global=me.ghost)
    ps._global_offsets = fes.offsets
    own_off, gh_off = ps.offsets()
    # ---- halo plan
    neigh_recv = {}
    for f in fes.field_order:
        own_of_ghost = owners[f][me.ghost[f]]
        for s in np.unique(own_of_ghost):
            sel = np.nonzero(own_of_ghost == s)[0]
            neigh_recv.setdefault(int(s), []).append(gh_off[f] + sel)
    neigh_send = {}
    for s in range(nparts):
        if s == rank:
            continue
        other = local_sets(fes, cell_part, owners, s)
        for f in fes.field_order:
            g = other.ghost[f]
            mine = g[owners[f][g] == rank]
            if len(mine):
                pos = np.searchsorted(me.owned[f], mine)
                neigh_send.setdefault(s, []).append(own_off[f] + pos)
    neigh = sorted(set(neigh_recv) | set(neigh_send))
    sp, si, rp, ri = [0], [], [0], []
    for s in neigh:
        a = np.concatenate(neigh_send.get(s, [np.zeros(0, np.int64)]))
        b = np.concatenate(neigh_recv.get(s, [np.zeros(0, np.int64)]))
        si.append(a)
        ri.append(b)
        sp.append(sp[-1] + len(a))
        rp.append(rp[-1] + len(b))
    ps.neigh = np.array(neigh, dtype=np.int32)
    ps.send_ptr = np.array(sp, dtype=np.int64)
    ps.recv_ptr = np.array(rp, dtype=np.int64)
    ps.send_idx = (np.concatenate(si) if si else np.zeros(0)).astype(np.int32)
    ps.recv_idx = (np.concatenate(ri) if ri else np.zeros(0)).astype(np.int32)
    return ps


def hunt_cell_partition(mesh: HexMesh, np_xy) -> np.ndarray:
    """(px,py,1) block partition of the Hunt mesh (hunt_mesher.jl:116-118)."""
    return cartesian_partition(mesh.grid_shape, (np_xy[0], np_xy[1], 1))


def default_cell_partition(mesh: HexMesh, nparts: int, np_xy=None) -> np.ndarray:
    """cell -> part: the (px,py,1) block partition of structured Hunt meshes (hunt_mesher.jl:116-118) when `np_xy` is
    given, else recursive coordinate bisection of the cell centroids -- the stand-in for the METIS partition GridapGmsh
    produces for the Expansion meshes (expansion.jl:278); any other cell -> part array can be passed to
    `distribute_operator(cell_part=...)` directly."""
    if np_xy is not None and mesh.grid_shape is not None:
        return hunt_cell_partition(mesh, np_xy)
    from .mesh import rcb_partition

    return rcb_partition(mesh.cell_coords().mean(axis=1), nparts)


def distribute_operator(fes_global: FESpaces, params, np_xy, rank: int, world: int, dist=None, use_peer_memory=True, cell_part=None):
    """Create this rank's `B200FEOperator` (owned rows, ghost-cell redundant integration), install the halo plan and
    bring up the NCCL communicator of the library (the unique id travels through torch.distributed / MPI).
    `cell_part` [ncells] -> rank: an arbitrary partition (METIS-style, expansion.jl:278); default: see
    `default_cell_partition`."""
    import ctypes as C

    from .. import lib as L
    from ..feoperator import B200FEOperator, B200H1H1FEOperator

    if cell_part is None:
        cell_part = default_cell_partition(fes_global.mesh, world, np_xy)
    cell_part = np.asarray(cell_part)
    assert cell_part.shape == (fes_global.mesh.ncells,) and cell_part.min() >= 0 and cell_part.max() < max(world, 1)
    ps = partition_fespaces(fes_global, cell_part, rank)
    lib = L.load()
    if world > 1:
        import torch

        idbuf = torch.zeros(128, dtype=torch.uint8)
        if rank == 0:
            raw = (C.c_ubyte * 128)()
            L.check(lib.mhd_comm_get_unique_id(raw))
            idbuf = torch.tensor(list(raw), dtype=torch.uint8)
        dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
        idbuf = idbuf.to(dev)
        dist.broadcast(idbuf, src=0)
        raw = (C.c_ubyte * 128)(*idbuf.cpu().tolist())
        L.check(lib.mhd_comm_init(rank, world, raw))
    cls = B200H1H1FEOperator if isinstance(fes_global, H1H1Spaces) else B200FEOperator
    op = cls(ps.fes, params["fluid"], nowned=ps.nowned)
    L.check(lib.mhd_operator_set_halo(op.handle, len(ps.neigh), L.ptr(ps.neigh), L.ptr(ps.send_ptr), L.ptr(ps.send_idx),
                                      L.ptr(ps.recv_ptr), L.ptr(ps.recv_idx)))
    if world > 1 and len(ps.neigh) > 0 and use_peer_memory:
        op.allocate_jacobian()  # the fused kernel needs the pattern (rows with ghost columns)
        # fused SpMV + halo over NVLink peer memory: exchange the CUDA IPC handles of the inboxes through the host
        raw = (C.c_ubyte * 64)()
        L.check(lib.mhd_operator_halo_ipc_export(op.handle, raw))
        mine = {"handle": bytes(raw), "neigh": ps.neigh.tolist(), "recv_ptr": ps.recv_ptr.tolist(),
                "recv_idx": ps.recv_idx.tolist(), "nrows": ps.nrows, "ncols": ps.ncols}
        allinfo = [None] * world
        dist.all_gather_object(allinfo, mine)
        handles = b"".join(allinfo[s]["handle"] for s in ps.neigh)
        slot = np.array([allinfo[s]["neigh"].index(rank) for s in ps.neigh], dtype=np.int32)
        nghost = np.array([allinfo[s]["ncols"] - allinfo[s]["nrows"] for s in ps.neigh], dtype=np.int64)
        # ghost slot on the neighbour of every value this rank sends (both lists are ordered identically)
        dst = []
        for k, s_ in enumerate(ps.neigh):
            info = allinfo[s_]
            kk = info["neigh"].index(rank)
            seg = np.array(info["recv_idx"][info["recv_ptr"][kk] : info["recv_ptr"][kk + 1]], dtype=np.int64) - info["nrows"]
            assert len(seg) == ps.send_ptr[k + 1] - ps.send_ptr[k]
            dst.append(seg)
        send_dst = np.concatenate(dst).astype(np.int32) if dst else np.zeros(0, np.int32)
        hbuf = (C.c_ubyte * len(handles)).from_buffer_copy(handles)
        L.check(lib.mhd_operator_halo_ipc_connect(op.handle, hbuf, L.ptr(send_dst), L.ptr(slot), L.ptr(nghost)))
        dist.barrier()
    return op, ps
