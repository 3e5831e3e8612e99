"""Multi-GPU host logic on CPU: partition of the FE spaces, halo plan and the distributed SpMV/dot data flow with
world_size=2 over gloo (the GPU path replaces the gloo calls by NCCL inside libmhdb200)."""
import os
import socket

import numpy as np
import pytest

from gridapmhd_jl_b200.applications import hunt_params, setup_spaces
from gridapmhd_jl_b200.host.partition import dof_owners, hunt_cell_partition, partition_fespaces


def _case():
    params = hunt_params(nc=(4, 3), B=(0.0, 20.0, 0.0), solver="badia2024")
    return params, setup_spaces(params)


def test_partition_is_a_partition_and_plans_match():
    params, fes = _case()
    for np_xy in ((2, 1), (2, 2)):
        cell_part = hunt_cell_partition(fes.mesh, np_xy)
        nparts = np_xy[0] * np_xy[1]
        parts = [partition_fespaces(fes, cell_part, r) for r in range(nparts)]
        # every free dof owned exactly once; owned cells partition the mesh
        for f in ("u", "p", "j", "phi"):
            allown = np.concatenate([p.own_global[f] for p in parts])
            assert len(allown) == fes.nfree[f] and len(np.unique(allown)) == fes.nfree[f]
        assert sum(p.nowned_cells for p in parts) == fes.mesh.ncells
        # send list r->s and receive list s<-r describe the same global dofs in the same order
        for r, pr in enumerate(parts):
            gr = pr.local_vector_ids()
            for k, s in enumerate(pr.neigh):
                ps = parts[s]
                gs = ps.local_vector_ids()
                ks = list(ps.neigh).index(r)
                sent = gr[pr.send_idx[pr.send_ptr[k] : pr.send_ptr[k + 1]]]
                recv = gs[ps.recv_idx[ps.recv_ptr[ks] : ps.recv_ptr[ks + 1]]]
                assert np.array_equal(sent, recv)
            assert np.all(pr.send_idx < pr.nrows) and np.all(pr.recv_idx >= pr.nrows)
            # all ghosts are received exactly once
            assert len(np.unique(pr.recv_idx)) == pr.ncols - pr.nrows == len(pr.recv_idx)


def test_arbitrary_partition_of_the_reference_expansion_mesh():
    """An arbitrary (non-Cartesian) cell partition -- recursive coordinate bisection of the reference's Expansion_6k mesh into
    2, 4 and 8 parts, the stand-in for the METIS partition of expansion.jl:278: every dof owned once, every ghost received
    exactly once from its owner, send and receive lists pair up."""
    from gridapmhd_jl_b200.applications import expansion_params
    from gridapmhd_jl_b200.host import mesh as M
    from gridapmhd_jl_b200.host.partition import default_cell_partition

    m = M.load_mesh_npz(os.path.join(os.path.dirname(__file__), "golden", "expansion_6k_mesh.npz"))
    params = expansion_params(Ha=100.0, N=3740.0, mesh=m, solver="badia2024")
    fes = setup_spaces(params)
    for nparts in (2, 4, 8):
        cell_part = default_cell_partition(fes.mesh, nparts)
        counts = np.bincount(cell_part, minlength=nparts)
        assert counts.min() >= 0.9 * m.ncells / nparts and counts.max() <= 1.1 * m.ncells / nparts
        parts = [partition_fespaces(fes, cell_part, r) for r in range(nparts)]
        for f in ("u", "p", "j", "phi"):
            allown = np.concatenate([p.own_global[f] for p in parts])
            assert len(allown) == fes.nfree[f] and len(np.unique(allown)) == fes.nfree[f]
        assert sum(p.nowned_cells for p in parts) == m.ncells
        for r, pr in enumerate(parts):
            gr = pr.local_vector_ids()
            for k, s in enumerate(pr.neigh):
                ps = parts[s]
                ks = list(ps.neigh).index(r)
                sent = gr[pr.send_idx[pr.send_ptr[k] : pr.send_ptr[k + 1]]]
                recv = ps.local_vector_ids()[ps.recv_idx[ps.recv_ptr[ks] : ps.recv_ptr[ks + 1]]]
                assert np.array_equal(sent, recv)
            assert len(np.unique(pr.recv_idx)) == pr.ncols - pr.nrows == len(pr.recv_idx)


def test_local_rows_reproduce_global_rows():
    """Rows assembled by a rank (owned + ghost cells) equal the same rows of the global matrix: no entry has to cross
    ranks ("fully assembled rows")."""
    from oracle import mhd_oracle as O

    params, fes = _case()
    fl = params["fluid"]
    prm = O.FluidParams(fl.alpha, fl.beta, fl.gamma, fl.sigma, fl.zeta_u, fl.zeta_j, fl.B, fl.f, fl.g, fl.convection)
    x = np.random.default_rng(0).random(fes.ndofs)
    A = O.jacobian(fes, x, prm)
    cell_part = hunt_cell_partition(fes.mesh, (2, 1))
    for r in range(2):
        ps = partition_fespaces(fes, cell_part, r)
        gl = ps.local_vector_ids()
        xl = x[gl]
        # local assembly with the oracle on the local spaces, laid out as the library does: [owned | ghosts]
        lf = ps.fes
        own_off, gh_off = ps.offsets()
        cols = []
        for f in ("u", "p", "j", "phi"):
            ids = lf.cell_dofs[f]
            no = ps.nowned[f]
            g = np.where(ids > 0, np.where(ids <= no, own_off[f] + ids - 1, gh_off[f] + ids - 1 - no), -1)
            cols.append(g)
        gids = np.concatenate(cols, axis=1)
        st = np.concatenate([np.where(lf.cell_dofs[f] > 0, xl[np.where(gids[:, s0:s1] >= 0, gids[:, s0:s1], 0)],
                                      lf.dirichlet_values[f][np.where(lf.cell_dofs[f] < 0, -lf.cell_dofs[f] - 1, 0)] if len(lf.dirichlet_values[f]) else 0.0)
                             for f, (s0, s1) in zip(("u", "p", "j", "phi"), ((0, 81), (81, 85), (85, 121), (121, 129)))], axis=1)
        K = O.cell_jacobians(lf.tables, lf.mesh.cell_coords(), st, lf.j_sign, prm)
        rows_gids = np.where(gids < ps.nrows, gids, -1)  # drop non-owned rows
        import scipy.sparse as sp

        mask = O.touched_mask()
        li, lj = np.nonzero(mask)
        rr, cc, vv = rows_gids[:, li], gids[:, lj], K[:, li, lj]
        ok = (rr >= 0) & (cc >= 0)
        Al = sp.coo_matrix((vv[ok], (rr[ok], cc[ok])), shape=(ps.nrows, ps.ncols)).tocsr()
        Aref = A[gl[: ps.nrows]][:, gl]
        assert abs(Al - Aref).max() < 1e-12 * abs(Aref).max()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, ret):
    import torch
    import torch.distributed as dist

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import mhd_oracle as O

    params, fes = _case()
    fl = params["fluid"]
    prm = O.FluidParams(fl.alpha, fl.beta, fl.gamma, fl.sigma, fl.zeta_u, fl.zeta_j, fl.B, fl.f, fl.g, fl.convection)
    x = np.random.default_rng(0).random(fes.ndofs)
    v = np.random.default_rng(1).standard_normal(fes.ndofs)
    A = O.jacobian(fes, x, prm)
    cell_part = hunt_cell_partition(fes.mesh, (2, 1))
    ps = partition_fespaces(fes, cell_part, rank)
    gl = ps.local_vector_ids()
    # local vector: owned values known, ghosts poisoned; halo exchange (the library does pack / ncclSend+Recv / unpack)
    vl = np.full(ps.ncols, np.nan)
    vl[: ps.nrows] = v[gl[: ps.nrows]]
    reqs, bufs = [], []
    for k, s in enumerate(ps.neigh):
        snd = torch.from_numpy(vl[ps.send_idx[ps.send_ptr[k] : ps.send_ptr[k + 1]]].copy())
        rcv = torch.empty(int(ps.recv_ptr[k + 1] - ps.recv_ptr[k]), dtype=torch.float64)
        reqs.append(dist.isend(snd, int(s)))
        reqs.append(dist.irecv(rcv, int(s)))
        bufs.append((k, rcv, snd))
    for r in reqs:
        r.wait()
    for k, rcv, _ in bufs:
        vl[ps.recv_idx[ps.recv_ptr[k] : ps.recv_ptr[k + 1]]] = rcv.numpy()
    ok_halo = bool(np.array_equal(vl, v[gl]))
    # distributed SpMV on owned rows + all-reduced dot
    Aloc = A[gl[: ps.nrows]][:, gl]
    yl = Aloc @ vl
    ok_spmv = bool(np.abs(yl - (A @ v)[gl[: ps.nrows]]).max() < 1e-12 * np.abs(A @ v).max())
    d = torch.tensor([float(yl @ vl[: ps.nrows])], dtype=torch.float64)
    dist.all_reduce(d)
    ok_dot = bool(abs(d.item() - (A @ v) @ v) < 1e-10 * abs((A @ v) @ v))
    ret[rank] = (ok_halo, ok_spmv, ok_dot)
    dist.destroy_process_group()


def test_halo_exchange_and_allreduce_world2_gloo():
    import torch.multiprocessing as mp

    port = _free_port()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(2, port, ret), nprocs=2, join=True)
    assert ret[0] == (True, True, True) and ret[1] == (True, True, True)


def test_h1h1_spaces_partition_and_local_rows_reproduce_global_rows():
    """the same decomposition for the H1-H1 spaces (u, p, continuous Q3 phi): every free dof owned once, the halo plans
    of neighbours match, and the rows a rank assembles from its owned + ghost cells are the global rows"""
    import scipy.sparse as sp

    from oracle import mhd_oracle as O
    from oracle import mhd_oracle_h1h1 as H

    params = hunt_params(nc=(4, 3), B=(0.0, 20.0, 0.0), current_disc="H1")
    fes = setup_spaces(params)
    fl = params["fluid"]
    prm = O.FluidParams(fl.alpha, fl.beta, fl.gamma, fl.sigma, fl.zeta_u, fl.zeta_j, fl.B, fl.f, fl.g, fl.convection)
    x = np.random.default_rng(0).random(fes.ndofs)
    A = H.jacobian(fes, x, prm)
    cell_part = hunt_cell_partition(fes.mesh, (2, 2))
    parts = [partition_fespaces(fes, cell_part, r) for r in range(4)]
    for f in ("u", "p", "phi"):
        allown = np.concatenate([p.own_global[f] for p in parts])
        assert len(allown) == fes.nfree[f] and len(np.unique(allown)) == fes.nfree[f]
    for r, pr in enumerate(parts):
        gr = pr.local_vector_ids()
        for k, s in enumerate(pr.neigh):
            ps = parts[s]
            ks = list(ps.neigh).index(r)
            sent = gr[pr.send_idx[pr.send_ptr[k] : pr.send_ptr[k + 1]]]
            recv = ps.local_vector_ids()[ps.recv_idx[ps.recv_ptr[ks] : ps.recv_ptr[ks + 1]]]
            assert np.array_equal(sent, recv)
        assert len(np.unique(pr.recv_idx)) == pr.ncols - pr.nrows == len(pr.recv_idx)
        # local assembly on the local spaces in the library layout [owned | ghosts]
        lf = pr.fes
        own_off, gh_off = pr.offsets()
        cols = []
        for f in ("u", "p", "phi"):
            ids, no = lf.cell_dofs[f], pr.nowned[f]
            cols.append(np.where(ids > 0, np.where(ids <= no, own_off[f] + ids - 1, gh_off[f] + ids - 1 - no), -1))
        gids = np.concatenate(cols, axis=1)
        xl = x[gr]
        st = np.where(gids >= 0, xl[np.where(gids >= 0, gids, 0)], 0.0)
        for f, (s0, s1) in zip(("u", "p", "phi"), ((0, 81), (81, 85), (85, 149))):
            ids = lf.cell_dofs[f]
            if len(lf.dirichlet_values[f]):
                st[:, s0:s1] = np.where(ids < 0, lf.dirichlet_values[f][np.where(ids < 0, -ids - 1, 0)], st[:, s0:s1])
        K = H.cell_jacobians(lf.tables, lf.mesh.cell_coords(), st, prm)
        li, lj = np.nonzero(H.touched_mask())
        rr, cc, vv = np.where(gids < pr.nrows, gids, -1)[:, li], gids[:, lj], K[:, li, lj]
        ok = (rr >= 0) & (cc >= 0)
        Al = sp.coo_matrix((vv[ok], (rr[ok], cc[ok])), shape=(pr.nrows, pr.ncols)).tocsr()
        Aref = A[gr[: pr.nrows]][:, gr]
        assert abs(Al - Aref).max() < 1e-12 * abs(Aref).max()
