"""Multi-GPU parity check, run under torchrun (one rank per GPU):

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
      tests/multigpu_check.py

Every rank assembles its owned rows (owned + ghost cells) on its GPU and the results are compared with the oracle's
global matrix/residual; then the distributed SpMV (NCCL halo exchange), dot (NCCL all-reduce) and a short distributed
FGMRES are compared with the global ones.  Prints "MULTIGPU_OK <world>" on rank 0 when every rank passed.

`MHD_CHECK_FORMULATION=h1h1` runs the same checks for the H1-H1 formulation (u, p, continuous Q3 phi; Jacobi-preconditioned
FGMRES).  `MHD_CHECK_CASE=expansion6k` runs them on the reference's Expansion_6k mesh (4 320 non-affine hexes, fixture
tests/golden/expansion_6k_mesh.npz) with an ARBITRARY cell partition (recursive coordinate bisection: the stand-in for the
METIS partition of expansion.jl:278).  `MHD_CHECK_STRESS=N` appends N back-to-back fused SpMV + halo products (arrival
counters, double-buffered inboxes and the acquire ordering under load; N = 10000 at 8 GPUs is the stress run).
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist

    import gridapmhd_jl_b200  # noqa: F401
    from gridapmhd_jl_b200 import lib as L
    from gridapmhd_jl_b200.applications import hunt_params, setup_spaces
    from gridapmhd_jl_b200.feoperator import B200LinearSolver, B200SolverOptions
    from gridapmhd_jl_b200.host.partition import distribute_operator
    from oracle import mhd_oracle as O

    rank, world, lrank = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(lrank)
    L.init(lrank)
    dist.init_process_group("nccl", device_id=torch.device("cuda", lrank))
    np_xy = {1: (1, 1), 2: (2, 1), 4: (2, 2), 8: (4, 2)}[world]
    h1h1 = os.environ.get("MHD_CHECK_FORMULATION", "hdiv") == "h1h1"
    case = os.environ.get("MHD_CHECK_CASE", "hunt")
    cell_part = None
    if case == "expansion6k":
        from gridapmhd_jl_b200.applications import expansion_params
        from gridapmhd_jl_b200.host import mesh as M
        from gridapmhd_jl_b200.host.partition import default_cell_partition

        OH = O
        mesh = M.load_mesh_npz(os.path.join(ROOT, "tests", "golden", "expansion_6k_mesh.npz"))
        params = expansion_params(Ha=100.0, N=3740.0, zeta_u=10.0, zeta_j=10.0, mesh=mesh, solver="badia2024")
        cell_part = default_cell_partition(mesh, world)
        np_xy = None
    elif h1h1:
        from oracle import mhd_oracle_h1h1 as OH

        params = hunt_params(nc=(6, 4), B=(0.0, 20.0, 0.0), zeta_u=1.0, current_disc="H1")
    else:
        OH = O
        params = hunt_params(nc=(6, 4), B=(0.0, 20.0, 0.0), solver="badia2024", zeta_u=1.0, zeta_j=1.0)
    fes = setup_spaces(params)
    fl = params["fluid"]
    prm = O.FluidParams(fl.alpha, fl.beta, fl.gamma, fl.sigma, fl.zeta_u, fl.zeta_j, fl.B, fl.f, fl.g, fl.convection)
    x = np.random.default_rng(0).random(fes.ndofs)
    v = np.random.default_rng(1).standard_normal(fes.ndofs)
    Ag = OH.jacobian(fes, x, prm)
    rg = OH.residual(fes, x, prm)

    op, ps = distribute_operator(fes, params, np_xy, rank, world, dist, cell_part=cell_part)
    gl = ps.local_vector_ids()
    A = op.allocate_jacobian()
    assert op.nrows == ps.nrows and op.ncols == ps.ncols
    b = np.empty(op.nrows)
    op.residual_and_jacobian_b(b, A, x[gl])
    Aloc = A.to_scipy()
    Aref = Ag[gl[: ps.nrows]][:, gl].tocsr()
    Aref.sort_indices()
    ok = True
    err_a = abs(Aloc - Aref).max() / abs(Aref).max()
    err_r = np.abs(b - rg[gl[: ps.nrows]]).max() / np.abs(rg).max()
    ok &= err_a < 1e-12 and err_r < 1e-12
    # distributed SpMV with poisoned ghosts: the halo exchange must fill them
    vl = torch.full((op.ncols,), float("nan"), dtype=torch.float64, device="cuda")
    vl[: op.nrows] = torch.from_numpy(v[gl[: op.nrows]]).cuda()
    y = op.spmv(vl)
    yref = (Ag @ v)[gl[: op.nrows]]
    err_y = np.abs(y.cpu().numpy() - yref).max() / np.abs(Ag @ v).max()
    import ctypes as C

    fused, tmo = C.c_int32(), C.c_int32()
    L.check(L.load().mhd_operator_halo_status(op.handle, C.byref(fused), C.byref(tmo)))
    want_fused = os.environ.get("MHD_HALO_NCCL") is None and world > 1
    ok &= err_y < 1e-12 and tmo.value == 0 and (fused.value == 1) == want_fused
    if not fused.value:
        ok &= bool(torch.isfinite(vl).all())  # NCCL path: the exchange fills the ghost section of x
    # repeated products (double-buffered inboxes, monotone arrival counters)
    for rep in range(7):
        vr = torch.from_numpy((v * (rep + 2))[gl[: op.nrows]]).cuda()
        vfull = torch.zeros(op.ncols, dtype=torch.float64, device="cuda")
        vfull[: op.nrows] = vr
        yr = op.spmv(vfull)
        ok &= np.abs(yr.cpu().numpy() - (rep + 2) * yref).max() / np.abs(Ag @ v).max() < 1e-11
    nstress = int(os.environ.get("MHD_CHECK_STRESS", "0"))
    if nstress:
        vfull = torch.zeros(op.ncols, dtype=torch.float64, device="cuda")
        vfull[: op.nrows] = torch.from_numpy(v[gl[: op.nrows]]).cuda()
        yy = torch.empty(op.nrows, dtype=torch.float64, device="cuda")
        bad = 0
        for rep in range(nstress):  # enqueued back to back: neighbours run ahead / behind each other by whole products
            op.spmv(vfull, yy)
            if rep % 97 == 0 or rep == nstress - 1:
                bad += int(np.abs(yy.cpu().numpy() - yref).max() / np.abs(Ag @ v).max() >= 1e-12)
        L.check(L.load().mhd_operator_halo_status(op.handle, C.byref(fused), C.byref(tmo)))
        ok &= bad == 0 and tmo.value == 0
        print(f"[rank {rank}] stress: {nstress} fused products, {bad} wrong, timed_out={tmo.value}", flush=True)
    d = op.dot(vl, vl)  # all-reduced over ranks (owned entries only)
    err_d = abs(d - v @ v) / (v @ v)
    ok &= err_d < 1e-13
    # distributed FGMRES (inner Jacobi-GMRES): same residual history on every rank, decreasing
    ns = B200LinearSolver(B200SolverOptions(m=20, maxiter=20, rtol=1e-12, atol=0.0, uj_inner_its=20, uj_inner_restart=20,
                                            precond="jacobi" if h1h1 else "block_tri")).symbolic_setup(A).numerical_setup()
    dx = np.zeros(op.nrows)
    ns.solve_b(dx, -b)
    h = torch.tensor(ns.history, dtype=torch.float64, device="cuda")
    hmax, hmin = h.clone(), h.clone()
    dist.all_reduce(hmax, op=dist.ReduceOp.MAX)
    dist.all_reduce(hmin, op=dist.ReduceOp.MIN)
    ok &= bool(torch.equal(hmax, hmin)) and ns.history[-1] < (1.0 if (h1h1 or case != "hunt") else 0.2) * ns.history[0]
    # true residual of the distributed solution against the global matrix
    xs = np.zeros(fes.ndofs)
    xs[gl[: op.nrows]] = dx
    t = torch.from_numpy(xs).cuda()
    dist.all_reduce(t)
    true_res = np.linalg.norm(Ag @ t.cpu().numpy() + rg)
    ok &= abs(true_res - ns.resnorm) < 1e-6 * ns.history[0]
    print(f"[rank {rank}] rows {op.nrows} cols {op.ncols} nnz {op.nnz} errA {err_a:.1e} errR {err_r:.1e} errY {err_y:.1e} "
          f"errDot {err_d:.1e} fused={fused.value} fgmres {ns.history[0]:.3e}->{ns.history[-1]:.3e} true {true_res:.3e} ok={ok}", flush=True)
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    ns.destroy()
    op.destroy()
    L.load().mhd_comm_finalize()
    dist.destroy_process_group()
    L.finalize()
    if rank == 0:
        print(("MULTIGPU_OK %d" % world) if flag.item() == 1 else "MULTIGPU_FAIL", flush=True)
    sys.exit(0 if flag.item() == 1 else 1)


if __name__ == "__main__":
    main()
