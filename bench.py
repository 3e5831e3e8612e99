#!/usr/bin/env python
"""bench.py -- headline benchmark of the hot path: MHD Jacobian (+residual) assembly Mcells/s and Krylov SpMV GB/s.

  python bench.py --gpus N --steps K --warmup W            # our arm (CUDA, C ABI)
  python bench.py --impl reference --gpus N --steps K ...  # CPU arm: the oracle restatement on the host cores

Workload at N=1: BASELINE.json configs[1] -- Hunt duct nc=(64,64), Ha=1000 (12 288 cells, 732 690 dofs,
146 981 976 nnz), Q2/P1disc/RT1/Q1disc, 27-point Gauss, convection :newton with a seeded random state
(mirrors main.jl:137-141).  N>1: weak scaling, 64x64x3 cells per GPU, Cartesian (px,py,1) partition
(hunt_mesher.jl:116-118), ghost-cell redundant integration, halo exchange + all-reduce over NCCL.

One step = one linearisation of the nonlinear problem at a given state: residual_and_jacobian!(b,A,op,x)
(what every Newton iteration of solve!(xh,solver,op) needs, main.jl:275), one fused kernel + the nzval memset.  `value` = cells/s with x resident in
HBM; `e2e` = the same through the host-buffer API (H2D of x, D2H of the residual inside the timed region; the
matrix stays on the device behind the handle, as a PETSc Mat would).  The SpMV leg is reported under "spmv".
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")  # stdout carries the single JSON line only (NCCL's banner goes to stderr)

import numpy as np  # noqa: E402

FLOP_PER_CELL_NEWTON = 0.86e6  # structure-exploiting count, FMA = 2, Newton convection, fused residual excluded
NC_PER_GPU = (64, 64)
HA = 1000.0
PARTS = {1: (1, 1), 2: (2, 1), 4: (2, 2), 8: (4, 2)}


def algorithmic_bytes_jacobian(ncells, nnz, nentries):
    """8 B per stored value (written once) + 4 B per scattered entry (int32 scatter map)
    + 1740 B per cell (8 nodes x 24 B + 129 ids x 4 B + 129 state x 8 B)."""
    return 8 * nnz + 4 * nentries + 1740 * ncells


def algorithmic_bytes_spmv(nrows, ncols, nnz):
    """12 B/nnz (value + int32 column) + y + x once + rowptr."""
    return 12 * nnz + 8 * nrows + 8 * ncols + 4 * (nrows + 1)


class ClockSampler:
    """nvidia-smi clocks during the timed region (B200_PROFILING.md recipe)."""

    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index=0):
        self.samples = []
        self.proc = None
        self.index = index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.index), "-lms", "20"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.samples.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for s in self.samples:
            p = [t.strip() for t in s.split(",")]
            if len(p) < 6:
                continue
            try:
                sm.append(float(p[0]))
                mx = float(p[1])
            except ValueError:
                continue
            for n, v in zip(names, p[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


def workload_string(nc):
    """`config.workload`, identical in both arms"""
    return (f"Hunt duct nc=({nc[0]},{nc[1]},3) Ha={HA:g} Q2/P1disc/RT1/Q1disc 27-pt Gauss, convection newton, seeded random state; "
            f"step = residual_and_jacobian! (one Newton linearisation)")


def oracle_params(fl):
    from oracle import mhd_oracle as O

    return O.FluidParams(fl.alpha, fl.beta, fl.gamma, fl.sigma, fl.zeta_u, fl.zeta_j, fl.B, fl.f, fl.g, fl.convection)


NC_GLOBAL = None  # --nc-global NX NY: fixed global mesh (strong scaling over N GPUs; a large per-GPU point at N = 1)


def build_case(nparts, rank):
    """Hunt cfg2-per-GPU mesh (weak scaling, the default) or a fixed global mesh (--nc-global), FE spaces, operator inputs."""
    import gridapmhd_jl_b200  # noqa: F401
    from gridapmhd_jl_b200.applications import hunt_params, setup_spaces

    px, py = PARTS[nparts]
    nc = (NC_PER_GPU[0] * px, NC_PER_GPU[1] * py) if NC_GLOBAL is None else tuple(NC_GLOBAL)
    params = hunt_params(nc=nc, B=(0.0, HA, 0.0), solver="badia2024", convection="newton")
    fes = setup_spaces(params)
    return params, fes, nc


def run_reference(args):
    """CPU arm: the oracle's C restatement (oracle/mhd_oracle.c, OpenMP over cells/rows) on ALL host cores of the box, on the
    mesh of the N-GPU run (same `config.workload` as our arm).  A step integrates a bounded SAMPLE of the workload -- cells
    [0, ref_cells) -- and `ms_per_step` is the MEASURED time of that sample (nothing is extrapolated); `value` = sampled
    cells / that time.  Under torchrun only rank 0 works (torchrun exports OMP_NUM_THREADS=1: overridden here)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    ncores = int(os.environ.get("MHD_REF_THREADS", os.cpu_count() or 1))
    os.environ["OMP_NUM_THREADS"] = str(ncores)  # before the OpenMP runtime of the oracle library starts
    import gridapmhd_jl_b200  # noqa: F401
    from oracle import mhd_oracle as O
    from oracle.c_oracle import COracle

    params, fes, nc = build_case(args.gpus, 0)
    prm = oracle_params(params["fluid"])
    ncells = fes.mesh.ncells
    nsample = min(ncells, args.ref_cells)
    co = COracle(fes, prm, cells=np.arange(nsample), threads=ncores)
    # pattern restricted to the sampled cells keeps the symbolic cost bounded
    gids = fes.cell_global_ids()[:nsample]
    rp, cv = O.symbolic_csr(gids, fes.ndofs)
    x = np.random.default_rng(1234).random(fes.ndofs)
    nz = np.zeros(len(cv))
    times = []
    for it in range(args.warmup + args.steps):
        nz[:] = 0.0
        t0 = time.perf_counter()
        co.residual(x)
        co.jacobian_values(x, rp, cv, out=nz)
        dt = time.perf_counter() - t0
        if it >= args.warmup:
            times.append(dt)
    t = float(np.mean(times))
    val = nsample / t / 1e6
    # SpMV / dot / axpy on the sampled matrix
    v = np.random.default_rng(1).standard_normal(fes.ndofs)
    ts = []
    for it in range(3 + 10):
        t0 = time.perf_counter()
        co.spmv(rp, cv, nz, v)
        if it >= 3:
            ts.append(time.perf_counter() - t0)
    spmv_gbs = (16 * len(cv) + 8 * 2 * fes.ndofs + 8 * (fes.ndofs + 1)) / float(np.mean(ts)) / 1e9
    sample = (f"cells [0,{nsample}) of {ncells}: residual + Jacobian re-assembly into sorted CSR, all {co.threads} host threads "
              f"(C/OpenMP restatement of the algorithm -- a port, not Gridap/Julia); ms_per_step is the measured time of this sample")
    line = {
        "impl": "reference", "metric": "mhd_assembly_jacobian_plus_residual", "value": val, "unit": "Mcells/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": t * 1e3,
        "higher_is_better": True, "scaling": "weak" if NC_GLOBAL is None else "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_string(nc), "ncells": ncells, "ndofs": fes.ndofs, "sample_cells_per_step": nsample},
        "cpu_baseline": {"value": val, "unit": "Mcells/s", "cores": co.threads, "kind": "port", "sample": sample},
        "spmv": {"value": spmv_gbs, "unit": "GB/s", "note": "int64 CSR of the sampled cells, OpenMP"},
        "e2e": {"value": val, "unit": "Mcells/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit_json(line)


def cpu_baseline_leg(fes, params, seconds=12.0):
    """Bounded CPU sample for the cpu_baseline object (rank 0, N=1)."""
    from oracle import mhd_oracle as O
    from oracle.c_oracle import COracle

    prm = oracle_params(params["fluid"])
    nsample = min(fes.mesh.ncells, 2048)
    co = COracle(fes, prm, cells=np.arange(nsample), threads=int(os.environ.get("MHD_REF_THREADS", os.cpu_count() or 1)))
    gids = fes.cell_global_ids()[:nsample]
    rp, cv = O.symbolic_csr(gids, fes.ndofs)
    x = np.random.default_rng(1234).random(fes.ndofs)
    nz = np.zeros(len(cv))
    co.jacobian_values(x, rp, cv, 0, min(nsample, 256), out=nz)  # warm-up
    reps, t_total = 0, 0.0
    while t_total < seconds and reps < 20:
        nz[:] = 0.0
        t0 = time.perf_counter()
        co.residual(x)
        co.jacobian_values(x, rp, cv, out=nz)
        t_total += time.perf_counter() - t0
        reps += 1
    return {"value": nsample * reps / t_total / 1e6, "unit": "Mcells/s", "cores": co.threads, "kind": "port",
            "sample": f"{reps} x cells [0,{nsample}) of {fes.mesh.ncells}: residual + Jacobian re-assembly into sorted "
                      f"CSR, C/OpenMP restatement (oracle/mhd_oracle.c), {t_total:.1f} s of CPU work"}


def h1h1_extra_leg(L, nc, reps=10):
    """H1-H1 formulation (u Q2, p P1disc, phi Q3 continuous) on the bench mesh: kernel times by CUDA events."""
    import torch

    from gridapmhd_jl_b200.applications import hunt_params, make_operator, setup_spaces

    p = hunt_params(nc=(nc[0], nc[1]), B=(0.0, HA, 0.0), current_disc="H1")
    fes = setup_spaces(p)
    op = make_operator(fes, p["fluid"])
    op.allocate_jacobian()
    x = torch.from_numpy(np.random.default_rng(1234).random(fes.ndofs)).cuda()
    y, r = torch.empty_like(x), torch.empty_like(x)
    for _ in range(3):
        op.jacobian(x), op.residual_b(r, x), op.spmv(x, y)
    L.load().mhd_profile_enable(1)
    L.load().mhd_profile_reset()
    for _ in range(reps):
        op.jacobian(x), op.residual_b(r, x), op.spmv(x, y)
    jac, nj = L.profile_get("jacobian")
    res, nr = L.profile_get("residual")
    spmv, ns = L.profile_get("spmv")
    L.load().mhd_profile_enable(0)
    nent, _ = op.scatter_stats()
    ncells = fes.mesh.ncells
    alg = 8.0 * op.nnz + 2.0 * nent + ncells * (8 * 24 + 149 * 4 + 149 * 8 + 149 * 8)  # values + u16 map + coords/ids/state/rowstarts
    out = {"workload": f"Hunt nc=({nc[0]},{nc[1]},3) Ha={HA:g}, H1-H1: Q2/P1disc/Q3, newton convection", "ncells": ncells,
           "ndofs": fes.ndofs, "nnz": op.nnz, "jacobian_kernel_ms": jac / nj, "residual_kernel_ms": res / nr, "spmv_kernel_ms": spmv / ns,
           "jacobian_Mcells_s": ncells / (jac / nj) / 1e3, "jacobian_algorithmic_bytes": alg, "jacobian_GBs": alg / (jac / nj) / 1e6,
           "spmv_GBs": (12.0 * op.nnz + 20.0 * op.nrows) / (spmv / ns) / 1e6}
    op.destroy()
    return out


def expansion6k_extra_leg(L, reps=10):
    """BASELINE configs 3/4 stand-in: the reference's own Expansion_6k.msh (4 320 non-affine hexes; fixture
    tests/golden/expansion_6k_mesh.npz), Expansion parameterisation (expansion.jl:40-181), Newton convection, inlet Dirichlet
    data: fused residual + Jacobian assembly and SpMV, with the parity of every row against the C oracle."""
    import torch

    from gridapmhd_jl_b200.applications import expansion_params, setup_spaces
    from gridapmhd_jl_b200.feoperator import B200FEOperator
    from gridapmhd_jl_b200.host import mesh as M
    from oracle import parity as PAR

    m = M.load_mesh_npz(os.path.join(ROOT, "tests", "golden", "expansion_6k_mesh.npz"))
    params = expansion_params(Ha=100.0, N=3740.0, zeta_u=10.0, zeta_j=10.0, mesh=m)
    fes = setup_spaces(params)
    op = B200FEOperator(fes, params["fluid"])
    A = op.allocate_jacobian()
    xh = np.random.default_rng(6).random(fes.ndofs)
    x = torch.from_numpy(xh).cuda()
    r, y = torch.empty_like(x), torch.empty_like(x)
    for _ in range(3):
        op.residual_and_jacobian_b(r, A, x), op.spmv(x, y)
    L.load().mhd_profile_enable(1)
    L.load().mhd_profile_reset()
    for _ in range(reps):
        op.residual_and_jacobian_b(r, A, x), op.spmv(x, y)
    jac, nj = L.profile_get("jacobian")
    spmv, ns = L.profile_get("spmv")
    L.load().mhd_profile_enable(0)
    rowptr, colval = A.pattern()
    par = PAR.assembly_parity(fes, oracle_params(params["fluid"]), xh, rowptr, colval, A.nzval(), r.cpu().numpy(), op.nrows, ncells=m.ncells)
    out = {"workload": "Expansion_6k.msh (reference mesh, 4320 non-affine hexes), Ha=100 N=3740 zeta=10, newton convection, inlet profile",
           "ncells": m.ncells, "ndofs": fes.ndofs, "nnz": op.nnz, "kernel_version": op.kernel_version,
           "residual_and_jacobian_kernel_ms": jac / nj, "Mcells_s": m.ncells / (jac / nj) / 1e3, "spmv_kernel_ms": spmv / ns,
           "spmv_GBs": (12.0 * op.nnz + 20.0 * op.nrows) / (spmv / ns) / 1e6,
           "parity": {k: par[k] for k in ("jac_rel", "res_rel", "csr_bitexact", "rows_checked")}}
    op.destroy()
    return out


def solve_extra_leg(L, nc):
    """BASELINE config 2 "assembly + GMRES with block preconditioner", converging: Hunt nc=(64,64), Ha=1000 on the reference's
    default (boundary-layer adapted) mesh, augmented Lagrangian zeta = 100 (the reference's tests use 10..100), one Newton step
    from x = 0: FGMRES(30) + Badia2024 block-triangular preconditioner, (u,j) block = inner GMRES(30) preconditioned by the
    vertex-patch block-Jacobi smoother -- all on the device.  Reports iterations, wall time and the TRUE relative residual."""
    from gridapmhd_jl_b200.applications import hunt_params, setup_spaces
    from gridapmhd_jl_b200.feoperator import B200FEOperator, B200LinearSolver, B200SolverOptions

    p = hunt_params(nc=(nc[0], nc[1]), B=(0.0, HA, 0.0), zeta_u=100.0, zeta_j=100.0, solver="badia2024")
    fes = setup_spaces(p)
    op = B200FEOperator(fes, p["fluid"])
    A = op.allocate_jacobian()
    b = np.empty(op.nrows)
    op.residual_and_jacobian_b(b, A, np.zeros(fes.ndofs))
    t0 = time.perf_counter()
    opts = B200SolverOptions(m=30, maxiter=120, rtol=1e-8, atol=0.0, precond="block_tri", uj_solver="gmres_patch", uj_inner_its=30,
                             uj_inner_restart=30, patch_its=1, patch_omega=1.0)
    ns = B200LinearSolver(opts).symbolic_setup(A).numerical_setup()
    L.check(L.load().mhd_device_synchronize())
    t_setup = time.perf_counter() - t0
    dx = np.zeros(op.nrows)
    t0 = time.perf_counter()
    ns.solve_b(dx, -b)
    t_solve = time.perf_counter() - t0
    true = float(np.linalg.norm(op.spmv(dx) + b) / np.linalg.norm(b))
    out = {"workload": f"Hunt nc=({nc[0]},{nc[1]},3) Ha={HA:g}, zeta_u=zeta_j=100, linear solve of the first Newton step, ndofs {fes.ndofs}",
           "solver": "FGMRES(30) + block-triangular preconditioner (badia2024.jl); (u,j) block: GMRES(30) + vertex-patch smoother (gmg.jl:62-81); "
                     "p, phi blocks: exact cell mass inverses", "iterations": int(ns.iters), "setup_ms": t_setup * 1e3, "solve_ms": t_solve * 1e3,
           "true_relative_residual": true, "estimated_relative_residual": float(ns.history[-1] / ns.history[0]),
           "converged_to_1e-8": bool(true <= 2e-8), "patch_inverse_bytes": 8 * ns.patch_entries}
    ns.destroy()
    op.destroy()
    return out


def patch_extra_leg(L, op, A, fes, reps=10):
    """Vertex-patch block-Jacobi smoother of the (u,j) block on the bench matrix: setup = gather + blocked
    Gauss-Jordan inversion of every patch, apply = one additive sweep (streams the explicit inverses once)."""
    import torch

    from gridapmhd_jl_b200.feoperator import B200LinearSolver, B200SolverOptions

    if tuple(fes.field_order[:2]) != ("u", "j"):
        return {"skipped": "field order %s has no leading (u,j) block" % (fes.field_order,)}
    ns = B200LinearSolver(B200SolverOptions(precond="block_tri", uj_solver="gmres_patch")).symbolic_setup(A).numerical_setup()
    L.load().mhd_profile_enable(1)
    L.load().mhd_profile_reset()
    for _ in range(2):
        ns.numerical_setup_b(A)
    setup, nset = L.profile_get("patch_setup")
    nuj = fes.nfree["u"] + fes.nfree["j"]
    r = torch.from_numpy(np.random.default_rng(0).standard_normal(nuj)).cuda()
    for _ in range(3):
        ns.patch_apply(r)
    L.load().mhd_profile_reset()
    for _ in range(reps):
        ns.patch_apply(r)
    app, napp = L.profile_get("patch_apply")
    L.load().mhd_profile_enable(0)
    out = {"npatches": ns.npatches, "inverse_bytes": 8 * ns.patch_entries, "setup_kernel_ms": setup / nset, "apply_kernel_ms": app / napp,
           "apply_GBs": 8 * ns.patch_entries / (app / napp) / 1e6}
    try:
        # one FGMRES(15) cycle with the block-triangular preconditioner and the inner patch-GMRES(30) on the bench matrix
        # (Jacobian at a random state, zeta = 0, no coarse level): what the Krylov machinery costs with a real preconditioner
        bb = np.random.default_rng(7).standard_normal(A.op.nrows)
        xx = np.zeros(A.op.nrows)
        t0 = time.perf_counter()
        ns.solve_b(xx, bb)
        out["solve"] = {"solver": "FGMRES(15) + block-triangular preconditioner, (u,j) block = GMRES(30) with the patch solver",
                        "wall_ms": (time.perf_counter() - t0) * 1e3, "iterations": int(ns.iters),
                        "residual_reduction": float(ns.history[-1] / ns.history[0]) if len(ns.history) else None}
    except Exception as e:
        out["solve"] = {"error": repr(e)}
    ns.destroy()
    return out


def run_ours(args):
    import torch
    import torch.distributed as dist

    import gridapmhd_jl_b200  # noqa: F401
    from gridapmhd_jl_b200 import lib as L
    from gridapmhd_jl_b200.feoperator import B200FEOperator

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    assert world == args.gpus, f"--gpus {args.gpus} but WORLD_SIZE={world}"
    torch.cuda.set_device(local_rank)
    L.init(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    params, fes_global, nc = build_case(world, rank)
    if world > 1:
        from gridapmhd_jl_b200.host.partition import distribute_operator

        op, part = distribute_operator(fes_global, params, PARTS[world], rank, world, dist)
        fes = part.fes
    else:
        op = B200FEOperator(fes_global, params["fluid"])
        fes = fes_global
    t0 = time.perf_counter()
    A = op.allocate_jacobian()
    L.check(L.load().mhd_device_synchronize())
    t_symbolic = time.perf_counter() - t0
    nentries, nexcl = op.scatter_stats()
    ncells_owned = fes.mesh.ncells if world == 1 else part.nowned_cells
    ncells_local = fes.mesh.ncells
    ncells_global = fes_global.mesh.ncells

    rng = np.random.default_rng(1234 + rank)
    x_host = torch.from_numpy(rng.random(op.ncols)).pin_memory()
    r_host = torch.empty(op.nrows, dtype=torch.float64).pin_memory()
    x_dev = x_host.cuda()
    r_dev = torch.empty(op.nrows, dtype=torch.float64, device="cuda")
    v_dev = torch.from_numpy(rng.standard_normal(op.ncols)).cuda()
    y_dev = torch.empty(op.nrows, dtype=torch.float64, device="cuda")

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step_device():
        op.residual_and_jacobian_b(r_dev, A, x_dev)

    def step_host():
        op.residual_and_jacobian_b(r_host.numpy(), A, x_host.numpy())

    def timed(fn, steps, warmup):
        for _ in range(warmup):
            fn()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item()) / steps

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        t_wait = time.perf_counter()
        while not sampler.samples and time.perf_counter() - t_wait < 5.0:  # nvidia-smi needs ~1 s to produce its first line
            time.sleep(0.02)
    L.load().mhd_profile_enable(1)
    L.load().mhd_profile_reset()
    l0 = L.launch_count()
    ms_step = timed(step_device, args.steps, args.warmup)
    launches = (L.launch_count() - l0) // (args.steps + args.warmup) * args.steps
    jac_ms, jac_n = L.profile_get("jacobian")
    L.load().mhd_profile_reset()
    for _ in range(3):  # separate residual! kernel, for reference
        op.residual_b(r_dev, x_dev)
    res_ms, res_n = L.profile_get("residual")
    L.load().mhd_profile_reset()
    ms_spmv = timed(lambda: op.spmv(v_dev, y_dev), max(args.steps, 20), max(args.warmup, 3))
    spmv_ms, spmv_n = L.profile_get("spmv")
    L.load().mhd_profile_enable(0)
    if os.environ.get("MHD_HALO_DEBUG") and world > 1:  # spin / interface statistics of the fused exchange (stderr)
        import ctypes as C
        fused, tmo = C.c_int32(0), C.c_int32(0)
        L.check(L.load().mhd_operator_halo_status(op.handle, C.byref(fused), C.byref(tmo)))
    # Krylov leg (single GPU): one FGMRES(15) cycle of the reference's configuration (badia2024.jl:40: m = maxiter = 15) on
    # the assembled matrix, enqueued as a whole on the stream (SpMV, fused CGS2 Gram-Schmidt, Givens on the device; the
    # host reads the scalars once per cycle).  Point-Jacobi preconditioner: the block preconditioner's (u,j) solve is
    # not scalable yet, so this times the Krylov machinery, not a converged solve.
    # Runs at every N: the (m+1)-double all-reduces and the fused peer-memory halo are part of the iteration when N > 1.
    krylov = None
    try:
        from gridapmhd_jl_b200.feoperator import B200LinearSolver, B200SolverOptions

        ns = B200LinearSolver(B200SolverOptions(m=15, maxiter=15, rtol=1e-30, atol=0.0, precond="jacobi")).symbolic_setup(A).numerical_setup()
        bb = torch.from_numpy(np.random.default_rng(7 + rank).standard_normal(op.nrows)).cuda()
        xx = torch.zeros(op.nrows, dtype=torch.float64, device="cuda")
        ns.solve_b(xx, bb)  # warm-up (captures the CUDA graph of the restart cycle on one GPU)
        reps = 5

        def one_cycle():
            xx.zero_()
            ns.solve_b(xx, bb)

        ms_cycle = timed(one_cycle, reps, 2)
        krylov = {"solver": "FGMRES(15), point-Jacobi, 15 iterations = one restart cycle, device vectors; CGS2 fused into 4 launches per "
                            "Arnoldi step, cycle replayed as a CUDA graph on one GPU (MHD_KRYLOV_GRAPH=0 disables)",
                  "ms_per_cycle": ms_cycle, "ms_per_iteration": ms_cycle / 15, "iterations": int(ns.iters),
                  "residual_reduction": float(ns.history[-1] / ns.history[0]) if len(ns.history) else None}
        ns.destroy()
    except Exception as e:  # keep the contract line alive
        krylov = {"error": repr(e)}
    # dot / axpy / fused multi-dot bandwidth (BASELINE metric names them): vectors of the operator's size, 31 of them for m = 15
    blas1 = None
    try:
        nv = op.nrows
        Vb = torch.from_numpy(np.random.default_rng(3).standard_normal((16, nv))).cuda()
        wv = torch.from_numpy(np.random.default_rng(4).standard_normal(nv)).cuda()
        hbuf = torch.zeros(16, dtype=torch.float64, device="cuda")
        lib = L.load()
        ms_dot = timed(lambda: L.check(lib.mhd_dot(op.handle, L.ptr(Vb[0]), L.ptr(wv), L.ptr(hbuf))), 50, 5)
        ms_axpy = timed(lambda: L.check(lib.mhd_axpy(op.handle, 0.5, L.ptr(Vb[0]), L.ptr(wv))), 50, 5)
        ms_gs = timed(lambda: L.check(lib.mhd_multi_dot_axpy(op.handle, 16, L.ptr(Vb), nv, L.ptr(wv), L.ptr(hbuf))), 20, 3)
        blas1 = {"n": int(nv), "note": "vectors are L2-resident at this size (5.9 MB each): GB/s are L2, not HBM, figures",
                 "dot_ms": ms_dot, "dot_GBs": 16 * nv / ms_dot / 1e6, "axpy_ms": ms_axpy, "axpy_GBs": 24 * nv / ms_axpy / 1e6,
                 "multi_dot_axpy16_ms": ms_gs, "multi_dot_axpy16_GBs": (17 * 8 + 18 * 8) * nv / ms_gs / 1e6}
    except Exception as e:
        blas1 = {"error": repr(e)}
    # host wall-clock for the e2e leg (host buffers; copies inside): CUDA events bracket it too
    ms_e2e = timed(step_host, args.steps, args.warmup)
    clocks = sampler.stop() if rank == 0 else None
    # second end-to-end figure: the assembled values ALSO return to a pinned host buffer every step (what a host-side
    # direct solver such as the reference's :julia LU would need; julia/GridapMHDB200.jl `jacobian!` with a host matrix)
    big = op.nnz > int(os.environ.get("MHD_BENCH_BIG_NNZ", "600000000"))  # large-mesh point (--nc-global): no 19 GB pinned host copy of the matrix, no host-side parity pass
    nz_host = None if big else torch.empty(op.nnz, dtype=torch.float64).pin_memory()

    def step_host_matrix():
        op.residual_and_jacobian_b(r_host.numpy(), A, x_host.numpy())
        L.check(L.load().mhd_get_nzval(op.handle, L.ptr(nz_host.numpy())))

    ms_e2e_mat = None if big else timed(step_host_matrix, 3, 1)
    # parity of what was just timed (this rank's owned rows) against the C oracle on a block of >= 2048 cells, and of the
    # SpMV (incl. the ghost exchange) against a host product with the device's own matrix -- after every timed region
    parity = None
    if big and not args.no_parity:
        # the matrix stays on the device: only the entries of the sampled rows (a >= 2048-cell block) travel to the host
        import ctypes as C

        from oracle import parity as PAR

        op.residual_and_jacobian_b(r_dev, A, x_dev)
        torch.cuda.synchronize()
        pr, pc, pv = C.c_void_p(), C.c_void_p(), C.c_void_p()
        L.check(L.load().mhd_operator_device_ptrs(op.handle, C.byref(pr), C.byref(pc), C.byref(pv)))

        def dev_view(ptr, n, typestr, dtype):
            holder = type("DevArray", (), {"__cuda_array_interface__": {"shape": (n,), "typestr": typestr, "data": (ptr.value, False), "version": 3}})()
            return torch.as_tensor(holder, device="cuda", dtype=dtype)

        rp_d = dev_view(pr, op.nrows + 1, "<i8", torch.int64)
        cv_d = dev_view(pc, op.nnz, "<i4", torch.int32)
        nz_d = dev_view(pv, op.nnz, "<f8", torch.float64)

        def gather(idx):
            it = torch.from_numpy(np.ascontiguousarray(idx)).cuda()
            return cv_d[it].cpu().numpy(), nz_d[it].cpu().numpy()

        gl = part.local_vector_ids() if world > 1 else np.arange(op.ncols)
        v_np = np.cos(0.37 * gl)
        parity = PAR.assembly_parity(fes, oracle_params(params["fluid"]), x_dev.cpu().numpy(), rp_d.cpu().numpy(), None, None,
                                     r_dev.cpu().numpy(), op.nrows, nowned=part.nowned if world > 1 else None, ncells=2048,
                                     gather=gather, v_lib=v_np)
        v_chk = torch.from_numpy(v_np).cuda()
        if world > 1:
            v_chk[op.nrows:] = float("nan")
        y_chk = op.spmv(v_chk).cpu().numpy()
        rows_lib, y_rows = parity.pop("rows_lib"), parity.pop("y_rows")
        parity["spmv_rel"] = float(np.abs(y_chk[rows_lib] - y_rows).max() / np.abs(y_rows).max())
        # whole-matrix properties (every entry takes part): all values finite; device SpMV against a device-side product
        # assembled from the same CSR with torch (fp64 index_add over all nnz)
        y_t, finite = None, True
        if world == 1:
            y_t = torch.zeros(op.nrows, dtype=torch.float64, device="cuda")
        CH = 1 << 27  # entries per pass: bounded scratch memory
        for e0 in range(0, op.nnz, CH):
            e1 = min(op.nnz, e0 + CH)
            finite = finite and bool(torch.isfinite(nz_d[e0:e1]).all().item())
            if y_t is not None:
                rows_of = torch.searchsorted(rp_d, torch.arange(e0, e1, device="cuda"), right=True) - 1
                y_t.index_add_(0, rows_of, nz_d[e0:e1] * v_chk[cv_d[e0:e1].long()])
                del rows_of
        parity["all_finite"] = finite
        if y_t is not None:
            parity["spmv_full_rel"] = float(((op.spmv(v_chk) - y_t).abs().max() / y_t.abs().max()).item())
        if world > 1:
            t = torch.tensor([parity["jac_rel"], parity["res_rel"], parity["spmv_rel"], 0.0 if parity["csr_bitexact"] and parity["all_finite"] else 1.0],
                             dtype=torch.float64, device="cuda")
            t = torch.nan_to_num(t, nan=float("inf"))
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            parity.update(jac_rel=float(t[0]), res_rel=float(t[1]), spmv_rel=float(t[2]), csr_bitexact=bool(t[3] == 0.0))
        parity["checked_on"] = ("every rank: complete owned rows of a >= 2048-cell block vs oracle/mhd_oracle.c, entries gathered on the device "
                                "(matrix of %d nnz stays in HBM); SpMV on those rows vs a host product, and on ALL rows vs a torch fp64 product on the device" % op.nnz)
        parity["tolerance"] = 1e-12
        parity["ok"] = bool(parity["csr_bitexact"] and parity["all_finite"] and
                            max(parity["jac_rel"], parity["res_rel"], parity["spmv_rel"], parity.get("spmv_full_rel", 0.0)) <= 1e-12)
    elif not args.no_parity:
        from oracle import parity as PAR

        op.residual_and_jacobian_b(r_dev, A, x_dev)
        torch.cuda.synchronize()
        rowptr_h, colval_h = A.pattern(index_bytes=4 if op.nnz < 2**31 - 1 else 8)
        nz_np = nz_host.numpy()
        L.check(L.load().mhd_get_nzval(op.handle, L.ptr(nz_np)))
        parity = PAR.assembly_parity(fes, oracle_params(params["fluid"]), x_dev.cpu().numpy(), rowptr_h, colval_h, nz_np,
                                     r_dev.cpu().numpy(), op.nrows, nowned=part.nowned if world > 1 else None, ncells=2048)
        gl = part.local_vector_ids() if world > 1 else np.arange(op.ncols)
        v_np = np.cos(0.37 * gl)  # a function of the GLOBAL dof id: ghost values agree with their owners
        v_chk = torch.from_numpy(v_np).cuda()
        if world > 1:
            v_chk[op.nrows:] = float("nan")  # the exchange must deliver the ghosts
        y_chk = op.spmv(v_chk)
        parity["spmv_rel"] = PAR.spmv_parity(rowptr_h, colval_h, nz_np, v_np, y_chk.cpu().numpy())
        if world > 1:
            t = torch.tensor([parity["jac_rel"], parity["res_rel"], parity["spmv_rel"], 0.0 if parity["csr_bitexact"] else 1.0],
                             dtype=torch.float64, device="cuda")
            t = torch.nan_to_num(t, nan=float("inf"))
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            parity.update(jac_rel=float(t[0]), res_rel=float(t[1]), spmv_rel=float(t[2]), csr_bitexact=bool(t[3] == 0.0))
        parity["checked_on"] = "every rank: complete owned rows of a >= 2048-cell block vs oracle/mhd_oracle.c; max over ranks"
        parity["tolerance"] = 1e-12
        parity["ok"] = bool(parity["csr_bitexact"] and max(parity["jac_rel"], parity["res_rel"], parity["spmv_rel"]) <= 1e-12)
        del rowptr_h, colval_h
    peaks = {}
    pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(pk):
        peaks = json.load(open(pk))
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"

    import ctypes
    fp64 = {}
    for kind, name in ((0, "dmma_tflops"), (1, "dfma_tflops")):
        v = ctypes.c_double()
        L.check(L.load().mhd_fp64_peak(kind, ctypes.byref(v)))
        fp64[name] = v.value
    # extra legs (single GPU, after every number of the main line has been taken): the H1-H1 formulation on the same mesh
    # and the vertex-patch smoother of the (u,j) block.  Reported beside the main line, never part of `value`.
    h1h1_leg = patch_leg = exp6k_leg = solve_leg = None
    if world == 1 and not args.no_extra:
        try:
            exp6k_leg = expansion6k_extra_leg(L)
        except Exception as e:
            exp6k_leg = {"error": repr(e)}
        try:
            solve_leg = solve_extra_leg(L, nc)
        except Exception as e:
            solve_leg = {"error": repr(e)}
        try:
            h1h1_leg = h1h1_extra_leg(L, nc)
        except Exception as e:  # keep the contract line alive
            h1h1_leg = {"error": repr(e)}
        try:
            patch_leg = patch_extra_leg(L, op, A, fes)
        except Exception as e:
            patch_leg = {"error": repr(e)}
    jac_kernel_ms = jac_ms / max(jac_n, 1)
    jac_bytes = algorithmic_bytes_jacobian(ncells_local, op.nnz, nentries)
    jac_gbs = jac_bytes / (jac_kernel_ms * 1e-3) / 1e9
    spmv_kernel_ms = spmv_ms / max(spmv_n, 1)
    spmv_bytes = algorithmic_bytes_spmv(op.nrows, op.ncols, op.nnz)
    spmv_gbs = spmv_bytes / (spmv_kernel_ms * 1e-3) / 1e9

    value = ncells_global / (ms_step * 1e-3) / 1e6
    e2e = ncells_global / (ms_e2e * 1e-3) / 1e6
    line = {
        "metric": "mhd_assembly_jacobian_plus_residual", "value": value, "unit": "Mcells/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True,
        "scaling": "weak" if NC_GLOBAL is None else "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_string(nc),
                   "ncells": ncells_global, "ncells_per_gpu": ncells_owned, "ndofs_local": op.ncols, "nnz_local": op.nnz,
                   "partition": list(PARTS[world]), "l2_policy": "working set >> L2 (nzval+map = %.2f GB per GPU)" % ((8 * op.nnz + 2 * nentries) / 1e9),
                   "symbolic_s": t_symbolic, "scatter_entries": nentries, "exclusive_entries": nexcl},
        "e2e": {"value": e2e, "unit": "Mcells/s", "ms_per_step": ms_e2e, "h2d_bytes_per_step": 8 * op.ncols,
                "d2h_bytes_per_step": 8 * op.nrows, "note": "matrix stays on the device behind the handle (device-resident solve)"},
        "e2e_with_matrix_d2h": None if ms_e2e_mat is None else {"value": ncells_global / (ms_e2e_mat * 1e-3) / 1e6, "unit": "Mcells/s", "ms_per_step": ms_e2e_mat,
                                "h2d_bytes_per_step": 8 * op.ncols, "d2h_bytes_per_step": 8 * op.nrows + 8 * op.nnz,
                                "note": "nzval also copied to pinned host memory every step (host-side direct solver)"},
        "parity": parity,
        "gpu_launches": int(launches),
        "roofline": {"kernel": {7: "hdiv_v7_jacobian_kernel<CONV=newton,RES=1> (sum-factorised, fused residual_and_jacobian!; zero_shared in its own launch)",
                                5: "jacobian_kernel<CONV=newton,RES=1> (fused residual_and_jacobian!)"}[op.kernel_version], "bound": "hbm", "achieved": jac_gbs, "peak": hbm_peak, "unit": "GB/s",
                     "frac": jac_gbs / hbm_peak, "traffic": None, "peak_source": peak_src, "kernel_ms": jac_kernel_ms,
                     "algorithmic_bytes": jac_bytes, "jacobian_only_Mcells_s": ncells_local / (jac_kernel_ms * 1e-3) / 1e6,
                     # second roofline of the same kernel: FP64 work.  Algorithmic count of DESIGN.md 4.2
                     # (0.86 MFLOP per fluid cell with Newton convection) against the DMMA peak measured on this device
                     "fp64": {"flop_per_cell": FLOP_PER_CELL_NEWTON, "achieved_tflops": FLOP_PER_CELL_NEWTON * ncells_local / (jac_kernel_ms * 1e-3) / 1e12,
                              "peak_tflops": fp64["dmma_tflops"], "dfma_peak_tflops": fp64["dfma_tflops"],
                              "frac": FLOP_PER_CELL_NEWTON * ncells_local / (jac_kernel_ms * 1e-3) / 1e12 / fp64["dmma_tflops"],
                              "peak_source": "measured by mhd_fp64_peak (mma.sync m8n8k4 f64 chains on all SMs)"}},
        "residual": {"kernel_ms": res_ms / max(res_n, 1)},
        "spmv": {"value": spmv_gbs * world, "unit": "GB/s", "ms": ms_spmv, "kernel_ms": spmv_kernel_ms,
                 "roofline": {"bound": "hbm", "achieved": spmv_gbs, "peak": hbm_peak, "unit": "GB/s", "frac": spmv_gbs / hbm_peak,
                              "traffic": None, "algorithmic_bytes": spmv_bytes}},
        "krylov": krylov,
        "blas1": blas1,
        "solve": solve_leg,
        "expansion6k": exp6k_leg,
        "h1h1": h1h1_leg,
        "patch_smoother": patch_leg,
        "clocks": clocks,
    }
    if rank == 0:
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline_leg(fes_global, params)
        traffic_file = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(traffic_file):
            tr = json.load(open(traffic_file))
            line["roofline"]["traffic"] = tr.get("jacobian_kernel_dram_bytes")
            line["spmv"]["roofline"]["traffic"] = tr.get("spmv_dram_bytes")
            line["roofline"]["traffic_source"] = line["spmv"]["roofline"]["traffic_source"] = "profiles/traffic.json (static, from the committed ncu capture)"
        if world == 1 and NC_GLOBAL is None and not args.no_extra:
            live = measure_traffic_with_ncu()  # DRAM bytes of THIS build on THIS box; None when ncu is not usable here
            if live:
                if "jacobian" in live:
                    line["roofline"]["traffic"] = live["jacobian"]
                    line["roofline"]["traffic_source"] = live["source"]
                if "spmv" in live:
                    line["spmv"]["roofline"]["traffic"] = live["spmv"]
                    line["spmv"]["roofline"]["traffic_source"] = live["source"]
        emit_json(line)
    barrier()  # nobody tears its inbox down while a neighbour may still push into it
    op.destroy()
    if world > 1:
        L.load().mhd_comm_finalize()
        dist.destroy_process_group()
    L.finalize()


def measure_traffic_with_ncu():
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the fused Jacobian kernel and of the SpMV, from an ncu run of a
    short child bench (outside every timed region; the child's numbers are discarded).  Returns None if ncu cannot profile here."""
    import shutil
    import tempfile

    ncu = shutil.which("ncu") or ("/usr/local/cuda/bin/ncu" if os.path.exists("/usr/local/cuda/bin/ncu") else None)
    if ncu is None or os.environ.get("MHD_BENCH_NO_NCU"):
        return None
    if any(k in os.environ for k in ("CUDA_INJECTION64_PATH", "NV_COMPUTE_PROFILER_PERFWORKS_DIR", "NSYS_PROFILING_SESSION_ID")):
        return None  # this process is itself running under a profiler: do not nest another one
    with tempfile.TemporaryDirectory() as tmp:
        log = os.path.join(tmp, "dram.csv")
        cmd = [ncu, "--metrics", "dram__bytes_read.sum,dram__bytes_write.sum", "--clock-control", "none", "--print-units", "base",
               "-k", "regex:hdiv_v7_jacobian_kernel|jacobian_kernel|spmv_warp_row", "-c", "12", "--csv", "--log-file", log,
               sys.executable, os.path.abspath(__file__), "--steps", "2", "--warmup", "1", "--no-cpu-baseline", "--no-parity", "--no-extra"]
        env = dict(os.environ, MHD_BENCH_NO_NCU="1")
        try:
            subprocess.run(cmd, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL, timeout=150, env=env, check=False)
            out = parse_ncu_dram_csv(log)
        except Exception:
            return None
    if out:
        out["source"] = "ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum --clock-control none on a child run of this bench (same build, same box)"
    return out or None


def parse_ncu_dram_csv(path):
    """{"jacobian": bytes, "spmv": bytes} from an ncu --csv log holding dram__bytes_read.sum / dram__bytes_write.sum per launch"""
    import csv

    rows = [r for r in csv.reader(open(path)) if len(r) > 6]
    if not rows or "Kernel Name" not in rows[0]:
        return None
    h = rows[0]
    ki, mi, vi, ii = h.index("Kernel Name"), h.index("Metric Name"), h.index("Metric Value"), h.index("ID")
    per = {}
    for r in rows[1:]:
        try:
            per.setdefault((r[ii], r[ki]), {})[r[mi]] = float(r[vi].replace(",", ""))
        except ValueError:
            continue
    out = {}
    for (_, name), m in per.items():  # later launches overwrite earlier ones: the last fused-kernel / SpMV launch counts
        if "dram__bytes_read.sum" not in m or "dram__bytes_write.sum" not in m:
            continue
        tot = int(m["dram__bytes_read.sum"] + m["dram__bytes_write.sum"])
        if "spmv_warp_row" in name:
            out["spmv"] = tot
        elif "jacobian_kernel" in name and (", 1>" in name or "hdiv_v7" not in name):  # MODE 1 = fused residual + Jacobian
            out["jacobian"] = tot
    return out or None


_JSON_FD = None


def _claim_stdout():
    """stdout carries the single JSON line and nothing else: everything libraries print to fd 1 from here on (NCCL banners of
    the library's own dlopen'ed copy, torchrun notes, ...) is sent to stderr; emit_json() writes to the saved descriptor."""
    global _JSON_FD
    if _JSON_FD is None:
        sys.stdout.flush()
        _JSON_FD = os.dup(1)
        os.dup2(2, 1)


def emit_json(line):
    data = (json.dumps(line) + "\n").encode()
    if _JSON_FD is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_JSON_FD, data)


def main():
    _claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--ref-cells", type=int, default=1536, help="cells per step of the bounded CPU sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-parity", action="store_true", help="skip the post-timing parity check against the oracle")
    ap.add_argument("--no-extra", action="store_true", help="skip the extra legs (solve, expansion6k, H1-H1, patch smoother)")
    ap.add_argument("--nc-global", type=int, nargs=2, default=None, metavar=("NX", "NY"),
                    help="fixed GLOBAL mesh (NX,NY,3) instead of 64x64x3 cells per GPU: strong scaling over --gpus, or a large single-GPU point")
    args = ap.parse_args()
    global NC_GLOBAL
    NC_GLOBAL = args.nc_global
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
