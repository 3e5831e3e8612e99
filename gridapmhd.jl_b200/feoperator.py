"""Host-side mirror of the reference's operator / solver seams for the hot path, on top of the C ABI.

Reference interfaces mirrored (GridapMHD.jl / Gridap API as used by it):
  * `_fe_operator(U,V,params)` -> `FEOperator(res,jac,U,V,assem)`            src/main.jl:207-233
  * `residual(op,xh)`, `jacobian(op,xh)`                                       src/main.jl:158,163
  * `residual!(b,op,x)`, `jacobian!(A,op,x)`, `allocate_jacobian`              Gridap NonlinearOperator API used by
                                                                               `solve!(xh,solver,op)`, src/main.jl:275
  * `symbolic_setup / numerical_setup / numerical_setup! / solve!`             Gridap LinearSolver API behind
                                                                               `_solver`, `get_block_solver`
                                                                               (src/main.jl:181-190, Solvers/gridap.jl:2-3)
  * `GridapSolvers.NewtonSolver(ls;maxiter,atol,rtol)`                         src/main.jl:183, Solvers/badia2024.jl:46

Julia's `f!` is spelled `f_b` ("bang").  All arithmetic happens in libmhdb200.so on the GPU; this module only
marshals arrays (numpy or torch) across the C ABI.
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass

import numpy as np

from . import lib as L
from .host.fespaces import FIELDS, FESpaces


def _as_f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _is_torch(a):
    return hasattr(a, "data_ptr")


@dataclass
class FluidParams:
    """`params[:fluid]` after `params_fluid` (src/parameters.jl:693-723): alpha=α, beta=β, gamma=γ, sigma=σ,
    zeta_u=ζᵤ, zeta_j=ζⱼ; convection in {"none","picard","newton"} (default :newton, parameters.jl:717)."""

    alpha: float = 1.0
    beta: float = 1.0
    gamma: float = 1.0
    sigma: float = 1.0
    zeta_u: float = 0.0
    zeta_j: float = 0.0
    B: tuple = (0.0, 1.0, 0.0)
    f: tuple = (0.0, 0.0, 0.0)
    g: tuple = (0.0, 0.0, 0.0)
    convection: str = "newton"

    def to_c(self) -> L.mhd_params_t:
        p = L.mhd_params_t()
        p.alpha, p.beta, p.gamma, p.sigma = self.alpha, self.beta, self.gamma, self.sigma
        p.zeta_u, p.zeta_j = self.zeta_u, self.zeta_j
        for i in range(3):
            p.B[i], p.f[i], p.g[i] = float(self.B[i]), float(self.f[i]), float(self.g[i])
        p.convection = L.CONVECTION[self.convection]
        return p


class B200Matrix:
    """Device-resident CSR Jacobian (the `SparseMatrixCSR{0,Float64,Int}` the assembler would return).
    The pattern is fetched to the host lazily; values stay on the GPU unless `nzval()`/`to_scipy()` is called."""

    def __init__(self, op: "B200FEOperator"):
        self.op = op
        self._pattern = None

    @property
    def shape(self):
        return (self.op.nrows, self.op.ncols)

    @property
    def nnz(self):
        return self.op.nnz

    def pattern(self, index_bytes=8, base=0):
        key = (index_bytes, base)
        if self._pattern is None or self._pattern[0] != key:
            dt = np.int64 if index_bytes == 8 else np.int32
            rowptr = np.empty(self.op.nrows + 1, dtype=dt)
            colval = np.empty(self.op.nnz, dtype=dt)
            L.check(L.load().mhd_operator_get_csr(self.op.handle, L.ptr(rowptr), L.ptr(colval), index_bytes, base))
            self._pattern = (key, rowptr, colval)
        return self._pattern[1], self._pattern[2]

    def nzval(self) -> np.ndarray:
        v = np.empty(self.op.nnz)
        L.check(L.load().mhd_get_nzval(self.op.handle, L.ptr(v)))
        return v

    def to_scipy(self):
        import scipy.sparse as sp

        rowptr, colval = self.pattern()
        return sp.csr_matrix((self.nzval(), colval, rowptr), shape=self.shape)

    def mul(self, x, y=None):
        """mul!(y,A,x)"""
        return self.op.spmv(x, y)


class B200FEOperator:
    """`FEOperator` whose residual/Jacobian are integrated and assembled on the B200 (C ABI: mhd_operator_*)."""

    def __init__(self, fes: FESpaces, fluid: FluidParams, nowned: dict | None = None):
        self.fes = fes
        self.fluid = fluid
        lib = L.load()
        m = fes.mesh
        T = fes.tables
        self._keep = []  # numpy arrays borrowed by the C structs during the create call

        def hold(a, dt):
            a = np.ascontiguousarray(a, dtype=dt)
            self._keep.append(a)
            return a.ctypes.data

        has_solid = fes.cell_solid is not None and bool(np.any(fes.cell_solid))
        mesh = L.mhd_mesh_t(m.coords.shape[0], hold(m.coords, np.float64), m.ncells, hold(m.cell_nodes, np.int32), 0,
                            hold(fes.cell_solid, np.uint8) if has_solid else None,
                            hold(fes.cell_sigma, np.float64) if has_solid else None)
        tab = L.mhd_tables_t(T.nq, hold(T.w, np.float64), hold(T.geo_grad, np.float64), hold(T.nu, np.float64),
                             hold(T.dnu, np.float64), hold(T.pp, np.float64), hold(T.psi, np.float64),
                             hold(T.dpsi, np.float64), hold(T.chi, np.float64))
        lay = L.mhd_layout_t()
        for f in FIELDS:
            i = L.FIELD_IDS[f]
            lay.cell_dofs[i] = hold(fes.cell_dofs[f], np.int32)
            lay.nfree[i] = fes.nfree[f]
            lay.nowned[i] = fes.nfree[f] if nowned is None else nowned[f]
            lay.ndir[i] = fes.ndir[f]
            dv = fes.dirichlet_values[f]
            lay.dir_values[i] = hold(dv, np.float64) if len(dv) else None
        lay.j_sign = hold(fes.j_sign, np.int8)
        for k, f in enumerate(fes.field_order):
            lay.field_order[k] = L.FIELD_IDS[f]
        prm = fluid.to_c()
        h = C.c_void_p()
        L.check(lib.mhd_operator_create(C.byref(mesh), C.byref(tab), C.byref(lay), C.byref(prm), C.byref(h)))
        self.handle = h
        self._keep = []
        self.nrows = self.ncols = self.nnz = None
        self._A = None

    # -- lifetime ------------------------------------------------------------------------------
    def destroy(self):
        if getattr(self, "handle", None):
            L.load().mhd_operator_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.destroy()
        except Exception:
            pass

    @property
    def kernel_version(self) -> int:
        """7: sum-factorised kernel (tables recognised as tensor products), 5: generic tensor-core kernel."""
        v = C.c_int32()
        L.check(L.load().mhd_operator_get_kernel_version(self.handle, C.byref(v)))
        return v.value

    def set_deterministic(self, on: bool = True) -> int:
        """Coloured assembly: bit-identical results from run to run (returns the number of colours)."""
        n = C.c_int32()
        L.check(L.load().mhd_operator_set_deterministic(self.handle, 1 if on else 0, C.byref(n)))
        return n.value

    def set_fluid(self, fluid: FluidParams):
        self.fluid = fluid
        prm = fluid.to_c()
        L.check(L.load().mhd_operator_set_params(self.handle, C.byref(prm)))

    # -- symbolic ------------------------------------------------------------------------------
    def allocate_jacobian(self) -> B200Matrix:
        if self._A is None:
            nr, ncol, nnz = C.c_int64(), C.c_int64(), C.c_int64()
            L.check(L.load().mhd_operator_symbolic(self.handle, C.byref(nr), C.byref(ncol), C.byref(nnz)))
            self.nrows, self.ncols, self.nnz = nr.value, ncol.value, nnz.value
            self._A = B200Matrix(self)
        return self._A

    def scatter_stats(self):
        self.allocate_jacobian()
        a, b = C.c_int64(), C.c_int64()
        L.check(L.load().mhd_operator_get_scatter_stats(self.handle, C.byref(a), C.byref(b)))
        return a.value, b.value

    # -- numeric -------------------------------------------------------------------------------
    def jacobian_b(self, A: B200Matrix, x):
        """jacobian!(A,op,x): x is a numpy vector (host buffers, copies inside) or a CUDA torch tensor."""
        assert A is self._A
        xx = x if _is_torch(x) else _as_f64(x)
        L.check(L.load().mhd_jacobian(self.handle, L.ptr(xx), None))
        return A

    def jacobian(self, x) -> B200Matrix:
        return self.jacobian_b(self.allocate_jacobian(), x)

    def residual_and_jacobian_b(self, b, A: B200Matrix, x):
        """residual_and_jacobian!(b,A,op,x): fused kernel (the per-cell geometry/basis preparation is shared)."""
        assert A is self._A
        xx = x if _is_torch(x) else _as_f64(x)
        L.check(L.load().mhd_residual_and_jacobian(self.handle, L.ptr(xx), L.ptr(b)))
        return b, A

    def residual_b(self, b, x):
        """residual!(b,op,x)"""
        xx = x if _is_torch(x) else _as_f64(x)
        L.check(L.load().mhd_residual(self.handle, L.ptr(xx), L.ptr(b)))
        return b

    def residual(self, x):
        self.allocate_jacobian()
        if _is_torch(x):
            import torch

            b = torch.empty(self.nrows, dtype=torch.float64, device=x.device)
        else:
            b = np.empty(self.nrows)
        return self.residual_b(b, x)

    # -- post-processing (hunt.jl:239-260) ----------------------------------------------------------
    def hunt_error_norms(self, x, T6, Ha, nsums, u0=1.0, jscale=1.0, a=1.0, mu=1.0, sigma=1.0, grad_pz=-1.0):
        """eu_l2, eu_h1, ej_l2 against the analytical Hunt series and uh_l2, uh_h1, jh_l2, integrated on the device with
        the degree 2*(order+1) tables `T6` (`reffe.make_tables(6)`)."""
        keep = [np.ascontiguousarray(v, dtype=np.float64) for v in (T6.w, T6.geo_grad, T6.nu, T6.dnu, T6.pp, T6.psi, T6.dpsi, T6.chi)]
        tab = L.mhd_tables_t(T6.nq, *[v.ctypes.data for v in keep])
        prm = L.mhd_hunt_post_t(a, mu, sigma, grad_pz, Ha, int(nsums), 0, u0, jscale)
        out = (C.c_double * 6)()
        xx = x if _is_torch(x) else _as_f64(x)
        L.check(L.load().mhd_hunt_error_norms(self.handle, L.ptr(xx), C.byref(tab), C.byref(prm), out))
        return dict(zip(("eu_l2", "eu_h1", "ej_l2", "uh_l2", "uh_h1", "jh_l2"), list(out)))

    # -- Krylov building blocks ------------------------------------------------------------------
    def spmv(self, x, y=None):
        self.allocate_jacobian()
        if y is None:
            if _is_torch(x):
                import torch

                y = torch.empty(self.nrows, dtype=torch.float64, device=x.device)
            else:
                y = np.empty(self.nrows)
        xx = x if _is_torch(x) else _as_f64(x)
        L.check(L.load().mhd_spmv(self.handle, L.ptr(xx), L.ptr(y)))
        return y

    def dot(self, x, y) -> float:
        out = C.c_double()
        xx = x if _is_torch(x) else _as_f64(x)
        yy = y if _is_torch(y) else _as_f64(y)
        L.check(L.load().mhd_dot(self.handle, L.ptr(xx), L.ptr(yy), C.addressof(out)))
        return out.value

    def axpy(self, a: float, x, y):
        xx = x if _is_torch(x) else _as_f64(x)
        L.check(L.load().mhd_axpy(self.handle, float(a), L.ptr(xx), L.ptr(y)))
        return y

    def multi_dot_axpy(self, V, w):
        """h = V w ; w -= V^T h  (fused Gram-Schmidt step). V: [k, n]"""
        k, ldv = V.shape
        h = np.empty(k)
        L.check(L.load().mhd_multi_dot_axpy(self.handle, k, L.ptr(V), ldv, L.ptr(w), L.ptr(h)))
        return h


class B200H1H1FEOperator(B200FEOperator):
    """`FEOperator` of the H1-H1 formulation (`weak_form_h1_h1`, src/weakforms.jl:344-355; spaces (u,p,phi) of
    src/fespaces.jl:32-41): same handle type and methods as `B200FEOperator` (jacobian, residual, spmv, dot, ...), created
    through `mhd_h1h1_operator_create`.  `fes` is an `H1H1Spaces` (host/fespaces_h1h1.py)."""

    H1H1_FIELD_IDS = {"u": 0, "p": 1, "phi": 2}

    def __init__(self, fes, fluid: FluidParams, nowned: dict | None = None):
        self.fes = fes
        self.fluid = fluid
        lib = L.load()
        m = fes.mesh
        T = fes.tables
        keep = []

        def hold(a, dt):
            a = np.ascontiguousarray(a, dtype=dt)
            keep.append(a)
            return a.ctypes.data

        has_solid = fes.cell_solid is not None and bool(np.any(fes.cell_solid))
        mesh = L.mhd_mesh_t(m.coords.shape[0], hold(m.coords, np.float64), m.ncells, hold(m.cell_nodes, np.int32), 0,
                            hold(fes.cell_solid, np.uint8) if has_solid else None,
                            hold(np.ones(m.ncells), np.float64) if has_solid else None)
        tab = L.mhd_tables_h1h1_t(T.nq, hold(T.w, np.float64), hold(T.geo_grad, np.float64), hold(T.nu, np.float64),
                                  hold(T.dnu, np.float64), hold(T.pp, np.float64), hold(T.dphi3, np.float64))
        lay = L.mhd_layout_h1h1_t()
        for f, i in self.H1H1_FIELD_IDS.items():
            lay.cell_dofs[i] = hold(fes.cell_dofs[f], np.int32)
            lay.nfree[i] = fes.nfree[f]
            lay.nowned[i] = fes.nfree[f] if nowned is None else nowned[f]
            lay.ndir[i] = fes.ndir[f]
            dv = fes.dirichlet_values[f]
            lay.dir_values[i] = hold(dv, np.float64) if len(dv) else None
        for k, f in enumerate(fes.field_order):
            lay.field_order[k] = self.H1H1_FIELD_IDS[f]
        prm = fluid.to_c()
        h = C.c_void_p()
        L.check(lib.mhd_h1h1_operator_create(C.byref(mesh), C.byref(tab), C.byref(lay), C.byref(prm), C.byref(h)))
        self.handle = h
        self.nrows = self.ncols = self.nnz = None
        self._A = None


# ----------------------------------------------------------------------------------------------
# linear solver seam


@dataclass
class B200SolverOptions:
    """`default_solver_params(Val(:badia2024))` values relevant to the linear solve (src/parameters.jl:259-271)."""

    m: int = 15  # niter_ls
    maxiter: int = 15
    rtol: float = 1e-7  # nl_rtol/10 (badia2024.jl:37)
    atol: float = 1e-8
    precond: str = "block_tri"  # "none" | "jacobi" | "block_tri" (H1-HDiv, badia2024.jl) | "h1h1_blocks" (H1-H1, h1h1blocks.jl)
    uj_inner_its: int = 30
    uj_inner_restart: int = 30
    # (u,j) block: "gmres_jacobi" | "dense_lu" (exact, small problems) | "gmres_patch" (vertex-patch block-Jacobi smoother of
    # gmg_block_jacobi_smoothers, src/Solvers/gmg.jl:62-81, inside the inner GMRES; any size, one GPU)
    uj_solver: str = "gmres_jacobi"
    patch_its: int = 1  # Richardson sweeps per application (the reference smoother: niter = 10)
    patch_omega: float = 1.0  # damping (the reference smoother: w = 0.2)


class B200LinearSolver:
    """Gridap `LinearSolver` backed by the device FGMRES (`FGMRESSolver(m,P;...)`, badia2024.jl:40)."""

    def __init__(self, opts: B200SolverOptions | None = None):
        self.opts = opts or B200SolverOptions()

    def symbolic_setup(self, A: B200Matrix):
        return B200SymbolicSetup(self, A)


class B200SymbolicSetup:
    def __init__(self, solver, A):
        self.solver, self.A = solver, A

    def numerical_setup(self, A: B200Matrix | None = None):
        A = A or self.A
        return B200NumericalSetup(self.solver, A)


class B200NumericalSetup:
    def __init__(self, solver: B200LinearSolver, A: B200Matrix):
        self.solver, self.A = solver, A
        o = solver.opts
        fl = A.op.fluid
        c = L.mhd_solver_opts_t()
        L.check(L.load().mhd_solver_default_opts(C.byref(c)))
        c.m, c.maxiter, c.rtol, c.atol = o.m, o.maxiter, o.rtol, o.atol
        c.precond = L.PRECOND[o.precond]
        c.uj_inner_its, c.uj_inner_restart = o.uj_inner_its, o.uj_inner_restart
        c.uj_solver = L.UJ_SOLVER[o.uj_solver]
        c.patch_its, c.patch_omega = o.patch_its, o.patch_omega
        c.alpha_p = -1.0 / (fl.beta + fl.zeta_u)  # badia2024.jl:11
        c.alpha_phi = -1.0 / (1.0 + fl.zeta_j)  # badia2024.jl:12
        h = C.c_void_p()
        L.check(L.load().mhd_solver_create(A.op.handle, C.byref(c), C.byref(h)))
        self.handle = h
        if o.precond == "h1h1_blocks" and o.uj_solver == "gmres_patch":
            from .host.patches import vertex_patches_h1h1

            (pu, du), (pf, df) = vertex_patches_h1h1(A.op.fes)
            keep = [np.ascontiguousarray(pu, dtype=np.int64), np.ascontiguousarray(du, dtype=np.int32),
                    np.ascontiguousarray(pf, dtype=np.int64), np.ascontiguousarray(df, dtype=np.int32)]
            L.check(L.load().mhd_solver_set_patches(h, len(pu) - 1, L.ptr(keep[0]), L.ptr(keep[1])))
            L.check(L.load().mhd_solver_set_phi_patches(h, len(pf) - 1, L.ptr(keep[2]), L.ptr(keep[3])))
            self.npatches = len(pu) - 1
            self.patch_entries = int((np.diff(pu) ** 2).sum() + (np.diff(pf) ** 2).sum())
        if o.precond == "block_tri" and o.uj_solver == "gmres_patch":
            # PatchTopology(ReferenceFE{0}, model) of the host (gmg.jl:69): vertex-star dof lists of the (u,j) block
            from .host.patches import vertex_patches

            ptr, dofs = vertex_patches(A.op.fes)
            ptr = np.ascontiguousarray(ptr, dtype=np.int64)
            dofs = np.ascontiguousarray(dofs, dtype=np.int32)
            L.check(L.load().mhd_solver_set_patches(h, len(ptr) - 1, L.ptr(ptr), L.ptr(dofs)))
            self.npatches, self.patch_entries = len(ptr) - 1, int((np.diff(ptr) ** 2).sum())
        self.iters = 0
        self.resnorm = float("nan")
        self.history = np.zeros(0)
        self.numerical_setup_b(A)

    def numerical_setup_b(self, A: B200Matrix | None = None):
        """numerical_setup!(ns,A): refresh preconditioner data after jacobian!"""
        L.check(L.load().mhd_solver_setup(self.handle))
        return self

    def patch_apply(self, r, omega=1.0):
        """solve!(z, BlockJacobiSolver, r): one additive sweep of the vertex-patch solver on the (u,j) block"""
        rr = r if _is_torch(r) else _as_f64(r)
        if _is_torch(r):
            import torch

            z = torch.empty_like(r)
        else:
            z = np.empty(len(rr))
        L.check(L.load().mhd_solver_patch_apply(self.handle, L.ptr(rr), L.ptr(z), float(omega)))
        return z

    def solve_b(self, x, b, raise_on_maxiter=False):
        """solve!(x,ns,b): x holds the initial guess on entry and the solution on exit."""
        it, rn = C.c_int32(), C.c_double()
        hist = np.zeros(4 * 64 + 2)
        bb = b if _is_torch(b) else _as_f64(b)
        rc = L.load().mhd_solve(self.handle, L.ptr(bb), L.ptr(x), C.byref(it), C.byref(rn), L.ptr(hist))
        self.iters, self.resnorm = it.value, rn.value
        self.history = hist[: it.value + 1].copy()
        if rc == -6 and not raise_on_maxiter:
            return x
        L.check(rc)
        return x

    def destroy(self):
        if getattr(self, "handle", None):
            L.load().mhd_solver_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.destroy()
        except Exception:
            pass


class NewtonSolver:
    """`GridapSolvers.NewtonSolver(ls;maxiter,atol,rtol)` as GridapMHD builds it (main.jl:183, badia2024.jl:46):
    stop when |r| <= max(atol, rtol*|r0|).  The nonlinear loop is host logic in the reference too."""

    def __init__(self, ls: B200LinearSolver, maxiter=10, rtol=1e-6, atol=0.0, verbose=False):
        self.ls, self.maxiter, self.rtol, self.atol, self.verbose = ls, maxiter, rtol, atol, verbose
        self.log = []

    def solve_b(self, x: np.ndarray, op: B200FEOperator):
        A = op.allocate_jacobian()
        b = np.empty(op.nrows)
        op.residual_and_jacobian_b(b, A, x)
        r0 = float(np.linalg.norm(b))
        self.log = [r0]
        ns = None
        for it in range(self.maxiter):
            if r0 == 0.0:
                break
            ns = self.ls.symbolic_setup(A).numerical_setup() if ns is None else ns.numerical_setup_b(A)
            dx = np.zeros(op.nrows)
            ns.solve_b(dx, -b)
            x += dx
            op.residual_and_jacobian_b(b, A, x)
            rn = float(np.linalg.norm(b))
            self.log.append(rn)
            if self.verbose:
                print(f"  newton {it+1}: |r|={rn:.3e} (rel {rn/r0:.3e}), linear its {ns.iters}, lin res {ns.resnorm:.3e}")
            if rn <= max(self.atol, self.rtol * r0):
                break
        if ns is not None:
            ns.destroy()
        return x
