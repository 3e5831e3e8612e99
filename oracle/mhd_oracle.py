"""CPU ORACLE (test infrastructure only -- never imported by the product path).

NumPy restatement of the reference's hot path: cell-wise integration of the H1-HDiv inductionless
MHD residual/Jacobian and assembly into CSR, plus the Krylov building blocks.

What it follows in /root/reference (file:line):
  * integrands and signs .......... src/weakforms.jl:255-281 (`res_fluid_h1_hdiv`),
                                     src/weakforms.jl:283-312 (`jac_fluid_h1_hdiv`),
                                     src/weakforms.jl:670 (`conv`), :672-681 (`local_projection_operator`)
  * derived fields ................ src/weakforms.jl:36-43 (`setup_variable`)
  * measure / quadrature .......... src/geometry.jl:116-128, src/parameters.jl:381-389,617-639
  * assembly semantics ............ Gridap `SparseMatrixAssembler` as used at src/main.jl:222-223:
                                     every touched (row,col) of the 8 touched blocks is inserted
                                     (explicit zeros kept), Dirichlet rows/cols dropped, indices sorted.
  * Krylov ........................ GridapSolvers FGMRES as configured at src/Solvers/badia2024.jl:36-40
                                     (right-preconditioned flexible GMRES, modified Gram-Schmidt, Givens).

The arithmetic of this path lives in un-vendored, un-pinned Julia dependencies (Gridap 0.19/0.20,
GridapSolvers 0.6/0.7, PartitionedArrays; `Project.toml:29-46`, no root Manifest) and Julia is not
available in the build container, so the reference itself cannot be executed here.

PARITY PINNING: the reference's tests hold no matrix/residual-level golden values ("parity unpinned" at
the entry level).  The oracle is pinned at SOLUTION level against the reference's published runs
(`analysis/gadi/results/2023_04/**/summary.csv`): exact DOF counts and 16-digit discrete-solution norms
(uh_l2, uh_h1, jh_l2) of the Hunt benchmark -- see tests/test_oracle_pins.py -- plus known-answer tests
(manufactured in-space fields of src/Applications/transient.jl:262-270, FD-Jacobian, block identities).
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla

NU, NP, NJ, NPHI = 81, 4, 36, 8
OFF_U, OFF_P, OFF_J, OFF_PHI = 0, 81, 85, 121
NLOC = 129


@dataclass
class FluidParams:
    """`retrieve_fluid_params` tuple (src/weakforms.jl:71-83) for constant coefficients."""

    alpha: float = 1.0
    beta: float = 1.0
    gamma: float = 1.0
    sigma: float = 1.0
    zeta_u: float = 0.0
    zeta_j: float = 0.0
    B: tuple = (0.0, 1.0, 0.0)
    f: tuple = (0.0, 0.0, 0.0)
    g: tuple = (0.0, 0.0, 0.0)
    convection: str = "newton"  # "none" | "picard" | "newton"  (src/parameters.jl:717 default :newton)


def touched_mask() -> np.ndarray:
    """129x129 bool: the 8 blocks the weak form touches (uu,up,uj,pu,jj,ju,j-phi,phi-j)."""
    m = np.zeros((NLOC, NLOC), dtype=bool)
    u, p, j, ph = slice(0, 81), slice(81, 85), slice(85, 121), slice(121, 129)
    for r, c in ((u, u), (u, p), (u, j), (p, u), (j, j), (j, u), (j, ph), (ph, j)):
        m[r, c] = True
    return m


# ----------------------------------------------------------------------------
# per-cell geometry and mapped bases


def cell_geometry(T, X):
    """X [nc,8,3] -> J [nc,nq,3,3] (J[i,k]=dx_i/dxi_k), det [nc,nq], invJ [nc,nq,3,3] (invJ[k,i]=dxi_k/dx_i)."""
    J = np.einsum("cvi,qvk->cqik", X, T.geo_grad)
    det = np.linalg.det(J)
    invJ = np.linalg.inv(J)
    return J, det, invJ


def mapped_bases(T, X, j_sign):
    J, det, invJ = cell_geometry(T, X)
    w = T.w[None, :] * np.abs(det)
    gradN = np.einsum("qak,cqki->cqai", T.dnu, invJ)  # physical gradients of the scalar Q2 basis
    s = j_sign.astype(float)
    psi = np.einsum("cqik,qmk->cqmi", J, T.psi) / det[:, :, None, None] * s[:, None, :, None]  # Piola
    dpsi = T.dpsi[None, :, :] / det[:, :, None] * s[:, None, :]
    return w, gradN, psi, dpsi


def _cross(a, B):
    """a[...,3] x B[3]"""
    B = np.asarray(B, dtype=float)
    return np.stack(
        [a[..., 1] * B[2] - a[..., 2] * B[1], a[..., 2] * B[0] - a[..., 0] * B[2], a[..., 0] * B[1] - a[..., 1] * B[0]],
        axis=-1,
    )


def cell_jacobians(T, X, state, j_sign, prm: FluidParams, solid=None, sigma_c=None):
    """Dense cell matrices [nc,129,129] of `jac_fluid_h1_hdiv` (src/weakforms.jl:283-312).
    rows = test functions, cols = trial functions; local order u(a+27c), p, j, phi.
    `solid` [nc] bool marks cells of `jac_solid_h1_hdiv` (:327-338): only jj, j-phi (sigma = sigma_c of the cell) and
    phi-j with the OPPOSITE sign (+phi div j) are present; their u/p dofs are absent and dropped at assembly."""
    nc = X.shape[0]
    w, gN, psi, dpsi = mapped_bases(T, X, j_sign)
    N, Pp, Chi = T.nu, T.pp, T.chi
    K = np.zeros((nc, NLOC, NLOC))
    # --- uu: beta grad(du) : grad(v)
    S = np.einsum("cq,cqai,cqbi->cab", w, gN, gN)
    Kuu = np.zeros((nc, 3, 27, 3, 27))  # [c, comp_row, a, comp_col, b]
    for c in range(3):
        Kuu[:, c, :, c, :] += prm.beta * S
    if prm.convection in ("picard", "newton"):
        ustate = state[:, :81].reshape(nc, 3, 27)
        uq = np.einsum("qa,cia->cqi", N, ustate)  # u at quadrature points
        ugN = np.einsum("cqi,cqbi->cqb", uq, gN)  # u . grad N_b
        C = np.einsum("cq,qa,cqb->cab", w, N, ugN)
        for c in range(3):
            Kuu[:, c, :, c, :] += prm.alpha * C
        if prm.convection == "newton":
            gu = np.einsum("cqbd,cib->cqdi", gN, ustate)  # gu[d,i] = d_d u_i
            Kuu += prm.alpha * np.einsum("cq,qa,qb,cqdi->ciadb", w, N, N, gu)
    # --- div tables for u: div(N_a e_c) = d_c N_a
    if prm.zeta_u != 0.0:
        D = np.einsum("cq,qk,cqai->ckia", w, Pp, gN).reshape(nc, 4, 81)  # D[k,(c,a)]
        Mp = np.einsum("cq,qk,ql->ckl", w, Pp, Pp)
        E = np.linalg.solve(Mp, D)
        Kuu += prm.zeta_u * np.einsum("cki,ckj->cij", D, E).reshape(nc, 3, 27, 3, 27)
    K[:, :81, :81] = Kuu.reshape(nc, 81, 81)
    # --- up / pu
    Kup = -np.einsum("cq,qk,cqai->ciak", w, Pp, gN).reshape(nc, 81, 4)
    K[:, :81, 81:85] = Kup
    K[:, 81:85, :81] = np.transpose(Kup, (0, 2, 1))
    # --- uj / ju
    pxB = _cross(psi, prm.B)  # [c,q,m,3]
    Kuj = -prm.gamma * np.einsum("cq,qa,cqmi->ciam", w, N, pxB).reshape(nc, 81, 36)
    Kju = prm.sigma * np.einsum("cq,qb,cqmd->cmdb", w, N, pxB).reshape(nc, 36, 81)
    K[:, :81, 85:121] = Kuj
    K[:, 85:121, :81] = Kju
    # --- jj
    Kjj = np.einsum("cq,cqmi,cqni->cmn", w, psi, psi)
    if prm.zeta_j != 0.0:
        Kjj += prm.zeta_j * np.einsum("cq,cqm,cqn->cmn", w, dpsi, dpsi)
    K[:, 85:121, 85:121] = Kjj
    # --- j-phi / phi-j
    sig = np.full(nc, prm.sigma) if solid is None else np.where(solid, sigma_c, prm.sigma)
    sgn = np.ones(nc) if solid is None else np.where(solid, -1.0, 1.0)
    JF = np.einsum("cq,ql,cqm->cml", w, Chi, dpsi)
    K[:, 85:121, 121:129] = -sig[:, None, None] * JF
    K[:, 121:129, 85:121] = -sgn[:, None, None] * np.transpose(JF, (0, 2, 1))
    if solid is not None:
        K[solid, :85, :] = 0.0
        K[solid, :, :85] = 0.0
    return K


def cell_residuals(T, X, state, j_sign, prm: FluidParams, solid=None, sigma_c=None):
    """Cell vectors [nc,129] of `res_fluid_h1_hdiv` (src/weakforms.jl:255-281); `solid` cells follow
    `res_solid_h1_hdiv` (:314-325): j.s + zeta div j div s - sigma_c phi div s + phi-test * div j - s.g."""
    nc = X.shape[0]
    w, gN, psi, dpsi = mapped_bases(T, X, j_sign)
    N, Pp, Chi = T.nu, T.pp, T.chi
    us = state[:, :81].reshape(nc, 3, 27)
    ps = state[:, 81:85]
    js = state[:, 85:121]
    fs = state[:, 121:129]
    uq = np.einsum("qa,cia->cqi", N, us)
    gu = np.einsum("cqbd,cib->cqdi", gN, us)  # gu[d,i] = d_d u_i
    divu = np.einsum("cqii->cq", gu)
    pq = np.einsum("qk,ck->cq", Pp, ps)
    jq = np.einsum("cqmi,cm->cqi", psi, js)
    divj = np.einsum("cqm,cm->cq", dpsi, js)
    fq = np.einsum("ql,cl->cq", Chi, fs)
    B = np.asarray(prm.B, dtype=float)
    # body forces: constants (what the C ABI carries) or, in the oracle only, functions of the physical point
    # (TimeSpaceFunction forcing of src/Applications/transient.jl:331-344) evaluated at the quadrature points
    xq = np.einsum("qv,cvi->cqi", T.geo_val, X) if (callable(prm.f) or callable(prm.g)) else None
    f = prm.f(xq) if callable(prm.f) else np.broadcast_to(np.asarray(prm.f, dtype=float), (nc, len(T.w), 3))
    g = prm.g(xq) if callable(prm.g) else np.broadcast_to(np.asarray(prm.g, dtype=float), (nc, len(T.w), 3))
    R = np.zeros((nc, NLOC))
    # u rows
    ru = prm.beta * np.einsum("cq,cqdi,cqad->cia", w, gu, gN)
    if prm.convection != "none":
        conv = np.einsum("cqd,cqdi->cqi", uq, gu)
        ru += prm.alpha * np.einsum("cq,qa,cqi->cia", w, N, conv)
    if prm.zeta_u != 0.0:
        Mp = np.einsum("cq,qk,ql->ckl", w, Pp, Pp)
        rhs = np.einsum("cq,qk,cq->ck", w, Pp, divu)
        coef = np.linalg.solve(Mp, rhs[..., None])[..., 0]
        proj = np.einsum("qk,ck->cq", Pp, coef)
        ru += prm.zeta_u * np.einsum("cq,cq,cqai->cia", w, proj, gN)
    ru -= np.einsum("cq,cq,cqai->cia", w, pq, gN)
    jxB = _cross(jq, B)
    ru -= prm.gamma * np.einsum("cq,qa,cqi->cia", w, N, jxB)
    ru -= np.einsum("cq,qa,cqi->cia", w, N, f)
    R[:, :81] = ru.reshape(nc, 81)
    # p rows
    R[:, 81:85] = -np.einsum("cq,qk,cq->ck", w, Pp, divu)
    # j rows
    uxB = _cross(uq, B)
    rj = np.einsum("cq,cqi,cqmi->cm", w, jq, psi)
    if prm.zeta_j != 0.0:
        rj += prm.zeta_j * np.einsum("cq,cq,cqm->cm", w, divj, dpsi)
    sig = np.full(nc, prm.sigma) if solid is None else np.where(solid, sigma_c, prm.sigma)
    sgn = np.ones(nc) if solid is None else np.where(solid, -1.0, 1.0)
    rj -= sig[:, None] * np.einsum("cq,cq,cqm->cm", w, fq, dpsi)
    rj -= prm.sigma * np.einsum("cq,cqi,cqmi->cm", w, uxB, psi)
    rj -= np.einsum("cq,cqi,cqmi->cm", w, g, psi)
    R[:, 85:121] = rj
    # phi rows
    R[:, 121:129] = -sgn[:, None] * np.einsum("cq,ql,cq->cl", w, Chi, divj)
    if solid is not None:
        R[solid, :85] = 0.0
    return R


# ----------------------------------------------------------------------------
# assembly (Gridap SparseMatrixAssembler semantics)


def symbolic_csr(gids: np.ndarray, n: int):
    """CSR pattern from [nc,129] 0-based global ids (-1 = Dirichlet): every touched pair inserted.
    Returns rowptr[int64 n+1], colval[int64 nnz] with sorted columns."""
    mask = touched_mask()
    li, lj = np.nonzero(mask)
    P = None
    chunk = 8192
    for s in range(0, gids.shape[0], chunk):
        g = gids[s : s + chunk]
        r = g[:, li]
        c = g[:, lj]
        ok = (r >= 0) & (c >= 0)
        Pi = sp.coo_matrix((np.ones(int(ok.sum()), dtype=np.int8), (r[ok], c[ok])), shape=(n, n)).tocsr()
        Pi.sum_duplicates()
        Pi.data[:] = 1
        P = Pi if P is None else P + Pi
    P.sum_duplicates()
    P.sort_indices()
    return P.indptr.astype(np.int64), P.indices.astype(np.int64)


def assemble_matrix(K: np.ndarray, gids: np.ndarray, n: int, pattern=None) -> sp.csr_matrix:
    """Scatter-add dense cell matrices into CSR on the symbolic pattern (explicit zeros kept)."""
    if pattern is None:
        pattern = symbolic_csr(gids, n)
    rowptr, colval = pattern
    keyP = np.repeat(np.arange(n, dtype=np.int64), np.diff(rowptr)) * n + colval
    mask = touched_mask()
    li, lj = np.nonzero(mask)
    data = np.zeros(len(colval))
    chunk = 2048
    for s in range(0, gids.shape[0], chunk):
        g = gids[s : s + chunk]
        r = g[:, li]
        c = g[:, lj]
        v = K[s : s + chunk][:, li, lj]
        ok = (r >= 0) & (c >= 0)
        pos = np.searchsorted(keyP, r[ok] * n + c[ok])
        np.add.at(data, pos, v[ok])
    return sp.csr_matrix((data, colval.copy(), rowptr.copy()), shape=(n, n))


def assemble_vector(R: np.ndarray, gids: np.ndarray, n: int) -> np.ndarray:
    ok = gids >= 0
    out = np.zeros(n)
    np.add.at(out, gids[ok], R[ok])
    return out


def jacobian(fes, x, prm: FluidParams, chunk: int = 2048, pattern=None) -> sp.csr_matrix:
    """`jacobian(op,xh)` (src/main.jl:163): assembled Jacobian in CSR (0-based, sorted, explicit zeros kept)."""
    X = fes.mesh.cell_coords()
    st = fes.cell_state(x)
    gids = fes.cell_global_ids()
    n = fes.ndofs
    if pattern is None:
        pattern = symbolic_csr(gids, n)
    rowptr, colval = pattern
    data = np.zeros(len(colval))
    for s in range(0, X.shape[0], chunk):
        sl = slice(s, s + chunk)
        solid = None if fes.cell_solid is None else fes.cell_solid[sl]
        sigc = None if fes.cell_sigma is None else fes.cell_sigma[sl]
        K = cell_jacobians(fes.tables, X[sl], st[sl], fes.j_sign[sl], prm, solid, sigc)
        data += assemble_matrix(K, gids[sl], n, pattern).data
    return sp.csr_matrix((data, colval.copy(), rowptr.copy()), shape=(n, n))


def residual(fes, x, prm: FluidParams, chunk: int = 4096) -> np.ndarray:
    """`residual(op,xh)` (src/main.jl:158)."""
    X = fes.mesh.cell_coords()
    st = fes.cell_state(x)
    gids = fes.cell_global_ids()
    n = fes.ndofs
    out = np.zeros(n)
    for s in range(0, X.shape[0], chunk):
        sl = slice(s, s + chunk)
        solid = None if fes.cell_solid is None else fes.cell_solid[sl]
        sigc = None if fes.cell_sigma is None else fes.cell_sigma[sl]
        R = cell_residuals(fes.tables, X[sl], st[sl], fes.j_sign[sl], prm, solid, sigc)
        out += assemble_vector(R, gids[sl], n)
    return out


# ----------------------------------------------------------------------------
# Krylov building blocks and solvers


def spmv(rowptr, colval, nzval, x):
    n = len(rowptr) - 1
    return sp.csr_matrix((nzval, colval, rowptr), shape=(n, len(x))) @ x


def fgmres(A, b, x0=None, M=None, m=15, maxiter=15, rtol=1e-7, atol=1e-8):
    """Right-preconditioned flexible GMRES(m), modified Gram-Schmidt + Givens rotations
    (GridapSolvers `FGMRESSolver(m,P;maxiter,rtol,atol)` as built at src/Solvers/badia2024.jl:36-40).
    M(v) applies the preconditioner. Returns (x, iters, residual history)."""
    n = len(b)
    x = np.zeros(n) if x0 is None else x0.copy()
    M = M or (lambda v: v)
    r = b - A @ x
    beta = np.linalg.norm(r)
    hist = [beta]
    tol = max(atol, rtol * beta)
    it = 0
    while beta > tol and it < maxiter:
        V = np.zeros((m + 1, n))
        Z = np.zeros((m, n))
        H = np.zeros((m + 1, m))
        cs, sn = np.zeros(m), np.zeros(m)
        g = np.zeros(m + 1)
        V[0] = r / beta
        g[0] = beta
        j = 0
        while j < m and beta > tol and it < maxiter:
            Z[j] = M(V[j])
            wv = A @ Z[j]
            for i in range(j + 1):
                H[i, j] = wv @ V[i]
                wv -= H[i, j] * V[i]
            H[j + 1, j] = np.linalg.norm(wv)
            if H[j + 1, j] > 0:
                V[j + 1] = wv / H[j + 1, j]
            for i in range(j):
                t = cs[i] * H[i, j] + sn[i] * H[i + 1, j]
                H[i + 1, j] = -sn[i] * H[i, j] + cs[i] * H[i + 1, j]
                H[i, j] = t
            d = np.hypot(H[j, j], H[j + 1, j])
            cs[j], sn[j] = H[j, j] / d, H[j + 1, j] / d
            H[j, j] = d
            H[j + 1, j] = 0.0
            g[j + 1] = -sn[j] * g[j]
            g[j] = cs[j] * g[j]
            beta = abs(g[j + 1])
            hist.append(beta)
            j += 1
            it += 1
        y = np.linalg.solve(np.triu(H[:j, :j]), g[:j])
        x += Z[:j].T @ y
        r = b - A @ x
        beta = np.linalg.norm(r)
    return x, it, np.array(hist)


def newton_lu(fes, prm: FluidParams, x0=None, maxiter=10, rtol=1e-6, verbose=False, min_iters=1):
    """`_solver(::Val{:julia})` (src/main.jl:181-186): Newton with sparse LU, rtol 1e-6 on the residual norm."""
    x = np.zeros(fes.ndofs) if x0 is None else x0.copy()
    b = residual(fes, x, prm)
    r0 = np.linalg.norm(b)
    hist = [r0]
    for it in range(maxiter):
        A = jacobian(fes, x, prm)
        dx = spla.splu(A.tocsc()).solve(-b)
        x += dx
        b = residual(fes, x, prm)
        rn = np.linalg.norm(b)
        hist.append(rn)
        if verbose:
            print(f"  newton it {it+1}: |r| = {rn:.3e} (rel {rn/r0:.3e})")
        if (rn <= rtol * r0 or rn < 1e-14) and it + 1 >= min_iters:  # extra iterations = iterative refinement
            break
    return x, hist


# ----------------------------------------------------------------------------
# post-processing norms (src/Applications/hunt.jl:247-260)


def solution_norms(fes, x, T6, u0=1.0, jscale=1.0):
    """uh_l2, uh_h1, jh_l2 with a degree-2*(order+1) quadrature table T6 (hunt.jl:247,258-260)."""
    X = fes.mesh.cell_coords()
    st = fes.cell_state(x)
    nc = X.shape[0]
    w, gN, psi, _ = mapped_bases(T6, X, fes.j_sign)
    us = st[:, :81].reshape(nc, 3, 27) * u0
    js = st[:, 85:121] * jscale
    uq = np.einsum("qa,cia->cqi", T6.nu, us)
    gu = np.einsum("cqbd,cib->cqdi", gN, us)
    jq = np.einsum("cqmi,cm->cqi", psi, js)
    uu = np.einsum("cq,cqi,cqi->", w, uq, uq)
    gg = np.einsum("cq,cqdi,cqdi->", w, gu, gu)
    jj = np.einsum("cq,cqi,cqi->", w, jq, jq)
    return {"uh_l2": np.sqrt(uu), "uh_h1": np.sqrt(gg + uu), "jh_l2": np.sqrt(jj)}


# ----------------------------------------------------------------------------
# analytical Hunt solution and the error norms of the reference's post-processing


def analytical_hunt(xy, a=1.0, b=1.0, mu=1.0, sigma=1.0, grad_pz=-1.0, Ha=50.0, n=500):
    """`analytical_hunt_u` / `analytical_hunt_j` (src/Applications/hunt.jl:372-457) at points xy [npts,2], the
    Fourier series truncated after k = 0..n, plus the derivatives of u_z the H1 error norm needs (the reference gets
    them from ForwardDiff).  Returns u_z, du_z/dx, du_z/dy, j_x, j_y (zero outside the duct, hunt.jl:385-387)."""
    xy = np.asarray(xy, dtype=float)
    ll = b / a
    xi, eta = xy[:, 0] / a, xy[:, 1] / a
    inside = (np.abs(xi) <= 1.0) & (np.abs(eta) <= 1.0)
    k = np.arange(n + 1, dtype=float)[None, :]
    al = (k + 0.5) * np.pi / ll
    N = np.sqrt(Ha**2 + 4.0 * al**2)
    r1, r2 = 0.5 * (Ha + N), 0.5 * (-Ha + N)
    sgn = np.where(np.arange(n + 1) % 2 == 0, 1.0, -1.0)[None, :]
    X, E = xi[:, None], eta[:, None]
    e1m, e1p = np.exp(-r1 * (1 - E)), np.exp(-r1 * (1 + E))
    e2m, e2p = np.exp(-r2 * (1 - E)), np.exp(-r2 * (1 + E))
    d1, d2 = 1.0 + np.exp(-2.0 * r1), 1.0 + np.exp(-2.0 * r2)
    V2, V3 = (r2 / N) * (e1m + e1p) / d1, (r1 / N) * (e2m + e2p) / d2
    V2e, V3e = (r2 / N) * r1 * (e1m - e1p) / d1, (r1 / N) * r2 * (e2m - e2p) / d2
    ck, sk = np.cos(al * X), np.sin(al * X)
    cu = 2.0 * sgn / (ll * al**3)
    scale_u = (a**2 / mu) * (-grad_pz)
    uz = scale_u * np.sum(cu * ck * (1.0 - V2 - V3), axis=1)
    uz_x = scale_u / a * np.sum(cu * (-al * sk) * (1.0 - V2 - V3), axis=1)
    uz_y = scale_u / a * np.sum(cu * ck * (-V2e - V3e), axis=1)
    H2, H3 = (r2 / N) * (e1m - e1p) / d1, (r1 / N) * (e2m - e2p) / d2
    H2y, H3y = (r2 / N) * (r1 / a) * (e1m + e1p) / d1, (r1 / N) * (r2 / a) * (e2m + e2p) / d2
    H_dx = np.sum(-2.0 * sgn * sk / (a * ll * al**2) * (H2 - H3), axis=1)
    H_dy = np.sum(2.0 * sgn * ck / (ll * al**3) * (H2y - H3y), axis=1)
    scale_j = a**2 * np.sqrt(sigma) / np.sqrt(mu) * (-grad_pz)
    jx, jy = scale_j * H_dy, scale_j * (-H_dx)
    z = np.zeros_like(uz)
    return tuple(np.where(inside, v, z) for v in (uz, uz_x, uz_y, jx, jy))


def hunt_error_norms(fes, x, T6, Ha, nsums, u0=1.0, jscale=1.0, a=1.0, mu=1.0, sigma=1.0, grad_pz=-1.0, chunk=4096):
    """eu_l2, eu_h1, ej_l2 of the reference's post-processing (hunt.jl:239-256): errors against the analytical Hunt
    solution with the degree 2*(order+1) quadrature table T6."""
    X = fes.mesh.cell_coords()
    st = fes.cell_state(x)
    nc = X.shape[0]
    w, gN, psi, _ = mapped_bases(T6, X, fes.j_sign)
    # physical coordinates of the quadrature points (trilinear map of the 8 vertices)
    xq = np.einsum("qv,cvi->cqi", T6.geo_val, X)
    us = st[:, :81].reshape(nc, 3, 27) * u0
    js = st[:, 85:121] * jscale
    uq = np.einsum("qa,cia->cqi", T6.nu, us)
    gu = np.einsum("cqbd,cib->cqdi", gN, us)  # [d,i] = d_d u_i
    jq = np.einsum("cqmi,cm->cqi", psi, js)
    pts = xq.reshape(-1, 3)[:, :2]
    outs = [np.empty(len(pts)) for _ in range(5)]
    for s in range(0, len(pts), chunk):
        r = analytical_hunt(pts[s : s + chunk], a=a, b=a, mu=mu, sigma=sigma, grad_pz=grad_pz, Ha=Ha, n=nsums)
        for o, v in zip(outs, r):
            o[s : s + chunk] = v
    uz, uz_x, uz_y, jx, jy = (o.reshape(nc, -1) for o in outs)
    eu = -uq.copy()
    eu[:, :, 2] += uz
    ge = -gu.copy()
    ge[:, :, 0, 2] += uz_x
    ge[:, :, 1, 2] += uz_y
    ej = -jq.copy()
    ej[:, :, 0] += jx
    ej[:, :, 1] += jy
    l2 = np.einsum("cq,cqi,cqi->", w, eu, eu)
    h1 = np.einsum("cq,cqdi,cqdi->", w, ge, ge)
    jl2 = np.einsum("cq,cqi,cqi->", w, ej, ej)
    return {"eu_l2": np.sqrt(l2), "eu_h1": np.sqrt(h1 + l2), "ej_l2": np.sqrt(jl2)}
