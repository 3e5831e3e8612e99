"""CPU ORACLE for the H1-H1 formulation (test infrastructure only -- never imported by the product path).

NumPy restatement of `res_fluid_h1_h1` / `jac_fluid_h1_h1` (/root/reference/src/weakforms.jl:415-438, :440-466) and of
the solid-cell forms `res_solid_h1_h1` / `jac_solid_h1_h1` (:468-478): u Q2, p P1disc, phi Q3 continuous; the current
j = sigma (u x B - grad phi) is eliminated.  Terms (fluid):

    beta grad u : grad v + gamma (u x B).(v x B) [+ zeta_u Pi_p(u) div v] [+ alpha v.conv(u, grad u)]
    - p div v - div u q + grad phi . grad w - gamma grad phi . (v x B) - (u x B) . grad w - f.v

with conv(u, grad u) = (grad u)'.u (weakforms.jl:670) and Pi_p the cell-wise L2 projection of `div` onto the pressure
space (:672-681).  Touched blocks: uu, up, u-phi, pu, phi-u, phi-phi (pp, p-phi, phi-p are never inserted).
Assembly semantics and quadrature are those of mhd_oracle.py (Gridap SparseMatrixAssembler; Quadrature(HEX,5)).

PARITY PINNING: as for H1-HDiv, the reference holds no entry-level golden values and publishes no H1-H1 norms
("parity unpinned" at the entry level).  Pins used in tests/test_oracle_h1h1.py: FD-Jacobian, in-space manufactured
solution (residual = 0), block identities (K_pu = K_up', K_{phi u} = K_{u phi}'/gamma, K_{phi phi} SPSD with the
constants in its kernel), and the Hunt solution at SOLUTION level: the discrete H1-H1 velocity converges to the same
analytical Hunt series as the (published-norm-pinned) H1-HDiv solution.
"""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla

from .mhd_oracle import FluidParams, _cross, analytical_hunt, cell_geometry

NU, NP, NPHI = 81, 4, 64
OFF_U, OFF_P, OFF_PHI = 0, 81, 85
NLOC = 149


def touched_mask() -> np.ndarray:
    m = np.zeros((NLOC, NLOC), dtype=bool)
    u, p, ph = slice(0, 81), slice(81, 85), slice(85, 149)
    for r, c in ((u, u), (u, p), (u, ph), (p, u), (ph, u), (ph, ph)):
        m[r, c] = True
    return m


def mapped_bases(T, X):
    _, det, invJ = cell_geometry(T, X)
    w = T.w[None, :] * np.abs(det)
    gN = np.einsum("qak,cqki->cqai", T.dnu, invJ)
    gF = np.einsum("qlk,cqki->cqli", T.dphi3, invJ)
    return w, gN, gF


def cell_jacobians(T, X, state, prm: FluidParams):
    """Dense cell matrices [nc,149,149] of `jac_fluid_h1_h1` (weakforms.jl:440-466).  Solid cells need no special case:
    their u/p dofs are absent (dropped at assembly) and the phi-phi block is the same Laplacian (:474-478)."""
    nc = X.shape[0]
    w, gN, gF = mapped_bases(T, X)
    N, Pp = T.nu, T.pp
    B = np.asarray(prm.B, dtype=float)
    K = np.zeros((nc, NLOC, NLOC))
    S = np.einsum("cq,cqai,cqbi->cab", w, gN, gN)
    M = np.einsum("cq,qa,qb->cab", w, N, N)
    Kuu = np.zeros((nc, 3, 27, 3, 27))
    for c in range(3):
        Kuu[:, c, :, c, :] += prm.beta * S
    # gamma (du x B).(v x B) = gamma N_a N_b (|B|^2 delta_cd - B_c B_d)
    L = prm.gamma * (B @ B * np.eye(3) - np.outer(B, B))
    Kuu += np.einsum("cab,id->ciadb", M, L)
    if prm.convection in ("picard", "newton"):
        us = state[:, :81].reshape(nc, 3, 27)
        uq = np.einsum("qa,cia->cqi", N, us)
        ugN = np.einsum("cqi,cqbi->cqb", uq, gN)
        C = np.einsum("cq,qa,cqb->cab", w, N, ugN)
        for c in range(3):
            Kuu[:, c, :, c, :] += prm.alpha * C
        if prm.convection == "newton":
            gu = np.einsum("cqbd,cib->cqdi", gN, us)  # gu[d,i] = d_d u_i
            Kuu += prm.alpha * np.einsum("cq,qa,qb,cqdi->ciadb", w, N, N, gu)
    if prm.zeta_u != 0.0:
        D = np.einsum("cq,qk,cqai->ckia", w, Pp, gN).reshape(nc, 4, 81)
        Mp = np.einsum("cq,qk,ql->ckl", w, Pp, Pp)
        E = np.linalg.solve(Mp, D)
        Kuu += prm.zeta_u * np.einsum("cki,ckj->cij", D, E).reshape(nc, 3, 27, 3, 27)
    K[:, :81, :81] = Kuu.reshape(nc, 81, 81)
    Kup = -np.einsum("cq,qk,cqai->ciak", w, Pp, gN).reshape(nc, 81, 4)
    K[:, :81, 81:85] = Kup
    K[:, 81:85, :81] = np.transpose(Kup, (0, 2, 1))
    # u-phi: -gamma grad(dphi).(v x B), v = N_a e_c: (e_c x B).g = (B x g)_c
    BxG = -_cross(gF, B)  # B x grad phi_l  [c,q,l,3]
    V = np.einsum("cq,qa,cqli->cial", w, N, BxG).reshape(nc, 81, 64)
    K[:, :81, 85:] = -prm.gamma * V
    # phi-u: -(du x B).grad w
    K[:, 85:, :81] = -np.transpose(V, (0, 2, 1))
    K[:, 85:, 85:] = np.einsum("cq,cqli,cqmi->clm", w, gF, gF)
    return K


def cell_residuals(T, X, state, prm: FluidParams):
    """Cell vectors [nc,149] of `res_fluid_h1_h1` (weakforms.jl:415-438), divg = 0."""
    nc = X.shape[0]
    w, gN, gF = mapped_bases(T, X)
    N, Pp = T.nu, T.pp
    B = np.asarray(prm.B, dtype=float)
    f = np.asarray(prm.f, dtype=float)
    us = state[:, :81].reshape(nc, 3, 27)
    ps = state[:, 81:85]
    fs = state[:, 85:]
    uq = np.einsum("qa,cia->cqi", N, us)
    gu = np.einsum("cqbd,cib->cqdi", gN, us)
    divu = np.einsum("cqii->cq", gu)
    pq = np.einsum("qk,ck->cq", Pp, ps)
    gphi = np.einsum("cqli,cl->cqi", gF, fs)
    uB = _cross(uq, B)
    R = np.zeros((nc, NLOC))
    ru = prm.beta * np.einsum("cq,cqdi,cqad->cia", w, gu, gN)
    # gamma (u x B).(v x B) = gamma N_a [B x (u x B)]_c
    ru += prm.gamma * np.einsum("cq,qa,cqi->cia", w, N, -_cross(uB, B))
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
    ru -= prm.gamma * np.einsum("cq,qa,cqi->cia", w, N, -_cross(gphi, B))  # grad phi.(v x B) = N_a (B x grad phi)_c
    ru -= np.einsum("cq,qa,i->cia", w, N, f)
    R[:, :81] = ru.reshape(nc, 81)
    R[:, 81:85] = -np.einsum("cq,qk,cq->ck", w, Pp, divu)
    R[:, 85:] = np.einsum("cq,cqi,cqli->cl", w, gphi - uB, gF)
    return R


def symbolic_csr(gids: np.ndarray, n: int):
    li, lj = np.nonzero(touched_mask())
    P = None
    for s in range(0, gids.shape[0], 4096):
        g = gids[s : s + 4096]
        r, c = g[:, li], g[:, lj]
        ok = (r >= 0) & (c >= 0)
        Pi = sp.coo_matrix((np.ones(int(ok.sum()), dtype=np.int8), (r[ok], c[ok])), shape=(n, n)).tocsr()
        Pi.sum_duplicates()
        Pi.data[:] = 1
        P = Pi if P is None else P + Pi
    P.sum_duplicates()
    P.sort_indices()
    return P.indptr.astype(np.int64), P.indices.astype(np.int64)


def _assemble(K, gids, n, pattern):
    rowptr, colval = pattern
    keyP = np.repeat(np.arange(n, dtype=np.int64), np.diff(rowptr)) * n + colval
    li, lj = np.nonzero(touched_mask())
    r, c = gids[:, li], gids[:, lj]
    ok = (r >= 0) & (c >= 0)
    data = np.zeros(len(colval))
    np.add.at(data, np.searchsorted(keyP, r[ok] * n + c[ok]), K[:, li, lj][ok])
    return data


def jacobian(fes, x, prm: FluidParams, chunk: int = 1024, pattern=None) -> sp.csr_matrix:
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
        data += _assemble(cell_jacobians(fes.tables, X[sl], st[sl], prm), gids[sl], n, pattern)
    return sp.csr_matrix((data, colval.copy(), rowptr.copy()), shape=(n, n))


def residual(fes, x, prm: FluidParams, chunk: int = 4096) -> np.ndarray:
    X = fes.mesh.cell_coords()
    st = fes.cell_state(x)
    gids = fes.cell_global_ids()
    out = np.zeros(fes.ndofs)
    for s in range(0, X.shape[0], chunk):
        sl = slice(s, s + chunk)
        R = cell_residuals(fes.tables, X[sl], st[sl], prm)
        ok = gids[sl] >= 0
        np.add.at(out, gids[sl][ok], R[ok])
    return out


def newton_lu(fes, prm: FluidParams, x0=None, maxiter=10, rtol=1e-6, verbose=False, min_iters=1):
    """`_solver(::Val{:julia})` (src/main.jl:181-186): Newton with sparse LU."""
    x = np.zeros(fes.ndofs) if x0 is None else x0.copy()
    b = residual(fes, x, prm)
    r0 = np.linalg.norm(b)
    hist = [r0]
    pattern = symbolic_csr(fes.cell_global_ids(), fes.ndofs)
    for it in range(maxiter):
        A = jacobian(fes, x, prm, pattern=pattern)
        x = x + spla.splu(A.tocsc()).solve(-b)
        b = residual(fes, x, prm)
        hist.append(np.linalg.norm(b))
        if verbose:
            print(f"  newton it {it+1}: |r| = {hist[-1]:.3e} (rel {hist[-1]/r0:.3e})")
        if (hist[-1] <= rtol * r0 or hist[-1] < 1e-14) and it + 1 >= min_iters:
            break
    return x, hist


def hunt_norms(fes, x, T6, Bbar, Ha, nsums, u0=1.0, jscale=1.0, a=1.0, mu=1.0, sigma=1.0, grad_pz=-1.0, chunk=4096):
    """Post-processing of `hunt` for `current_disc == :H1` (src/Applications/hunt.jl:218-226,247-260):
    jh = jscale (ubar_h x Bbar - grad phibar_h); errors against the analytical series."""
    X = fes.mesh.cell_coords()
    st = fes.cell_state(x)
    nc = X.shape[0]
    w, gN, gF = mapped_bases(T6, X)
    xq = np.einsum("qv,cvi->cqi", T6.geo_val, X)
    us = st[:, :81].reshape(nc, 3, 27)
    uq = np.einsum("qa,cia->cqi", T6.nu, us)
    gu = np.einsum("cqbd,cib->cqdi", gN, us) * u0
    gphi = np.einsum("cqli,cl->cqi", gF, st[:, 85:])
    jq = jscale * (_cross(uq, Bbar) - gphi)
    uq = uq * u0
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
    uu = np.einsum("cq,cqi,cqi->", w, uq, uq)
    gg = np.einsum("cq,cqdi,cqdi->", w, gu, gu)
    return {"eu_l2": np.sqrt(l2), "eu_h1": np.sqrt(h1 + l2), "ej_l2": np.sqrt(np.einsum("cq,cqi,cqi->", w, ej, ej)),
            "uh_l2": np.sqrt(uu), "uh_h1": np.sqrt(gg + uu), "jh_l2": np.sqrt(np.einsum("cq,cqi,cqi->", w, jq, jq))}
