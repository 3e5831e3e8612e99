"""Design study (CPU, test infrastructure): the WHOLE H1-HDiv cell Jacobian from 1-D factors of the reference tables.

1. discovers the tensor-product structure of the Q2 / RT1 / Q1 tables numerically (rank-1 factorisation + grid completion),
2. evaluates every block as sum_q F(q) a(q) b(q) with a, b products of 1-D factors and F a per-cell coefficient field,
3. compares with oracle.cell_jacobians.
The CUDA kernel (csrc/hdiv7_cell.h) contracts the same expressions one direction at a time."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def rank1(t):
    """t[3,3,3] (q1,q2,q3) -> scale, f1, f2, f3 with max-abs of every factor = 1 at the arg-max; deviation"""
    p = np.unravel_index(np.argmax(np.abs(t)), t.shape)
    s = t[p]
    f1, f2, f3 = t[:, p[1], p[2]] / s, t[p[0], :, p[2]] / s, t[p[0], p[1], :] / s
    dev = np.abs(t - s * np.einsum("i,j,k->ijk", f1, f2, f3)).max()
    return s, (f1, f2, f3), dev


def classify(fs, tol=1e-10):
    """cluster 1-D functions up to sign: returns class id and sign of each"""
    reps, ids, sg = [], [], []
    for f in fs:
        for i, r in enumerate(reps):
            if np.abs(f - r).max() < tol:
                ids.append(i); sg.append(1.0); break
            if np.abs(f + r).max() < tol:
                ids.append(i); sg.append(-1.0); break
        else:
            reps.append(f); ids.append(len(reps) - 1); sg.append(1.0)
    return np.array(ids), np.array(sg), reps


def discover_scalar(tab, dtab=None):
    """tab[27 q, n] scalar basis on the tensor rule (q = q1 + 3 q2 + 9 q3).  Returns idx[n,3] (class per direction), per-direction
    tables V[d][class][q] such that tab[q,f] = prod_d V[d][idx[f,d]][q_d] EXACTLY in structure (scales folded into direction 0
    through grid completion), and the derivative tables D[d] if dtab[27,n,3] is given."""
    n = tab.shape[1]
    T = tab.T.reshape(n, 3, 3, 3).transpose(0, 3, 2, 1)  # [f, q1, q2, q3]
    sc, F = [], [[], [], []]
    for f in range(n):
        s, fs, dev = rank1(T[f])
        assert dev < 1e-12 * abs(s), dev
        sc.append(s)
        for d in range(3):
            F[d].append(fs[d])
    idx = np.zeros((n, 3), dtype=int)
    ncls = []
    for d in range(3):
        ids, sg, reps = classify(F[d])
        idx[:, d] = ids
        ncls.append(len(reps))
    # the functions must fill the grid of classes exactly once
    assert n == ncls[0] * ncls[1] * ncls[2], (n, ncls)
    assert len({tuple(r) for r in idx}) == n
    # grid completion: tables from the "axis" functions through function 0
    f0 = 0
    base = idx[f0]
    V = [np.zeros((ncls[d], 3)) for d in range(3)]
    find = {tuple(r): f for f, r in enumerate(idx)}
    q0 = [int(np.argmax(np.abs(F[d][f0]))) for d in range(3)]  # a point where every factor of f0 is 1
    for d in range(3):
        for c in range(ncls[d]):
            key = list(base); key[d] = c
            f = find[tuple(key)]
            # line through q0 along direction d
            sl = [q0[0], q0[1], q0[2]]; sl[d] = slice(None)
            line = T[f][tuple(sl)]
            V[d][c] = line / (T[f0][tuple(q0)] if d > 0 else 1.0)
    rec = np.einsum("fi,fj,fk->fijk", V[0][idx[:, 0]], V[1][idx[:, 1]], V[2][idx[:, 2]])
    assert np.abs(rec - T).max() < 1e-12 * np.abs(T).max(), np.abs(rec - T).max()
    D = None
    if dtab is not None:
        DT = dtab.transpose(1, 2, 0).reshape(n, 3, 3, 3, 3).transpose(0, 1, 4, 3, 2)  # [f, dir, q1, q2, q3]
        D = [np.zeros((ncls[d], 3)) for d in range(3)]
        for d in range(3):
            for c in range(ncls[d]):
                key = list(base); key[d] = c
                f = find[tuple(key)]
                sl = [q0[0], q0[1], q0[2]]; sl[d] = slice(None)
                others = np.prod([V[e][idx[f, e]][q0[e]] for e in range(3) if e != d])
                D[d][c] = DT[f, d][tuple(sl)] / others
        for d in range(3):
            facs = [V[e][idx[:, e]] for e in range(3)]
            facs[d] = D[d][idx[:, d]]
            rec = np.einsum("fi,fj,fk->fijk", *facs)
            assert np.abs(rec - DT[:, d]).max() < 1e-11 * np.abs(DT).max()
    return idx, V, D


def discover_rt(psi, dpsi):
    """psi[27,36,3], dpsi[27,36]: per component k the 12 functions with that single non-zero component."""
    comp = np.argmax(np.abs(psi).max(axis=0), axis=1)
    for m in range(36):
        off = [k for k in range(3) if k != comp[m]]
        assert np.abs(psi[:, m, off]).max() < 1e-13
    out = []
    for k in range(3):
        ms = np.nonzero(comp == k)[0]
        assert len(ms) == 12
        d3 = np.zeros((27, 12, 3))
        d3[:, :, k] = dpsi[:, ms]
        idx, V, D = discover_scalar(psi[:, ms, k], d3)
        assert [len(v) for v in V] == [3 if d == k else 2 for d in range(3)]
        out.append((ms, idx, V, D[k]))
    return out


def cell_jacobian_factored(T, X, state, j_sign, prm, st):
    """Dense 129x129 cell matrix from the factored tables `st` (one cell)."""
    q_idx = np.array([[q % 3, (q // 3) % 3, q // 9] for q in range(27)])
    (uidx, LV, LD), rt, (xidx, XV) = st
    N = np.prod([LV[d][uidx[:, d]][:, q_idx[:, d]] for d in range(3)], axis=0).T  # [q, a]
    dN = np.zeros((27, 27, 3))
    for k in range(3):
        fac = [LV[d][uidx[:, d]][:, q_idx[:, d]] for d in range(3)]
        fac[k] = LD[k][uidx[:, k]][:, q_idx[:, k]]
        dN[:, :, k] = np.prod(fac, axis=0).T
    psh = np.zeros((27, 36))  # scalar psi-hat of each dof (its component = comp[m])
    dph = np.zeros((27, 36))
    comp = np.zeros(36, dtype=int)
    for k, (ms, idx, V, Dk) in enumerate(rt):
        comp[ms] = k
        psh[:, ms] = np.prod([V[d][idx[:, d]][:, q_idx[:, d]] for d in range(3)], axis=0).T
        fac = [V[d][idx[:, d]][:, q_idx[:, d]] for d in range(3)]
        fac[k] = Dk[idx[:, k]][:, q_idx[:, k]]
        dph[:, ms] = np.prod(fac, axis=0).T
    chi = np.prod([XV[d][xidx[:, d]][:, q_idx[:, d]] for d in range(3)], axis=0).T
    J = np.einsum("vi,qvk->qik", X, T.geo_grad)
    det = np.linalg.det(J)
    inv = np.linalg.inv(J)  # inv[q][k][i] = d xi_k / d x_i
    W = T.w * np.abs(det)
    B = np.asarray(prm.B, float)
    K = np.zeros((129, 129))
    sg = j_sign.astype(float)
    us = state[:81].reshape(3, 27)
    # ---- uu
    G = prm.beta * np.einsum("q,qmi,qni->qmn", W, inv, inv)
    Kb = np.einsum("qmn,qam,qbn->ab", G, dN, dN)
    Kuu = np.zeros((3, 27, 3, 27))
    if prm.convection != "none":
        uq = N @ us.T  # [q, i]
        U = prm.alpha * np.einsum("q,qni,qi->qn", W, inv, uq)
        Kb += np.einsum("qn,qa,qbn->ab", U, N, dN)
        if prm.convection == "newton":
            gur = np.einsum("qak,ca->qkc", dN, us)  # reference gradient
            M = prm.alpha * np.einsum("q,qkd,qkc->qcd", W, inv, gur)  # M[c,d] = alpha W d_d u_c
            Kuu += np.einsum("qcd,qa,qb->cadb", M, N, N)
    for c in range(3):
        Kuu[c, :, c, :] += Kb
    # ---- up via fields  F[k,k',c] = W pi_k inv[k'][c]
    Dm = np.einsum("q,qk,qlc,qal->kca", W, T.pp, inv, dN).reshape(4, 81)
    if prm.zeta_u != 0.0:
        Mp = np.einsum("q,qk,ql->kl", W, T.pp, T.pp)
        E = np.linalg.solve(Mp, Dm)
        Kuu += prm.zeta_u * (Dm.T @ E).reshape(3, 27, 3, 27)
    K[:81, :81] = Kuu.reshape(81, 81)
    K[:81, 81:85] = -Dm.T
    K[81:85, :81] = -Dm
    # ---- uj / ju: V[c][a][m] = sum_q E_{c,k(m)} N_a psh_m
    E_ck = np.zeros((27, 3, 3))
    for c in range(3):
        c1, c2 = (c + 1) % 3, (c + 2) % 3
        E_ck[:, c, :] = (W / det)[:, None] * (J[:, c1, :] * B[c2] - J[:, c2, :] * B[c1])
    V = np.einsum("qcm,qa,qm->cam", E_ck[:, :, comp], N, psh) * sg[None, None, :]
    K[:81, 85:121] = -prm.gamma * V.reshape(81, 36)
    K[85:121, :81] = prm.sigma * V.reshape(81, 36).T
    # ---- jj
    H = np.einsum("q,qik,qil->qkl", W / det**2, J, J)
    Kjj = np.einsum("qmn,qm,qn->mn", H[:, comp][:, :, comp], psh, psh)
    if prm.zeta_j != 0.0:
        Kjj += prm.zeta_j * np.einsum("q,qm,qn->mn", W / det**2, dph, dph)
    K[85:121, 85:121] = Kjj * np.outer(sg, sg)
    JF = np.einsum("q,qm,ql->ml", W / det, dph, chi) * sg[:, None]
    K[85:121, 121:129] = -prm.sigma * JF
    K[121:129, 85:121] = -JF.T
    return K


if __name__ == "__main__":
    import gridapmhd_jl_b200  # noqa: F401
    from gridapmhd_jl_b200.applications import u_inlet_parabolic
    from gridapmhd_jl_b200.host import mesh as M
    from gridapmhd_jl_b200.host.fespaces import setup_fe_spaces
    from oracle import mhd_oracle as O

    m = M.expansion_generate_mesh(0, perturb=0.2, seed=1)
    fes = setup_fe_spaces(m, u_tags=("inlet", "wall"), u_values=(u_inlet_parabolic(), None), j_tags=("wall", "inlet", "outlet"))
    T = fes.tables
    st = (discover_scalar(T.nu, T.dnu), discover_rt(T.psi, T.dpsi), discover_scalar(T.chi)[:2])
    print("Q2 classes", [len(v) for v in st[0][1]], "chi", [len(v) for v in st[2][1]])
    prm = O.FluidParams(alpha=0.5, beta=0.01, gamma=3.0, sigma=0.7, zeta_u=2.0, zeta_j=2.0, B=(0.2, 1.0, -0.3), convection="newton")
    x = np.random.default_rng(2).random(fes.ndofs)
    Ko = O.cell_jacobians(T, m.cell_coords(), fes.cell_state(x), fes.j_sign, prm)
    Xc, S = m.cell_coords(), fes.cell_state(x)
    err = 0.0
    for c in range(min(10, m.ncells)):
        K = cell_jacobian_factored(T, Xc[c], S[c], fes.j_sign[c], prm, st)
        err = max(err, np.abs(K - Ko[c]).max() / np.abs(Ko[c]).max())
    print("max rel err vs oracle", err)
