"""Design study for the next Jacobian kernel: the uu block by SUM FACTORISATION.

The Q2 basis and the 27-point Gauss rule are tensor products, N_a(q) = l_i(q1) l_j(q2) l_k(q3), so every uu contribution
    K[(a,c),(b,d)] = sum_q  D^m N_a(q)  D^n N_b(q)  C^{mn}_{cd}(q)          (D^0 = value, D^1..3 = reference derivatives)
can be contracted one direction at a time with the 3 x 3 x 3 tables  P^{mn}_x[i,i',q] = D^m l_i(q) D^n l_i'(q):
    T1[k,k',q1,q2] = sum_q3 P_z[k,k',q3] C[q1,q2,q3]     (243 FMA per coefficient field)
    T2[j,j',k,k',q1] = sum_q2 P_y[j,j',q2] T1[k,k',q1,q2]  (729 FMA)
    K[i,i',j,j',k,k'] = sum_q1 P_x[i,i',q1] T2[...]         (2187 FMA)
i.e. 3 159 FMA per coefficient field instead of 729 x 27 = 19 683.  Coefficient fields of the uu block (all built from
point data the kernel has anyway: w|det J|, J^{-1}, u, grad u):
    mass-type   N_a N_b M_cd            9 fields (Lorentz gamma(|B|^2 d_cd - B_c B_d) + Newton alpha d_d u_c)   -> 9 x 3 159
    stiffness   d_m N_a d_n N_b G^{mn}  9 fields G = beta w|det| J^{-1} J^{-T} (same for the 3 diagonal components) -> 9 x 3 159
    convection  N_a d_n N_b U^n         3 fields U = alpha w|det| J^{-1} u                                        -> 3 x 3 159
This script checks the algebra against the oracle's dense cell matrices on a non-affine mesh and prints the FMA counts.
Run: python tools_sumfac_study.py   (CPU only; test infrastructure, not product code)"""
import numpy as np

import gridapmhd_jl_b200  # noqa: F401
from gridapmhd_jl_b200.host import mesh as M
from gridapmhd_jl_b200.host.fespaces import setup_fe_spaces
from gridapmhd_jl_b200.host.reffe import Q2_NODE_IJK, _lagrange_1d, gauss_legendre_01
from oracle import mhd_oracle as O


def tables_1d():
    x, _ = gauss_legendre_01(3)
    v, d = _lagrange_1d(np.array([0.0, 0.5, 1.0]), x)  # [3 basis, 3 points]
    return v, d


def sumfac(Px, Py, Pz, C):
    """K[i,i',j,j',k,k'] = sum_q Px[i,i',q1] Py[j,j',q2] Pz[k,k',q3] C[q1,q2,q3]; C indexed [q1,q2,q3] (x fastest in the rule)"""
    T1 = np.einsum("kKc,abc->kKab", Pz, C)
    T2 = np.einsum("jJb,kKab->jJkKa", Py, T1)
    return np.einsum("iIa,jJkKa->iIjJkK", Px, T2)


def main():
    m = M.expansion_generate_mesh(0, perturb=0.2, seed=1)
    fes = setup_fe_spaces(m, u_tags=("inlet", "wall"), u_values=(None, None), j_tags=("wall",))
    T = fes.tables
    prm = O.FluidParams(alpha=0.7, beta=0.3, gamma=50.0, sigma=1.0, B=(0.2, 1.0, -0.1), convection="newton")
    x = np.random.default_rng(0).random(fes.ndofs)
    X = m.cell_coords()
    st = fes.cell_state(x)
    Kref = O.cell_jacobians(T, X, st, fes.j_sign, prm)[:, :81, :81]
    v, d = tables_1d()
    # tensor structure of the tables: N_a(q) = v[i,q1] v[j,q2] v[k,q3] with q = q1 + 3 q2 + 9 q3 and a <-> (i,j,k) = Q2_NODE_IJK[a]
    N = np.einsum("ia,jb,kc->ijkcba", v, v, v).reshape(3, 3, 3, 27)  # [i,j,k,q] with q3 slowest
    assert np.abs(N[Q2_NODE_IJK[:, 0], Q2_NODE_IJK[:, 1], Q2_NODE_IJK[:, 2]].T - T.nu).max() < 1e-14
    D = [v, d]  # D[0] values, D[1] derivatives
    P = {(mm, nn): np.einsum("iq,Iq->iIq", D[mm], D[nn]) for mm in (0, 1) for nn in (0, 1)}
    _, det, invJ = O.cell_geometry(T, X)
    w = T.w[None, :] * np.abs(det)
    nc = X.shape[0]
    us = st[:, :81].reshape(nc, 3, 27)
    uq = np.einsum("qa,cia->cqi", T.nu, us)
    gN = np.einsum("qak,cqki->cqai", T.dnu, invJ)
    gu = np.einsum("cqbd,cib->cqdi", gN, us)  # [d,i] = d_d u_i
    B = np.asarray(prm.B)
    ijk = Q2_NODE_IJK
    err = 0.0
    for c in range(nc):
        K = np.zeros((3, 27, 3, 27))
        fld = lambda a: a.reshape(3, 3, 3).transpose(2, 1, 0)  # [q] (q3 slowest) -> [q1,q2,q3]
        # mass-type fields (note: H1-HDiv has no gamma (u x B).(v x B) term in uu; the Newton term alone)
        for ci in range(3):
            for di in range(3):
                Kt = sumfac(P[0, 0], P[0, 0], P[0, 0], fld(w[c] * prm.alpha * gu[c, :, di, ci]))
                K[ci, :, di, :] += Kt[ijk[:, 0][:, None], ijk[:, 0][None, :], ijk[:, 1][:, None], ijk[:, 1][None, :], ijk[:, 2][:, None], ijk[:, 2][None, :]]
        # stiffness: G^{mn} = beta w sum_i invJ[m,i] invJ[n,i]
        G = prm.beta * np.einsum("q,qmi,qni->qmn", w[c], invJ[c], invJ[c])
        S = np.zeros((27, 27))
        for mm in range(3):
            for nn in range(3):
                dm = [1 if ax == mm else 0 for ax in range(3)]
                dn = [1 if ax == nn else 0 for ax in range(3)]
                Kt = sumfac(P[dm[0], dn[0]], P[dm[1], dn[1]], P[dm[2], dn[2]], fld(G[:, mm, nn]))
                S += Kt[ijk[:, 0][:, None], ijk[:, 0][None, :], ijk[:, 1][:, None], ijk[:, 1][None, :], ijk[:, 2][:, None], ijk[:, 2][None, :]]
        # convection (Picard part): U^n = alpha w sum_i invJ[n,i] u_i
        U = prm.alpha * np.einsum("q,qni,qi->qn", w[c], invJ[c], uq[c])
        for nn in range(3):
            dn = [1 if ax == nn else 0 for ax in range(3)]
            Kt = sumfac(P[0, dn[0]], P[0, dn[1]], P[0, dn[2]], fld(U[:, nn]))
            S += Kt[ijk[:, 0][:, None], ijk[:, 0][None, :], ijk[:, 1][:, None], ijk[:, 1][None, :], ijk[:, 2][:, None], ijk[:, 2][None, :]]
        for ci in range(3):
            K[ci, :, ci, :] += S
        err = max(err, np.abs(K.reshape(81, 81) - Kref[c]).max() / np.abs(Kref[c]).max())
    per_field = 9 * 9 * 3 + 9 * 9 * 3 * 3 + 9 * 9 * 9 * 3
    direct = 729 * 27 * (9 + 3 + 1)
    print(f"uu block, {nc} non-affine cells: max relative deviation from the oracle {err:.2e}")
    print(f"FMA per cell: sum-factorised {21 * per_field} (21 coefficient fields x {per_field}) vs direct {direct} "
          f"(729 pairs x 27 points x 13) -> {direct / (21 * per_field):.1f}x fewer")
    print("operand traffic: the 1-D tables are 4 x 27 doubles (registers / constant bank); T1, T2 intermediates 81 + 243 doubles per field")
    assert err < 1e-12


if __name__ == "__main__":
    main()
