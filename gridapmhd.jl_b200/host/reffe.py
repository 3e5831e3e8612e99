"""Reference elements and quadrature tabulated for the H1-HDiv MHD hot path.

Host-side mirror of the element selection in the reference
(`src/parameters.jl:436-441` Q2 vector Lagrangian + discontinuous P1,
`src/parameters.jl:521-525` Raviart-Thomas order 1 + discontinuous Q1,
`src/parameters.jl:381-389,617-639` quadrature degree q=5 -> 3x3x3 Gauss).

In production these tables come from Gridap (the C ABI is basis-agnostic: it
only consumes tabulated values at the quadrature points, `include/mhdb200.h`
`mhd_tables_t`).  In this repository the Python host generates them.

Conventions: HEX reference cell [0,1]^3, vertices
lexicographic with x fastest; faces ordered z=0, z=1, y=0, y=1, x=0, x=1;
edges: 4 along x, 4 along y, 4 along z.
"""
from __future__ import annotations

import itertools
from dataclasses import dataclass

import numpy as np

# ----------------------------------------------------------------------------
# HEX topology (local numbering)

HEX_VERTS = np.array([[i, j, k] for k in (0, 1) for j in (0, 1) for i in (0, 1)], dtype=np.int64)
# edges as pairs of local vertices: along x, along y, along z
HEX_EDGES = np.array(
    [[0, 1], [2, 3], [4, 5], [6, 7], [0, 2], [1, 3], [4, 6], [5, 7], [0, 4], [1, 5], [2, 6], [3, 7]],
    dtype=np.int64,
)
# faces as 4 local vertices in the face-lexicographic order
HEX_FACES = np.array(
    [[0, 1, 2, 3], [4, 5, 6, 7], [0, 1, 4, 5], [2, 3, 6, 7], [0, 2, 4, 6], [1, 3, 5, 7]], dtype=np.int64
)
# (normal axis, side) per face, outward normal = (2*side-1) e_axis
HEX_FACE_AXIS = np.array([2, 2, 1, 1, 0, 0], dtype=np.int64)
HEX_FACE_SIDE = np.array([0, 1, 0, 1, 0, 1], dtype=np.int64)


def gauss_legendre_01(n: int):
    """n-point Gauss-Legendre rule on [0,1]."""
    x, w = np.polynomial.legendre.leggauss(n)
    return 0.5 * (x + 1.0), 0.5 * w


def tensor_quadrature(n: int):
    """Tensor Gauss rule on [0,1]^3, x fastest. Returns (points[nq,3], weights[nq])."""
    x, w = gauss_legendre_01(n)
    pts = np.array([[x[i], x[j], x[k]] for k in range(n) for j in range(n) for i in range(n)])
    wts = np.array([w[i] * w[j] * w[k] for k in range(n) for j in range(n) for i in range(n)])
    return pts, wts


def quadrature_for_degree(q: int):
    """Gridap `Quadrature(HEX,q)`: tensor Gauss with ceil((q+1)/2) points per direction."""
    return tensor_quadrature((q + 2) // 2)


# ----------------------------------------------------------------------------
# 1D Lagrange helpers


def _lagrange_1d(nodes: np.ndarray, x: np.ndarray):
    """Values and derivatives of the Lagrange basis on `nodes` at points x: ([n,len(x)], [n,len(x)])."""
    n = len(nodes)
    vals = np.ones((n, len(x)))
    ders = np.zeros((n, len(x)))
    for i in range(n):
        for j in range(n):
            if j != i:
                vals[i] *= (x - nodes[j]) / (nodes[i] - nodes[j])
        for j in range(n):
            if j == i:
                continue
            term = np.full(len(x), 1.0 / (nodes[i] - nodes[j]))
            for k in range(n):
                if k != i and k != j:
                    term = term * (x - nodes[k]) / (nodes[i] - nodes[k])
            ders[i] += term
    return vals, ders


# ----------------------------------------------------------------------------
# Q1 (trilinear) - geometry map and the discontinuous phi space


def q1_tabulate(pts: np.ndarray):
    """Trilinear vertex basis: values [nq,8], reference gradients [nq,8,3]."""
    nodes = np.array([0.0, 1.0])
    v = [_lagrange_1d(nodes, pts[:, d]) for d in range(3)]
    val = np.empty((len(pts), 8))
    grad = np.empty((len(pts), 8, 3))
    for a, (i, j, k) in enumerate(HEX_VERTS):
        val[:, a] = v[0][0][i] * v[1][0][j] * v[2][0][k]
        grad[:, a, 0] = v[0][1][i] * v[1][0][j] * v[2][0][k]
        grad[:, a, 1] = v[0][0][i] * v[1][1][j] * v[2][0][k]
        grad[:, a, 2] = v[0][0][i] * v[1][0][j] * v[2][1][k]
    return val, grad


# ----------------------------------------------------------------------------
# Q2 scalar Lagrangian: node order = 8 vertices, 12 edges, 6 faces, 1 interior


def q2_node_grid_indices():
    """(ix,iy,iz) in {0,1,2} (1 = mid point) for the 27 nodes in local order."""
    idx = []
    for v in HEX_VERTS:
        idx.append(tuple(2 * v))
    for e in HEX_EDGES:
        a, b = HEX_VERTS[e[0]], HEX_VERTS[e[1]]
        idx.append(tuple(a + b))
    for f in HEX_FACES:
        s = HEX_VERTS[f].sum(axis=0) // 2
        idx.append(tuple(s))
    idx.append((1, 1, 1))
    return np.array(idx, dtype=np.int64)


Q2_NODE_IJK = q2_node_grid_indices()
Q2_NODE_XI = Q2_NODE_IJK * 0.5  # reference coordinates of the nodes


def q2_tabulate(pts: np.ndarray):
    """Scalar Q2 basis: values [nq,27], reference gradients [nq,27,3]."""
    nodes = np.array([0.0, 0.5, 1.0])
    v = [_lagrange_1d(nodes, pts[:, d]) for d in range(3)]
    val = np.empty((len(pts), 27))
    grad = np.empty((len(pts), 27, 3))
    for a, (i, j, k) in enumerate(Q2_NODE_IJK):
        val[:, a] = v[0][0][i] * v[1][0][j] * v[2][0][k]
        grad[:, a, 0] = v[0][1][i] * v[1][0][j] * v[2][0][k]
        grad[:, a, 1] = v[0][0][i] * v[1][1][j] * v[2][0][k]
        grad[:, a, 2] = v[0][0][i] * v[1][0][j] * v[2][1][k]
    return val, grad


# ----------------------------------------------------------------------------
# discontinuous P1 in reference coordinates: span{1, xi, eta, zeta}, nodal at the simplex vertices


def p1_tabulate(pts: np.ndarray):
    val = np.empty((len(pts), 4))
    val[:, 0] = 1.0 - pts[:, 0] - pts[:, 1] - pts[:, 2]
    val[:, 1] = pts[:, 0]
    val[:, 2] = pts[:, 1]
    val[:, 3] = pts[:, 2]
    return val


# ----------------------------------------------------------------------------
# Raviart-Thomas RT1 on the HEX: Q(2,1,1) e1 + Q(1,2,1) e2 + Q(1,1,2) e3, dim 36.
# dofs: 6 faces x 4 normal-flux moments against the bilinear vertex functions of the face,
#       then 12 interior moments against Q(0,1,1)e1 + Q(1,0,1)e2 + Q(1,1,0)e3 (bilinear nodal).


def _rt1_monomials():
    """List of (comp, (a,b,c)) exponents of the 36 prebasis monomials."""
    mons = []
    for comp in range(3):
        ranges = [range(2), range(2), range(2)]
        ranges[comp] = range(3)
        for c in ranges[2]:
            for b in ranges[1]:
                for a in ranges[0]:
                    mons.append((comp, (a, b, c)))
    return mons


def _eval_monomials(mons, pts):
    """values [npts, 36, 3] and divergence [npts, 36] of the prebasis."""
    n = len(pts)
    val = np.zeros((n, len(mons), 3))
    div = np.zeros((n, len(mons)))
    x, y, z = pts[:, 0], pts[:, 1], pts[:, 2]
    for m, (comp, (a, b, c)) in enumerate(mons):
        val[:, m, comp] = x**a * y**b * z**c
        e = (a, b, c)[comp]
        if e > 0:
            ee = [a, b, c]
            ee[comp] -= 1
            div[:, m] = e * x ** ee[0] * y ** ee[1] * z ** ee[2]
    return val, div


def _bilinear(s, t):
    """4 bilinear nodal functions at (0,0),(1,0),(0,1),(1,1): [npts,4]."""
    return np.stack([(1 - s) * (1 - t), s * (1 - t), (1 - s) * t, s * t], axis=1)


_RT1_COEFFS = None


def rt1_coefficients():
    """Coefficient matrix C[36 mon, 36 basis] with basis_j = sum_m C[m,j] mon_m (dual to the moments)."""
    global _RT1_COEFFS
    if _RT1_COEFFS is not None:
        return _RT1_COEFFS
    mons = _rt1_monomials()
    g, gw = gauss_legendre_01(4)
    M = np.zeros((36, 36))  # M[dof, mon]
    # face moments
    S, T = np.meshgrid(g, g, indexing="ij")
    s, t = S.ravel(), T.ravel()
    w2 = (gw[:, None] * gw[None, :]).ravel()
    for f in range(6):
        ax, side = HEX_FACE_AXIS[f], HEX_FACE_SIDE[f]
        others = [d for d in range(3) if d != ax]
        pts = np.zeros((len(s), 3))
        pts[:, ax] = float(side)
        pts[:, others[0]] = s
        pts[:, others[1]] = t
        val, _ = _eval_monomials(mons, pts)
        vn = val[:, :, ax] * (2.0 * side - 1.0)
        q = _bilinear(s, t)
        M[4 * f : 4 * f + 4, :] = np.einsum("p,pk,pm->km", w2, q, vn)
    # interior moments
    pts3, w3 = tensor_quadrature(4)
    val, _ = _eval_monomials(mons, pts3)
    for comp in range(3):
        others = [d for d in range(3) if d != comp]
        q = _bilinear(pts3[:, others[0]], pts3[:, others[1]])
        M[24 + 4 * comp : 24 + 4 * comp + 4, :] = np.einsum("p,pk,pm->km", w3, q, val[:, :, comp])
    _RT1_COEFFS = np.linalg.inv(M)
    return _RT1_COEFFS


def rt1_tabulate(pts: np.ndarray):
    """Reference RT1 basis: values [nq,36,3] and reference divergence [nq,36]."""
    mons = _rt1_monomials()
    C = rt1_coefficients()
    val, div = _eval_monomials(mons, pts)
    return np.einsum("pmc,mj->pjc", val, C), div @ C


# ----------------------------------------------------------------------------


@dataclass
class Tables:
    """Reference-element tables at the cell quadrature points (what `mhd_tables_t` carries)."""

    nq: int
    xi: np.ndarray  # [nq,3]
    w: np.ndarray  # [nq]
    geo_val: np.ndarray  # [nq,8]   trilinear vertex functions (geometry map)
    geo_grad: np.ndarray  # [nq,8,3]
    nu: np.ndarray  # [nq,27]  scalar Q2 basis (u = nu (x) e_c, dof = a + 27 c)
    dnu: np.ndarray  # [nq,27,3]
    pp: np.ndarray  # [nq,4]   P1disc
    psi: np.ndarray  # [nq,36,3] reference RT1
    dpsi: np.ndarray  # [nq,36]
    chi: np.ndarray  # [nq,8]   Q1disc


def make_tables(qdegree: int = 5) -> Tables:
    xi, w = quadrature_for_degree(qdegree)
    gv, gg = q1_tabulate(xi)
    nu, dnu = q2_tabulate(xi)
    psi, dpsi = rt1_tabulate(xi)
    return Tables(
        nq=len(w), xi=xi, w=w, geo_val=gv, geo_grad=gg, nu=nu, dnu=dnu, pp=p1_tabulate(xi), psi=psi, dpsi=dpsi, chi=gv.copy()
    )


NDOF_U, NDOF_P, NDOF_J, NDOF_PHI = 81, 4, 36, 8
NDOF_CELL = NDOF_U + NDOF_P + NDOF_J + NDOF_PHI  # 129
