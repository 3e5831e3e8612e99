"""Pins the CPU oracle (CPU-only tests).

The reference's own tests hold no numeric assertions for this path, so the oracle is pinned with
  K1  exact DOF counts published with the reference's Gadi runs (analysis/gadi/results/2023_04/**/summary.csv),
  K5  the 16-digit discrete-solution norms of the Hunt benchmark from the same CSVs (solution-level pin),
  K2  manufactured in-space fields of src/Applications/transient.jl:262-270 (residual == 0),
  K3  Jacobian == finite-difference of the residual (Newton convection),
  K4  block identities of the weak form (src/weakforms.jl:293-311),
  and the independent C restatement (oracle/mhd_oracle.c) against the NumPy one.
"""
import numpy as np
import pytest

from gridapmhd_jl_b200.host import fespaces as F
from gridapmhd_jl_b200.host import mesh as M
from gridapmhd_jl_b200.host import reffe
from oracle import mhd_oracle as O


def hunt_fes(nc, Ha, **kw):
    m = M.hunt_generate_base_mesh((nc, nc), Ha=Ha, **kw)
    return F.setup_fe_spaces(m)


# K1 ------------------------------------------------------------------------------------------
@pytest.mark.parametrize(
    "nc,expected",
    [
        (4, dict(u=882, p=192, j=1152, phi=384)),  # SURVEY.md section 8, cfg1 (2 610 dofs)
        (10, dict(u=6498, p=1200, j=7200, phi=2400)),  # hconv_ha00050ns500/summary.csv:7 (17 298 dofs)
        (64, dict(u=290322, p=49152, j=294912, phi=98304)),  # weak_ls16ns200/summary.csv:4 (732 690 dofs)
    ],
)
def test_k1_dof_counts_match_published(nc, expected):
    fes = hunt_fes(nc, 50.0)
    assert fes.nfree == expected
    assert fes.mesh.ncells == 3 * nc * nc


def test_k1_nnz_counts_cfg1():
    fes = hunt_fes(4, 10.0)
    rp, cv = O.symbolic_csr(fes.cell_global_ids(), fes.ndofs)
    assert len(cv) == 381336  # SURVEY.md section 8 [calc]
    assert np.all(np.diff(rp) > 0)
    # sorted, unique columns inside each row
    for i in (0, 17, 2609):
        c = cv[rp[i] : rp[i + 1]]
        assert np.all(np.diff(c) > 0)


# K5 ------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def hunt_ha50_nc10_solution():
    """Discrete solution of the published run ha00050cx010 (nc=10, Ha=50, kmap=1 = unstretched mesh), sparse LU Newton."""
    Ha = 50.0
    fes = hunt_fes(10, Ha, BL_adapted=False)
    prm = O.FluidParams(alpha=1.0, beta=1.0, gamma=Ha**2, sigma=1.0, B=(0, 1, 0), f=(0, 0, 1), convection="newton")
    x, hist = O.newton_lu(fes, prm)
    assert hist[-1] < 1e-10 * hist[0]
    return fes, x, Ha


def test_k5_hunt_ha50_nc10_published_norms(hunt_ha50_nc10_solution):
    """analysis/gadi/results/2023_04/eaab9d14.../hconv_ha00050ns500/summary.csv:7 (nc=10, Ha=50): uh_l2, uh_h1, jh_l2.
    The 2023 runs used the kmap=1 power map, i.e. the unstretched mesh (BL_adapted=False, kmap=1 here)."""
    fes, x, Ha = hunt_ha50_nc10_solution
    nr = O.solution_norms(fes, x, reffe.make_tables(6), u0=1.0, jscale=Ha)  # jh = sigma*u0*B0*jbar (hunt.jl:217)
    pins = dict(uh_l2=0.001125968494949451, uh_h1=0.009386206346670825, jh_l2=0.019669872964491745)
    for k, v in pins.items():
        assert abs(nr[k] - v) / v < 1e-9, (k, nr[k], v)


def test_k6_hunt_error_norms_against_the_analytical_solution(hunt_ha50_nc10_solution):
    """Same published row: the errors against the analytical Hunt series (hunt.jl:239-256,372-457) with nsums = 500
    (eu_l2, eu_h1, ej_l2) and 2*nsums = 1000 terms (the *_ref columns), degree-6 quadrature.  Reproduced to 1e-9
    relative (observed 4e-13): pins the mesh, the spaces, the solve, the RT/Piola evaluation and the series at once."""
    fes, x, Ha = hunt_ha50_nc10_solution
    T6 = reffe.make_tables(6)
    pins = {500: dict(eu_l2=6.274034420523594e-5, eu_h1=0.0021489743289400043, ej_l2=0.0011055397108926523),
            1000: dict(eu_l2=6.273968018318573e-5, eu_h1=0.002148559357584549, ej_l2=0.0011055397108926512)}
    for nsums, pin in pins.items():
        e = O.hunt_error_norms(fes, x, T6, Ha, nsums, u0=1.0, jscale=Ha)
        for k, v in pin.items():
            assert abs(e[k] - v) / v < 1e-9, (nsums, k, e[k], v)


@pytest.mark.slow
def test_k6_second_published_row_nc18():
    """hconv_ha00050ns500/summary.csv, run ha00050cx018 (nc=18, 972 cells, 57 042 dofs): all six norms (~100 s of sparse LU)."""
    Ha = 50.0
    fes = hunt_fes(18, Ha, BL_adapted=False)
    assert fes.ndofs == 57042
    prm = O.FluidParams(alpha=1.0, beta=1.0, gamma=Ha**2, sigma=1.0, B=(0, 1, 0), f=(0, 0, 1), convection="newton")
    x, hist = O.newton_lu(fes, prm)
    T6 = reffe.make_tables(6)
    got = O.hunt_error_norms(fes, x, T6, Ha, 500, u0=1.0, jscale=Ha)
    got.update(O.solution_norms(fes, x, T6, u0=1.0, jscale=Ha))
    pins = dict(eu_l2=1.4099417236975484e-5, eu_h1=0.0008916378834878792, ej_l2=0.0006661061529286713,
                uh_l2=0.0011232728005488165, uh_h1=0.009581884653926384, jh_l2=0.019654422812794527)
    for k, v in pins.items():
        assert abs(got[k] - v) / v < 1e-9, (k, got[k], v)


# K2 ------------------------------------------------------------------------------------------
def _interpolate(fes, ufun, jconst, pconst, phiconst):
    """Interpolant of u (nodal), constant j (RT moments of a constant field), constant p and phi."""
    x = np.zeros(fes.ndofs)
    off = fes.offsets
    # u: nodal values
    ids = fes.cell_dofs["u"]
    vals = ufun(fes.u_node_coords.reshape(-1, 3)).reshape(fes.mesh.ncells, 27, 3)
    for c in range(3):
        idc = ids[:, 27 * c : 27 * (c + 1)]
        free = idc > 0
        x[off["u"] + idc[free] - 1] = vals[:, :, c][free]
    x[off["p"] : off["p"] + fes.nfree["p"]] = pconst
    x[off["phi"] : off["phi"] + fes.nfree["phi"]] = phiconst
    # j: L2-project the constant field cell by cell (exactly representable on affine cells)
    T = fes.tables
    X = fes.mesh.cell_coords()
    w, _, psi, _ = O.mapped_bases(T, X, fes.j_sign)
    Mjj = np.einsum("cq,cqmi,cqni->cmn", w, psi, psi)
    rhs = np.einsum("cq,cqmi,i->cm", w, psi, np.asarray(jconst, float))
    coef = np.linalg.solve(Mjj, rhs[..., None])[..., 0]
    idj = fes.cell_dofs["j"]
    free = idj > 0
    x[off["j"] + idj[free] - 1] = coef[free]
    return x, coef


def test_k2_manufactured_inspace_fields_zero_residual():
    """In-space manufactured solution in the spirit of src/Applications/transient.jl:262-270,326-344, restricted to
    constant forcing (the C ABI carries constant f, g): u = u0 (constant), j = j0 (constant), p = a.x, phi = b.x with
        f = grad p - gamma j0 x B           (momentum, weakforms.jl:280: -p div v - gamma (j x B).v - f.v)
        g = j0 + sigma grad phi - sigma u0 x B   (Ohm,  weakforms.jl:280: j.s - sigma phi div s - sigma (u x B).s - g.s)
    u and j.n are imposed strongly on the whole boundary, so no face term is needed.  All four residual blocks
    must vanish at the free rows; exercises both signs of the Lorentz coupling."""
    m = M.hunt_generate_base_mesh((3, 2), Ha=10.0, BL_adapted=False, periodic_z=False)
    M.tag_from_boundary_faces(m, "allwalls", m.face_ncells == 1)
    u0 = np.array([0.3, -0.2, 0.5])
    j0 = np.array([0.7, 0.1, -0.4])
    a = np.array([0.2, -0.6, 0.9])
    b = np.array([-0.5, 0.4, 0.3])
    B = np.array([0.3, 1.0, -0.2])
    alpha, beta, gamma, sigma = 0.5, 0.3, 7.0, 1.9
    uex = lambda X: np.tile(u0, (len(X), 1))
    fes = F.setup_fe_spaces(m, u_tags=("allwalls",), u_values=(uex,), j_tags=("allwalls",))
    f = a - gamma * np.cross(j0, B)
    g = j0 + sigma * b - sigma * np.cross(u0, B)
    X = fes.mesh.cell_coords()
    x, coef = _interpolate(fes, uex, j0, 0.0, 0.0)
    off = fes.offsets
    # p: nodal values at the reference simplex vertices (vertex 0, 1, 2, 4 of the hex); phi: at the 8 vertices
    pv = np.einsum("cvi,i->cv", X[:, [0, 1, 2, 4], :], a)
    x[off["p"] : off["p"] + fes.nfree["p"]] = pv.reshape(-1)
    fv = np.einsum("cvi,i->cv", X, b)
    x[off["phi"] : off["phi"] + fes.nfree["phi"]] = fv.reshape(-1)
    idj = fes.cell_dofs["j"]
    dirv = np.zeros(fes.ndir["j"])
    dmask = idj < 0
    dirv[-idj[dmask] - 1] = coef[dmask]
    fes.dirichlet_values["j"] = dirv
    for conv in ("none", "newton"):  # (u0.grad)u0 = 0 for a constant field
        prm = O.FluidParams(alpha=alpha, beta=beta, gamma=gamma, sigma=sigma, zeta_u=3.0, zeta_j=2.0, B=tuple(B), f=tuple(f),
                            g=tuple(g), convection=conv)
        r = fes.split(O.residual(fes, x, prm))
        for k in ("u", "p", "j", "phi"):
            assert np.abs(r[k]).max() < 1e-10, (conv, k, np.abs(r[k]).max())
    # and the residual is NOT zero if a sign of the coupling is flipped (guards against a vacuous test)
    bad = O.FluidParams(alpha=alpha, beta=beta, gamma=gamma, sigma=sigma, B=tuple(B), f=tuple(a + gamma * np.cross(j0, B)),
                        g=tuple(g), convection="none")
    assert np.abs(fes.split(O.residual(fes, x, bad))["u"]).max() > 1e-3


def test_k2b_reference_manufactured_solution_with_nonconstant_velocity():
    """The reference's own in-FE-space manufactured solution (src/Applications/transient.jl:262-270, `:stationary_fespace`):
        u = (y, x, 0),  j = (y, 1, 0),  p = 0,  phi = 1
    with the forcing of `_transient_solution_f` / `_transient_solution_fj` (:331-344)
        f(x) = alpha (u.grad)u - beta lap u + grad p - gamma j x B = alpha (x, y, 0) - gamma j x B,
        g(x) = j + sigma grad phi - sigma u x B.
    (u.grad)u = (x, y, 0) is NOT zero here, so this pins the convection term `alpha v.((grad u)' u)` of
    res_fluid_h1_hdiv (weakforms.jl:270, `conv` :670) independently of Hunt (fully developed: (u.grad)u = 0) and of the
    Jacobian-vs-residual check K3.  For that field grad u is symmetric (a transposed gradient would give the same vector), so a
    second divergence-free field with a non-symmetric gradient, u = (y, x^2, 0), is checked as well."""
    m = M.hunt_generate_base_mesh((3, 2), Ha=10.0, BL_adapted=False, periodic_z=False)
    M.tag_from_boundary_faces(m, "allwalls", m.face_ncells == 1)
    B = np.array([0.3, 1.0, -0.2])
    alpha, beta, gamma, sigma = 0.5, 0.3, 7.0, 1.9
    cases = {
        # name: (u, (u.grad)u with Gridap's convention sum_i u_i d_i u_c, laplacian of u)
        "reference": (lambda X: np.stack([X[:, 1], X[:, 0], 0 * X[:, 0]], axis=1),
                      lambda X: np.stack([X[..., 0], X[..., 1], 0 * X[..., 0]], axis=-1),
                      lambda X: 0 * X),
        # u = (y, x^2, 0): div u = 0, grad u not symmetric: (u.grad)u = (u_y d_y u_x, u_x d_x u_y, 0) = (x^2, 2 x y, 0), lap u = (0, 2, 0)
        "nonsymmetric": (lambda X: np.stack([X[:, 1], X[:, 0] ** 2, 0 * X[:, 0]], axis=1),
                         lambda X: np.stack([X[..., 0] ** 2, 2 * X[..., 0] * X[..., 1], 0 * X[..., 0]], axis=-1),
                         lambda X: np.stack([0 * X[..., 0], 2 + 0 * X[..., 0], 0 * X[..., 0]], axis=-1)),
    }
    jfun = lambda X: np.stack([X[..., 1], 1 + 0 * X[..., 0], 0 * X[..., 0]], axis=-1)
    for name, (uex, convex, lapex) in cases.items():
        fes = F.setup_fe_spaces(m, u_tags=("allwalls",), u_values=(uex,), j_tags=("allwalls",))
        ufield = lambda X: uex(X.reshape(-1, 3)).reshape(X.shape)
        f = lambda X: alpha * convex(X) - beta * lapex(X) - gamma * np.cross(jfun(X), B)
        g = lambda X: jfun(X) - sigma * np.cross(ufield(X), B)
        # interpolate: u nodal, p = 0, phi = 1, j = cell-wise L2 projection of the (in-space) field (y, 1, 0)
        x, _ = _interpolate(fes, uex, (0.0, 0.0, 0.0), 0.0, 1.0)
        T = fes.tables
        X = fes.mesh.cell_coords()
        w, _, psi, _ = O.mapped_bases(T, X, fes.j_sign)
        xq = np.einsum("qv,cvi->cqi", T.geo_val, X)
        Mjj = np.einsum("cq,cqmi,cqni->cmn", w, psi, psi)
        rhs = np.einsum("cq,cqmi,cqi->cm", w, psi, jfun(xq))
        coef = np.linalg.solve(Mjj, rhs[..., None])[..., 0]
        idj = fes.cell_dofs["j"]
        off = fes.offsets
        free = idj > 0
        x[off["j"] + idj[free] - 1] = coef[free]
        dirv = np.zeros(fes.ndir["j"])
        dirv[-idj[idj < 0] - 1] = coef[idj < 0]
        fes.dirichlet_values["j"] = dirv
        prm = O.FluidParams(alpha=alpha, beta=beta, gamma=gamma, sigma=sigma, zeta_u=3.0, zeta_j=2.0, B=tuple(B), f=f, g=g, convection="newton")
        r = fes.split(O.residual(fes, x, prm))
        for k in ("u", "p", "j", "phi"):
            assert np.abs(r[k]).max() < 1e-10, (name, k, np.abs(r[k]).max())
        # the test is not vacuous: without the convection term, or with the transposed gradient, the momentum rows do not vanish
        off_prm = O.FluidParams(alpha=alpha, beta=beta, gamma=gamma, sigma=sigma, B=tuple(B), f=f, g=g, convection="none")
        assert np.abs(fes.split(O.residual(fes, x, off_prm))["u"]).max() > 1e-3, name
    # the transposed convention (grad u)u instead of (grad u)'u differs for the non-symmetric field
    uex, convex, _ = cases["nonsymmetric"]
    Xs = np.array([[0.3, 0.7, 0.1]])
    gu = np.array([[0.0, 2 * Xs[0, 0], 0.0], [1.0, 0.0, 0.0], [0.0, 0.0, 0.0]])  # gu[i, c] = d_i u_c at Xs
    u = uex(Xs)[0]
    assert np.allclose(gu.T @ u, convex(Xs)[0]) and not np.allclose(gu @ u, convex(Xs)[0])


# K3 ------------------------------------------------------------------------------------------
@pytest.mark.parametrize("conv", ["newton", "none"])
def test_k3_jacobian_is_derivative_of_residual(conv):
    fes = hunt_fes(2, 10.0)
    prm = O.FluidParams(alpha=0.8, beta=0.6, gamma=30.0, sigma=1.2, zeta_u=2.0, zeta_j=1.5, B=(0.2, 1.0, -0.1), f=(0.1, 0.2, 1.0),
                        g=(0.3, 0.0, 0.1), convection=conv)
    rng = np.random.default_rng(0)
    x = rng.random(fes.ndofs)
    A = O.jacobian(fes, x, prm)
    d = rng.standard_normal(fes.ndofs)
    eps = 1e-6
    fd = (O.residual(fes, x + eps * d, prm) - O.residual(fes, x - eps * d, prm)) / (2 * eps)
    assert np.abs(fd - A @ d).max() / np.abs(A @ d).max() < 1e-7


def test_k3_picard_jacobian_drops_the_newton_term():
    fes = hunt_fes(2, 10.0)
    x = np.random.default_rng(1).random(fes.ndofs)
    base = dict(alpha=1.0, beta=1.0, gamma=10.0, B=(0, 1, 0))
    An = O.jacobian(fes, x, O.FluidParams(convection="newton", **base))
    Ap = O.jacobian(fes, x, O.FluidParams(convection="picard", **base))
    A0 = O.jacobian(fes, x, O.FluidParams(convection="none", **base))
    assert np.array_equal(An.indices, Ap.indices) and np.array_equal(An.indices, A0.indices)  # same pattern, zeros kept
    assert abs(An - Ap).max() > 1e-3 and abs(Ap - A0).max() > 1e-3


# K4 ------------------------------------------------------------------------------------------
def test_k4_block_identities():
    """K_pu = K_up^T, K_phi_j = K_j_phi^T / sigma, K_ju = -(sigma/gamma) K_uj^T, jj SPD, constant pressure in the kernel."""
    fes = hunt_fes(3, 20.0)
    prm = O.FluidParams(alpha=1.0, beta=1.0, gamma=400.0, sigma=1.7, zeta_j=0.5, B=(0, 1, 0), convection="none")
    A = O.jacobian(fes, np.zeros(fes.ndofs), prm).tocsr()
    off = fes.offsets
    sl = {f: slice(off[f], off[f] + fes.nfree[f]) for f in off}
    blk = lambda r, c: A[sl[r], sl[c]].toarray()
    assert np.abs(blk("p", "u") - blk("u", "p").T).max() < 1e-14
    assert np.abs(blk("phi", "j") - blk("j", "phi").T / prm.sigma).max() < 1e-13
    assert np.abs(blk("j", "u") + (prm.sigma / prm.gamma) * blk("u", "j").T).max() < 1e-12
    jj = blk("j", "j")
    assert np.abs(jj - jj.T).max() < 1e-13 * np.abs(jj).max() and np.linalg.eigvalsh(jj).min() > 0
    # Hunt: u=0 on all walls and z periodic => K_up * 1 = 0 (constant-pressure null mode)
    assert np.abs(blk("u", "p") @ np.ones(fes.nfree["p"])).max() < 1e-13
    for r, c in (("u", "phi"), ("p", "p"), ("p", "j"), ("p", "phi"), ("j", "p"), ("phi", "u"), ("phi", "p"), ("phi", "phi")):
        assert A[sl[r], sl[c]].nnz == 0  # untouched blocks are never inserted


def test_explicit_zeros_are_kept():
    fes = hunt_fes(2, 10.0)
    A = O.jacobian(fes, np.zeros(fes.ndofs), O.FluidParams(convection="none"))
    assert (A.data == 0.0).sum() > 0  # uu off-diagonal component blocks are structural entries with value 0


# C restatement ---------------------------------------------------------------------------------
@pytest.mark.parametrize("conv,zu,zj", [("newton", 0.0, 0.0), ("none", 10.0, 5.0), ("picard", 3.0, 0.0)])
def test_c_oracle_matches_numpy_oracle(conv, zu, zj):
    from oracle.c_oracle import COracle

    fes = hunt_fes(3, 10.0)
    prm = O.FluidParams(alpha=0.7, beta=0.9, gamma=100.0, sigma=1.3, zeta_u=zu, zeta_j=zj, B=(0.1, 1, 0.2), f=(0.3, 0.1, 1),
                        g=(0.1, 0.2, 0.3), convection=conv)
    x = np.random.default_rng(0).random(fes.ndofs)
    A = O.jacobian(fes, x, prm)
    co = COracle(fes, prm)
    nz = co.jacobian_values(x, A.indptr, A.indices)
    assert np.abs(nz - A.data).max() / np.abs(A.data).max() < 1e-13
    assert np.abs(co.residual(x) - O.residual(fes, x, prm)).max() < 1e-12 * np.abs(O.residual(fes, x, prm)).max()
    v = np.random.default_rng(1).standard_normal(fes.ndofs)
    assert np.abs(co.spmv(A.indptr.astype(np.int64), A.indices.astype(np.int64), A.data, v) - A @ v).max() < 1e-11 * np.abs(A @ v).max()


def test_fgmres_oracle_converges_like_direct_solve():
    fes = hunt_fes(2, 10.0)
    prm = O.FluidParams(alpha=1.0, beta=1.0, gamma=100.0, B=(0, 1, 0), f=(0, 0, 1), convection="none")
    A = O.jacobian(fes, np.zeros(fes.ndofs), prm)
    b = -O.residual(fes, np.zeros(fes.ndofs), prm)
    import scipy.sparse.linalg as spla

    lu = spla.splu(A.tocsc())
    xd = lu.solve(b)
    x, it, hist = O.fgmres(A, b, M=lu.solve, m=5, maxiter=5, rtol=1e-10, atol=0.0)
    # the Hunt Jacobian is singular (constant-pressure mode): compare u and j, and the residual
    su, sd = fes.split(x), fes.split(xd)
    assert it <= 2
    assert np.abs(su["u"] - sd["u"]).max() < 1e-8 * np.abs(sd["u"]).max()
    assert np.abs(su["j"] - sd["j"]).max() < 1e-8 * np.abs(sd["j"]).max()
    assert np.linalg.norm(A @ x - b) < 1e-8 * np.linalg.norm(b)


def test_solid_subdomain_oracle_fd_and_structure():
    """Solid walls (weakforms.jl:314-338): u/p rows exist only for fluid cells, the phi-row sign flips on solid cells,
    per-cell sigma; the oracle Jacobian is the derivative of its residual."""
    from gridapmhd_jl_b200.applications import hunt_params, setup_spaces

    params = hunt_params(nc=(12, 12), B=(0.0, 20.0, 0.0), tw=0.2, BL_adapted=False, kmap_x=3, kmap_y=3, zeta_j=2.0)
    fes = setup_spaces(params)
    fl = params["fluid"]
    prm = O.FluidParams(fl.alpha, fl.beta, fl.gamma, fl.sigma, fl.zeta_u, fl.zeta_j, fl.B, fl.f, fl.g, fl.convection)
    assert fes.cell_solid.sum() == 132 and set(np.unique(fes.cell_sigma)) == {0.1, 1.0, 10.0}
    assert (fes.cell_dofs["u"][fes.cell_solid] == 0).all() and (fes.cell_dofs["p"][fes.cell_solid] == 0).all()
    assert fes.nfree["p"] == 4 * int((~fes.cell_solid).sum())
    rng = np.random.default_rng(0)
    x, d = rng.random(fes.ndofs), rng.standard_normal(fes.ndofs)
    A = O.jacobian(fes, x, prm)
    eps = 1e-6
    fd = (O.residual(fes, x + eps * d, prm) - O.residual(fes, x - eps * d, prm)) / (2 * eps)
    assert np.abs(fd - A @ d).max() / np.abs(A @ d).max() < 1e-7
    # phi-j block: K_phij = +K_jphi^T/sigma on fluid cells (both carry a minus sign), -K_jphi^T/sigma_c on solid cells
    off = fes.offsets
    K = O.cell_jacobians(fes.tables, fes.mesh.cell_coords(), fes.cell_state(x), fes.j_sign, prm, fes.cell_solid, fes.cell_sigma)
    c_s, c_f = int(np.nonzero(fes.cell_solid)[0][0]), int(np.nonzero(~fes.cell_solid)[0][0])
    assert np.allclose(K[c_s, 121:129, 85:121], -K[c_s, 85:121, 121:129].T / fes.cell_sigma[c_s])
    assert np.allclose(K[c_f, 121:129, 85:121], +K[c_f, 85:121, 121:129].T / prm.sigma)
    assert np.abs(K[c_s, 121:129, 85:121]).max() > 0
