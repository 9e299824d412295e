"""ON_BFACES assembly (the best-approximation Dirichlet data of boundarydata.jl:297-347): boundary mass matrix
`DiscreteSymmetricBilinearForm([Identity, Identity], [FE, FE]; AT = ON_BFACES, regions)` and boundary right-hand side
`DiscreteLinearForm([Identity], [FE], fdot_action(data); AT = ON_BFACES, regions)` on the device against the oracle, bit for bit
(forms on boundary-face items run on the bit-exact path), then the best-approximation problem itself."""
import numpy as np
import pytest

import grmp_b200 as G
import oracle as O
from parity import oracle_blf

pytestmark = pytest.mark.gpu


def _grid(dim, level, jitter=False):
    g = G.uniform_refine(G.grid_unitsquare() if dim == 2 else G.grid_unitcube(), level)
    return G.perturb_interior_nodes(g, 0.15) if jitter else g


CASES = [
    ("P1 2D", 2, 3, lambda: G.H1P1(1), [0]),
    ("P1x2 2D regions 1,3", 2, 2, lambda: G.H1P1(2), [1, 3]),
    ("P2 2D", 2, 3, lambda: G.H1P2(1, 2), [0]),
    ("P2x2 2D regions 2,4", 2, 2, lambda: G.H1P2(2, 2), [2, 4]),
    ("P1 3D", 3, 2, lambda: G.H1P1(1), [0]),
    ("P2 3D regions 1,5,6", 3, 1, lambda: G.H1P2(1, 3), [1, 5, 6]),
    ("P2x3 3D", 3, 1, lambda: G.H1P2(3, 3), [0]),
]


@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
@pytest.mark.parametrize("apt", ["symmetric", "general"])
def test_boundary_mass_matrix_bit_equal(case, apt):
    _, dim, level, fef, regions = case
    g = _grid(dim, level)
    s = G.FESpace(fef(), g)
    ctor = G.DiscreteSymmetricBilinearForm if apt == "symmetric" else G.DiscreteBilinearForm
    AP = ctor([G.Identity, G.Identity], [s, s], regions=regions, AT="ON_BFACES")
    cp, rv, nz = G.assemble_csc(AP, 1.25)
    assert AP.AM is not None and G.blf_stats(AP).path == G._lib.PATH_GENERIC
    ocp, orv, onz = oracle_blf(AP, 1.25)
    assert np.array_equal(cp, ocp) and np.array_equal(rv, orv)
    assert np.array_equal(nz, onz), f"max abs diff {np.abs(nz - onz).max():.3e}"
    # size-independent property: 1' M 1 = ncomp * measure of the selected boundary part
    bg = g.bface_grid()
    sel = np.ones(bg.ncells, bool) if regions == [0] else np.isin(bg.cellregions, regions)
    assert abs(nz.sum() - 1.25 * s.fetype.ncomponents * bg.cellvolumes[sel].sum()) < 1e-12


def _xq(bg, qf):
    x = bg.coords
    cn = bg.cellnodes.astype(np.int64) - 1
    xq = np.repeat(x[cn[:, 0]][:, None, :], len(qf), axis=1).copy()
    for j in range(bg.dim):
        xq += (x[cn[:, j + 1]] - x[cn[:, 0]])[:, None, :] * qf.xref[None, :, j, None]
    return xq


@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_boundary_linearform_bit_equal(case):
    _, dim, level, fef, regions = case
    g = _grid(dim, level, jitter=True)
    s = G.FESpace(fef(), g)
    nc = s.fetype.ncomponents
    data = G.DataFunction(lambda x: np.stack([np.cos(x[0] + k) * x[1] + (x[2] ** 2 if len(x) > 2 else 0.5) for k in range(nc)]), [nc, dim],
                          bonus_quadorder=3)
    AP = G.DiscreteLinearForm([G.Identity], [s], G.fdot_action(data), regions=regions, AT="ON_BFACES")
    b = G.FEVector([s])
    b.entries[:] = -0.5
    G.assemble(b[1], AP, factor=0.75)
    bs = s.on_bfaces()
    bg = bs.xgrid
    P = AP.AM
    qo = P.quadorder
    assert qo == s.fetype.polynomialorder(bg.dim) + 3
    table = np.asarray(data.kernel(_xq(bg, P.qf).reshape(-1, dim).T), dtype=np.float64).reshape(nc, -1).T.reshape(bg.ncells, len(P.qf), nc)
    ob = np.full(s.ndofs, -0.5)
    O.qrule_override(bg.dim, qo, P.qf.xref, P.qf.w)
    try:
        O.lf_assemble(ob, bg, bs, O.OP_ID, fsrc=O.F_QP_TABLE, fdata=table, regions=regions, factor=0.75, bonus_quadorder=3)
    finally:
        O.qrule_override(bg.dim, qo)
    assert np.array_equal(b.entries, ob), f"max abs diff {np.abs(b.entries - ob).max():.3e}"
    # interior dofs are untouched
    touched = np.zeros(s.ndofs, bool)
    touched[bs.celldofs.astype(np.int64).ravel() - 1] = True
    assert np.all(b.entries[~touched] == -0.5)


@pytest.mark.parametrize("dim", [2, 3])
def test_best_approximation_reproduces_quadratic_boundary_data(dim):
    """boundarydata.jl:297-347: M_bnd u = b_bnd on the boundary dofs; P2 reproduces a quadratic exactly"""
    import scipy.sparse as sp
    import scipy.sparse.linalg as spla
    g = _grid(dim, 2 if dim == 2 else 1, jitter=True)
    s = G.FESpace(G.H1P2(1, dim), g)
    u = lambda x: 1.0 + x[0] * x[1] - 2.0 * x[dim - 1] ** 2 + 0.5 * x[0]
    data = G.DataFunction(lambda x: np.stack([u(x)]), [1, dim], bonus_quadorder=2)
    A = G.DiscreteSymmetricBilinearForm([G.Identity, G.Identity], [s, s], AT="ON_BFACES")
    cp, rv, nz = G.assemble_csc(A, 1.0)
    b = G.FEVector([s])
    G.assemble(b[1], G.DiscreteLinearForm([G.Identity], [s], G.fdot_action(data), AT="ON_BFACES"))
    M = sp.csc_matrix((nz, rv - 1, cp - 1), shape=(s.ndofs, s.ndofs))
    bd = np.unique(s.bfacedofs.astype(np.int64).ravel() - 1)
    sol = spla.spsolve(M[bd][:, bd].tocsc(), b.entries[bd])
    # nodal values: vertices, then edge midpoints (face dofs in 2D, edge dofs in 3D)
    en = (g.facenodes if dim == 2 else g.edgenodes).astype(np.int64) - 1
    xdof = np.concatenate([g.coords, 0.5 * (g.coords[en[:, 0]] + g.coords[en[:, 1]])])
    exact = u(xdof[bd].T)
    assert np.abs(sol - exact).max() < 1e-11


def test_only_identity_of_h1_spaces_is_admitted():
    g = _grid(2, 1)
    s = G.FESpace(G.H1P1(1), g)
    with pytest.raises(G._lib.GrmpError):
        G.assemble_csc(G.DiscreteSymmetricBilinearForm([G.Gradient, G.Gradient], [s, s], AT="ON_BFACES"))
    with pytest.raises(G._lib.GrmpError):           # Hdiv spaces on boundary faces: NormalFlux only
        G.assemble_csc(G.DiscreteSymmetricBilinearForm([G.Identity, G.Identity], [G.FESpace(G.HDIVRT0(2), g)] * 2, AT="ON_BFACES"))
    with pytest.raises(G._lib.GrmpError):           # NormalFlux lives on boundary faces
        G.assemble_csc(G.DiscreteSymmetricBilinearForm([G.NormalFlux, G.NormalFlux], [G.FESpace(G.HDIVRT0(2), g)] * 2))
    with pytest.raises(NotImplementedError):
        G.assemble_csc(G.DiscreteSymmetricBilinearForm([G.Identity, G.Identity], [G.FESpace(G.H1BR(2), g)] * 2, AT="ON_BFACES"))
    with pytest.raises(NotImplementedError):
        G.DiscreteSymmetricBilinearForm([G.Identity, G.Identity], [s, s], AT="ON_FACES")


@pytest.mark.parametrize("dim", [2, 3])
def test_poisson_with_mixed_dirichlet_data_end_to_end(dim):
    """solve! flow of the reference (solvers.jl:600-668) with every Dirichlet type of boundarydata.jl: assemble on the device,
    boundarydata -> fixed dofs, penalties on the device-resident matrix, host solve, residual on the device.  P2 + quadratic
    solution: every step is exact up to rounding, so the nodal values come back."""
    import scipy.sparse as sp
    import scipy.sparse.linalg as spla
    g = _grid(dim, 2 if dim == 2 else 1, jitter=True)
    s = G.FESpace(G.H1P2(1, dim), g)
    u = lambda x: x[0] ** 2 + x[0] * x[1] - (x[dim - 1] ** 2 if dim == 3 else 0.0) + 1.0        # -Laplace u = -2 (2D), 0 (3D)
    udata = G.DataFunction(lambda x: np.stack([u(x)]), [1, dim], bonus_quadorder=2)
    A = G.DiscreteSymmetricBilinearForm([G.Gradient, G.Gradient], [s, s])
    cp, rv, _ = G.assemble_csc(A, 1.0)
    rhs = G.FEVector([s])
    G.assemble(rhs[1], G.DiscreteLinearForm([G.Identity], [s], G.fdot_action(G.DataFunction([-2.0 if dim == 2 else 0.0]))))
    sol = G.FEVector([s])
    nreg = int(g.bfaceregions.max())
    O = [G.BoundaryData(G.InterpolateDirichletBoundary, data=udata, regions=[1]),
         G.BoundaryData(G.BestapproxDirichletBoundary, data=udata, regions=list(range(2, nreg + 1)))]
    fixed = G.boundarydata(sol[1], O)
    assert np.array_equal(np.sort(fixed), np.unique(s.bfacedofs))
    penalty = 1e60
    G.apply_penalties(A, fixed, penalty)
    rhs.entries[fixed - 1] = penalty * sol.entries[fixed - 1]
    nz = G.fetch_values(A)
    M = sp.csc_matrix((nz, rv - 1, cp - 1), shape=(s.ndofs, s.ndofs))
    x = spla.spsolve(M, rhs.entries)
    en = (g.facenodes if dim == 2 else g.edgenodes).astype(np.int64) - 1
    xdof = np.concatenate([g.coords, 0.5 * (g.coords[en[:, 0]] + g.coords[en[:, 1]])])
    assert np.abs(x - u(xdof.T)).max() < 1e-10
    _, nrm = G.residual(A, x, rhs.entries, fixed_dofs=fixed, want_vector=False)
    assert nrm < 1e-20


def test_homogeneous_boundary_and_order_of_fixed_dofs():
    g = _grid(2, 1)
    s = G.FESpace(G.H1P1(2), g)
    t = G.FEVector([s])
    t.entries[:] = 7.0
    O = [G.BoundaryData(G.HomogeneousDirichletBoundary, regions=[1, 3])]
    fixed = G.boundarydata(t[1], O)
    sel = np.isin(g.bfaceregions, [1, 3])
    expect = s.bfacedofs[sel].astype(np.int64).ravel()
    _, first = np.unique(expect, return_index=True)
    assert np.array_equal(fixed, expect[np.sort(first)])           # Base.unique order (boundarydata.jl:246)
    assert np.all(t.entries[fixed - 1] == 0) and np.count_nonzero(t.entries == 7.0) == s.ndofs - fixed.size


# ---- Hdiv boundary data: NormalFlux of HDIVRT0 / HDIVBDM1 on boundary faces (boundarydata.jl:301-302, 321-323; Example302's space) ----
HDIV_CASES = [("RT0 2D", 2, 3, lambda: G.HDIVRT0(2), [0]), ("BDM1 2D regions 1,2", 2, 2, lambda: G.HDIVBDM1(2), [1, 2]),
              ("RT0 3D regions 3,4", 3, 2, lambda: G.HDIVRT0(3), [3, 4]), ("BDM1 3D", 3, 1, lambda: G.HDIVBDM1(3), [0])]


@pytest.mark.parametrize("case", HDIV_CASES, ids=[c[0] for c in HDIV_CASES])
def test_normalflux_boundary_mass_and_rhs_bit_equal(case):
    _, dim, level, fef, regions = case
    g = _grid(dim, level, jitter=True)
    s = G.FESpace(fef(), g)
    AP = G.DiscreteSymmetricBilinearForm([G.NormalFlux, G.NormalFlux], [s, s], regions=regions, AT="ON_BFACES")
    cp, rv, nz = G.assemble_csc(AP, 0.5)
    ocp, orv, onz = oracle_blf(AP, 0.5)
    assert np.array_equal(cp, ocp) and np.array_equal(rv, orv) and np.array_equal(nz, onz)
    data = G.DataFunction(lambda x: np.stack([np.sin(x[0]) + x[1] * x[k % dim] for k in range(dim)]), [dim, dim], bonus_quadorder=2)
    L = G.DiscreteLinearForm([G.NormalFlux], [s], G.fdotn_action(data, g), regions=regions, AT="ON_BFACES")
    b = G.FEVector([s])
    G.assemble(b[1], L, factor=2.0)
    bs, P = s.on_bfaces(), L.AM
    bg = bs.xgrid
    vals = np.asarray(data.kernel(_xq(bg, P.qf).reshape(-1, dim).T), dtype=np.float64).reshape(dim, -1).T.reshape(bg.ncells, len(P.qf), dim)
    nrm = g.facenormals[g.bfacefaces.astype(np.int64) - 1]
    table = np.ascontiguousarray((vals * nrm[:, None, :]).sum(axis=2)[:, :, None])
    ob = np.zeros(s.ndofs)
    O.qrule_override(bg.dim, P.quadorder, P.qf.xref, P.qf.w)
    try:
        O.lf_assemble(ob, bg, bs, O.OP_NORMALFLUX, fsrc=O.F_QP_TABLE, fdata=table, regions=regions, factor=2.0, bonus_quadorder=2)
    finally:
        O.qrule_override(bg.dim, P.quadorder)
    assert np.array_equal(b.entries, ob)


@pytest.mark.parametrize("dim", [2, 3])
@pytest.mark.parametrize("fam", ["RT0", "BDM1"])
def test_hdiv_best_approximation_boundary_data_equals_face_moments(dim, fam):
    """the face bases are dual to the interpolation functionals (hdiv_rt0.jl:37-52, hdiv_bdm1.jl:43-66): for data whose normal flux lies in the
    trace space the best approximation returns  int_F u.n,  int_F u.n (xref_1 - 1/d),  int_F u.n (xref_2 - 1/d)"""
    g = _grid(dim, 2 if dim == 2 else 1, jitter=True)
    fe = G.HDIVRT0(dim) if fam == "RT0" else G.HDIVBDM1(dim)
    s = G.FESpace(fe, g)
    if fam == "RT0":
        u = lambda x: np.stack([np.full_like(x[0], 0.75 - 0.25 * k) for k in range(dim)])
    else:
        u = lambda x: np.stack([0.5 + x[k] - 2.0 * x[(k + 1) % dim] for k in range(dim)])
    t = G.FEVector([s])
    O_ = [G.BoundaryData(G.BestapproxDirichletBoundary, data=G.DataFunction(u, [dim, dim], bonus_quadorder=1), regions=list(range(1, int(g.bfaceregions.max()) + 1)))]
    fixed = G.boundarydata(t[1], O_)
    bs = s.on_bfaces()
    assert np.array_equal(np.sort(fixed), np.unique(bs.celldofs))
    # moments by quadrature on the faces (face node order = FaceNodes order)
    bg = bs.xgrid
    qf = G.QuadratureRule("Edge1D" if dim == 2 else "Triangle2D", 3)
    xq = _xq(bg, qf)
    nrm = g.facenormals[g.bfacefaces.astype(np.int64) - 1]
    un = (np.moveaxis(u(xq.reshape(-1, dim).T).reshape(dim, bg.ncells, len(qf)), 0, 2) * nrm[:, None, :]).sum(axis=2)       # [face, q]
    vol = bg.cellvolumes
    mom = [vol * (un * qf.w).sum(axis=1)]
    if fam == "BDM1":
        for j in range(dim - 1):
            mom.append(vol * (un * qf.w * (qf.xref[:, j] - 1.0 / dim)).sum(axis=1))
    dofs = bs.celldofs.astype(np.int64) - 1
    for k, m in enumerate(mom):
        assert np.abs(t.entries[dofs[:, k]] - m).max() < 1e-12, (k, np.abs(t.entries[dofs[:, k]] - m).max())


@pytest.mark.parametrize("dim", [2, 3])
def test_boundary_item_integrators(dim):
    """ItemIntegrator / L2ErrorIntegrator with AT = ON_BFACES (itemintegrator.jl:18-21, 33-78): boundary integrals of a P2 function and the
    boundary L2 error, item by item against the oracle (bit-equal) and against the closed form"""
    g = _grid(dim, 2 if dim == 2 else 1, jitter=True)
    s = G.FESpace(G.H1P2(1, dim), g)
    u = lambda x: 1.0 + x[0] * x[dim - 1] + 0.5 * x[1] ** 2
    v = G.FEVector([s])
    en = (g.facenodes if dim == 2 else g.edgenodes).astype(np.int64) - 1
    xdof = np.concatenate([g.coords, 0.5 * (g.coords[en[:, 0]] + g.coords[en[:, 1]])])
    v.entries[:] = u(xdof.T)                                   # nodal interpolation = u itself (P2 reproduces quadratics)
    bs = s.on_bfaces()
    bg = bs.xgrid
    II = G.ItemIntegrator([G.Identity], regions=[1, 2], AT="ON_BFACES")
    b = np.zeros((bg.ncells, 1))
    G.evaluate_itemwise(b, II, v[1])
    ob, _ = O.ii_evaluate(bg, bs, O.OP_ID, v.entries, regions=[1, 2], itemwise=True)
    assert np.array_equal(b, ob)
    assert np.all(b[~np.isin(bg.cellregions, [1, 2])] == 0)
    tot = G.evaluate(II, v[1])
    assert abs(tot - b.sum()) <= 1e-12 * abs(b).sum()
    # closed form on the unit square: regions 1 (y = 0) and 2 (x = 1):  int_0^1 1 dx + int_0^1 (1 + y + y^2 / 2) dy = 1 + 5/3
    if dim == 2:
        assert abs(tot - (1.0 + 1.0 + 0.5 + 1.0 / 6.0)) < 1e-12
    err = G.evaluate(G.L2ErrorIntegrator(G.DataFunction(lambda x: np.stack([u(x)]), [1, dim], bonus_quadorder=2), G.Identity, quadorder=4, AT="ON_BFACES"), v[1])
    assert 0.0 <= err < 1e-24
    nrm = G.evaluate(G.L2NormIntegrator(1, G.Identity, quadorder=4, AT="ON_BFACES"), v[1])
    assert nrm > 1.0


# ---- the reference's own KAT: test/runtests.jl:481-509 "H1-Bestapproximations" ------------------------------------------------------
def _exact(dim, order):
    """exact_function2D / exact_function3D of test/runtests.jl:38-80 and their gradients (row-major: d_k u_c at [c * dim + k])"""
    if dim == 2:
        u = lambda x: np.stack([x[0] ** order + 2 * x[1] ** order + 1, 3 * x[0] ** order - x[1] ** order - 1])
        dp = lambda t: order * t ** (order - 1) if order > 0 else 0.0 * t
        du = lambda x: np.stack([dp(x[0]), 2 * dp(x[1]), 3 * dp(x[0]), -dp(x[1])])
    else:
        u = lambda x: np.stack([2 * x[2] ** order - x[1] ** order - 1, x[0] ** order + 2 * x[1] ** order + 1, 3 * x[0] ** order - x[1] ** order - 1])
        dp = lambda t: order * t ** (order - 1) if order > 0 else 0.0 * t
        z = lambda x: 0.0 * x[0]
        du = lambda x: np.stack([z(x), -dp(x[1]), 2 * dp(x[2]), dp(x[0]), 2 * dp(x[1]), z(x), 3 * dp(x[0]), -dp(x[1]), z(x)])
    return u, du


@pytest.mark.parametrize("dim,fe,order", [(2, "P1", 1), (2, "P2", 2), (3, "P1", 1), (3, "P2", 2)])
def test_reference_kat_h1_bestapproximation_with_bestapprox_boundary(dim, fe, order):
    """H1BestapproximationProblem(grad u, u; bestapprox_boundary_regions = [1, 2]) (pdeprototypes.jl:170-201) on testgrid (runtests.jl:14-19),
    FETypes of TestCatalog3D (H1P1{3}: order 1, H1P2{3,3}: order 2) and their 2D twins: LaplaceOperator + LinearForm(Gradient, grad u) assembled on the
    device, boundary data by boundarydata (ON_BFACES forms on the device), penalties on the device, host solve, L2ErrorIntegrator on the device;
    the reference asserts sqrt(error) < 6e-12 (runtests.jl:23, 503)"""
    import scipy.sparse as sp
    import scipy.sparse.linalg as spla
    g = G.uniform_refine(G.grid_unitsquare() if dim == 2 else G.grid_unitcube(), 1)
    s = G.FESpace(G.H1P1(dim) if fe == "P1" else G.H1P2(dim, dim), g)
    u, du = _exact(dim, order)
    udata = G.DataFunction(u, [dim, dim], bonus_quadorder=order)
    gdata = G.DataFunction(du, [dim * dim, dim], bonus_quadorder=max(order - 1, 0))
    A = G.DiscreteSymmetricBilinearForm([G.Gradient, G.Gradient], [s, s])                    # LaplaceOperator()
    cp, rv, _ = G.assemble_csc(A, 1.0)
    rhs = G.FEVector([s])
    G.assemble(rhs[1], G.DiscreteLinearForm([G.Gradient], [s], G.fdot_action(gdata)))          # LinearForm(Gradient, grad u)
    sol = G.FEVector([s])
    fixed = G.boundarydata(sol[1], [G.BoundaryData(G.BestapproxDirichletBoundary, data=udata, regions=[1, 2])])
    assert fixed.size > 0
    penalty = 1e60                                                                              # solvers.jl:632-652
    G.apply_penalties(A, fixed, penalty)
    rhs.entries[fixed - 1] = penalty * sol.entries[fixed - 1]
    M = sp.csc_matrix((G.fetch_values(A), rv - 1, cp - 1), shape=(s.ndofs, s.ndofs))
    sol.entries[:] = spla.spsolve(M, rhs.entries)
    err2 = G.evaluate(G.L2ErrorIntegrator(udata, G.Identity, quadorder=order), sol[1])
    assert np.all(np.asarray(err2) >= -1e-30)
    assert np.sqrt(np.abs(np.asarray(err2)).sum()) < 6e-12
