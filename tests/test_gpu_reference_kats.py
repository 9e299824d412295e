"""The reference's own solve-level tests (test/runtests.jl) replayed through the device path, under the reference's tolerance (6e-12, runtests.jl:23):
"L2-Bestapproximations" (runtests.jl:355-447) for every FEType of its catalogue that is on the path.  Assembly (AUTO path: column kernels, LinearForm
gather kernels) and the L2ErrorIntegrator run on the device, the solve on the host like in the reference.  The H1 twin with best-approximation boundary data
is tests/test_gpu_bfaces.py::test_reference_kat_h1_bestapproximation_with_bestapprox_boundary; the oracle-only twins are in tests/test_oracle_kat.py."""
import numpy as np
import pytest
import scipy.sparse as sp
import scipy.sparse.linalg as spla

import grmp_b200 as G
from test_oracle_kat import L2_CATALOG, catalog_fetype, exact_function

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("dim,name,order", L2_CATALOG, ids=["%s{%d} order %d" % (n, d, o) for d, n, o in L2_CATALOG])
def test_reference_l2_bestapproximation_on_device(dim, name, order):
    g = G.uniform_refine(G.grid_unitsquare() if dim == 2 else G.grid_unitcube(), 1)
    s = G.FESpace(catalog_fetype(name, dim), g)
    udata = G.DataFunction(exact_function(dim, order), [dim, dim], bonus_quadorder=order)
    A = G.DiscreteBilinearForm([G.Identity, G.Identity], [s, s])                                   # ReactionOperator()
    cp, rv, nz = G.assemble_csc(A, 1.0)
    rhs = G.FEVector([s])
    G.assemble(rhs[1], G.DiscreteLinearForm([G.Identity], [s], G.fdot_action(udata)))              # LinearForm(Identity, u)
    sol = G.FEVector([s])
    sol.entries[:] = spla.spsolve(sp.csc_matrix((nz, rv - 1, cp - 1), shape=(s.ndofs, s.ndofs)), rhs.entries)
    err2 = G.evaluate(G.L2ErrorIntegrator(udata, G.Identity, quadorder=order), sol[1])
    assert np.sqrt(np.abs(np.asarray(err2)).sum()) < 6e-12


@pytest.mark.parametrize("dim", [2, 3])
def test_reference_stokes_taylor_hood_on_device(dim):
    """"Stokes-FEM" (runtests.jl:606-672) for the Taylor-Hood pairs of its catalogues: [H1P2{2,2}, H1P1{1}] on Triangle2D and [H1P2{3,3}, H1P1{1}] on
    Tetrahedron3D, orders (2, 1): IncompressibleNavierStokesProblem(dim; nonlinear = false) = LaplaceOperator + LagrangeMultiplier(Divergence) with the
    transposed block (pdeprototypes.jl), BestapproxDirichletBoundary for the velocity on all boundary regions, LinearForm(Identity, rhs), pressure
    with zero integral mean; then max(errorV, errorP) < tolerance.  Operators, boundary forms and error integrators run on the device."""
    g = G.uniform_refine(G.grid_unitsquare() if dim == 2 else G.grid_unitcube(), 1)
    sv, sq = G.FESpace(G.H1P2(dim, dim), g), G.FESpace(G.H1P1(1), g)
    ov, op = 2, 1
    if dim == 2:       # exact_functions_stokes2D (runtests.jl:522-545)
        u = lambda x: np.stack([x[1] ** ov + 1, x[0] ** ov - 1])
        p = lambda x: np.stack([x[0] ** op + x[1] ** op - 2.0 / (op + 1)])
        f = lambda x: np.stack([-ov * (ov - 1) * x[1] ** (ov - 2) + op * x[0] ** (op - 1), -ov * (ov - 1) * x[0] ** (ov - 2) + op * x[1] ** (op - 1)])
    else:              # exact_functions_stokes3D (runtests.jl:547-576)
        u = lambda x: np.stack([x[2] ** ov + 1, x[0] ** ov - 1, x[1] ** ov])
        p = lambda x: np.stack([x[0] ** op + x[1] ** op + x[2] ** op - 3.0 / (op + 1)])
        f = lambda x: np.stack([-ov * (ov - 1) * x[2] ** (ov - 2) + op * x[0] ** (op - 1), -ov * (ov - 1) * x[0] ** (ov - 2) + op * x[1] ** (op - 1),
                                -ov * (ov - 1) * x[1] ** (ov - 2) + op * x[2] ** (op - 1)])
    udata = G.DataFunction(u, [dim, dim], bonus_quadorder=ov)
    pdata = G.DataFunction(p, [1, dim], bonus_quadorder=op)
    fdata = G.DataFunction(f, [dim, dim], bonus_quadorder=max(0, op - 1))
    A = G.FEMatrix([sv, sq])
    G.assemble_operator(A[1, 1], G.LaplaceOperator(1.0))
    G.assemble_operator(A[1, 2], G.LagrangeMultiplier(G.Divergence), At=A[2, 1])
    rhs = G.FEVector([sv, sq])
    G.assemble_operator(rhs[1], G.LinearForm(G.Identity, fdata))
    sol = G.FEVector([sv, sq])
    fixed = G.boundarydata(sol[1], [G.BoundaryData(G.BestapproxDirichletBoundary, data=udata, regions=list(range(1, int(g.bfaceregions.max()) + 1)))])
    assert np.array_equal(np.sort(fixed), np.unique(sv.bfacedofs))
    # penalties for the velocity boundary dofs and for one pressure dof (FixedIntegralMean: fix, solve, shift; globalconstraints.jl)
    M = A.tocsc().tolil()
    b = rhs.entries.copy()
    penalty = 1e60
    for j in fixed - 1:
        M[j, j] = penalty
        b[j] = penalty * sol.entries[j]
    jp = sv.ndofs
    M[jp, jp] = penalty
    b[jp] = 0.0
    # the penalised rows are scaled back to O(1) before the factorisation: the reference's `\` (UMFPACK) equilibrates the 1e60 rows of this
    # indefinite system, SuperLU's default does not (same equations, same solution)
    d = np.ones(M.shape[0])
    d[fixed - 1] = 1.0 / penalty
    d[jp] = 1.0 / penalty
    sol.entries[:] = spla.spsolve((sp.diags(d) @ M.tocsr()).tocsc(), d * b)
    mean = G.evaluate(G.ItemIntegrator([G.Identity]), sol[2]) / g.cellvolumes.sum()
    sol.entries[sv.ndofs:] -= mean
    errV = np.sqrt(np.abs(np.asarray(G.evaluate(G.L2ErrorIntegrator(udata, G.Identity, quadorder=ov), sol[1]))).sum())
    errP = np.sqrt(abs(G.evaluate(G.L2ErrorIntegrator(pdata, G.Identity, quadorder=op), sol[2])))
    assert max(errV, errP) < 6e-12, (errV, errP)


from test_oracle_kat import RECON_CATALOG, br_boundary_values, stokes_exact  # noqa: E402


@pytest.mark.parametrize("dim,recon,ov,op", RECON_CATALOG, ids=["BR{%d} x P0 R=%s orders %d,%d" % (d, r, a, b) for d, r, a, b in RECON_CATALOG])
def test_reference_pressure_robust_stokes_on_device(dim, recon, ov, op):
    """"Reconstruction-Operators" (runtests.jl:674-723): Bernardi-Raugel x P0 Stokes with the right-hand side tested by R v (R: BR -> RT0 | BDM1) and a cubic
    pressure: the discrete velocity is exact (errorV < tolerance on R u_h).  Laplace and divergence blocks by the column kernels, the reconstructed
    LinearForm by the gather kernels (2D) / the bit-exact path (3D), the L2ErrorIntegrator with the reconstruction operator on the device."""
    g = G.uniform_refine(G.grid_unitsquare() if dim == 2 else G.grid_unitcube(), 1)
    sv, sq = G.FESpace(G.H1BR(dim), g), G.FESpace(G.L2P0(1), g)
    u, p, f = stokes_exact(dim, ov, op)
    R = G.ReconstructionIdentity(G.HDIVRT0(dim) if recon == "RT0" else G.HDIVBDM1(dim))
    udata = G.DataFunction(u, [dim, dim], bonus_quadorder=ov)
    fdata = G.DataFunction(f, [dim, dim], bonus_quadorder=max(0, op - 1))
    A = G.FEMatrix([sv, sq])
    G.assemble_operator(A[1, 1], G.LaplaceOperator(1.0))
    G.assemble_operator(A[1, 2], G.LagrangeMultiplier(G.Divergence), At=A[2, 1])
    rhs = G.FEVector([sv, sq])
    G.assemble_operator(rhs[1], G.LinearForm(R, fdata))
    fixed, target = br_boundary_values(sv, u)
    n = sv.ndofs
    M = A.tocsc().tolil()
    b = rhs.entries.copy()
    penalty = 1e60
    d = np.ones(M.shape[0])
    for j in list(fixed) + [n]:
        M[j, j] = penalty
        b[j] = penalty * (target[j] if j < n else 0.0)
        d[j] = 1.0 / penalty
    sol = G.FEVector([sv, sq])
    sol.entries[:] = spla.spsolve((sp.diags(d) @ M.tocsr()).tocsc(), d * b)
    err2 = G.evaluate(G.L2ErrorIntegrator(udata, R, quadorder=ov), sol[1])
    assert np.sqrt(np.abs(np.asarray(err2)).sum()) < 6e-12


@pytest.mark.parametrize("recon", ["RT0", "BDM1"])
def test_example222_hydrostatic_problem_on_device(recon):
    """Example222_PressureRobustness2D.test() (runtests.jl:829-831: < 1e-14) = BASELINE configuration C4 on the device: Bernardi-Raugel Laplacian and
    divergence block by the column kernels, LinearForm(ReconstructionIdentity{RT0 | BDM1}, grad p) with the 9-point Stroud rule by the gather kernel,
    L2ErrorIntegrator on the device; u = 0, p = x^3 + y^3 - 1/2.  The pressure-robust velocity vanishes to rounding, the classical one does not."""
    g = G.uniform_refine(G.grid_unitsquare(), 2)
    sv, sq = G.FESpace(G.H1BR(2), g), G.FESpace(G.L2P0(1), g)
    fdata = G.DataFunction(lambda x: np.stack([3 * x[0] ** 2, 3 * x[1] ** 2]), [2, 2], bonus_quadorder=2)
    zero = G.DataFunction([0.0, 0.0])
    R = G.ReconstructionIdentity(G.HDIVRT0(2) if recon == "RT0" else G.HDIVBDM1(2))
    A = G.FEMatrix([sv, sq])
    G.assemble_operator(A[1, 1], G.LaplaceOperator(1.0))
    G.assemble_operator(A[1, 2], G.LagrangeMultiplier(G.Divergence), At=A[2, 1])
    fixed, _ = br_boundary_values(sv, lambda x: np.zeros((2, x.shape[1])))
    n = sv.ndofs
    M = A.tocsc().tolil()
    penalty = 1e60
    d = np.ones(M.shape[0])
    for j in list(fixed) + [n]:
        M[j, j] = penalty
        d[j] = 1.0 / penalty
    Ms = (sp.diags(d) @ M.tocsr()).tocsc()
    errs = {}
    for name, op in (("robust", R), ("classical", G.Identity)):
        rhs = G.FEVector([sv, sq])
        Lf = G.LinearForm(op, fdata)
        G.assemble_operator(rhs[1], Lf)
        if name == "robust":
            assert len(Lf._pattern.AM.qf) == 9 and G.blf_stats(Lf._pattern).path == G._lib.PATH_COLUMNS      # Stroud rule, gather kernel
        b = rhs.entries.copy()
        b[list(fixed) + [n]] = 0.0
        sol = G.FEVector([sv, sq])
        sol.entries[:] = spla.spsolve(Ms, d * b)
        errs[name] = np.sqrt(np.abs(np.asarray(G.evaluate(G.L2ErrorIntegrator(zero, G.Identity, quadorder=0), sol[1]))).sum())
    assert errs["robust"] < 1e-14, errs
    assert errs["classical"] > 1e-3, errs


@pytest.mark.parametrize("fam", ["RT0", "BDM1"])
def test_example302_hdiv_bestapproximation_on_device(fam):
    """Example302_BestapproximationHdiv3D = BASELINE configuration C5 on the device: RT0 / BDM1 mass matrix with the Piola map (column kernels), the
    Hdiv x P0 divergence block with its transposed copy, both right-hand sides, and the per-cell ItemIntegrator of Divergence(u_h).  The example's claim
    "the divergence of the approximation equals the piecewise integral mean of the exact divergence" holds cell by cell."""
    g = G.uniform_refine(G.reference_domain("Tetrahedron3D"), 2)
    sv, sq = G.FESpace((G.HDIVRT0 if fam == "RT0" else G.HDIVBDM1)(3), g), G.FESpace(G.L2P0(1), g)
    udata = G.DataFunction(lambda x: np.stack([x[0] ** 3 + x[2] ** 2, -x[0] ** 2 + x[1] + 1, x[0] * x[1]]), [3, 3], bonus_quadorder=3)
    ddata = G.DataFunction(lambda x: np.stack([3 * x[0] ** 2 + 1.0]), [1, 3], bonus_quadorder=2)
    A = G.FEMatrix([sv, sq])
    G.assemble_operator(A[1, 1], G.ReactionOperator(1.0))
    G.assemble_operator(A[1, 2], G.LagrangeMultiplier(G.Divergence), At=A[2, 1])
    rhs = G.FEVector([sv, sq])
    G.assemble_operator(rhs[1], G.LinearForm(G.Identity, udata))
    # the transposed copy of a LagrangeMultiplier block carries the opposite sign (bilinearform.jl:354-360: "sign is changed in case nonzero rhs
    # data is applied to LagrangeMultiplier"), so the constraint row reads (div u_h, q) = (div u, q) with the example's plain right-hand side
    G.assemble_operator(rhs[2], G.LinearForm(G.Identity, ddata))
    b2 = rhs.entries[sv.ndofs:].copy()
    sol = G.FEVector([sv, sq])
    sol.entries[:] = spla.spsolve(A.tocsc(), rhs.entries)
    cell_div = np.zeros((g.ncells, 1))
    G.evaluate_itemwise(cell_div, G.ItemIntegrator([G.Divergence]), sol[1])
    assert np.abs(cell_div[:, 0] - b2).max() < 1e-14
    err = np.sqrt(np.abs(np.asarray(G.evaluate(G.L2ErrorIntegrator(udata, G.Identity, quadorder=3), sol[1]))).sum())
    assert 1e-4 < err < 0.2                                                        # u is cubic: a genuine approximation error remains


@pytest.mark.parametrize("level", [2, 3])
def test_example301_poisson3d_on_device_with_the_metric_kernel(level):
    """Example301_Poisson3D = BASELINE configuration C2, the form of the headline metric, solved end to end: the P2 Laplacian by the ring-walk kernel
    (GRMP_PATH_AUTO -> PATH_FAST), LinearForm(Identity, Laplace u; factor = -1), BestapproxDirichletBoundary on all six regions (ON_BFACES forms),
    penalties on the device-resident matrix, host solve, L2 and H1 error integrators on the device.  u = x (z - y) + y^2 lies in H1P2{1,3}: both errors
    vanish to rounding (the example itself uses H1P1 and reports a convergence history)."""
    g = G.uniform_refine(G.grid_unitcube(), level)
    s = G.FESpace(G.H1P2(1, 3), g)
    udata = G.DataFunction(lambda x: np.stack([x[0] * (x[2] - x[1]) + x[1] * x[1]]), [1, 3], bonus_quadorder=2)
    gdata = G.DataFunction(lambda x: np.stack([x[2] - x[1], -x[0] + 2 * x[1], x[0]]), [3, 3], bonus_quadorder=1)
    A = G.DiscreteSymmetricBilinearForm([G.Gradient, G.Gradient], [s, s])
    cp, rv, _ = G.assemble_csc(A, 1.0)
    assert G.blf_stats(A).path == G._lib.PATH_FAST
    rhs = G.FEVector([s])
    G.assemble(rhs[1], G.DiscreteLinearForm([G.Identity], [s], G.fdot_action(G.DataFunction([2.0]))), factor=-1)
    sol = G.FEVector([s])
    fixed = G.boundarydata(sol[1], [G.BoundaryData(G.BestapproxDirichletBoundary, data=udata, regions=[1, 2, 3, 4, 5, 6])])
    G.apply_penalties(A, fixed, 1e60)
    rhs.entries[fixed - 1] = 1e60 * sol.entries[fixed - 1]
    M = sp.csc_matrix((G.fetch_values(A), rv - 1, cp - 1), shape=(s.ndofs, s.ndofs))
    sol.entries[:] = spla.spsolve(M, rhs.entries)
    e0 = np.sqrt(abs(G.evaluate(G.L2ErrorIntegrator(udata, G.Identity), sol[1])))
    e1 = np.sqrt(abs(G.evaluate(G.L2ErrorIntegrator(gdata, G.Gradient), sol[1])))
    assert e0 < 6e-12 and e1 < 6e-11, (e0, e1)
    _, nrm = G.residual(A, sol.entries, rhs.entries, fixed_dofs=fixed, want_vector=False)
    assert nrm < 1e-20
