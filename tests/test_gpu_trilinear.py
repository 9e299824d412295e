"""Trilinear forms with one coefficient (FEB) argument on the device (first slice): the Picard-linearised convection
term of ConvectionOperator (pdeoperators.jl:435-510) through assemble!(A, AP, FEB; fixed_arguments = [1]) (bilinearform.jl:235-257).
Generic path: pattern and values bit-identical to the oracle."""
import numpy as np
import pytest

import grmp_b200 as G
import oracle as O

pytestmark = pytest.mark.gpu


def grid(dim, L, perturbed):
    g = G.uniform_refine(G.grid_unitsquare("Triangle2D") if dim == 2 else G.grid_unitcube("Tetrahedron3D"), L)
    return G.perturb_interior_nodes(g) if perturbed else g


CASES = [
    ("P2 velocity, P2 coefficient, tri", 2, 2, True, lambda d: G.H1P2(d, d), lambda d: G.H1P2(d, d)),
    ("P2 velocity, P1 coefficient, tri (axis aligned)", 2, 3, False, lambda d: G.H1P2(d, d), lambda d: G.H1P1(d)),
    ("BR velocity, BR coefficient, tri", 2, 2, True, lambda d: G.H1BR(d), lambda d: G.H1BR(d)),
    ("P2 velocity, P2 coefficient, tet", 3, 1, True, lambda d: G.H1P2(d, d), lambda d: G.H1P2(d, d)),
    ("P1 velocity, P1 coefficient, tet", 3, 1, False, lambda d: G.H1P1(d), lambda d: G.H1P1(d)),
]


@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_convection_operator_parity(case):
    _, dim, L, pert, fev, fea = case
    g = grid(dim, L, pert)
    sv = G.FESpace(fev(dim), g)
    sa = sv if fea(dim).__class__ is fev(dim).__class__ else G.FESpace(fea(dim), g)
    sol = G.FEVector([sa] if sa is sv else [sa, sv])
    rng = np.random.default_rng(11)
    sol.entries[:] = rng.standard_normal(sol.entries.size)
    A = G.FEMatrix([sv])
    Op = G.ConvectionOperator(1, G.Identity, dim, dim, factor=0.75)
    AP = G.assemble_operator(A[1, 1], Op, sol)
    assert G.blf_stats(AP).path == G._lib.PATH_GENERIC
    cp, rv, nz = AP.AM.colptr, AP.AM.rowval, G.fetch_values(AP)
    qo = G.quadrature_order(AP)
    P = AP.AM
    O.qrule_override(dim, qo, P.qf.xref, P.qf.w)
    try:
        OA = O.OracleMatrix(sv.ndofs, sv.ndofs)
        O.blf_assemble(OA, g, sv, sv, O.OP_GRAD, O.OP_ID, action=O.ACT_CONVECTION, transposed_assembly=True, factor=0.75,
                       fixed=(sa, O.OP_ID, sol.entries[: sa.ndofs]))
        ocp, orv, onz = OA.csc()
    finally:
        O.qrule_override(dim, qo)
    assert np.array_equal(cp, ocp) and np.array_equal(rv, orv), "pattern differs"
    assert np.array_equal(nz, onz), f"values not bit-identical, max abs diff {np.abs(nz - onz).max():.3e}"
    # next Picard iteration on the frozen pattern: new coefficients, skip_preps = true
    sol.entries[: sa.ndofs] *= -0.5
    G.assemble_operator(A[1, 1], Op, sol, Pattern=AP, skip_preps=True)
    nz2 = G.fetch_values(AP)
    O.qrule_override(dim, qo, P.qf.xref, P.qf.w)
    try:
        OA.fill_zero()
        O.blf_assemble(OA, g, sv, sv, O.OP_GRAD, O.OP_ID, action=O.ACT_CONVECTION, transposed_assembly=True, factor=0.75,
                       fixed=(sa, O.OP_ID, sol.entries[: sa.ndofs]))
        onz2 = OA.csc()[2]
    finally:
        O.qrule_override(dim, qo)
    assert np.array_equal(nz2, onz2)
    assert np.array_equal(nz2, -0.5 * nz)            # the form is linear in a, and scaling by -0.5 is exact


def test_convection_needs_its_coefficient():
    g = grid(2, 1, False)
    sv = G.FESpace(G.H1P2(2, 2), g)
    AP = G.DiscreteBilinearForm([G.Identity, G.Gradient, G.Identity], [sv, sv, sv], G.ConvectionAction(2, 2))
    with pytest.raises(G._lib.GrmpError):
        G.assemble_csc(AP, 1.0)                      # no fixed argument set: GRMP_ESTATE, nothing is assembled


LF_FEB_CASES = [
    ("P2 test, P1 coefficient, id-id, tri", 2, 3, True, lambda d: G.H1P2(1, d), G.Identity, lambda d: G.H1P1(1), G.Identity),
    ("P2{2} test, grad of P2 scalar as coefficient, tri", 2, 2, True, lambda d: G.H1P2(d, d), G.Identity, lambda d: G.H1P2(1, d), G.Gradient),
    ("RT0 test, P1{3} coefficient, tet", 3, 1, False, lambda d: G.HDIVRT0(d), G.Identity, lambda d: G.H1P1(d), G.Identity),
    ("P0 test, div of BR coefficient, tri", 2, 2, True, lambda d: G.L2P0(1), G.Identity, lambda d: G.H1BR(d), G.Divergence),
]


@pytest.mark.parametrize("case", LF_FEB_CASES, ids=[c[0] for c in LF_FEB_CASES])
@pytest.mark.parametrize("path", ["generic", "auto"])
def test_linearform_with_coefficient_argument(case, path):
    """assemble!(b, AP, FEB) with nFE = 2 and NoAction (linearform.jl:130-178): b[dof] += int op_a(FEB[1]) . op(v_dof)"""
    _, dim, L, pert, fet, opt, fea, opa = case
    g = grid(dim, L, pert)
    st, sa = G.FESpace(fet(dim), g), G.FESpace(fea(dim), g)
    sol = G.FEVector([sa, st])
    rng = np.random.default_rng(5)
    sol.entries[:] = rng.standard_normal(sol.entries.size)
    old = G.assembly.DEFAULT_PATH
    G.assembly.DEFAULT_PATH = G._lib.PATH_GENERIC if path == "generic" else G._lib.PATH_AUTO
    try:
        AP = G.DiscreteLinearForm([opa, opt], [sa, st])
        b = G.FEVector([sa, st])
        b.entries[:] = 0.5
        G.assemble(b[2], AP, [sol[1]], factor=1.5)
    finally:
        G.assembly.DEFAULT_PATH = old
    qo = G.quadrature_order(AP)
    P = AP.AM
    opt_c, opa_c = G.assembly._op(opt), G.assembly._op(opa)
    O.qrule_override(dim, qo, P.qf.xref, P.qf.w)
    try:
        table = O.feb_table(g, sa, opa_c.code, sol.entries[: sa.ndofs], qo)
        ob = np.full(st.ndofs, 0.5)
        bonus = qo - (st.fetype.polynomialorder(dim) - opt_c.needed_derivative)      # the oracle adds the test space's own order
        O.lf_assemble(ob, g, st, opt_c.code, fsrc=O.F_QP_TABLE, fdata=table, factor=1.5, bonus_quadorder=bonus)
    finally:
        O.qrule_override(dim, qo)
    assert np.all(b.entries[: sa.ndofs] == 0.5)
    got = b.entries[sa.ndofs:]
    if path == "generic":
        assert np.array_equal(got, ob), f"max abs diff {np.abs(got - ob).max():.3e}"
    else:
        assert np.abs(got - ob).max() <= 1e-12 * np.abs(ob - 0.5).max()


NLF_CASES = [
    ("P2 tri", 2, 2, True, lambda d: G.H1P2(d, d)),
    ("P1 tri axis aligned", 2, 3, False, lambda d: G.H1P1(d)),
    ("BR tri", 2, 2, True, lambda d: G.H1BR(d)),
    ("P2 tet", 3, 1, True, lambda d: G.H1P2(d, d)),
]


@pytest.mark.parametrize("case", NLF_CASES, ids=[c[0] for c in NLF_CASES])
def test_newton_convection_form_parity(case):
    """full_assemble!(A, b, AP, FEB) of ConvectionOperator(...; newton = true) (nonlinearform.jl:44-245): Jacobian matrix and right-hand
    side bit-identical to the oracle's restatement (which follows the reference's sparse-jacobian mul! order; a dense jacobian would go
    through BLAS gemv, whose rounding is build dependent -- 1e-12 is the meaningful bar against the reference itself)"""
    _, dim, L, pert, fe = case
    g = grid(dim, L, pert)
    sv = G.FESpace(fe(dim), g)
    sol = G.FEVector([sv])
    sol.entries[:] = np.random.default_rng(21).standard_normal(sv.ndofs)
    A = G.FEMatrix([sv])
    b = G.FEVector([sv])
    b.entries[:] = 0.25
    Op = G.ConvectionOperator(1, G.Identity, dim, dim, newton=True, factor=1.25)
    AP = G.full_assemble_operator(A[1, 1], b[1], Op, sol)
    assert G.blf_stats(AP).path == G._lib.PATH_GENERIC
    cp, rv, nz = AP.AM.colptr, AP.AM.rowval, G.fetch_values(AP)
    qo = G.quadrature_order(AP)
    P = AP.AM
    O.qrule_override(dim, qo, P.qf.xref, P.qf.w)
    try:
        OA = O.OracleMatrix(sv.ndofs, sv.ndofs)
        ob = np.full(sv.ndofs, 0.25)
        O.nlf_convection(OA, ob, g, sv, sol.entries, factor=1.25)
        ocp, orv, onz = OA.csc()
    finally:
        O.qrule_override(dim, qo)
    assert np.array_equal(cp, ocp) and np.array_equal(rv, orv), "pattern differs"
    assert np.array_equal(nz, onz), f"jacobian not bit-identical, max abs diff {np.abs(nz - onz).max():.3e}"
    assert np.array_equal(b.entries, ob), f"rhs not bit-identical, max abs diff {np.abs(b.entries - ob).max():.3e}"
    # Newton identity on the device result: A u = 2 (b - 0.25)
    import scipy.sparse as sp_
    M = sp_.csc_matrix((nz, rv - 1, cp - 1), shape=(sv.ndofs, sv.ndofs))
    assert np.abs(M @ sol.entries - 2 * (b.entries - 0.25)).max() < 1e-10 * np.abs(b.entries - 0.25).max()
