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
