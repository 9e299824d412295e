"""ItemIntegrator on the device (grmp_ii_*) vs the CPU oracle's restatement of itemintegrator.jl:160-360.
evaluate!(b, AP, FEB): per-item results bit-identical (same operation order, no FMA).  evaluate(AP, FEB): the reference adds every
(item, quadrature point) term to one running sum, the device adds per-item sums in a tree: 1e-12 relative."""
import numpy as np
import pytest

import grmp_b200 as G
import oracle as O

pytestmark = pytest.mark.gpu


def tri_grid(L, perturbed=False):
    g = G.uniform_refine(G.grid_unitsquare("Triangle2D"), L)
    return G.perturb_interior_nodes(g) if perturbed else g


def tet_grid(L, perturbed=False):
    g = G.uniform_refine(G.grid_unitcube("Tetrahedron3D"), L)
    return G.perturb_interior_nodes(g) if perturbed else g


CASES = [
    ("P2 tri id L2 error", lambda: tri_grid(3, True), lambda: G.H1P2(1, 2), G.Identity, "l2error", [0]),
    ("P2 tet grad L2 norm", lambda: tet_grid(1, True), lambda: G.H1P2(1, 3), G.Gradient, "l2norm", [0]),
    ("P1 tri id integral, region 2", lambda: tri_grid(3), lambda: G.H1P1(1), G.Identity, "none", [2]),
    ("P2{2} tri symgrad integral", lambda: tri_grid(2, True), lambda: G.H1P2(2, 2), G.SymmetricGradient(1), "none", [0]),
    ("BR tri div L2 norm", lambda: tri_grid(2, True), lambda: G.H1BR(2), G.Divergence, "l2norm", [0]),
    ("RT0 tet id L2 error", lambda: tet_grid(1), lambda: G.HDIVRT0(3), G.Identity, "l2error", [0]),
    ("BDM1 tet div integral", lambda: tet_grid(1, True), lambda: G.HDIVBDM1(3), G.Divergence, "none", [0]),
    ("BR tri recon BDM1 L2 error", lambda: tri_grid(2, True), lambda: G.H1BR(2), G.ReconstructionIdentity(G.HDIVBDM1(2)), "l2error", [0]),
    ("P0 tri id L2 norm", lambda: tri_grid(2), lambda: G.L2P0(1), G.Identity, "l2norm", [0]),
]


@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_itemintegrator_parity(case):
    _, gridf, fef, op, kind, regions = case
    g = gridf()
    if regions != [0]:
        g.cellregions[: g.ncells // 3] = regions[0]
    s = G.FESpace(fef(), g)
    dim = g.dim
    rng = np.random.default_rng(7)
    u = G.FEVector([s])
    u.entries[:] = rng.standard_normal(s.ndofs)
    opc = G.assembly._op(op)
    rdim = G.assembly._resultdim(type("A", (), {"FES": [s], "operators": [opc]})())
    factor = 1.0
    if kind == "l2error":
        data = G.DataFunction(lambda x: np.stack([np.cos(x[k % len(x)]) + 0.1 * j for j, k in enumerate(range(rdim))]), [rdim, dim], bonus_quadorder=2)
        factor = 0.5
        AP = G.L2ErrorIntegrator(data, op, factor=factor, regions=regions)
    elif kind == "l2norm":
        AP = G.L2NormIntegrator(rdim, op, quadorder=2, regions=regions)
    else:
        AP = G.ItemIntegrator([op], regions=regions)
    ard = rdim if kind == "none" else 1
    b = np.full((g.ncells, ard), 0.125)             # += semantics
    G.evaluate_itemwise(b, AP, u[1])
    total = G.evaluate(AP, u[1], skip_preps=True)
    # oracle on the same quadrature rule and the same host-evaluated table
    qo = G.quadrature_order(AP)
    P = AP.AM
    O.qrule_override(dim, qo, P.qf.xref, P.qf.w)
    try:
        table = None
        if kind == "l2error":
            xq = O.quadpoints(g, qo)
            flat = xq.reshape(-1, dim)
            vals = np.asarray(data.kernel(flat.T), dtype=np.float64).reshape(-1, flat.shape[0]).T
            table = vals.reshape(g.ncells, len(P.qf), -1)
        kw = dict(kind={"none": O.II_NONE, "l2norm": O.II_L2NORM, "l2error": O.II_L2ERROR}[kind], factor=factor, data=table, regions=regions,
                  bonus_quadorder=AP.action.bonus_quadorder)
        ob, otot = O.ii_evaluate(g, s, opc.code, u.entries, **kw)
        ob2, _ = O.ii_evaluate(g, s, opc.code, u.entries, b=np.full((g.ncells, ard), 0.125), **kw)   # += into the same start values
    finally:
        O.qrule_override(dim, qo)
    assert np.array_equal(b, ob2), f"per-item results not bit-identical: max abs diff {np.abs(b - ob2).max():.3e}"
    tot = np.atleast_1d(total)
    scale = np.abs(ob).sum(axis=0)                   # sum of |item values|: conditioning of the sum
    assert np.all(np.abs(tot - otot) <= 1e-12 * np.maximum(scale, 1e-300)), (tot, otot)
    if regions != [0]:
        assert np.all(b[g.ncells // 3:] == 0.125)


def test_l2error_of_interpolant_is_zero_and_repeatable():
    g = tet_grid(2)
    s = G.FESpace(G.H1P2(1, 3), g)
    f = lambda p: p[0] ** 2 - p[1] * p[2]            # noqa: E731
    en = g.edgenodes.astype(int) - 1
    x = np.concatenate([g.coords, (g.coords[en[:, 0]] + g.coords[en[:, 1]]) / 2])
    u = G.FEVector([s])
    u.entries[:] = [f(p) for p in x]
    AP = G.L2ErrorIntegrator(G.DataFunction(lambda x: x[0] ** 2 - x[1] * x[2], [1, 3], bonus_quadorder=2), G.Identity)
    e1 = G.evaluate(AP, u[1])
    e2 = G.evaluate(AP, u[1], skip_preps=True)
    assert e1 == e2 and 0 <= e1 < 1e-24
