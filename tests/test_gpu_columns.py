"""GPU parity tests of the kernel family behind GRMP_PATH_COLUMNS (owner-computes column kernels, csrc/colpath.cu) and of the
measured scatter alternatives GRMP_PATH_ATOMIC / GRMP_PATH_COLOURED (csrc/cellpath.cu): through the C ABI against the CPU
oracle on the same seeded inputs.  Pattern bit-identical, values within the two-tier 1e-12 tolerance of tests/parity.py."""
import numpy as np
import pytest

import grmp_b200 as G
import oracle as O
from parity import oracle_blf, oracle_scale, rel_err, tier_report
from test_gpu_parity import tet_grid, tri_grid

pytestmark = pytest.mark.gpu
RTOL = 1e-12
COLS, ATOM, COLR, GEN = G._lib.PATH_COLUMNS, G._lib.PATH_ATOMIC, G._lib.PATH_COLOURED, G._lib.PATH_GENERIC


def check(AP, factor=1.0, path=COLS, transposed=False):
    G.blf_set_path(AP, path)
    cp, rv, nz = G.assemble_csc(AP, factor, transposed_assembly=transposed)
    assert G.blf_stats(AP).path == path
    kw = {"transposed_assembly": True} if transposed else {}
    ocp, orv, onz = oracle_blf(AP, factor, **kw)
    assert np.array_equal(cp, ocp), "colptr differs"
    assert np.array_equal(rv, orv), "rowval differs"
    S = oracle_scale(AP, factor, **kw)
    assert rel_err(nz, onz, S) <= RTOL, tier_report(nz, onz, S)
    cp2, rv2, nz2 = G.assemble_csc(AP, -0.5 * factor, skip_preps=True, transposed_assembly=transposed)      # frozen pattern, other factor
    assert rel_err(nz2, -0.5 * onz, 0.5 * S) <= RTOL, tier_report(nz2, -0.5 * onz, 0.5 * S)
    return cp, rv, nz


SQUARE = [
    ("P1 tri Laplace", lambda: tri_grid(3), lambda: G.H1P1(1), (G.Gradient, G.Gradient)),
    ("P1 tri Laplace perturbed", lambda: tri_grid(3, True), lambda: G.H1P1(1), (G.Gradient, G.Gradient)),
    ("P1x2 tri Laplace", lambda: tri_grid(2, True), lambda: G.H1P1(2), (G.Gradient, G.Gradient)),
    ("P2 tri Laplace (C1)", lambda: tri_grid(4), lambda: G.H1P2(1, 2), (G.Gradient, G.Gradient)),
    ("P2 tri Laplace perturbed", lambda: tri_grid(3, True), lambda: G.H1P2(1, 2), (G.Gradient, G.Gradient)),
    ("P2 tri mass", lambda: tri_grid(3), lambda: G.H1P2(1, 2), (G.Identity, G.Identity)),
    ("P2x2 tri mass", lambda: tri_grid(2, True), lambda: G.H1P2(2, 2), (G.Identity, G.Identity)),
    ("P1 tet Laplace (C2)", lambda: tet_grid(2), lambda: G.H1P1(1), (G.Gradient, G.Gradient)),
    ("P1 tet Laplace perturbed", lambda: tet_grid(2, True), lambda: G.H1P1(1), (G.Gradient, G.Gradient)),
    ("P2 tet Laplace (C2*)", lambda: tet_grid(2), lambda: G.H1P2(1, 3), (G.Gradient, G.Gradient)),
    ("P2 tet Laplace perturbed", lambda: tet_grid(2, True), lambda: G.H1P2(1, 3), (G.Gradient, G.Gradient)),
    ("P2 tet mass", lambda: tet_grid(1), lambda: G.H1P2(1, 3), (G.Identity, G.Identity)),
    ("P2x3 tet Laplace", lambda: tet_grid(1, True), lambda: G.H1P2(3, 3), (G.Gradient, G.Gradient)),
    ("P0 tri mass", lambda: tri_grid(2), lambda: G.L2P0(1), (G.Identity, G.Identity)),
    ("RT0 tri mass", lambda: tri_grid(3), lambda: G.HDIVRT0(2), (G.Identity, G.Identity)),
    ("RT0 tri mass perturbed", lambda: tri_grid(3, True), lambda: G.HDIVRT0(2), (G.Identity, G.Identity)),
    ("BDM1 tri mass", lambda: tri_grid(3, True), lambda: G.HDIVBDM1(2), (G.Identity, G.Identity)),
    ("RT0 tet mass (C5)", lambda: tet_grid(1), lambda: G.HDIVRT0(3), (G.Identity, G.Identity)),
    ("RT0 tet mass perturbed", lambda: tet_grid(2, True), lambda: G.HDIVRT0(3), (G.Identity, G.Identity)),
    ("BDM1 tet mass (C5)", lambda: tet_grid(1, True), lambda: G.HDIVBDM1(3), (G.Identity, G.Identity)),
    ("BDM1 tet mass axis aligned", lambda: tet_grid(2), lambda: G.HDIVBDM1(3), (G.Identity, G.Identity)),
    ("RT0 tet div-div", lambda: tet_grid(1), lambda: G.HDIVRT0(3), (G.Divergence, G.Divergence)),
    ("BDM1 tri div-div", lambda: tri_grid(2, True), lambda: G.HDIVBDM1(2), (G.Divergence, G.Divergence)),
    ("BR tri Laplace (C4)", lambda: tri_grid(3), lambda: G.H1BR(2), (G.Gradient, G.Gradient)),
    ("BR tri Laplace perturbed", lambda: tri_grid(3, True), lambda: G.H1BR(2), (G.Gradient, G.Gradient)),
    ("BR tet Laplace", lambda: tet_grid(1, True), lambda: G.H1BR(3), (G.Gradient, G.Gradient)),
    ("BR tri mass", lambda: tri_grid(2), lambda: G.H1BR(2), (G.Identity, G.Identity)),
    ("BR tri recon RT0 mass (C4)", lambda: tri_grid(2), lambda: G.H1BR(2),
     (G.ReconstructionIdentity(G.HDIVRT0(2)), G.ReconstructionIdentity(G.HDIVRT0(2)))),
    ("BR tri recon RT0 mass perturbed", lambda: tri_grid(3, True), lambda: G.H1BR(2),
     (G.ReconstructionIdentity(G.HDIVRT0(2)), G.ReconstructionIdentity(G.HDIVRT0(2)))),
    ("BR tri recon BDM1 mass (C4)", lambda: tri_grid(2, True), lambda: G.H1BR(2),
     (G.ReconstructionIdentity(G.HDIVBDM1(2)), G.ReconstructionIdentity(G.HDIVBDM1(2)))),
    ("BR tri recon BDM1 mass axis aligned", lambda: tri_grid(3), lambda: G.H1BR(2),
     (G.ReconstructionIdentity(G.HDIVBDM1(2)), G.ReconstructionIdentity(G.HDIVBDM1(2)))),
]


@pytest.mark.parametrize("case", SQUARE, ids=[c[0] for c in SQUARE])
@pytest.mark.parametrize("apt", ["sym", "gen"])
def test_column_kernels_parity(case, apt):
    _, gridf, fef, ops = case
    g = gridf()
    s = G.FESpace(fef(), g)
    ctor = G.DiscreteSymmetricBilinearForm if apt == "sym" else G.DiscreteBilinearForm
    check(ctor(list(ops), [s, s]), factor=0.75)


@pytest.mark.parametrize("dim,fe", [(2, "P1"), (2, "P2"), (3, "P1"), (3, "P2")])
def test_column_kernels_hooke(dim, fe):
    g = tri_grid(3, True) if dim == 2 else tet_grid(1, True)
    s = G.FESpace(G.H1P1(dim) if fe == "P1" else G.H1P2(dim, dim), g)
    mu = 1000 / 1.4
    lam = 0.4 * mu / 0.2
    AP = G.DiscreteBilinearForm([G.SymmetricGradient(1), G.SymmetricGradient(1)], [s, s], G.HookeAction(dim, mu, lam))
    check(AP)
    g2 = tri_grid(2) if dim == 2 else tet_grid(1)         # axis-aligned: pattern hinges on exact zeros
    s2 = G.FESpace(G.H1P1(dim) if fe == "P1" else G.H1P2(dim, dim), g2)
    check(G.DiscreteBilinearForm([G.SymmetricGradient(1), G.SymmetricGradient(1)], [s2, s2], G.HookeAction(dim, mu, lam)))


@pytest.mark.parametrize("dim", [2, 3])
@pytest.mark.parametrize("transposed", [False, True])
def test_column_kernels_stokes_divergence_block(dim, transposed):
    """LagrangeMultiplier(Divergence) block BR x P0 (Example222) and its transposed assembly"""
    g = tri_grid(3, True) if dim == 2 else tet_grid(1, True)
    sv = G.FESpace(G.H1BR(dim), g)
    sp = G.FESpace(G.L2P0(1), g)
    check(G.DiscreteBilinearForm([G.Divergence, G.Identity], [sv, sp]), factor=-1.0, transposed=transposed)


@pytest.mark.parametrize("dim", [2, 3])
@pytest.mark.parametrize("transposed", [False, True])
def test_column_kernels_taylor_hood_divergence_block(dim, transposed):
    g = tri_grid(2, True) if dim == 2 else tet_grid(1, True)
    su = G.FESpace(G.H1P2(dim, dim), g)
    sp = G.FESpace(G.H1P1(1), g)
    check(G.DiscreteBilinearForm([G.Divergence, G.Identity], [su, sp]), transposed=transposed)


@pytest.mark.parametrize("fe", ["RT0", "BDM1"])
@pytest.mark.parametrize("dim", [2, 3])
def test_column_kernels_hdiv_divergence_block(dim, fe):
    g = tri_grid(2, True) if dim == 2 else tet_grid(1, True)
    sv = G.FESpace(G.HDIVRT0(dim) if fe == "RT0" else G.HDIVBDM1(dim), g)
    sp = G.FESpace(G.L2P0(1), g)
    check(G.DiscreteBilinearForm([G.Divergence, G.Identity], [sv, sp]))
    check(G.DiscreteBilinearForm([G.Divergence, G.Identity], [sv, sp]), transposed=True)


def test_column_kernels_regions_and_kappa():
    g = tri_grid(3)
    g.cellregions[::3] = 2
    s = G.FESpace(G.H1P2(1, 2), g)
    check(G.DiscreteSymmetricBilinearForm([G.Gradient, G.Gradient], [s, s], regions=[2]), factor=1e-3)


def test_column_kernels_deterministic_and_close_to_generic_on_large_grids():
    """size-independent checks at 786 432 tets / 1 048 576 triangles: two runs bit-equal, agreement with the bit-exact generic
    path, A 1 = 0 for stiffness matrices"""
    import scipy.sparse as sp_
    for name, g, fe, ops, act in (
        ("P1 tet L5", tet_grid(5), G.H1P1(1), (G.Gradient, G.Gradient), None),
        ("BDM1 tet L4", tet_grid(4), G.HDIVBDM1(3), (G.Identity, G.Identity), None),
        ("Hooke P2 tri L8", tri_grid(8), G.H1P2(2, 2), (G.SymmetricGradient(1), G.SymmetricGradient(1)), G.HookeAction(2, 714.0, 1428.0)),
    ):
        s = G.FESpace(fe, g)
        ctor = G.DiscreteBilinearForm if act is not None else G.DiscreteSymmetricBilinearForm
        AP = ctor(list(ops), [s, s], act)
        G.blf_set_path(AP, COLS)
        cp, rv, nz = G.assemble_csc(AP, 1.0)
        _, _, nz2 = G.assemble_csc(AP, 1.0, skip_preps=True)
        assert np.array_equal(nz, nz2), name
        APg = ctor(list(ops), [s, s], act)
        G.blf_set_path(APg, GEN)
        cpg, rvg, nzg = G.assemble_csc(APg, 1.0)
        assert np.array_equal(cp, cpg) and np.array_equal(rv, rvg), name
        assert rel_err(nz, nzg) <= RTOL, name + ": " + tier_report(nz, nzg)      # axis-aligned grids: pure 1e-12
        if ops[0] is not G.Identity:
            A = sp_.csc_matrix((nz, rv - 1, cp - 1), shape=(s.ndofs, s.ndofs))
            k = np.ones(s.ndofs)
            assert np.abs(A @ k).max() < 1e-10 * np.abs(nz).max(), name


def test_column_kernels_owned_columns_partition():
    """multi-GPU path on one device: every 'rank' assembles only the columns it owns with the column kernels"""
    g = tri_grid(4, True)
    s = G.FESpace(G.H1P2(2, 2), g)
    act = G.HookeAction(2, 2.0, 3.0)
    AP = G.DiscreteBilinearForm([G.SymmetricGradient(1), G.SymmetricGradient(1)], [s, s], act)
    G.blf_set_path(AP, GEN)
    cp, rv, nz = G.assemble_csc(AP, 1.0)
    blocks = []
    for r in range(3):
        lp = G.partition.partition(s, r, 3)
        APr = G.DiscreteBilinearForm([G.SymmetricGradient(1), G.SymmetricGradient(1)], [lp.space, lp.space], act)
        G.prepare_assembly(APr)
        G.blf_set_path(APr, COLS)
        G._lib.check(G._lib.lib().grmp_blf_set_owned_columns(APr.AM.h, lp.n_owned))
        lcp, lrv, lnz = G.assemble_csc(APr, 1.0, skip_preps=True)
        blocks.append(G.partition.owned_block_to_global(lp, lcp, lrv, lnz))
    mcp, mrv, mnz = G.partition.merge_owned_columns(s.ndofs, blocks)
    assert np.array_equal(mcp, cp) and np.array_equal(mrv, rv)
    S = oracle_scale(AP, 1.0)
    assert rel_err(mnz, nz, S) <= RTOL, tier_report(mnz, nz, S)


@pytest.mark.parametrize("nw", [1, 2, 8])
def test_column_kernels_tile_shapes(nw, monkeypatch):
    monkeypatch.setenv("GRMP_COL_NW", str(nw))
    g = tet_grid(2, True)
    s = G.FESpace(G.H1P2(1, 3), g)
    check(G.DiscreteSymmetricBilinearForm([G.Gradient, G.Gradient], [s, s]))


SCATTER = [c for c in SQUARE if c[0] in ("P1 tri Laplace perturbed", "P2 tri Laplace (C1)", "P2 tet Laplace perturbed", "P2x2 tri mass",
                                         "RT0 tet mass perturbed", "BDM1 tri mass", "BR tri Laplace (C4)", "P1 tet Laplace (C2)")]


@pytest.mark.parametrize("case", SCATTER, ids=[c[0] for c in SCATTER])
@pytest.mark.parametrize("path", [ATOM, COLR], ids=["atomic", "coloured"])
def test_cell_parallel_scatter_variants(case, path):
    _, gridf, fef, ops = case
    g = gridf()
    s = G.FESpace(fef(), g)
    AP = G.DiscreteSymmetricBilinearForm(list(ops), [s, s])
    _, _, nz = check(AP, factor=1.5, path=path)
    if path == COLR:       # colouring fixes the summation order
        _, _, nz2 = G.assemble_csc(AP, 1.5, skip_preps=True)
        _, _, nz3 = G.assemble_csc(AP, 1.5, skip_preps=True)
        assert np.array_equal(nz2, nz3)
        assert G.blf_stats(AP).kernel_launches >= 3


def test_cell_parallel_hooke():
    g = tri_grid(3, True)
    s = G.FESpace(G.H1P2(2, 2), g)
    for path in (ATOM, COLR):
        check(G.DiscreteBilinearForm([G.SymmetricGradient(1), G.SymmetricGradient(1)], [s, s], G.HookeAction(2, 2.0, 3.0)), path=path)


# ---- linear forms: owner-computes gather kernels (one thread per dof) vs the oracle ----------------------------------------------
from test_gpu_parity import LF_CASES  # noqa: E402


@pytest.mark.parametrize("case", LF_CASES, ids=[c[0] for c in LF_CASES])
def test_lf_gather_kernels(case):
    _, gridf, fef, op, kind, regions = case
    g = gridf()
    s = G.FESpace(fef(), g)
    dim = g.dim
    nc = s.fetype.ncomponents
    bonus = 0
    if kind == "const":
        data = G.DataFunction([1.0])
    elif kind == "none":
        data = None
    elif kind == "fun":
        data = G.DataFunction(lambda x: np.sin(x[0]) * x[1] + (x[2] if len(x) > 2 else 0.0), [1, dim], bonus_quadorder=2)
        bonus = 2
    elif kind == "vfun":
        data = G.DataFunction(lambda x: np.stack([x[k] ** 2 + x[(k + 1) % len(x)] for k in range(len(x))]), [nc, dim], bonus_quadorder=1)
        bonus = 1
    else:
        data = G.DataFunction(lambda x: np.stack([3 * x[k] ** 2 for k in range(len(x))]), [nc, dim], bonus_quadorder=2)
        bonus = 2
    expect_fast = not (isinstance(op, G.ReconstructionIdentity) and dim == 3)     # tetrahedral reconstruction stays generic
    old = G.assembly.DEFAULT_PATH
    G.assembly.DEFAULT_PATH = G._lib.PATH_AUTO
    try:
        Op = G.LinearForm(op, data, regions=regions, factor=2.0)
        b = G.FEVector([s])
        b.entries[:] = 0.25
        AP = G.assemble_operator(b[1], Op)
        assert G.blf_stats(AP).path == (COLS if expect_fast else GEN)
        first = b.entries.copy()
        b.entries[:] = 0.25
        G.assemble_operator(b[1], Op, Pattern=AP, skip_preps=True)
        assert np.array_equal(first, b.entries), "two runs differ"
    finally:
        G.assembly.DEFAULT_PATH = old
    qo = G.quadrature_order(AP)
    P = AP.AM
    O.qrule_override(dim, qo, P.qf.xref, P.qf.w)
    try:
        ob = np.full(s.ndofs, 0.25)
        mag = np.zeros(s.ndofs)
        if kind == "const":
            O.lf_assemble(ob, g, s, op.code, fsrc=O.F_CONST, fdata=[1.0], regions=regions, factor=2.0, bonus_quadorder=bonus)
        elif kind == "none":
            O.lf_assemble(ob, g, s, op.code, fsrc=O.F_NONE, regions=regions, factor=2.0)
        else:
            xq = O.quadpoints(g, qo)
            flat = xq.reshape(-1, dim)
            vals = np.asarray(data.kernel(flat.T), dtype=np.float64).reshape(-1, flat.shape[0]).T
            table = vals.reshape(g.ncells, len(P.qf), -1)
            O.lf_assemble(ob, g, s, op.code, fsrc=O.F_QP_TABLE, fdata=table, regions=regions, factor=2.0, bonus_quadorder=bonus)
    finally:
        O.qrule_override(dim, qo)
    # entries of b are sums over cells and quadrature points with mixed signs: tolerance relative to the largest entry of the
    # contribution (b - 0.25), i.e. 1e-12 * max|b| absolute plus 1e-12 relative
    scale = np.abs(ob - 0.25).max()
    assert np.abs(b.entries - ob).max() <= 1e-12 * max(scale, 1e-300), np.abs(b.entries - ob).max() / scale
