"""GPU parity tests (run with -m gpu on the B200 box): libgrmp_cuda through the C ABI vs the
CPU oracle on the same seeded inputs.

Bar (BASELINE.json north_star): colptr/rowval bit-identical, every nzval within 1e-12 relative.

  * generic path: replays the reference's operation AND summation order without FMA, so its
    values are compared for EXACT equality (stricter than the bar, covers every entry literally);
  * fast kernels (ring walk, column kernels, cell-parallel kernels; FMA, other summation orders): two tiers,
      |ref| >  1e-13 * max|A| :  |val - ref| <= 1e-12 * |ref|            (the bar, pure relative error)
      |ref| <= 1e-13 * max|A| :  |val - ref| <= 1e-15 * max|A|           (explicit zeros: the reference stores
    +-1e-17-sized rounding residue where contributions cancel exactly, because ExtendableSparse keeps explicit zeros;
    no evaluation order other than the reference's own reproduces those digits).  `rel_err` returns the larger of the
    tier-1 relative error and the tier-2 error scaled so that "<= 1e-12" means both tiers hold.
"""
import numpy as np
import pytest

import grmp_b200 as G
import oracle as O

pytestmark = pytest.mark.gpu
RTOL = 1e-12


from parity import oracle_blf, oracle_scale, rel_err, tier_report  # noqa: E402,F401


@pytest.fixture(autouse=True)
def _bit_exact_default_path():
    """tests in this module that do not choose a back end themselves compare bit for bit: pin the generic path"""
    old = G.assembly.DEFAULT_PATH
    G.assembly.DEFAULT_PATH = G._lib.PATH_GENERIC
    yield
    G.assembly.DEFAULT_PATH = old


def tri_grid(L, perturbed=False):
    g = G.uniform_refine(G.grid_unitsquare("Triangle2D"), L)
    return G.perturb_interior_nodes(g) if perturbed else g


def tet_grid(L, perturbed=False):
    g = G.uniform_refine(G.grid_unitcube("Tetrahedron3D"), L)
    return G.perturb_interior_nodes(g) if perturbed else g


def check_blf(AP, factor=1.0, exact=True, path=None):
    if path is not None:
        G.blf_set_path(AP, path)
    cp, rv, nz = G.assemble_csc(AP, factor)
    ocp, orv, onz = oracle_blf(AP, factor)
    assert np.array_equal(cp, ocp), "colptr differs"
    assert np.array_equal(rv, orv), "rowval differs"
    S = None if exact else oracle_scale(AP, factor)
    if exact:
        assert np.array_equal(nz, onz), f"nzval not bit-identical (max rel {rel_err(nz, onz):.3e})"
    else:
        assert rel_err(nz, onz, S) <= RTOL, tier_report(nz, onz, S)
    # reassembly on the frozen pattern with another factor (skip_preps = true, solvers.jl:556)
    cp2, rv2, nz2 = G.assemble_csc(AP, 0.5 * factor, skip_preps=True)
    assert cp2 is cp or np.array_equal(cp2, cp)
    assert rel_err(nz2, 0.5 * onz, None if S is None else 0.5 * S) <= RTOL
    return cp, rv, nz


CASES = [
    # (name, grid fn, fetype ctor, operator pair, apt)
    ("P1 tri Laplace", lambda: tri_grid(3), lambda g: G.H1P1(1), (G.Gradient, G.Gradient), "sym"),
    ("P2 tri Laplace (C1)", lambda: tri_grid(4), lambda g: G.H1P2(1, 2), (G.Gradient, G.Gradient), "sym"),
    ("P2 tri mass", lambda: tri_grid(3), lambda g: G.H1P2(1, 2), (G.Identity, G.Identity), "sym"),
    ("P1 tet Laplace (C2)", lambda: tet_grid(2), lambda g: G.H1P1(1), (G.Gradient, G.Gradient), "sym"),
    ("P2 tet Laplace (C2*)", lambda: tet_grid(2), lambda g: G.H1P2(1, 3), (G.Gradient, G.Gradient), "sym"),
    ("P2 tet Laplace perturbed", lambda: tet_grid(2, True), lambda g: G.H1P2(1, 3), (G.Gradient, G.Gradient), "sym"),
    ("P2 tet mass", lambda: tet_grid(1), lambda g: G.H1P2(1, 3), (G.Identity, G.Identity), "sym"),
    ("P2 tet general BLF", lambda: tet_grid(1), lambda g: G.H1P2(1, 3), (G.Gradient, G.Gradient), "gen"),
    ("P1 tri lumped mass", lambda: tri_grid(2), lambda g: G.H1P1(1), (G.Identity, G.Identity), "lump"),
    ("RT0 tri mass", lambda: tri_grid(3), lambda g: G.HDIVRT0(2), (G.Identity, G.Identity), "sym"),
    ("BDM1 tri mass", lambda: tri_grid(3, True), lambda g: G.HDIVBDM1(2), (G.Identity, G.Identity), "sym"),
    ("RT0 tet mass (C5)", lambda: tet_grid(1), lambda g: G.HDIVRT0(3), (G.Identity, G.Identity), "sym"),
    ("BDM1 tet mass (C5)", lambda: tet_grid(1, True), lambda g: G.HDIVBDM1(3), (G.Identity, G.Identity), "sym"),
    ("RT0 tet div-div", lambda: tet_grid(1), lambda g: G.HDIVRT0(3), (G.Divergence, G.Divergence), "sym"),
    ("BR tri Laplace (C4)", lambda: tri_grid(3), lambda g: G.H1BR(2), (G.Gradient, G.Gradient), "sym"),
    ("BR tet Laplace", lambda: tet_grid(1, True), lambda g: G.H1BR(3), (G.Gradient, G.Gradient), "sym"),
    ("BR tri mass", lambda: tri_grid(2), lambda g: G.H1BR(2), (G.Identity, G.Identity), "sym"),
    ("BR tri recon RT0 mass (C4)", lambda: tri_grid(2), lambda g: G.H1BR(2),
     (G.ReconstructionIdentity(G.HDIVRT0(2)), G.ReconstructionIdentity(G.HDIVRT0(2))), "sym"),
    ("BR tri recon BDM1 mass (C4)", lambda: tri_grid(2, True), lambda g: G.H1BR(2),
     (G.ReconstructionIdentity(G.HDIVBDM1(2)), G.ReconstructionIdentity(G.HDIVBDM1(2))), "sym"),
    ("BR tet recon BDM1 mass", lambda: tet_grid(0), lambda g: G.H1BR(3),
     (G.ReconstructionIdentity(G.HDIVBDM1(3)), G.ReconstructionIdentity(G.HDIVBDM1(3))), "sym"),
]


@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_blf_parity_generic(case):
    _, gridf, fef, ops, apt = case
    g = gridf()
    s = G.FESpace(fef(g), g)
    ctor = {"sym": G.DiscreteSymmetricBilinearForm, "gen": G.DiscreteBilinearForm, "lump": G.DiscreteLumpedBilinearForm}[apt]
    AP = ctor(list(ops), [s, s])
    check_blf(AP, factor=1.0, exact=True, path=G._lib.PATH_GENERIC)


def test_laplace_kappa_and_regions():
    g = tri_grid(3)
    g.cellregions[::3] = 2
    s = G.FESpace(G.H1P2(1, 2), g)
    AP = G.DiscreteSymmetricBilinearForm([G.Gradient, G.Gradient], [s, s], regions=[2])
    check_blf(AP, factor=1e-3, path=G._lib.PATH_GENERIC)


def test_hooke2d_parity():
    g = tri_grid(3, True)
    s = G.FESpace(G.H1P2(2, 2), g)
    mu = 1000 / 1.4
    lam = 0.4 * mu / 0.2
    AP = G.DiscreteBilinearForm([G.SymmetricGradient(1), G.SymmetricGradient(1)], [s, s], G.HookeAction(2, mu, lam))
    check_blf(AP, path=G._lib.PATH_GENERIC)
    g2 = tri_grid(2)          # axis-aligned: pattern hinges on exact zeros
    s2 = G.FESpace(G.H1P1(2), g2)
    AP2 = G.DiscreteBilinearForm([G.SymmetricGradient(1), G.SymmetricGradient(1)], [s2, s2], G.HookeAction(2, mu, lam))
    check_blf(AP2, path=G._lib.PATH_GENERIC)


def test_hooke3d_parity():
    g = tet_grid(1)
    s = G.FESpace(G.H1P1(3), g)
    AP = G.DiscreteBilinearForm([G.SymmetricGradient(1), G.SymmetricGradient(1)], [s, s], G.HookeAction(3, 2.0, 3.0))
    check_blf(AP, path=G._lib.PATH_GENERIC)


@pytest.mark.parametrize("dim", [2, 3])
def test_stokes_divergence_block_with_transpose_copy(dim):
    g = tri_grid(3) if dim == 2 else tet_grid(1)
    sv = G.FESpace(G.H1BR(dim), g)
    sp = G.FESpace(G.L2P0(1), g)
    A = G.FEMatrix([sv, sp])
    O_ = G.LagrangeMultiplier(G.Divergence)
    G.assemble_operator(A[1, 2], O_, At=A[2, 1])
    (ocp, orv, onz), (tcp, trv, tnz) = oracle_blf(O_._pattern, -1.0, transpose_copy=True)
    # compare the two blocks of the device-assembled FEMatrix with the oracle's B and transpose copy
    M = A.tocsc().toarray() if A.m < 3000 else None
    import scipy.sparse as sp_
    B = sp_.csc_matrix((onz, orv - 1, ocp - 1), shape=(sv.ndofs, sp.ndofs))
    Bt = sp_.csc_matrix((tnz, trv - 1, tcp - 1), shape=(sp.ndofs, sv.ndofs))
    full = A.tocsc()
    assert abs(full[: sv.ndofs, sv.ndofs:] - B).max() == 0
    assert abs(full[sv.ndofs:, : sv.ndofs] - Bt).max() == 0
    assert full.nnz == B.nnz + Bt.nnz


def test_rectangular_p2_p1_divergence_transposed_assembly():
    g = tri_grid(2, True)
    su = G.FESpace(G.H1P2(2, 2), g)
    sp = G.FESpace(G.H1P1(1), g)
    AP = G.DiscreteBilinearForm([G.Divergence, G.Identity], [su, sp])
    cp, rv, nz = G.assemble_csc(AP, 1.0)
    ocp, orv, onz = oracle_blf(AP, 1.0)
    assert np.array_equal(cp, ocp) and np.array_equal(rv, orv) and np.array_equal(nz, onz)
    cpT, rvT, nzT = G.assemble_csc(AP, 1.0, transposed_assembly=True)
    At = O.OracleMatrix(sp.ndofs, su.ndofs)
    O.blf_assemble(At, g, su, sp, O.OP_DIV, O.OP_ID, transposed_assembly=True)
    tcp, trv, tnz = At.csc()
    assert np.array_equal(cpT, tcp) and np.array_equal(rvT, trv) and np.array_equal(nzT, tnz)


LF_CASES = [
    ("P2 tri id f=1 region 1 (C1)", lambda: tri_grid(4), lambda: G.H1P2(1, 2), G.Identity, "const", [1]),
    ("P2 tet id f(x)", lambda: tet_grid(1), lambda: G.H1P2(1, 3), G.Identity, "fun", [0]),
    ("P1 tri id none", lambda: tri_grid(2), lambda: G.H1P1(1), G.Identity, "none", [0]),
    ("RT0 tet id f(x)", lambda: tet_grid(1), lambda: G.HDIVRT0(3), G.Identity, "vfun", [0]),
    ("BDM1 tri id f(x)", lambda: tri_grid(2), lambda: G.HDIVBDM1(2), G.Identity, "vfun", [0]),
    ("BR tri recon RT0 (C4)", lambda: tri_grid(3), lambda: G.H1BR(2), G.ReconstructionIdentity(G.HDIVRT0(2)), "vfun2", [0]),
    ("BR tri recon BDM1 (C4)", lambda: tri_grid(3, True), lambda: G.H1BR(2), G.ReconstructionIdentity(G.HDIVBDM1(2)), "vfun2", [0]),
    ("BR tet recon RT0", lambda: tet_grid(1), lambda: G.H1BR(3), G.ReconstructionIdentity(G.HDIVRT0(3)), "vfun2", [0]),
    ("BR tet recon BDM1", lambda: tet_grid(1, True), lambda: G.H1BR(3), G.ReconstructionIdentity(G.HDIVBDM1(3)), "vfun2", [0]),
]


@pytest.mark.parametrize("case", LF_CASES, ids=[c[0] for c in LF_CASES])
def test_lf_parity(case):
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
    else:   # gradient of x^3+y^3(-1/2): Example222-style right-hand side (bonus 2 -> Stroud rule in 2D)
        data = G.DataFunction(lambda x: np.stack([3 * x[k] ** 2 for k in range(len(x))]), [nc, dim], bonus_quadorder=2)
        bonus = 2
    Op = G.LinearForm(op, data, regions=regions, factor=2.0)
    b = G.FEVector([s])
    b.entries[:] = 0.25            # += semantics
    AP = G.assemble_operator(b[1], Op)
    # oracle on the same quadrature rule and the same host-evaluated table
    qo = G.quadrature_order(AP)
    P = AP.AM
    O.qrule_override(dim, qo, P.qf.xref, P.qf.w)
    try:
        ob = np.full(s.ndofs, 0.25)
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
    assert np.array_equal(b.entries, ob), f"max abs diff {np.abs(b.entries - ob).max():.3e}"


def test_lf_offset_into_block_vector():
    g = tri_grid(2)
    s1 = G.FESpace(G.H1P1(1), g)
    s2 = G.FESpace(G.H1P2(1, 2), g)
    b = G.FEVector([s1, s2])
    G.assemble_operator(b[2], G.LinearForm(G.Identity, G.DataFunction([3.0])))
    assert np.all(b.entries[: s1.ndofs] == 0)
    assert abs(b.entries[s1.ndofs:].sum() - 3.0) < 1e-13


def test_error_behaviour():
    g = tri_grid(1)
    s = G.FESpace(G.H1P1(1), g)
    sp = G.FESpace(G.HDIVRT0(2), g)
    with pytest.raises(G._lib.GrmpError):
        G.assemble_csc(G.DiscreteBilinearForm([G.Gradient, G.Gradient], [sp, sp]))      # Hdiv gradient is not ported
    with pytest.raises(G._lib.GrmpError):
        G.assemble_csc(G.DiscreteBilinearForm([G.Gradient, G.Identity], [s, s]))        # result dims differ
    with pytest.raises(NotImplementedError):
        G.Action(lambda r, i: None, [1, 1])


def test_determinism_two_runs_bit_equal():
    g = tet_grid(2)
    s = G.FESpace(G.H1P2(1, 3), g)
    AP = G.DiscreteSymmetricBilinearForm([G.Gradient, G.Gradient], [s, s])
    _, _, a = G.assemble_csc(AP, 1.0)
    _, _, b = G.assemble_csc(AP, 1.0, skip_preps=True)
    assert np.array_equal(a, b)


# ---- fast path: owner-computes P2-tet Laplace kernel (the metric kernel) -------------------------
@pytest.mark.parametrize("L,perturbed,apt", [(0, False, "sym"), (1, False, "sym"), (2, False, "sym"), (2, True, "sym"),
                                             (3, False, "sym"), (3, True, "sym"), (1, True, "gen")])
def test_fast_p2tet_laplace_parity(L, perturbed, apt):
    g = tet_grid(L, perturbed)
    s = G.FESpace(G.H1P2(1, 3), g)
    ctor = G.DiscreteSymmetricBilinearForm if apt == "sym" else G.DiscreteBilinearForm
    AP = ctor([G.Gradient, G.Gradient], [s, s])
    check_blf(AP, factor=0.75, exact=False, path=G._lib.PATH_FAST)
    st = G.blf_stats(AP)
    assert st.path == G._lib.PATH_FAST and st.kernel_launches == 2 and st.ntiles >= 1


def test_fast_p2tet_matches_generic_and_is_deterministic():
    g = tet_grid(3, True)
    s = G.FESpace(G.H1P2(1, 3), g)
    out = {}
    for path in (G._lib.PATH_GENERIC, G._lib.PATH_FAST):
        AP = G.DiscreteSymmetricBilinearForm([G.Gradient, G.Gradient], [s, s])
        G.blf_set_path(AP, path)
        cp, rv, nz = G.assemble_csc(AP, 2.0)
        _, _, nz2 = G.assemble_csc(AP, 2.0, skip_preps=True)
        assert np.array_equal(nz, nz2), "two runs differ"
        out[path] = (cp, rv, nz)
    a, b = out[G._lib.PATH_GENERIC], out[G._lib.PATH_FAST]
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
    S = oracle_scale(AP, 2.0)
    assert rel_err(b[2], a[2], S) <= RTOL, tier_report(b[2], a[2], S)
    # size-independent properties of a stiffness matrix: A*1 = 0, symmetry
    import scipy.sparse as sp_
    A = sp_.csc_matrix((b[2], b[1] - 1, b[0] - 1), shape=(s.ndofs, s.ndofs))
    assert np.abs(A @ np.ones(s.ndofs)).max() < 1e-11 * np.abs(b[2]).max()
    assert abs(A - A.T).max() < 1e-12 * np.abs(b[2]).max()


@pytest.mark.parametrize("grid", ["uniform L3", "perturbed L3", "delaunay"])
def test_fast_p2tet_device_build_equals_host_build(grid, monkeypatch):
    """the record build of the ring-walk kernel runs on the GPU (ring order per edge column, tiles packed in chunks); the host build
    (GRMP_FAST_HOST_BUILD=1) is kept for cross-validation: same pattern, same values bit for bit (the ring orders are the same, the
    tilings differ only by the cuts at chunk boundaries, and no value depends on the tiling)"""
    if grid == "delaunay":
        g = delaunay_tet_grid(300, 5)
    else:
        g = tet_grid(3, grid.startswith("perturbed"))
    s = G.FESpace(G.H1P2(1, 3), g)
    out = []
    for host in (False, True):
        if host:
            monkeypatch.setenv("GRMP_FAST_HOST_BUILD", "1")
        else:
            monkeypatch.delenv("GRMP_FAST_HOST_BUILD", raising=False)
        AP = G.DiscreteSymmetricBilinearForm([G.Gradient, G.Gradient], [s, s])
        G.blf_set_path(AP, G._lib.PATH_FAST)
        out.append(G.assemble_csc(AP, 1.5))
        assert G.blf_stats(AP).path == G._lib.PATH_FAST
    for a, b in zip(out[0], out[1]):
        assert np.array_equal(a, b)


def test_fast_p2tet_follows_geometry_updates():
    """the fast path keeps tile-blocked copies of the node coordinates: grmp_grid_update_geometry must refresh them"""
    g = tet_grid(2, True)
    s = G.FESpace(G.H1P2(1, 3), g)
    AP = G.DiscreteSymmetricBilinearForm([G.Gradient, G.Gradient], [s, s])
    G.blf_set_path(AP, G._lib.PATH_FAST)
    G.assemble_csc(AP, 1.0)
    # stretch the grid anisotropically (volumes scale by 6, the stiffness entries change non-uniformly)
    g2 = tet_grid(2, True)
    g2.coords[:, 0] *= 2.0
    g2.coords[:, 2] *= 3.0
    vol2 = np.ascontiguousarray(g.cellvolumes * 6.0)
    L = G._lib.lib()
    G._lib.check(L.grmp_grid_update_geometry(G.device_grid(g), G._lib.ptr(np.ascontiguousarray(g2.coords)), G._lib.ptr(vol2)))
    _, _, nz = G.assemble_csc(AP, 1.0, skip_preps=True)
    s2 = G.FESpace(G.H1P2(1, 3), g2)
    AP2 = G.DiscreteSymmetricBilinearForm([G.Gradient, G.Gradient], [s2, s2])
    G.blf_set_path(AP2, G._lib.PATH_GENERIC)
    _, _, ref = G.assemble_csc(AP2, 1.0)
    S = oracle_scale(AP2, 1.0)
    assert rel_err(nz, ref, S) <= RTOL, tier_report(nz, ref, S)


@pytest.mark.parametrize("nw,slot,kb", [(3, 256, 24), (4, 300, 40), (7, 800, 112), (5, 512, 64)])
def test_fast_p2tet_tile_shapes(nw, slot, kb, monkeypatch):
    """tiles / warp groups cut at different places (shared-memory budget, slot size, #consumer warps) give the same matrix"""
    monkeypatch.setenv("GRMP_FAST_NW", str(nw))
    monkeypatch.setenv("GRMP_FAST_SLOT", str(slot))
    monkeypatch.setenv("GRMP_FAST_SMEM_KB", str(kb))
    g = tet_grid(3, True)
    s = G.FESpace(G.H1P2(1, 3), g)
    AP = G.DiscreteSymmetricBilinearForm([G.Gradient, G.Gradient], [s, s])
    check_blf(AP, factor=1.25, exact=False, path=G._lib.PATH_FAST)
    assert G.blf_stats(AP).ntiles > 8


@pytest.mark.parametrize("path", ["fast", "generic"])
def test_assemble_host_one_call_matches_split_calls(path):
    """grmp_blf_assemble_host (uploads + assembly + download in one call, overlapped copies) == update_* + numeric"""
    g = tet_grid(2, True)
    s = G.FESpace(G.H1P2(1, 3), g)
    AP = G.DiscreteSymmetricBilinearForm([G.Gradient, G.Gradient], [s, s])
    G.blf_set_path(AP, G._lib.PATH_FAST if path == "fast" else G._lib.PATH_GENERIC)
    _, _, ref = G.assemble_csc(AP, 1.5)
    L = G._lib.lib()
    nz = np.zeros_like(ref)
    cn, dofs, vol = np.ascontiguousarray(g.cellnodes), np.ascontiguousarray(s.celldofs), np.ascontiguousarray(g.cellvolumes)
    x = np.ascontiguousarray(g.coords)
    G._lib.check(L.grmp_blf_assemble_host(AP.AM.h, 1.5, G._lib.ptr(x), G._lib.ptr(vol), G._lib.ptr(cn), G._lib.ptr(dofs), None, G._lib.ptr(nz)))
    assert np.array_equal(nz, ref)
    # scaled geometry through the same call: entries of the 3D stiffness matrix scale with the length
    x2, vol8 = np.ascontiguousarray(2.0 * x), np.ascontiguousarray(8.0 * vol)
    G._lib.check(L.grmp_blf_assemble_host(AP.AM.h, 1.5, G._lib.ptr(x2), G._lib.ptr(vol8), G._lib.ptr(cn), G._lib.ptr(dofs), None, G._lib.ptr(nz)))
    S = oracle_scale(AP, 1.5)
    assert rel_err(nz, 2.0 * ref, 2.0 * S) <= RTOL
    assert L.grmp_blf_assemble_host(AP.AM.h, 1.5, None, None, None, None, None, None) == -1


def test_fast_p2tet_level5_size_independent_properties():
    """786 432 cells / 2.97e7 non-zeros (one level below the BASELINE configuration): properties that need no reference
    matrix -- A 1 = 0, symmetry, the P1 relation of the vertex block -- plus agreement with the bit-exact generic path"""
    import scipy.sparse as sp_
    g = tet_grid(5)
    s = G.FESpace(G.H1P2(1, 3), g)
    AP = G.DiscreteSymmetricBilinearForm([G.Gradient, G.Gradient], [s, s])
    G.blf_set_path(AP, G._lib.PATH_FAST)
    cp, rv, nz = G.assemble_csc(AP, 1.0)
    _, _, nz2 = G.assemble_csc(AP, 1.0, skip_preps=True)
    assert np.array_equal(nz, nz2)
    assert G.blf_stats(AP).ntiles > 3000
    A = sp_.csc_matrix((nz, rv - 1, cp - 1), shape=(s.ndofs, s.ndofs))
    amax = np.abs(nz).max()
    assert np.abs(A @ np.ones(s.ndofs)).max() < 1e-11 * amax
    assert abs(A - A.T).max() < 1e-12 * amax
    nn = g.nnodes
    Avv = A[:nn, :nn].tocsc()
    assert np.abs(Avv.diagonal() - 3.0 * (np.asarray(Avv.sum(axis=0)).ravel() - Avv.diagonal())).max() < 1e-11 * amax
    APg = G.DiscreteSymmetricBilinearForm([G.Gradient, G.Gradient], [s, s])
    G.blf_set_path(APg, G._lib.PATH_GENERIC)
    cpg, rvg, nzg = G.assemble_csc(APg, 1.0)
    assert np.array_equal(cp, cpg) and np.array_equal(rv, rvg)
    assert rel_err(nz, nzg) <= RTOL, tier_report(nz, nzg)      # axis-aligned grid: no ill-conditioned entries, pure 1e-12


def delaunay_tet_grid(npts, seed):
    """unstructured conforming tetrahedral mesh of random points in the unit cube (scipy Delaunay; degenerate slivers
    removed, cells oriented positively): rings of very different lengths, many boundary (open) edge stars"""
    from scipy.spatial import Delaunay
    rng = np.random.default_rng(seed)
    corners = np.array([[i, j, k] for i in (0., 1.) for j in (0., 1.) for k in (0., 1.)])
    x = np.vstack([corners, rng.uniform(0.0, 1.0, size=(npts, 3))])
    cells = Delaunay(x).simplices.astype(np.int64)
    a, b, c = (x[cells[:, j]] - x[cells[:, 0]] for j in (1, 2, 3))
    det = np.einsum("ij,ij->i", a, np.cross(b, c))
    keep = np.abs(det) > 1e-13           # only exactly degenerate cells (none for these seeds; removing one would open the mesh)
    assert keep.all() or keep.sum() > 0
    cells = cells[keep]; det = det[keep]
    neg = det < 0
    cells[neg, 2], cells[neg, 3] = cells[neg, 3].copy(), cells[neg, 2].copy()
    return G.ExtendableGrid(x, cells + 1)


@pytest.mark.parametrize("npts,seed", [(60, 1), (400, 2), (2500, 3)])
def test_fast_p2tet_unstructured_delaunay_mesh(npts, seed):
    """ragged input: the ring walk on a mesh that is not a uniform refinement (ring lengths 3..15+, boundary chains).  The
    fast path must assemble it to the same matrix as the generic path"""
    g = delaunay_tet_grid(npts, seed)
    s = G.FESpace(G.H1P2(1, 3), g)
    APg = G.DiscreteSymmetricBilinearForm([G.Gradient, G.Gradient], [s, s])
    G.blf_set_path(APg, G._lib.PATH_GENERIC)
    cpg, rvg, nzg = G.assemble_csc(APg, 1.0)
    AP = G.DiscreteSymmetricBilinearForm([G.Gradient, G.Gradient], [s, s])
    G.blf_set_path(AP, G._lib.PATH_P2TET)          # a mesh the ring walk cannot order would raise with the reason
    cp, rv, nz = G.assemble_csc(AP, 1.0)
    assert np.array_equal(cp, cpg) and np.array_equal(rv, rvg)
    assert rel_err(nz, nzg) <= 1e-9                # slivers: the row-sum identities of the ring walk lose digits here ...
    # ... which is why AUTO does not take the ring walk on such meshes (cancellation guard) and still meets the bar
    APa = G.DiscreteSymmetricBilinearForm([G.Gradient, G.Gradient], [s, s])
    G.blf_set_path(APa, G._lib.PATH_AUTO)
    cpa, rva, nza = G.assemble_csc(APa, 1.0)
    assert G.blf_stats(APa).path == G._lib.PATH_COLUMNS
    assert np.array_equal(cpa, cpg) and np.array_equal(rva, rvg)
    S = oracle_scale(APa, 1.0)
    assert rel_err(nza, nzg, S) <= RTOL, tier_report(nza, nzg, S)


def test_fast_p2tet_region_filter_falls_back_correctly():
    g = tet_grid(1)
    g.cellregions[::2] = 2
    s = G.FESpace(G.H1P2(1, 3), g)
    AP = G.DiscreteSymmetricBilinearForm([G.Gradient, G.Gradient], [s, s], regions=[2])
    check_blf(AP, factor=1.0, exact=False, path=G._lib.PATH_AUTO)
    assert G.blf_stats(AP).path == G._lib.PATH_COLUMNS        # the ring walk has no region filter; the column kernels do


def test_auto_path_selection():
    """AUTO: ring walk for the metric form, column kernels for every other form with a kernel, generic otherwise"""
    g = tet_grid(1)
    s = G.FESpace(G.H1P2(1, 3), g)
    AP = G.DiscreteSymmetricBilinearForm([G.Gradient, G.Gradient], [s, s])
    G.blf_set_path(AP, G._lib.PATH_AUTO)
    G.assemble_csc(AP, 1.0)
    assert G.blf_stats(AP).path == G._lib.PATH_P2TET
    AP2 = G.DiscreteSymmetricBilinearForm([G.Identity, G.Identity], [s, s])
    G.blf_set_path(AP2, G._lib.PATH_AUTO)
    G.assemble_csc(AP2, 1.0)
    assert G.blf_stats(AP2).path == G._lib.PATH_COLUMNS
    AP3 = G.DiscreteLumpedBilinearForm([G.Identity, G.Identity], [s, s])          # no column kernel for lumped forms
    G.blf_set_path(AP3, G._lib.PATH_AUTO)
    G.assemble_csc(AP3, 1.0)
    assert G.blf_stats(AP3).path == G._lib.PATH_GENERIC
    with pytest.raises(G._lib.GrmpError):
        G.blf_set_path(AP3, G._lib.PATH_FAST)
        G.assemble_csc(AP3, 1.0)


def test_partitioned_assembly_matches_global():
    """multi-GPU path on one device: every 'rank' assembles its owned columns (fast path restricted by
    grmp_blf_set_owned_columns), the merged column blocks equal the single-rank matrix"""
    g = tet_grid(2, True)
    s = G.FESpace(G.H1P2(1, 3), g)
    AP = G.DiscreteSymmetricBilinearForm([G.Gradient, G.Gradient], [s, s])
    G.blf_set_path(AP, G._lib.PATH_GENERIC)
    cp, rv, nz = G.assemble_csc(AP, 1.0)
    world = 3
    blocks = []
    for r in range(world):
        lp = G.partition.partition(s, r, world)
        APr = G.DiscreteSymmetricBilinearForm([G.Gradient, G.Gradient], [lp.space, lp.space])
        G.prepare_assembly(APr)
        G.blf_set_path(APr, G._lib.PATH_AUTO)
        G._lib.check(G._lib.lib().grmp_blf_set_owned_columns(APr.AM.h, lp.n_owned))
        lcp, lrv, lnz = G.assemble_csc(APr, 1.0, skip_preps=True)
        assert G.blf_stats(APr).path == G._lib.PATH_FAST
        blocks.append(G.partition.owned_block_to_global(lp, lcp, lrv, lnz))
    mcp, mrv, mnz = G.partition.merge_owned_columns(s.ndofs, blocks)
    assert np.array_equal(mcp, cp) and np.array_equal(mrv, rv)
    S = oracle_scale(AP, 1.0)
    assert rel_err(mnz, nz, S) <= RTOL, tier_report(mnz, nz, S)


import glob
import os

GOLDEN = sorted(f for f in glob.glob(os.path.join(os.path.dirname(__file__), "golden", "*.npz")) if not os.path.basename(f).startswith("sg_"))


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p) for p in GOLDEN])
def test_gpu_against_committed_golden_fixtures(path):
    from golden.make_golden import CASES as GC, build_case
    d = np.load(path)
    name = os.path.basename(path)[:-4]
    grid, space, AP, factor = build_case(G, GC[name])
    G.blf_set_path(AP, G._lib.PATH_GENERIC)
    cp, rv, nz = G.assemble_csc(AP, factor)
    assert np.array_equal(cp, d["colptr"]) and np.array_equal(rv, d["rowval"]) and np.array_equal(nz, d["nzval"])


# ---- edge cases -------------------------------------------------------------------------------------
def test_region_filter_selecting_nothing_gives_empty_pattern():
    g = tri_grid(2)
    s = G.FESpace(G.H1P1(1), g)
    AP = G.DiscreteSymmetricBilinearForm([G.Gradient, G.Gradient], [s, s], regions=[7])
    cp, rv, nz = G.assemble_csc(AP, 1.0)
    assert rv.size == 0 and nz.size == 0 and np.all(cp == 1)
    ocp, orv, onz = oracle_blf(AP, 1.0)
    assert np.array_equal(cp, ocp) and orv.size == 0
    b = G.FEVector([s])
    G.assemble_operator(b[1], G.LinearForm(G.Identity, G.DataFunction([1.0]), regions=[7]))
    assert np.all(b.entries == 0)


@pytest.mark.parametrize("geo", ["Triangle2D", "Tetrahedron3D"])
def test_single_cell_grids_all_elements(geo):
    g = G.reference_domain(geo)
    dim = g.dim
    fes = [G.H1P1(1), G.H1P2(1, dim), G.H1P2(dim, dim), G.H1BR(dim), G.HDIVRT0(dim), G.HDIVBDM1(dim), G.L2P0(1)]
    for fe in fes:
        s = G.FESpace(fe, g)
        AP = G.DiscreteSymmetricBilinearForm([G.Identity, G.Identity], [s, s])
        check_blf(AP, factor=1.0, exact=True)
    s = G.FESpace(G.H1P2(1, dim), g)
    check_blf(G.DiscreteSymmetricBilinearForm([G.Gradient, G.Gradient], [s, s]), factor=3.0, exact=(dim == 2))


def test_zero_factor_gives_empty_pattern_like_addnz():
    # _addnz skips v == 0: with factor 0 nothing is ever inserted (fematrix.jl:54-58)
    g = tri_grid(1)
    s = G.FESpace(G.H1P1(1), g)
    AP = G.DiscreteSymmetricBilinearForm([G.Identity, G.Identity], [s, s])
    cp, rv, nz = G.assemble_csc(AP, 0.0)
    assert rv.size == 0
    ocp, orv, _ = oracle_blf(AP, 0.0)
    assert orv.size == 0 and np.array_equal(cp, ocp)


def test_c_abi_error_codes():
    import ctypes as C
    L = G._lib.lib()
    g = tri_grid(1)
    s = G.FESpace(G.H1P1(1), g)
    AP = G.DiscreteSymmetricBilinearForm([G.Gradient, G.Gradient], [s, s])
    G.prepare_assembly(AP)
    h = AP.AM.h
    assert L.grmp_blf_numeric(h, 1.0, None) == -5                 # GRMP_ESTATE: numeric before symbolic
    assert b"symbolic" in L.grmp_last_error()
    assert L.grmp_blf_get_pattern(h, None, None) == -1            # GRMP_EINVAL
    assert L.grmp_blf_set_path(h, 9) == -1
    hg = C.c_void_p()
    assert L.grmp_grid_create(G._lib.context(), 4, 0, None, 0, None, None, None, C.byref(hg)) == -1
    assert L.grmp_init(99, C.byref(hg)) == -1                     # device index out of range
    st = G._lib.Stats()
    assert L.grmp_blf_stats(h, C.byref(st)) == 0


def test_large_values_and_tiny_cells_scale_linearly():
    g = tet_grid(1)
    s = G.FESpace(G.H1P2(1, 3), g)
    AP = G.DiscreteSymmetricBilinearForm([G.Gradient, G.Gradient], [s, s])
    cp, rv, nz1 = G.assemble_csc(AP, 1.0)
    _, _, nz2 = G.assemble_csc(AP, 1e12, skip_preps=True)
    assert rel_err(nz2, 1e12 * nz1) <= RTOL
