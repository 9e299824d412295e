"""Value tolerance of the parity tests (BASELINE.json north_star: every stored value within 1e-12 relative).

With amax = max|A_ref| every stored entry is held to
    |ref| <= 1e-13 * amax :  |val - ref| <= 1e-15 * amax         explicit zeros: the reference stores the rounding residue of
                                                                  exactly cancelling contributions (fematrix.jl:54-65 keeps them)
    otherwise             :  |val - ref| <= 1e-12 * |ref|         the bar: pure relative error ...
... unless the entry is small BY CANCELLATION: S_ij = sum of |w_q a_k b_k| over all products that make up the entry (the oracle's
magnitude mode, oracle/grmp_oracle.cpp) bounds the rounding error of any evaluation order by (#ops) * eps * S_ij, and the
reference's own value carries that much noise.  On the BASELINE grids (uniform_refine of the unit square / cube) S/|ref| <= 14
for every entry, so the pure relative bar applies to all of them; on the jittered grids ~0.4 % of the entries of a P2 stiffness
matrix have S/|ref| between 1e3 and 1e5 (they are exact zeros of the unperturbed grid), and no summation order other than the
reference's reproduces their trailing digits.  Those entries (S > 16 |ref|) are held to 1e-12 * S / 16 = 6.25e-14 * S instead --
tighter, relative to the data that went into them, than 1e-12 relative is for a well-conditioned entry.
`rel_err(val, ref)` without S applies the pure relative bar to everything above the explicit-zero tier.
"""
import numpy as np

import grmp_b200 as G
import oracle as O

TIER_SPLIT = 1e-13
TIER2_ABS = 1e-15
COND_OK = 16.0

_APT = {G.assembly.APT_BilinearForm: O.APT_GENERAL, G.assembly.APT_SymmetricBilinearForm: O.APT_SYMMETRIC,
        G.assembly.APT_LumpedBilinearForm: O.APT_LUMPED}


def oracle_blf(AP, factor, transpose_copy=False, **kw):
    """oracle assemble! on the same quadrature table the host hands to the library (the eigen-generated
    Stroud rules agree between generators only to rounding; hard-coded rules are identical)"""
    s1, s2 = AP.item_space(0), AP.item_space(1)       # ON_BFACES: the boundary-face views (BFaceDofs, BFaceVolumes, BFaceRegions)
    A = O.OracleMatrix(s2.ndofs, s1.ndofs) if kw.get("transposed_assembly") else O.OracleMatrix(s1.ndofs, s2.ndofs)
    At = O.OracleMatrix(s2.ndofs, s1.ndofs) if transpose_copy else None
    act = AP.action
    dim = s1.xgrid.dim
    qo = G.quadrature_order(AP)
    qf = G.QuadratureRule({1: "Edge1D", 2: "Triangle2D", 3: "Tetrahedron3D"}[dim], qo)
    O.qrule_override(dim, qo, qf.xref, qf.w)
    try:
        O.blf_assemble(A, s1.xgrid, s1, s2, AP.operators[0].code, AP.operators[1].code, action=act.code, act_params=act.params,
                       apt=_APT[AP.APT], regions=AP.regions, factor=factor, transpose_copy=At, bonus_quadorder=act.bonus_quadorder, **kw)
    finally:
        O.qrule_override(dim, qo)
    return (A.csc(), At.csc()) if transpose_copy else A.csc()


def oracle_scale(AP, factor, **kw):
    """S_ij of every stored entry (same pattern as oracle_blf)"""
    O.set_magnitude_mode(True)
    try:
        return oracle_blf(AP, factor, **kw)[2]
    finally:
        O.set_magnitude_mode(False)


def tiers(val, ref, S=None):
    """(max relative error over the well-conditioned entries, max |err|/amax over the explicit zeros, max |err|/S over the
    entries that are small by cancellation, counts)"""
    if ref.size == 0:
        return 0.0, 0.0, 0.0, (0, 0, 0)
    amax = max(float(np.abs(ref).max()), 1e-300)
    err = np.abs(val - ref)
    zero = np.abs(ref) <= TIER_SPLIT * amax
    ill = np.zeros_like(zero) if S is None else (~zero & (np.abs(S) > COND_OK * np.abs(ref)))
    good = ~zero & ~ill
    r1 = float((err[good] / np.abs(ref[good])).max()) if good.any() else 0.0
    r2 = float(err[zero].max() / amax) if zero.any() else 0.0
    r3 = float((err[ill] / np.abs(S[ill])).max()) if ill.any() else 0.0
    return r1, r2, r3, (int(good.sum()), int(zero.sum()), int(ill.sum()))


def rel_err(val, ref, S=None):
    """<= 1e-12 iff every tier holds"""
    r1, r2, r3, _ = tiers(val, ref, S)
    return max(r1, r2 * (1e-12 / TIER2_ABS), r3 * COND_OK)


def tier_report(val, ref, S=None):
    r1, r2, r3, (n1, n2, n3) = tiers(val, ref, S)
    return (f"{n1} entries max rel {r1:.3e}; {n2} explicit zeros max abs/amax {r2:.3e}; "
            f"{n3} entries small by cancellation (S > {COND_OK:g}|ref|) max err/S {r3:.3e}")
