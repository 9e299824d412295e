"""Dirichlet boundary data (host mirror of src/boundarydata.jl:60-420) for H1P1 / H1P2 spaces (Identity trace) and HDIVRT0 / HDIVBDM1
spaces (NormalFlux trace, best approximation and homogeneous data).

`boundarydata(Target, O)` fills the boundary dofs of `Target` and returns `fixed_dofs` (1-based), in the reference's order:
interpolated regions, homogeneous regions, best-approximation regions.  What it costs on the device is the best-approximation
part -- the boundary mass matrix and right-hand side are ON_BFACES patterns assembled by libgrmp_cuda
(boundarydata.jl:313-326); dof enumeration, the interpolation of the (few) boundary dofs and the solve of the compressed
boundary system (boundarydata.jl:356-382, a direct solve in the reference as well) are host work.  The solver then fixes the
dofs with `apply_penalties` (fematrix.jl:349-355, solvers.jl:632-652) on the device-resident matrix.
"""
from __future__ import annotations

import numpy as np

from .assembly import (DataFunction, DiscreteLinearForm, DiscreteSymmetricBilinearForm, Identity, NormalFlux, assemble, assemble_csc, fdot_action,
                       fdotn_action)
from .fedefs import H1P1, H1P2, HDIVBDM1, HDIVRT0
from .fespace import FEVector, FEVectorBlock
from .quadrature import QuadratureRule

HomogeneousDirichletBoundary = "HomogeneousDirichletBoundary"
InterpolateDirichletBoundary = "InterpolateDirichletBoundary"
BestapproxDirichletBoundary = "BestapproxDirichletBoundary"


class BoundaryData:
    """BoundaryData(BDT; data, regions) (boundarydata.jl:41-58); component masks stay with the reference"""

    def __init__(self, btype, data=None, regions=(0,)):
        assert btype in (HomogeneousDirichletBoundary, InterpolateDirichletBoundary, BestapproxDirichletBoundary)
        assert btype == HomogeneousDirichletBoundary or isinstance(data, DataFunction)
        self.btype, self.data, self.bregions = btype, data, [int(r) for r in regions]
        self.bdofs = np.zeros(0, np.int64)


def _unique_in_order(a):
    """Base.unique: first occurrences, order kept"""
    a = np.asarray(a, dtype=np.int64)
    _, first = np.unique(a, return_index=True)
    return a[np.sort(first)]


def _data_at(data, x):
    """values [npts, ncomp] of a DataFunction at points x [npts, dim]"""
    if data.constant is not None:
        return np.broadcast_to(data.constant, (x.shape[0], data.constant.size))
    return np.asarray(data.kernel(x.T), dtype=np.float64).reshape(-1, x.shape[0]).T


def _interpolate_bfaces(Target, FES, data, bfaces):
    """interpolate!(Target, ON_BFACES, data; items) for H1P1 (h1_p1.jl:32-62: point evaluation at the nodes) and H1P2
    (h1_p2.jl:48-106: nodes, then every edge dof such that the edge mean of the data is preserved, interpolations.jl:176-240)"""
    g, nc = FES.xgrid, FES.ncomponents
    ent, off = Target.entries, Target.offset
    bn = g.bfacenodes.astype(np.int64)[bfaces] - 1
    nodes = np.unique(bn.ravel())
    vals = _data_at(data, g.coords[nodes])
    for c in range(nc):
        ent[off + c * FES.coffset + nodes] = vals[:, c]
    if not isinstance(FES.fetype, H1P2):
        return
    if g.dim == 2:
        edges = g.bfacefaces.astype(np.int64)[bfaces] - 1
        en = g.facenodes.astype(np.int64)[edges] - 1
    else:
        edges = np.unique(g.bfaceedges.astype(np.int64)[bfaces].ravel() - 1)
        en = g.edgenodes.astype(np.int64)[edges] - 1
    qf = QuadratureRule("Edge1D", data.bonus_quadorder + 1)
    xa, xb = g.coords[en[:, 0]], g.coords[en[:, 1]]
    mean = np.zeros((edges.size, nc))
    for i in range(len(qf)):
        mean += qf.w[i] * _data_at(data, xa + qf.xref[i, 0] * (xb - xa))[:, :nc]
    for c in range(nc):
        u1 = ent[off + c * FES.coffset + en[:, 0]]
        u2 = ent[off + c * FES.coffset + en[:, 1]]
        ent[off + c * FES.coffset + g.nnodes + edges] = 1.5 * (mean[:, c] - (u1 + u2) / 6.0)


def boundarydata(Target: FEVectorBlock, O, fixed_penalty=1e60):
    """boundarydata!(Target, O; fixed_penalty) -> fixed_dofs (1-based, the reference's order)"""
    FES = Target.FES
    hdiv = isinstance(FES.fetype, (HDIVRT0, HDIVBDM1))
    if not isinstance(FES.fetype, (H1P1, H1P2, HDIVRT0, HDIVBDM1)) or FES.broken:
        raise NotImplementedError("boundary data on the device path: H1P1 / H1P2 (Identity trace) and HDIVRT0 / HDIVBDM1 (NormalFlux); others stay "
                                  "with the reference")
    if hdiv and any(bd.btype == InterpolateDirichletBoundary for bd in O):
        raise NotImplementedError("interpolated Hdiv boundary data (face moments, hdiv_rt0.jl:37-52) stays with the reference; use the best approximation")
    Dbop = NormalFlux if hdiv else Identity            # DefaultDirichletBoundaryOperator4FE (boundarydata.jl:27-29)
    g = FES.xgrid
    bdm = FES.bfacedofs.astype(np.int64)
    breg = g.bfaceregions
    fixed = np.zeros(0, np.int64)

    def dofs_of(regions):
        sel = np.flatnonzero(np.isin(breg, regions))
        return sel, _unique_in_order(bdm[sel].ravel())

    for bd in O:                                            # boundarydata.jl:100-200
        if bd.btype == InterpolateDirichletBoundary:
            sel, bd.bdofs = dofs_of(bd.bregions)
            fixed = _unique_in_order(np.concatenate([fixed, bd.bdofs]))
            if sel.size:
                _interpolate_bfaces(Target, FES, bd.data, sel)
    for bd in O:                                            # boundarydata.jl:205-258
        if bd.btype == HomogeneousDirichletBoundary:
            _, bd.bdofs = dofs_of(bd.bregions)
            fixed = _unique_in_order(np.concatenate([fixed, bd.bdofs]))
            Target.entries[Target.offset + bd.bdofs - 1] = 0.0
    ba = [bd for bd in O if bd.btype == BestapproxDirichletBoundary]
    if ba:                                                  # boundarydata.jl:260-383
        import scipy.sparse as sp
        import scipy.sparse.linalg as spla
        b = FEVector([FES])
        badofs, baregions = np.zeros(0, np.int64), []
        for bd in ba:
            _, bd.bdofs = dofs_of(bd.bregions)
            baregions += bd.bregions
            action = fdotn_action(bd.data, g, bfaces=True) if hdiv else fdot_action(bd.data)          # boundarydata.jl:299-303
            rhs = DiscreteLinearForm([Dbop], [FES], action, regions=bd.bregions, AT="ON_BFACES", name="RHS bnd data bestapprox")
            assemble(b[1], rhs)
            badofs = _unique_in_order(np.concatenate([badofs, bd.bdofs]))
        lhs = DiscreteSymmetricBilinearForm([Dbop, Dbop], [FES, FES], regions=baregions, AT="ON_BFACES", name="LHS bnd data bestapprox")
        cp, rv, nz = assemble_csc(lhs, 1.0)
        A = sp.csc_matrix((nz, rv - 1, cp - 1), shape=(FES.ndofs, FES.ndofs)).tolil()
        rhsv = b.entries.copy()
        for j in fixed - 1:                                 # dofs already set by other boundary conditions
            A[j, j] = A[j, j] + fixed_penalty
            rhsv[j] = Target.entries[Target.offset + j] * fixed_penalty
        A = A.tocsc()
        keep = np.flatnonzero(np.diff(A.indptr) != 0)       # compress: drop the interior dofs (empty columns)
        sol = spla.spsolve(A[keep][:, keep].tocsc(), rhsv[keep])
        Target.entries[Target.offset + keep] = sol
        fixed = _unique_in_order(np.concatenate([fixed, badofs]))
    return fixed
