"""GPU tests of the 'next' rows N1 / N3: the matrix stays on the device after assemble! -- products
(addblock_matmul!, fematrix.jl:402-473), the residual check of solve_direct! (solvers.jl:661-668), Dirichlet penalties
(apply_penalties!, fematrix.jl:349-355) and the device CSC hand-off; plus the stale-pattern protection of the ABI."""
import ctypes as C

import numpy as np
import pytest

import grmp_b200 as G
from parity import oracle_blf
from test_gpu_parity import tet_grid, tri_grid

pytestmark = pytest.mark.gpu


def ref_matmul(cp, rv, nz, b, a, factor, transposed):
    """the reference's loop, literally (fematrix.jl:446-470): columns ascending, rows ascending, one term at a time"""
    a = a.copy()
    for col in range(cp.size - 1):
        for r in range(cp[col] - 1, cp[col + 1] - 1):
            row = rv[r] - 1
            if transposed:
                a[col] += nz[r] * b[row] * factor
            else:
                a[row] += nz[r] * b[col] * factor
    return a


@pytest.mark.parametrize("case", ["P2 tri Laplace", "BR x P0 divergence", "Hooke P1 tet"])
@pytest.mark.parametrize("path", [G._lib.PATH_GENERIC, G._lib.PATH_AUTO])
def test_matmul_bit_identical_to_reference_loop(case, path):
    rng = np.random.default_rng(7)
    if case == "P2 tri Laplace":
        g = tri_grid(3, True)
        s = G.FESpace(G.H1P2(1, 2), g)
        AP = G.DiscreteSymmetricBilinearForm([G.Gradient, G.Gradient], [s, s])
    elif case == "BR x P0 divergence":
        g = tri_grid(3, True)
        AP = G.DiscreteBilinearForm([G.Divergence, G.Identity], [G.FESpace(G.H1BR(2), g), G.FESpace(G.L2P0(1), g)])
    else:
        g = tet_grid(1, True)
        s = G.FESpace(G.H1P1(3), g)
        AP = G.DiscreteBilinearForm([G.SymmetricGradient(1), G.SymmetricGradient(1)], [s, s], G.HookeAction(3, 2.0, 3.0))
    G.blf_set_path(AP, path)
    cp, rv, nz = G.assemble_csc(AP, 1.0)
    nrows, ncols = AP.FES[0].ndofs, AP.FES[1].ndofs
    for transposed in (False, True):
        b = rng.standard_normal(nrows if transposed else ncols)
        a0 = rng.standard_normal(ncols if transposed else nrows)
        for factor in (1.0, -0.37):
            a = a0.copy()
            G.addblock_matmul(a, AP, b, factor=factor, transposed=transposed)
            assert np.array_equal(a, ref_matmul(cp, rv, nz, b, a0, factor, transposed))
    # against scipy on the ORACLE's matrix (independent of our nzval)
    import scipy.sparse as sp_
    ocp, orv, onz = oracle_blf(AP, 1.0)
    A = sp_.csc_matrix((onz, orv - 1, ocp - 1), shape=(nrows, ncols))
    x = rng.standard_normal(ncols)
    y = np.zeros(nrows)
    G.addblock_matmul(y, AP, x)
    assert np.abs(y - A @ x).max() <= 1e-12 * (np.abs(A) @ np.abs(x)).max()


def test_residual_and_penalties_like_solve_direct():
    """Poisson problem with homogeneous Dirichlet data assembled, penalised and checked entirely on the device; the solve
    itself (outside the path) is done by scipy on the downloaded matrix"""
    import scipy.sparse as sp_
    import scipy.sparse.linalg as spl
    g = tri_grid(4)
    s = G.FESpace(G.H1P2(1, 2), g)
    AP = G.DiscreteSymmetricBilinearForm([G.Gradient, G.Gradient], [s, s])
    cp, rv, _ = G.assemble_csc(AP, 1.0, fetch=False)
    b = G.FEVector([s])
    G.assemble_operator(b[1], G.LinearForm(G.Identity, G.DataFunction([1.0])))
    # boundary dofs of the unit square: nodes and edge midpoints on the boundary
    x = g.coords
    onb = lambda p: (np.abs(p[:, 0]) < 1e-12) | (np.abs(p[:, 0] - 1) < 1e-12) | (np.abs(p[:, 1]) < 1e-12) | (np.abs(p[:, 1] - 1) < 1e-12)  # noqa: E731
    dofx = np.zeros((s.ndofs, 2))
    dofs = s.celldofs.astype(np.int64) - 1
    cn = g.cellnodes.astype(np.int64) - 1
    for k in range(3):
        dofx[dofs[:, k]] = x[cn[:, k]]
        dofx[dofs[:, 3 + k]] = 0.5 * (x[cn[:, k]] + x[cn[:, (k + 1) % 3]])
    fixed = np.nonzero(onb(dofx))[0] + 1
    penalty = 1e60
    G.apply_penalties(AP, fixed, penalty)
    rhs = b.entries.copy()
    rhs[fixed - 1] = penalty * 0.0
    nz = G.fetch_values(AP)
    A = sp_.csc_matrix((nz, rv - 1, cp - 1), shape=(s.ndofs, s.ndofs))
    assert np.all(A.diagonal()[fixed - 1] == penalty)
    u = spl.spsolve(A, rhs)
    r, n2 = G.residual(AP, u, rhs, fixed)
    assert np.all(r[fixed - 1] == 0)
    assert np.sqrt(n2) <= 1e-12 * np.abs(rhs).max() * np.sqrt(s.ndofs)
    assert abs(n2 - float(np.sum(r * r))) <= 1e-12 * max(n2, 1e-300)
    assert 0.06 < u.max() < 0.08                     # max of the torsion function on the unit square: 0.0737
    # device CSC hand-off
    d = G.device_csc(AP)
    assert (d.nrows, d.ncols, d.nnz) == (s.ndofs, s.ndofs, rv.size) and d.colptr and d.rowval and d.nzval


def test_penalty_on_missing_diagonal_is_refused():
    g = tri_grid(2)
    AP = G.DiscreteBilinearForm([G.Divergence, G.Identity], [G.FESpace(G.H1BR(2), g), G.FESpace(G.L2P0(1), g)])
    G.assemble_csc(AP, 1.0, fetch=False)
    before = G.fetch_values(AP)
    miss = C.c_int64(0)
    fd = np.array([1, 2], dtype=np.int64)
    rc = G._lib.lib().grmp_blf_apply_penalties(AP.AM.h, G._lib.ptr(fd), 2, 1e60, C.byref(miss))
    assert rc == -2 and miss.value >= 1 and b"diagonal" in G._lib.lib().grmp_last_error()
    assert G.fetch_values(AP).shape == before.shape


def test_stale_pattern_is_detected():
    """uploading another topology after the symbolic pass must not silently scatter through the old maps"""
    L = G._lib.lib()
    g = tri_grid(2, True)
    s = G.FESpace(G.H1P1(1), g)
    AP = G.DiscreteSymmetricBilinearForm([G.Gradient, G.Gradient], [s, s])
    _, _, nz = G.assemble_csc(AP, 1.0)
    h = AP.AM.h
    cn = np.ascontiguousarray(g.cellnodes)
    dofs = np.ascontiguousarray(s.celldofs)
    x, vol = np.ascontiguousarray(g.coords), np.ascontiguousarray(g.cellvolumes)
    out = np.zeros_like(nz)
    # unchanged topology handed over: accepted, same matrix
    G._lib.check(L.grmp_blf_assemble_host(h, 1.0, G._lib.ptr(x), G._lib.ptr(vol), G._lib.ptr(cn), G._lib.ptr(dofs), None, G._lib.ptr(out)))
    assert np.array_equal(out, nz)
    # geometry only (topology trusted)
    G._lib.check(L.grmp_blf_assemble_host(h, 1.0, G._lib.ptr(x), G._lib.ptr(vol), None, None, None, G._lib.ptr(out)))
    assert np.array_equal(out, nz)
    # a different dof map in the same call: refused
    dofs2 = dofs.copy()
    dofs2[0, 0], dofs2[0, 1] = dofs[0, 1], dofs[0, 0]
    assert L.grmp_blf_assemble_host(h, 1.0, G._lib.ptr(x), G._lib.ptr(vol), G._lib.ptr(cn), G._lib.ptr(dofs2), None, G._lib.ptr(out)) == -5
    assert b"stale" in L.grmp_last_error()
    # re-uploaded dof map: numeric calls refuse until the symbolic pass has run again
    G._lib.check(L.grmp_space_update_dofs(G.device_space(s), G._lib.ptr(dofs2)))
    assert L.grmp_blf_numeric(h, 1.0, None) == -5
    nnz = C.c_int64(0)
    G._lib.check(L.grmp_blf_symbolic(h, 1.0, C.byref(nnz)))
    G._lib.check(L.grmp_blf_numeric(h, 1.0, G._lib.ptr(out)))


def test_transpose_copy_on_the_column_path():
    g = tri_grid(3, True)
    sv = G.FESpace(G.H1BR(2), g)
    sp = G.FESpace(G.L2P0(1), g)
    outs = {}
    for path in (G._lib.PATH_GENERIC, G._lib.PATH_COLUMNS):
        G.assembly.DEFAULT_PATH = path
        try:
            A = G.FEMatrix([sv, sp])
            G.assemble_operator(A[1, 2], G.LagrangeMultiplier(G.Divergence), At=A[2, 1])
            outs[path] = A.tocsc()
        finally:
            G.assembly.DEFAULT_PATH = G._lib.PATH_AUTO
    a, b = outs[G._lib.PATH_GENERIC], outs[G._lib.PATH_COLUMNS]
    assert np.array_equal(a.indptr, b.indptr) and np.array_equal(a.indices, b.indices)
    assert np.abs(a.data - b.data).max() <= 1e-12 * np.abs(a.data).max()
    assert abs(b[: sv.ndofs, sv.ndofs:] + b[sv.ndofs:, : sv.ndofs].T).max() <= 1e-15 * np.abs(a.data).max()   # B and -B^T


def test_handles_are_released():
    """device objects must not accumulate (a level-6 pattern holds ~10 GB)"""
    import gc
    import torch
    g = tet_grid(3)
    s = G.FESpace(G.H1P2(1, 3), g)
    AP = G.DiscreteSymmetricBilinearForm([G.Gradient, G.Gradient], [s, s])
    G.assemble_csc(AP, 1.0, fetch=False)
    torch.cuda.synchronize()
    free0 = torch.cuda.mem_get_info()[0]
    for _ in range(6):
        G.prepare_assembly(AP)            # replaces the device pattern
        G.assemble_csc(AP, 1.0, fetch=False)
    gc.collect()
    torch.cuda.synchronize()
    free1 = torch.cuda.mem_get_info()[0]
    assert free0 - free1 < 64 * 2**20, (free0, free1)
