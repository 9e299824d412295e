"""ctypes wrapper of the CPU oracle (oracle/grmp_oracle.cpp) -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / `--impl reference`
legs may import this module.  It deliberately does not import the product package: grid
and space arguments are duck-typed (attributes `coords, cellnodes, cellvolumes, ...` and
`fetype.code, ncomponents, ndofs, nd_cell, celldofs`).

parity unpinned: the reference is Julia and cannot run in this image; this oracle is a
restatement validated by the reference's analytic known-answer tests (tests/test_oracle_kat.py).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

OP_ID, OP_GRAD, OP_SYMGRAD, OP_DIV, OP_RECON_ID_RT0, OP_RECON_ID_BDM1, OP_NORMALFLUX = 1, 2, 3, 4, 5, 6, 7
ACT_NONE, ACT_HOOKE2D, ACT_HOOKE3D, ACT_CONVECTION = 0, 1, 2, 3
APT_GENERAL, APT_SYMMETRIC, APT_LUMPED = 0, 1, 2
F_NONE, F_CONST, F_QP_TABLE = 0, 1, 2
II_NONE, II_L2NORM, II_L2ERROR = 0, 1, 2


class _Grid(C.Structure):
    _fields_ = [("dim", C.c_int), ("xdim", C.c_int), ("nnodes", C.c_int64), ("ncells", C.c_int64), ("nfaces", C.c_int64),
                ("coords", C.c_void_p), ("cellnodes", C.c_void_p), ("cellvolumes", C.c_void_p), ("cellregions", C.c_void_p),
                ("cellfaces", C.c_void_p), ("cellfacesigns", C.c_void_p), ("cellfaceorient", C.c_void_p),
                ("facenormals", C.c_void_p), ("facevolumes", C.c_void_p)]


class _Space(C.Structure):
    _fields_ = [("fetype", C.c_int), ("ncomp", C.c_int), ("ndofs", C.c_int64), ("nd_cell", C.c_int), ("celldofs", C.c_void_p)]


def build(force: bool = False) -> str:
    so = os.path.join(_HERE, "libgrmp_oracle.so")
    src = os.path.join(_HERE, "grmp_oracle.cpp")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(build())
        _LIB.orc_last_error.restype = C.c_char_p
        _LIB.orc_matrix_create.restype = C.c_void_p
        _LIB.orc_matrix_create.argtypes = [C.c_int64, C.c_int64]
        _LIB.orc_matrix_nnz.restype = C.c_int64
        for f in ("orc_matrix_destroy", "orc_matrix_flush", "orc_matrix_nnz", "orc_matrix_fill_zero"):
            getattr(_LIB, f).argtypes = [C.c_void_p]
        _LIB.orc_matrix_get.argtypes = [C.c_void_p] * 4
        _LIB.orc_blf_assemble.argtypes = [C.c_void_p, C.POINTER(_Grid), C.POINTER(_Space), C.POINTER(_Space), C.c_int, C.c_int,
                                          C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_double, C.c_int, C.c_void_p,
                                          C.c_double, C.c_int64, C.c_int64, C.c_int]
        _LIB.orc_lf_assemble.argtypes = [C.c_void_p, C.POINTER(_Grid), C.POINTER(_Space), C.c_int, C.c_int, C.c_void_p,
                                         C.c_void_p, C.c_int, C.c_double, C.c_int64, C.c_int, C.c_void_p]
        _LIB.orc_ii_evaluate.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(_Grid), C.POINTER(_Space), C.c_int, C.c_int, C.c_void_p,
                                         C.c_double, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
        _LIB.orc_nlf_convection.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(_Grid), C.POINTER(_Space), C.c_int, C.c_int, C.c_int, C.c_void_p,
                                            C.c_void_p, C.c_int, C.c_double, C.c_int, C.c_int64, C.c_int64, C.c_int]
        _LIB.orc_feb_table.argtypes = [C.POINTER(_Grid), C.POINTER(_Space), C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        _LIB.orc_set_fixed_argument.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        _LIB.orc_quadpoints.argtypes = [C.POINTER(_Grid), C.c_int, C.c_void_p]
        _LIB.orc_qrule.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        _LIB.orc_reftables.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
    return _LIB


def _check(rc):
    if rc != 0:
        raise RuntimeError("oracle: " + lib().orc_last_error().decode())


def _p(a):
    return None if a is None else a.ctypes.data


class _Keep:
    """holds numpy arrays alive next to the ctypes struct that points into them"""


def _grid_struct(grid, need_faces):
    k = _Keep()
    k.coords = np.ascontiguousarray(grid.coords, dtype=np.float64)
    k.cellnodes = np.ascontiguousarray(grid.cellnodes, dtype=np.int32)
    k.vol = np.ascontiguousarray(grid.cellvolumes, dtype=np.float64)
    k.reg = np.ascontiguousarray(grid.cellregions, dtype=np.int32)
    k.cf = k.sg = k.ori = k.fn = k.fv = None
    nfaces = 0
    if need_faces:
        k.cf = np.ascontiguousarray(grid.cellfaces, dtype=np.int32)
        k.sg = np.ascontiguousarray(grid.cellfacesigns, dtype=np.int32)
        if grid.dim == 3:
            k.ori = np.ascontiguousarray(grid.cellfaceorientations, dtype=np.int32)
        k.fn = np.ascontiguousarray(grid.facenormals, dtype=np.float64)
        k.fv = np.ascontiguousarray(grid.facevolumes, dtype=np.float64)
        nfaces = k.fv.size
    k.s = _Grid(grid.dim, getattr(grid, "xdim", grid.dim), k.coords.shape[0], k.cellnodes.shape[0], nfaces, _p(k.coords), _p(k.cellnodes), _p(k.vol), _p(k.reg),
                _p(k.cf), _p(k.sg), _p(k.ori), _p(k.fn), _p(k.fv))
    return k


def _space_struct(space):
    k = _Keep()
    k.dofs = np.ascontiguousarray(space.celldofs, dtype=np.int32)
    k.s = _Space(space.fetype.code, space.fetype.ncomponents, space.ndofs, k.dofs.shape[1], _p(k.dofs))
    return k


def _needs_faces(*spaces):
    return any(s.fetype.code in (3, 4, 5) and not getattr(s.xgrid, "embedded", False) for s in spaces)


class OracleMatrix:
    """ExtendableSparseMatrix{Float64,Int64} stand-in living in the oracle library."""

    def __init__(self, m, n):
        self.m, self.n = m, n
        self.h = lib().orc_matrix_create(m, n)

    def __del__(self):
        if getattr(self, "h", None):
            lib().orc_matrix_destroy(self.h)
            self.h = None

    def flush(self):
        lib().orc_matrix_flush(self.h)

    def fill_zero(self):
        lib().orc_matrix_fill_zero(self.h)

    def csc(self):
        """(colptr, rowval, nzval): 1-based Int64 like SparseMatrixCSC{Float64,Int64}"""
        self.flush()
        nnz = lib().orc_matrix_nnz(self.h)
        colptr = np.zeros(self.n + 1, np.int64)
        rowval = np.zeros(nnz, np.int64)
        nzval = np.zeros(nnz, np.float64)
        lib().orc_matrix_get(self.h, _p(colptr), _p(rowval), _p(nzval))
        return colptr, rowval, nzval

    def toscipy(self):
        import scipy.sparse as sp
        cp, rv, nz = self.csc()
        return sp.csc_matrix((nz, rv - 1, cp - 1), shape=(self.m, self.n))


def blf_assemble(A: OracleMatrix, grid, space1, space2, op1, op2, *, action=ACT_NONE, act_params=None, apt=APT_GENERAL,
                 regions=(0,), factor=1.0, transposed_assembly=False, transpose_copy: OracleMatrix | None = None,
                 factor_transpose=None, offsetX=0, offsetY=0, bonus_quadorder=0, fixed=None):
    """assemble!(A, AP; factor, transposed_assembly, transpose_copy, offsetX, offsetY) -- bilinearform.jl:92-380.
    fixed = (space_a, op_a, coefficients): assemble!(A, AP, FEB; fixed_arguments = [1]) of a trilinear form (235-257), used with
    action = ACT_CONVECTION"""
    g = _grid_struct(grid, _needs_faces(space1, space2) or (fixed is not None and _needs_faces(fixed[0])))
    s1 = _space_struct(space1)
    s2 = s1 if space2 is space1 else _space_struct(space2)
    ap = None if act_params is None else np.ascontiguousarray(act_params, dtype=np.float64)
    rg = np.ascontiguousarray(regions, dtype=np.int32)
    ft = factor if factor_transpose is None else factor_transpose
    if fixed is not None:
        sa = _space_struct(fixed[0])
        ca = np.ascontiguousarray(fixed[2], dtype=np.float64)
        lib().orc_set_fixed_argument(C.byref(sa.s), int(fixed[1]), _p(ca))
    try:
        _check(lib().orc_blf_assemble(A.h, C.byref(g.s), C.byref(s1.s), C.byref(s2.s), op1, op2, action, _p(ap), apt, _p(rg), rg.size,
                                      float(factor), int(transposed_assembly), transpose_copy.h if transpose_copy else None,
                                      float(ft), offsetX, offsetY, bonus_quadorder))
    finally:
        if fixed is not None:
            lib().orc_set_fixed_argument(None, 0, None)


def set_magnitude_mode(on: bool):
    """test infrastructure: subsequent blf_assemble calls accumulate sum |w a b| per entry (same pattern); see grmp_oracle.cpp"""
    lib().orc_set_magnitude_mode(C.c_int(1 if on else 0))


def lf_nq(grid, space, op, bonus_quadorder=0):
    g = _grid_struct(grid, False)
    s = _space_struct(space)
    nq = C.c_int(0)
    rg = np.zeros(1, np.int32)
    _check(lib().orc_lf_assemble(None, C.byref(g.s), C.byref(s.s), op, F_NONE, None, _p(rg), 1, 1.0, 0, bonus_quadorder, C.byref(nq)))
    return nq.value


def lf_assemble(b: np.ndarray, grid, space, op, *, fsrc=F_NONE, fdata=None, regions=(0,), factor=1.0, offset=0, bonus_quadorder=0):
    """assemble!(b, AP; factor, offset) -- linearform.jl:47-237"""
    assert b.dtype == np.float64 and b.flags.c_contiguous
    g = _grid_struct(grid, _needs_faces(space))
    s = _space_struct(space)
    fd = None if fdata is None else np.ascontiguousarray(fdata, dtype=np.float64)
    rg = np.ascontiguousarray(regions, dtype=np.int32)
    _check(lib().orc_lf_assemble(_p(b), C.byref(g.s), C.byref(s.s), op, fsrc, _p(fd), _p(rg), rg.size, float(factor), offset,
                                 bonus_quadorder, None))


def qrule_override(edim, order, xref=None, w=None):
    """use caller-given points/weights for QuadratureRule(order) (None clears the override)"""
    if xref is None:
        lib().orc_qrule_override(-1, -1, 0, None, None)
        return
    x = np.ascontiguousarray(xref, dtype=np.float64)
    ww = np.ascontiguousarray(w, dtype=np.float64)
    lib().orc_qrule_override(C.c_int(edim), C.c_int(order), C.c_int(ww.size), C.c_void_p(_p(x)), C.c_void_p(_p(ww)))


def qrule(edim, order):
    nq = C.c_int(0)
    _check(lib().orc_qrule(edim, order, C.byref(nq), None, None, 0))
    x = np.zeros((nq.value, edim))
    w = np.zeros(nq.value)
    _check(lib().orc_qrule(edim, order, C.byref(nq), _p(x), _p(w), nq.value))
    return x, w


def reftables(fecode, ncomp, edim, xref, nd_all, ncomp_eff):
    """reference values / ForwardDiff-style jacobians of the oracle's own basis closures at xref"""
    x = np.ascontiguousarray(xref, dtype=np.float64)
    vals = np.zeros((x.shape[0], nd_all, ncomp_eff))
    der = np.zeros((x.shape[0], edim, nd_all * ncomp_eff))
    _check(lib().orc_reftables(fecode, ncomp, edim, x.shape[0], _p(x), _p(vals), _p(der)))
    return vals, der


def ii_evaluate(grid, space, op, coeffs, *, kind=II_NONE, factor=1.0, data=None, regions=(0,), bonus_quadorder=0, itemwise=True, b=None):
    """evaluate!(b, AP, FEB) / evaluate(AP, FEB) of a one-argument ItemIntegrator -- itemintegrator.jl:160-360.
    Returns (b[ncells, resultdim] or None, total[resultdim]); total is the reference's running sum over (item, qp);
    a caller-given b is updated in place (b[j,item] += ...)."""
    g = _grid_struct(grid, _needs_faces(space))
    s = _space_struct(space)
    rg = np.ascontiguousarray(regions, dtype=np.int32)
    nq, rd = C.c_int(0), C.c_int(0)
    _check(lib().orc_ii_evaluate(None, None, C.byref(g.s), C.byref(s.s), op, kind, None, 1.0, None, _p(rg), rg.size, bonus_quadorder,
                                 C.byref(nq), C.byref(rd)))
    c = np.ascontiguousarray(coeffs, dtype=np.float64)
    d = None if data is None else np.ascontiguousarray(data, dtype=np.float64)
    if b is None:
        b = np.zeros((g.cellnodes.shape[0], rd.value)) if itemwise else None
    else:
        assert b.dtype == np.float64 and b.flags.c_contiguous and b.shape == (g.cellnodes.shape[0], rd.value)
    total = np.zeros(rd.value)
    _check(lib().orc_ii_evaluate(_p(b), _p(total), C.byref(g.s), C.byref(s.s), op, kind, _p(c), float(factor), _p(d), _p(rg), rg.size,
                                 bonus_quadorder, None, None))
    return b, total


def nlf_convection(A: OracleMatrix, b, grid, space, coeffs, *, op_a=OP_ID, op_g=OP_GRAD, op_t=OP_ID, regions=(0,), factor=1.0,
                   transposed_assembly=True, offsetX=0, offsetY=0, bonus_quadorder=0):
    """full_assemble!(A, b, AP, FEB) of the Newton convection form (nonlinearform.jl:44-245, pdeoperators.jl:459-493)"""
    g = _grid_struct(grid, _needs_faces(space))
    s = _space_struct(space)
    c = np.ascontiguousarray(coeffs, dtype=np.float64)
    rg = np.ascontiguousarray(regions, dtype=np.int32)
    assert b is None or (b.dtype == np.float64 and b.flags.c_contiguous)
    _check(lib().orc_nlf_convection(A.h, _p(b), C.byref(g.s), C.byref(s.s), op_a, op_g, op_t, _p(c), _p(rg), rg.size, float(factor),
                                    int(transposed_assembly), offsetX, offsetY, bonus_quadorder))


def feb_table(grid, space, op, coeffs, order):
    """operator evaluation of the FE function at the quadrature points of the rule of this order: [ncells, nq, resultdim]"""
    g = _grid_struct(grid, _needs_faces(space))
    s = _space_struct(space)
    nq, rd = C.c_int(0), C.c_int(0)
    _check(lib().orc_feb_table(C.byref(g.s), C.byref(s.s), op, None, order, None, C.byref(nq), C.byref(rd)))
    c = np.ascontiguousarray(coeffs, dtype=np.float64)
    t = np.zeros((g.cellnodes.shape[0], nq.value, rd.value))
    _check(lib().orc_feb_table(C.byref(g.s), C.byref(s.s), op, _p(c), order, _p(t), None, None))
    return t


def quadpoints(grid, order):
    g = _grid_struct(grid, False)
    x, _ = qrule(grid.dim, order)
    xq = np.zeros((g.cellnodes.shape[0], x.shape[0], grid.dim))
    _check(lib().orc_quadpoints(C.byref(g.s), order, _p(xq)))
    return xq
