"""Loader for the raw arrays `baseline/run_reference.jl` writes where the REAL reference can run (Julia + GradientRobustMultiPhysics
v0.12 + ExtendableGrids / ExtendableSparse): grid arrays exactly as the Julia objects hold them (column-major, 1-based) and the
assembled SparseMatrixCSC.  Tests that find such a directory (GRMP_REFERENCE_DUMP or baseline/reference_dump/) run the oracle
and the library on the REFERENCE-GENERATED inputs and compare with the reference's own colptr / rowval / nzval: that turns
"parity unpinned"  into a bit-level pin.  Without a dump the tests are skipped."""
import os

import numpy as np

import grmp_b200 as G

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FILES = ("coords.f64", "cellnodes.i32", "cellvolumes.f64", "celldofs.i32", "colptr.i64", "rowval.i64", "nzval.f64")


def dump_dir():
    d = os.environ.get("GRMP_REFERENCE_DUMP", os.path.join(ROOT, "baseline", "reference_dump"))
    return d if all(os.path.exists(os.path.join(d, f)) for f in FILES) else None


def load(d, dim=3, nd=10):
    """-> (grid, space, (colptr, rowval, nzval)) built from the reference's arrays only"""
    rd = lambda n, t: np.fromfile(os.path.join(d, n), dtype=t)      # noqa: E731
    coords = rd("coords.f64", np.float64).reshape(-1, dim)              # Julia dim x nnodes, column-major
    cellnodes = rd("cellnodes.i32", np.int32).reshape(-1, dim + 1)      # (dim+1) x ncells
    vol = rd("cellvolumes.f64", np.float64)
    celldofs = rd("celldofs.i32", np.int32).reshape(-1, nd)             # FES[CellDofs].colentries
    g = G.ExtendableGrid(coords, cellnodes)
    assert vol.size == g.ncells and celldofs.shape[0] == g.ncells
    g._cache["vol"] = np.ascontiguousarray(vol)                         # CellVolumes are an input (bilinearform.jl:113)
    s = G.FESpace(G.H1P2(1, dim), g)
    own = s.celldofs.copy()                                             # this package's enumeration, for the report only
    s._celldofs = np.ascontiguousarray(celldofs)
    s.ndofs = int(celldofs.max())
    ref = (rd("colptr.i64", np.int64), rd("rowval.i64", np.int64), rd("nzval.f64", np.float64))
    return g, s, ref, bool(np.array_equal(own, celldofs))


def write(d, g, s, csc):
    """the same files run_reference.jl writes (used to exercise the loader without Julia)"""
    os.makedirs(d, exist_ok=True)
    np.ascontiguousarray(g.coords, np.float64).tofile(os.path.join(d, "coords.f64"))
    np.ascontiguousarray(g.cellnodes, np.int32).tofile(os.path.join(d, "cellnodes.i32"))
    np.ascontiguousarray(g.cellvolumes, np.float64).tofile(os.path.join(d, "cellvolumes.f64"))
    np.ascontiguousarray(s.celldofs, np.int32).tofile(os.path.join(d, "celldofs.i32"))
    for n, a, t in (("colptr.i64", csc[0], np.int64), ("rowval.i64", csc[1], np.int64), ("nzval.f64", csc[2], np.float64)):
        np.ascontiguousarray(a, t).tofile(os.path.join(d, n))
