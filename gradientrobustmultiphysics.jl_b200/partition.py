"""Multi-GPU sharding of the assembly path (one process per GPU).

The reference has no distributed code at all (single serial cell loop,
bilinearform.jl:226); the path shards naturally because cells are independent units of
work coupled only through shared dofs.  Scheme (owner-computes, variant B):

  * the cells are split into `world` contiguous ranges -- `uniform_refine` numbers the
    children of a coarse cell contiguously, so ranges are spatially compact.  The ranges are
    balanced by the cells a rank WALKS (own + halo), which is what its time is proportional to;
    the lower ranks own the dofs on the range interfaces and carry the larger halos;
  * a dof (matrix column/row) is owned by the rank that holds its lowest-numbered cell;
  * a rank assembles its cells plus the halo cells touching an owned dof, with the owned
    dofs numbered first (`grmp_blf_set_owned_columns`), so every owned column is complete
    and no numeric-phase exchange is needed;
  * the global CSC matrix is the column-wise concatenation of the owned column blocks after
    mapping local row numbers back to global ones (`merge_owned_columns`).

Only host logic lives here; it is covered by a world_size-2 gloo test on CPU.
"""
from __future__ import annotations

import numpy as np

from .fespace import FESpace
from .grid import ExtendableGrid


class LocalProblem:
    def __init__(self, grid, space, n_owned, local2global, cells):
        self.grid, self.space, self.n_owned, self.local2global, self.cells = grid, space, n_owned, local2global, cells


def cell_ranges(ncells: int, world: int):
    return [(ncells * r) // world for r in range(world + 1)]


def _first_cells(space: FESpace):
    nc = space.xgrid.ncells
    dofs = space.celldofs.astype(np.int64) - 1
    first_cell = np.full(space.ndofs, nc, dtype=np.int64)
    np.minimum.at(first_cell, dofs.ravel(), np.repeat(np.arange(nc, dtype=np.int64), dofs.shape[1]))
    return first_cell


def balanced_cell_ranges(space: FESpace, world: int):
    """cell range boundaries such that every rank owns about the same number of (dof, cell) visits: a dof belongs to its
    lowest-numbered cell and costs one visit per cell that contains it"""
    nc = space.xgrid.ncells
    if world <= 1 or nc == 0:
        return cell_ranges(nc, world)
    dofs = space.celldofs.astype(np.int64) - 1
    visits = np.bincount(dofs.ravel(), minlength=space.ndofs)            # cells per dof
    w = np.bincount(_first_cells(space), weights=visits, minlength=nc + 1)[:nc]
    cum = np.cumsum(w)
    bounds = [0]
    for r in range(1, world):
        bounds.append(int(np.searchsorted(cum, cum[-1] * r / world, side="left")) + 1)
    bounds.append(nc)
    for r in range(1, world + 1):                                          # keep the ranges non-empty and ordered
        bounds[r] = min(max(bounds[r], bounds[r - 1]), nc)
    return bounds


def halo_balanced_cell_ranges(space: FESpace, world: int, iterations: int = 3):
    """cell range boundaries such that every rank ASSEMBLES about the same number of cells, halo included.  Measured on the
    metric kernel (tools/emulate_ranks.py): a rank's time is proportional to the cells it walks -- halo cells cost as much as
    own cells, and the low ranks (which own the interface dofs) carry the larger halos."""
    nc = space.xgrid.ncells
    bounds = cell_ranges(nc, world)
    if world <= 1 or nc < world:
        return bounds
    dofs = space.celldofs.astype(np.int64) - 1
    first = _first_cells(space)
    for _ in range(iterations):
        owner = np.searchsorted(np.array(bounds[1:]), first, side="right")
        cell_owner_min = owner[dofs].min(axis=1)          # a cell is walked by every rank that owns one of its dofs
        cell_owner_max = owner[dofs].max(axis=1)
        own = np.diff(bounds).astype(np.float64)
        walked = np.zeros(world)
        for r in range(world):
            walked[r] = np.count_nonzero((cell_owner_min <= r) & (cell_owner_max >= r) & (owner[dofs] == r).any(axis=1))
        halo = walked - own
        target = (nc + halo.sum()) / world
        sizes = np.maximum(target - halo, 1.0)
        sizes *= nc / sizes.sum()
        nb = np.concatenate([[0], np.round(np.cumsum(sizes)).astype(np.int64)])
        nb[-1] = nc
        bounds = [int(min(max(b, 0), nc)) for b in nb]
        for r in range(1, world + 1):
            bounds[r] = max(bounds[r], bounds[r - 1])
    return bounds


_BALANCERS = {"work": balanced_cell_ranges, "halo": halo_balanced_cell_ranges, "cells": lambda space, world: cell_ranges(space.xgrid.ncells, world)}
DEFAULT_BALANCE = "halo"


def dof_owner(space: FESpace, world: int, bounds=None):
    bounds = np.array((bounds if bounds is not None else _BALANCERS[DEFAULT_BALANCE](space, world))[1:])
    return np.searchsorted(bounds, _first_cells(space), side="right")


def partition(space: FESpace, rank: int, world: int, balance: str | None = None, bounds=None) -> LocalProblem:
    """rank-local grid/space: own + halo cells, owned dofs first (local numbering, 1-based CellDofs)"""
    g = space.xgrid
    dofs = space.celldofs.astype(np.int64) - 1
    if bounds is None:
        bounds = _BALANCERS[balance or DEFAULT_BALANCE](space, world)
    owned = dof_owner(space, world, bounds) == rank
    cells = np.nonzero(owned[dofs].any(axis=1))[0]
    ldofs = dofs[cells]
    used = np.zeros(space.ndofs, bool)
    used[ldofs.ravel()] = True
    order = np.concatenate([np.nonzero(used & owned)[0], np.nonzero(used & ~owned)[0]])
    newid = np.full(space.ndofs, -1, dtype=np.int64)
    newid[order] = np.arange(order.size)
    n_owned = int((used & owned).sum())
    cn = g.cellnodes[cells].astype(np.int64) - 1
    nodes = np.unique(cn)
    nmap = np.full(g.nnodes, -1, dtype=np.int64)
    nmap[nodes] = np.arange(nodes.size)
    lg = ExtendableGrid(g.coords[nodes], nmap[cn] + 1, g.cellregions[cells])
    lg._cache["vol"] = np.ascontiguousarray(g.cellvolumes[cells])     # CellVolumes are an input, not recomputed
    ls = FESpace(space.fetype, lg)
    ls._celldofs = np.ascontiguousarray(newid[ldofs] + 1, dtype=np.int32)
    ls.ndofs = int(order.size)
    return LocalProblem(lg, ls, n_owned, order, cells)


def owned_block_to_global(lp: LocalProblem, colptr, rowval, nzval):
    """owned columns of a rank-local CSC (1-based) -> (global column ids, colptr, global rowval sorted, nzval)"""
    n = lp.n_owned
    cp = colptr[: n + 1]
    end = cp[-1] - 1
    rows_g = lp.local2global[rowval[:end] - 1] + 1
    vals = nzval[:end].copy()
    # rows must ascend per column in the global numbering
    cols = np.repeat(np.arange(n, dtype=np.int64), np.diff(cp))
    perm = np.lexsort((rows_g, cols))
    return lp.local2global[:n] + 1, cp.copy(), rows_g[perm], vals[perm]


def merge_owned_columns(ncols_global: int, blocks):
    """concatenate the owned column blocks of all ranks into one global CSC (1-based Int64)"""
    counts = np.zeros(ncols_global, dtype=np.int64)
    for gcols, cp, _, _ in blocks:
        counts[gcols - 1] = np.diff(cp)
    colptr = np.concatenate([[0], np.cumsum(counts)]).astype(np.int64) + 1
    rowval = np.zeros(colptr[-1] - 1, dtype=np.int64)
    nzval = np.zeros(colptr[-1] - 1)
    for gcols, cp, rv, nz in blocks:
        lens = np.diff(cp)
        dst0 = colptr[gcols - 1] - 1
        idx = np.repeat(dst0 - (cp[:-1] - 1), lens) + np.arange(rv.size)
        rowval[idx] = rv
        nzval[idx] = nz
    return colptr, rowval, nzval
