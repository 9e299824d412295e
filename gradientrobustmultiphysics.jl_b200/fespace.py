"""FESpace / CellDofs / FEVector / FEMatrix (host mirror of src/finiteelements.jl,
src/dofmaps.jl, src/fevector.jl, src/fematrix.jl).

`FESpace(fetype, grid)` counts dofs like `count_ndofs` (finiteelements.jl:199-253) and
generates `CellDofs` from the element's pattern string like
`init_dofmap_from_pattern!` (dofmaps.jl:201-363): per component node dofs ("N"), then
face ("F") / edge ("E") dofs, then single-component face dofs ("f"), interior ("I").
Broken spaces (always L2P0, finiteelements.jl:78-80) get the serial dofmap
(c-1)*nd+1 ... c*nd (dofmaps.jl:266-270).

The dof map is an *input* of the device library (include/grmp.h: grmp_space_create);
in production Julia supplies `FES[CellDofs].colentries`.
"""
from __future__ import annotations

import re

import numpy as np

from .fedefs import FEType
from .grid import ExtendableGrid


def parse_pattern(pattern: str):
    """dofmaps.jl:104-121 -> list of (type_char, each_component, ndofs)"""
    return [(m.group(1), m.group(1).isupper(), int(m.group(2))) for m in re.finditer(r"([NnFfEeIiCc])(\d+)", pattern)]


class FESpace:
    def __init__(self, fetype: FEType, xgrid: ExtendableGrid, name: str = "", broken: bool = False):
        self.fetype = fetype
        self.xgrid = xgrid
        self.broken = broken or fetype.broken
        self.name = name or (f"{fetype} (broken)" if self.broken else f"{fetype}")
        edim = xgrid.dim
        self.edim = edim
        self.ncomponents = fetype.ncomponents
        self.nd_cell = fetype.ndofs(edim)
        self.segments = parse_pattern(fetype.dofmap_pattern(edim))
        self.ndofs, self.coffset = self._count_ndofs()
        self._celldofs = None

    # finiteelements.jl:199-253
    def _count_ndofs(self):
        g, nc = self.xgrid, self.ncomponents
        nn_loc, nf_loc, ne_loc = g.dim + 1, g.dim + 1, (3 if g.dim == 2 else 6)
        tot = {"N": 0, "F": 0, "E": 0, "I": 0}
        per_c = {"N": 0, "F": 0, "E": 0, "I": 0}
        for ch, each, q in self.segments:
            key = ch.upper()
            if key not in tot:
                raise NotImplementedError(f"dof pattern segment {ch}")
            tot[key] += q * (nc if each else 1)
            if each:
                per_c[key] += q
        if self.broken:
            ndofs_cell = nn_loc * tot["N"] + nf_loc * tot["F"] + ne_loc * tot["E"] + tot["I"]
            return g.ncells * ndofs_cell, 0
        total = g.ncells * tot["I"]
        coffset = g.ncells * per_c["I"]
        total += g.nnodes * tot["N"]
        if tot["F"] > 0:
            total += g.nfaces * tot["F"]
        if tot["E"] > 0:
            total += g.nedges * tot["E"]
        coffset += g.nnodes * per_c["N"]
        if per_c["F"] > 0:
            coffset += g.nfaces * per_c["F"]
        if per_c["E"] > 0:
            coffset += g.nedges * per_c["E"]
        return total, coffset

    # dofmaps.jl:201-363 (vectorised over cells)
    @property
    def celldofs(self) -> np.ndarray:
        if self._celldofs is not None:
            return self._celldofs
        g = self.xgrid
        nc_cells = g.ncells
        if self.broken:
            nd = self.nd_cell
            dm = (np.arange(nc_cells, dtype=np.int64)[:, None] * nd + np.arange(1, nd + 1)[None, :])
            self._celldofs = np.ascontiguousarray(dm, dtype=np.int32)
            return self._celldofs
        cols = []
        nnodes = g.nnodes
        items = {"N": (g.cellnodes, nnodes)}
        if any(ch.upper() == "F" for ch, _, _ in self.segments):
            items["F"] = (g.cellfaces, g.nfaces)
        if any(ch.upper() == "E" for ch, _, _ in self.segments):
            items["E"] = (g.celledges, g.nedges)
        cell_ids = np.arange(1, nc_cells + 1, dtype=np.int64)
        for c in range(self.ncomponents):
            offset = c * self.coffset
            for ch, each, q in self.segments:
                if not each:
                    continue
                if ch == "I":
                    for m in range(q):
                        cols.append(cell_ids + offset)
                        offset += nc_cells
                    continue
                adj, nitems = items[ch]
                adj = adj.astype(np.int64)
                for n in range(adj.shape[1]):
                    for m in range(q):
                        cols.append(adj[:, n] + offset + m * nitems)
                offset += nitems * q
        offset = self.ncomponents * self.coffset
        for ch, each, q in self.segments:
            if each:
                continue
            key = ch.upper()
            if key == "I":
                for m in range(q):
                    cols.append(cell_ids + offset)
                    offset += nc_cells
                continue
            adj, nitems = items[key]
            adj = adj.astype(np.int64)
            for n in range(adj.shape[1]):
                for m in range(q):
                    cols.append(adj[:, n] + offset + m * nitems)
            offset += nitems * q
        dm = np.stack(cols, axis=1)
        assert dm.shape[1] == self.nd_cell, (dm.shape, self.nd_cell)
        assert dm.max() <= self.ndofs
        self._celldofs = np.ascontiguousarray(dm, dtype=np.int32)
        return self._celldofs

    # dofmaps.jl:201-363 with DM = BFaceDofs (patterns: h1_p1.jl:29 "N1", h1_p2.jl:37-38 "N1I1" / "N1E1")
    @property
    def bfacedofs(self) -> np.ndarray:
        if getattr(self, "_bfacedofs", None) is not None:
            return self._bfacedofs
        g = self.xgrid
        pattern = self.fetype.bface_dofmap_pattern(g.dim - 1)
        nb = g.bfacenodes.shape[0]
        cols = []
        segs = parse_pattern(pattern)
        for c in range(self.ncomponents):
            offset = c * self.coffset
            for ch, each, q in segs:
                if not each:
                    continue
                assert q == 1
                if ch == "N":
                    adj, nitems = g.bfacenodes.astype(np.int64), g.nnodes
                elif ch == "I":                      # interior dof of the face = face dof of the parent space
                    adj, nitems = g.bfacefaces.astype(np.int64)[:, None], g.nfaces
                elif ch == "E":
                    adj, nitems = g.bfaceedges.astype(np.int64), g.nedges
                else:
                    raise NotImplementedError(f"BFaceDofs pattern segment {ch}")
                for n in range(adj.shape[1]):
                    cols.append(adj[:, n] + offset)
                offset += nitems
        offset = self.ncomponents * self.coffset
        for ch, each, q in segs:                     # "i<q>": q face dofs not tied to a component (Hdiv normal-flux dofs), dofmaps.jl:338-343
            if each:
                continue
            if ch != "i":
                raise NotImplementedError(f"BFaceDofs pattern segment {ch}")
            for m in range(q):
                cols.append(g.bfacefaces.astype(np.int64) + offset)
                offset += g.nfaces
        dm = np.stack(cols, axis=1) if nb else np.zeros((0, len(cols)), np.int64)
        assert dm.max(initial=0) <= self.ndofs
        self._bfacedofs = np.ascontiguousarray(dm, dtype=np.int32)
        return self._bfacedofs

    def on_bfaces(self):
        """item view of this space for ON_BFACES assembly: same dofs, items = boundary faces, ItemDofs = BFaceDofs"""
        if getattr(self, "_bfspace", None) is None:
            self._bfspace = BFaceSpace(self)
        return self._bfspace

    def __repr__(self):
        return f"FESpace({self.name}, ndofs={self.ndofs})"


class BFaceSpace:
    """FES[BFaceDofs] + the face geometry (Dofmap4AssemblyType(ON_BFACES), dofmaps.jl:45)"""

    def __init__(self, parent: FESpace):
        if parent.broken or not hasattr(parent.fetype, "bface_dofmap_pattern"):
            raise NotImplementedError(f"ON_BFACES assembly: H1P1 / H1P2 / HDIVRT0 / HDIVBDM1 spaces, got {parent.name}")
        self.parent = parent
        self.fetype = parent.fetype
        # Hdiv face bases are moments with respect to the face's own node order (FaceNodes): the items must carry that order
        self.xgrid = parent.xgrid.bface_grid(face_order=bool(getattr(parent.fetype, "hdiv", False)))
        self.broken = False
        self.edim = self.xgrid.dim
        self.ncomponents = parent.ncomponents
        self.ndofs = parent.ndofs
        self.nd_cell = parent.fetype.ndofs(self.edim)
        self.name = parent.name + " [ON_BFACES]"

    @property
    def celldofs(self):
        return self.parent.bfacedofs


class FEVectorBlock:
    """src/fevector.jl:13-19 -- a view [offset+1 : last_index] into the shared entries"""

    def __init__(self, name, FES, offset, last_index, entries):
        self.name, self.FES, self.offset, self.last_index, self.entries = name, FES, offset, last_index, entries

    @property
    def view(self):
        return self.entries[self.offset:self.last_index]

    def fill(self, v):
        self.entries[self.offset:self.last_index] = v

    def __len__(self):
        return self.last_index - self.offset


class FEVector:
    def __init__(self, FES, name="auto"):
        if isinstance(FES, FESpace):
            FES = [FES]
        n = sum(f.ndofs for f in FES)
        self.entries = np.zeros(n)
        self.blocks = []
        off = 0
        for j, f in enumerate(FES):
            self.blocks.append(FEVectorBlock(f"{name}[{j + 1}]", f, off, off + f.ndofs, self.entries))
            off += f.ndofs

    def __getitem__(self, i):      # 1-based like Julia
        return self.blocks[i - 1]

    def __len__(self):
        return len(self.blocks)


class FEMatrixBlock:
    """src/fematrix.jl:14-23"""

    def __init__(self, name, FESX, FESY, offsetX, offsetY, parent):
        self.name, self.FESX, self.FESY = name, FESX, FESY
        self.offsetX, self.offsetY = offsetX, offsetY
        self.last_indexX, self.last_indexY = offsetX + FESX.ndofs, offsetY + FESY.ndofs
        self.parent = parent

    @property
    def shape(self):
        return (self.FESX.ndofs, self.FESY.ndofs)


class FEMatrix:
    """Block overlay over ONE sparse matrix (src/fematrix.jl:48-51, 186-213).

    `entries` is a scipy CSC matrix holding the union pattern of everything assembled so
    far -- the stand-in for ExtendableSparseMatrix{Float64,Int64}.cscmatrix after
    flush!.  Device-assembled operators are merged into it by `add_csc` (explicit zeros
    are kept, like ExtendableSparse's flush!)."""

    def __init__(self, FESX, FESY=None, name="auto"):
        if isinstance(FESX, FESpace):
            FESX = [FESX]
        if FESY is None:
            FESY = FESX
        elif isinstance(FESY, FESpace):
            FESY = [FESY]
        self.FESX, self.FESY = FESX, FESY
        self.m = sum(f.ndofs for f in FESX)
        self.n = sum(f.ndofs for f in FESY)
        self.blocks = {}
        ox = 0
        for j, fx in enumerate(FESX):
            oy = 0
            for k, fy in enumerate(FESY):
                self.blocks[(j + 1, k + 1)] = FEMatrixBlock(f"{name} [{j + 1},{k + 1}]", fx, fy, ox, oy, self)
                oy += fy.ndofs
            ox += fx.ndofs
        # CSC triplet store: union-pattern accumulation with kept explicit zeros
        self.colptr = np.ones(self.n + 1, dtype=np.int64)
        self.rowval = np.zeros(0, dtype=np.int64)
        self.nzval = np.zeros(0, dtype=np.float64)

    def __getitem__(self, jk):
        return self.blocks[jk]

    @property
    def nnz(self):
        return self.rowval.size

    def add_csc(self, colptr, rowval, nzval):
        """A += B for two 1-based CSC matrices of the full size, keeping every stored
        index (ExtendableSparse flush! semantics: csc = lnk + csc, zeros are not dropped)."""
        if self.rowval.size == 0:
            self.colptr, self.rowval, self.nzval = colptr.copy(), rowval.copy(), nzval.copy()
            return
        n = self.n
        ca = np.repeat(np.arange(n, dtype=np.int64), np.diff(self.colptr))
        cb = np.repeat(np.arange(n, dtype=np.int64), np.diff(colptr))
        key = np.concatenate([ca * (self.m + 1) + self.rowval, cb * (self.m + 1) + rowval])
        val = np.concatenate([self.nzval, nzval])
        uk, inv = np.unique(key, return_inverse=True)
        out = np.zeros(uk.size)
        # existing entries first, then the new ones (a + b per index)
        np.add.at(out, inv[: self.nzval.size], self.nzval)
        np.add.at(out, inv[self.nzval.size:], nzval)
        cols = uk // (self.m + 1)
        self.rowval = uk % (self.m + 1)
        self.nzval = out
        self.colptr = np.concatenate([[0], np.cumsum(np.bincount(cols, minlength=n))]).astype(np.int64) + 1

    def fill_block_zero(self, block: FEMatrixBlock):
        """fill!(B, 0): zero nzval inside the block, keep the pattern (fematrix.jl:220-232)"""
        cols = np.repeat(np.arange(self.n, dtype=np.int64), np.diff(self.colptr))
        sel = ((cols >= block.offsetY) & (cols < block.last_indexY)
               & (self.rowval > block.offsetX) & (self.rowval <= block.last_indexX))
        self.nzval[sel] = 0.0

    def tocsc(self):
        import scipy.sparse as sp
        return sp.csc_matrix((self.nzval, self.rowval - 1, self.colptr - 1), shape=(self.m, self.n))
