"""Host-side simplex grids: the input producer of the assembly path.

In production the Julia host (ExtendableGrids.jl, not vendored under /root/reference)
owns these arrays and hands them to libgrmp_cuda through the C-ABI (include/grmp.h).
This module is the in-container stand-in for that producer: it builds the same grid
*components* the reference reads on the hot path

  Coordinates, CellNodes, CellRegions, CellVolumes        (bilinearform.jl:113-114)
  CellFaces, CellFaceSigns, CellFaceOrientations          (hdiv_rt0.jl:106-116, hdiv_bdm1.jl:278-328)
  FaceNormals, FaceVolumes                                (h1v_br.jl:150-162, reconstructions.jl:27-30)
  CellEdges / EdgeNodes                                   (dofmaps.jl:232-240, "E" dofs of H1P2 in 3D)

for `grid_unitsquare(Triangle2D)`, `grid_unitcube(Tetrahedron3D)`,
`reference_domain(...)` and `uniform_refine` (examples/Example201_PoissonProblem2D.jl:32,
examples/Example301_Poisson3D.jl:39).

Conventions (all arrays are stored so that their memory is byte-identical to the
Julia column-major arrays):
  * indices are 1-based Int32 (ExtendableGrid{Float64,Int32}),
  * `coords[node, :]`      == Julia `Coordinates[:, node]`,
  * `cellnodes[cell, :]`   == Julia `CellNodes[:, cell]`, etc.

Enumeration of faces/edges (first encounter while looping cells, then local
faces/edges) and the child ordering of red refinement are restated from the published
behaviour of ExtendableGrids and are *unpinned* (see DESIGN.md "parity unpinned").
Everything is vectorised numpy so that the 6.3 M-cell benchmark grid builds in seconds.
"""
from __future__ import annotations

import numpy as np

# ----------------------------------------------------------------------------------
# local enumeration rules pinned by the reference tree
# ----------------------------------------------------------------------------------
# Triangle2D faces [1 2],[2 3],[3 1]           (h1_p2.jl:215-217)
TRI_FACENODES = np.array([[0, 1], [1, 2], [2, 0]], dtype=np.int64)
# Tetrahedron3D faces [1 3 2],[1 2 4],[2 3 4],[1 4 3]   (hdiv_bdm1.jl FACE1..FACE4 comments)
TET_FACENODES = np.array([[0, 2, 1], [0, 1, 3], [1, 2, 3], [0, 3, 2]], dtype=np.int64)
# Tetrahedron3D edges [1 2],[1 3],[1 4],[2 3],[2 4],[3 4]   (h1_p2.jl:231-236)
TET_EDGENODES = np.array([[0, 1], [0, 2], [0, 3], [1, 2], [1, 3], [2, 3]], dtype=np.int64)

# red refinement rules (child -> local node ids; ids >= nn are edge/face midpoints)
# triangle: 4=m12, 5=m23, 6=m31 (midpoints in local face order)
TRI_REFINE = np.array([[0, 3, 5], [3, 1, 4], [5, 4, 2], [4, 5, 3]], dtype=np.int64)
# tetrahedron: 4=m12 5=m13 6=m14 7=m23 8=m24 9=m34 (midpoints in local edge order);
# the inner octahedron is cut along the m12-m34 diagonal (4,9)
TET_REFINE = np.array(
    [[0, 4, 5, 6], [4, 1, 7, 8], [5, 7, 2, 9], [6, 8, 9, 3],
     [9, 4, 7, 8], [9, 4, 8, 6], [9, 4, 6, 5], [9, 4, 5, 7]], dtype=np.int64)


def _first_encounter_unique(keys2d: np.ndarray):
    """Unique rows of an (n,k) integer array, numbered in order of first appearance.

    Returns (ids, first) with ids[i] = 0-based id of row i and first[j] = index of the
    first row carrying id j."""
    n, k = keys2d.shape
    # reduce columns pairwise to a single int64 key without overflow
    key = keys2d[:, 0].astype(np.int64)
    for c in range(1, k):
        m = int(keys2d[:, c].max()) + 1 if n else 1
        if int(key.max() if n else 0) < (2**62) // max(m, 1):
            key = key * m + keys2d[:, c]
        else:  # compress the running key first
            _, inv = np.unique(key, return_inverse=True)
            key = inv.astype(np.int64) * m + keys2d[:, c]
    _, first, inv = np.unique(key, return_index=True, return_inverse=True)
    order = np.argsort(first, kind="stable")          # sorted-id -> encounter rank
    rank = np.empty_like(order)
    rank[order] = np.arange(order.size)
    return rank[inv], first[order]


class ExtendableGrid:
    """Simplex grid with the components the assembly path reads (1-based Int32)."""

    def __init__(self, coords, cellnodes, cellregions=None, bfacenodes=None, bfaceregions=None):
        self.coords = np.ascontiguousarray(coords, dtype=np.float64)
        self.cellnodes = np.ascontiguousarray(cellnodes, dtype=np.int32)
        self.dim = self.coords.shape[1]
        assert self.cellnodes.shape[1] == self.dim + 1, "simplex grids only"
        nc = self.cellnodes.shape[0]
        self.cellregions = (np.ones(nc, np.int32) if cellregions is None
                            else np.ascontiguousarray(cellregions, dtype=np.int32))
        self.bfacenodes = (np.zeros((0, self.dim), np.int32) if bfacenodes is None
                           else np.ascontiguousarray(bfacenodes, dtype=np.int32))
        self.bfaceregions = (np.zeros(self.bfacenodes.shape[0], np.int32) if bfaceregions is None
                             else np.ascontiguousarray(bfaceregions, dtype=np.int32))
        self._cache = {}

    # -- sizes -------------------------------------------------------------------
    @property
    def nnodes(self):
        return self.coords.shape[0]

    @property
    def ncells(self):
        return self.cellnodes.shape[0]

    @property
    def nfaces(self):
        return self.facenodes.shape[0]

    @property
    def nedges(self):
        return self.edgenodes.shape[0]

    # -- lazily instantiated components (like ExtendableGrids' instantiate) -------
    @property
    def cellvolumes(self):
        if "vol" not in self._cache:
            x = self.coords
            cn = self.cellnodes.astype(np.int64) - 1
            a = x[cn[:, 1]] - x[cn[:, 0]]
            b = x[cn[:, 2]] - x[cn[:, 0]]
            if self.dim == 2:
                det = a[:, 0] * b[:, 1] - a[:, 1] * b[:, 0]
                self._cache["vol"] = np.abs(det) / 2
            else:
                c = x[cn[:, 3]] - x[cn[:, 0]]
                det = (a[:, 0] * (b[:, 1] * c[:, 2] - b[:, 2] * c[:, 1])
                       - a[:, 1] * (b[:, 0] * c[:, 2] - b[:, 2] * c[:, 0])
                       + a[:, 2] * (b[:, 0] * c[:, 1] - b[:, 1] * c[:, 0]))
                self._cache["vol"] = np.abs(det) / 6
        return self._cache["vol"]

    def _build_faces(self):
        cn = self.cellnodes.astype(np.int64)
        rule = TRI_FACENODES if self.dim == 2 else TET_FACENODES
        nf_loc = rule.shape[0]
        nc = cn.shape[0]
        fn_all = cn[:, rule].reshape(nc * nf_loc, self.dim)   # cell-major, local face minor
        ids, first = _first_encounter_unique(np.sort(fn_all, axis=1))
        nfaces = first.size
        facenodes = fn_all[first]                            # node order as seen from first cell
        cellfaces = (ids.reshape(nc, nf_loc) + 1).astype(np.int32)
        inst = np.arange(nc * nf_loc)
        is_first = first[ids] == inst
        signs = np.where(is_first, 1, -1).astype(np.int32).reshape(nc, nf_loc)
        facecells = np.zeros((nfaces, 2), np.int32)
        facecells[ids[is_first], 0] = (inst[is_first] // nf_loc) + 1
        facecells[ids[~is_first], 1] = (inst[~is_first] // nf_loc) + 1
        c = self._cache
        c["facenodes"] = facenodes.astype(np.int32)
        c["cellfaces"] = cellfaces
        c["cellfacesigns"] = signs
        c["facecells"] = facecells
        x = self.coords
        f0 = facenodes - 1
        if self.dim == 2:
            t = x[f0[:, 1]] - x[f0[:, 0]]
            length = np.sqrt(t[:, 0] * t[:, 0] + t[:, 1] * t[:, 1])
            normals = np.stack([t[:, 1] / length, -t[:, 0] / length], axis=1)
            c["facevolumes"] = length
            c["facenormals"] = np.ascontiguousarray(normals)
        else:
            a = x[f0[:, 1]] - x[f0[:, 0]]
            b = x[f0[:, 2]] - x[f0[:, 0]]
            n = np.stack([a[:, 1] * b[:, 2] - a[:, 2] * b[:, 1],
                          a[:, 2] * b[:, 0] - a[:, 0] * b[:, 2],
                          a[:, 0] * b[:, 1] - a[:, 1] * b[:, 0]], axis=1)
            nn = np.sqrt(n[:, 0] * n[:, 0] + n[:, 1] * n[:, 1] + n[:, 2] * n[:, 2])
            c["facevolumes"] = nn / 2
            c["facenormals"] = np.ascontiguousarray(n / nn[:, None])
            # orientation of the local face node order relative to the global FaceNodes order
            # (g1,g2,g3): 1 = identical, 2 = (g3,g2,g1), 3 = (g2,g1,g3), 4 = (g1,g3,g2);
            # this is the convention under which the BDM1 subset tables
            # (hdiv_bdm1.jl: shift4orientation1/2) give a normal-continuous space.
            g = facenodes[ids]                                # (nc*4, 3) global order per instance
            l = fn_all
            orient = np.zeros(nc * nf_loc, np.int32)
            same = (l == g).all(axis=1)
            orient[same] = 1
            orient[(l[:, 0] == g[:, 2]) & (l[:, 1] == g[:, 1]) & (l[:, 2] == g[:, 0])] = 2
            orient[(l[:, 0] == g[:, 1]) & (l[:, 1] == g[:, 0]) & (l[:, 2] == g[:, 2])] = 3
            orient[(l[:, 0] == g[:, 0]) & (l[:, 1] == g[:, 2]) & (l[:, 2] == g[:, 1])] = 4
            assert (orient > 0).all(), "inconsistent face orientation (non-manifold grid?)"
            c["cellfaceorientations"] = orient.reshape(nc, nf_loc)

    def _build_edges(self):
        cn = self.cellnodes.astype(np.int64)
        nc = cn.shape[0]
        if self.dim == 2:   # edges == faces in 2D
            self._cache["edgenodes"] = self.facenodes
            self._cache["celledges"] = self.cellfaces
            return
        en_all = cn[:, TET_EDGENODES].reshape(nc * 6, 2)
        ids, first = _first_encounter_unique(np.sort(en_all, axis=1))
        self._cache["edgenodes"] = en_all[first].astype(np.int32)
        self._cache["celledges"] = (ids.reshape(nc, 6) + 1).astype(np.int32)

    def _get(self, name, builder):
        if name not in self._cache:
            builder()
        return self._cache[name]

    facenodes = property(lambda s: s._get("facenodes", s._build_faces))
    cellfaces = property(lambda s: s._get("cellfaces", s._build_faces))
    cellfacesigns = property(lambda s: s._get("cellfacesigns", s._build_faces))
    cellfaceorientations = property(lambda s: s._get("cellfaceorientations", s._build_faces))
    facecells = property(lambda s: s._get("facecells", s._build_faces))
    facenormals = property(lambda s: s._get("facenormals", s._build_faces))
    facevolumes = property(lambda s: s._get("facevolumes", s._build_faces))
    edgenodes = property(lambda s: s._get("edgenodes", s._build_edges))
    celledges = property(lambda s: s._get("celledges", s._build_edges))

    @property
    def bfacefaces(self):
        """global face number of every boundary face (BFaceFaces)."""
        if "bfacefaces" not in self._cache:
            fn = np.sort(self.facenodes.astype(np.int64), axis=1)
            bn = np.sort(self.bfacenodes.astype(np.int64), axis=1)
            both = np.concatenate([fn, bn])
            ids, _ = _first_encounter_unique(both)
            self._cache["bfacefaces"] = (ids[fn.shape[0]:] + 1).astype(np.int32)
        return self._cache["bfacefaces"]


    @property
    def bfacevolumes(self):
        """BFaceVolumes: length (2D) / area (3D) of every boundary face"""
        if "bfacevolumes" not in self._cache:
            x = self.coords
            bn = self.bfacenodes.astype(np.int64) - 1
            a = x[bn[:, 1]] - x[bn[:, 0]]
            if self.dim == 2:
                self._cache["bfacevolumes"] = np.sqrt((a * a).sum(axis=1))
            else:
                n = np.cross(a, x[bn[:, 2]] - x[bn[:, 0]])
                self._cache["bfacevolumes"] = np.sqrt((n * n).sum(axis=1)) / 2
        return self._cache["bfacevolumes"]

    @property
    def bfaceedges(self):
        """FaceEdges[:, BFaceFaces] in 3D: the edges (1,2), (2,3), (3,1) of every boundary triangle, nodes in BFaceNodes order"""
        if "bfaceedges" not in self._cache:
            assert self.dim == 3
            en = np.sort(self.edgenodes.astype(np.int64), axis=1)
            bn = self.bfacenodes.astype(np.int64)
            pairs = np.sort(bn[:, [[0, 1], [1, 2], [2, 0]]].reshape(-1, 2), axis=1)
            ids, _ = _first_encounter_unique(np.concatenate([en, pairs]))
            be = ids[en.shape[0]:]
            assert be.max(initial=-1) < en.shape[0], "boundary face with an edge that no cell has"
            self._cache["bfaceedges"] = (be.reshape(-1, 3) + 1).astype(np.int32)
        return self._cache["bfaceedges"]

    def bface_grid(self, face_order=False):
        """the boundary faces as assembly items (AT = ON_BFACES); face_order: item nodes in FaceNodes order instead of BFaceNodes order"""
        key = "bface_grid_f" if face_order else "bface_grid"
        if key not in self._cache:
            self._cache[key] = BFaceGrid(self, face_order)
        return self._cache[key]


class BFaceGrid:
    """Item view for ON_BFACES assembly (assemblypatterns.jl:400-440 with GridComponent*4AssemblyType(ON_BFACES)): items = boundary
    faces (Edge1D in 2D, Triangle2D in 3D) with BFaceNodes / BFaceVolumes / BFaceRegions in the roles of the cell components;
    coordinates keep the dimension of the parent grid."""
    embedded = True

    def __init__(self, parent: "ExtendableGrid", face_order=False):
        self.parent = parent
        self.dim = parent.dim - 1
        self.xdim = parent.dim
        self.coords = parent.coords
        self.cellnodes = (np.ascontiguousarray(parent.facenodes[parent.bfacefaces.astype(np.int64) - 1]) if face_order
                          else parent.bfacenodes)
        self.cellregions = parent.bfaceregions

    @property
    def cellvolumes(self):
        return self.parent.bfacevolumes

    @property
    def nnodes(self):
        return self.coords.shape[0]

    @property
    def ncells(self):
        return self.cellnodes.shape[0]


# ----------------------------------------------------------------------------------
# generators
# ----------------------------------------------------------------------------------
def reference_domain(geometry: str) -> ExtendableGrid:
    if geometry == "Triangle2D":
        return ExtendableGrid([[0, 0], [1, 0], [0, 1]], [[1, 2, 3]],
                              bfacenodes=[[1, 2], [2, 3], [3, 1]], bfaceregions=[1, 2, 3])
    if geometry == "Tetrahedron3D":
        return ExtendableGrid([[0, 0, 0], [1, 0, 0], [0, 1, 0], [0, 0, 1]], [[1, 2, 3, 4]],
                              bfacenodes=[[1, 3, 2], [1, 2, 4], [2, 3, 4], [1, 4, 3]],
                              bfaceregions=[1, 2, 3, 4])
    raise ValueError(geometry)


def grid_unitsquare(geometry: str = "Triangle2D") -> ExtendableGrid:
    """4 triangles around the centre node, boundary regions 1..4 (bottom,right,top,left)."""
    assert geometry == "Triangle2D"
    coords = [[0, 0], [1, 0], [1, 1], [0, 1], [0.5, 0.5]]
    cells = [[1, 2, 5], [2, 3, 5], [3, 4, 5], [4, 1, 5]]
    return ExtendableGrid(coords, cells, bfacenodes=[[1, 2], [2, 3], [3, 4], [4, 1]],
                          bfaceregions=[1, 2, 3, 4])


def grid_unitcube(geometry: str = "Tetrahedron3D") -> ExtendableGrid:
    """15 nodes (8 corners, 6 face centres, body centre), 24 positively oriented tets."""
    assert geometry == "Tetrahedron3D"
    corners = [[0, 0, 0], [1, 0, 0], [1, 1, 0], [0, 1, 0], [0, 0, 1], [1, 0, 1], [1, 1, 1], [0, 1, 1]]
    quads = [[1, 2, 3, 4], [1, 2, 6, 5], [2, 3, 7, 6], [3, 4, 8, 7], [4, 1, 5, 8], [5, 6, 7, 8]]
    coords = [list(map(float, c)) for c in corners]
    for q in quads:
        coords.append(list(np.mean([corners[i - 1] for i in q], axis=0)))
    coords.append([0.5, 0.5, 0.5])
    x = np.array(coords)
    cells, bfn, bfr = [], [], []
    for qi, q in enumerate(quads):
        fc = 9 + qi
        for k in range(4):
            a, b = q[k], q[(k + 1) % 4]
            tet = [a, b, fc, 15]
            p = x[[t - 1 for t in tet]]
            if np.linalg.det(p[1:] - p[0]) < 0:
                tet = [b, a, fc, 15]
            cells.append(tet)
            # boundary triangle with outward normal (pointing away from the body centre)
            tri = [tet[0], tet[2], tet[1]]
            bfn.append(tri)
            bfr.append(qi + 1)
    return ExtendableGrid(x, cells, bfacenodes=bfn, bfaceregions=bfr)


def uniform_refine(grid: ExtendableGrid, nrefinements: int = 1) -> ExtendableGrid:
    """Red refinement: old nodes keep their numbers, midpoints are appended in global
    face (2D) / edge (3D) order, the children of cell c occupy a contiguous block."""
    for _ in range(nrefinements):
        grid = _refine_once(grid)
    return grid


def _refine_once(g: ExtendableGrid) -> ExtendableGrid:
    x = g.coords
    cn = g.cellnodes.astype(np.int64)
    nn = g.nnodes
    en = g.edgenodes.astype(np.int64) - 1
    mid = (x[en[:, 0]] + x[en[:, 1]]) / 2
    coords = np.concatenate([x, mid])
    ce = g.celledges.astype(np.int64)
    loc = np.concatenate([cn, ce + nn], axis=1)              # local id -> global node (1-based)
    rule = TRI_REFINE if g.dim == 2 else TET_REFINE
    children = loc[:, rule]                                  # (nc, nchild, dim+1)
    nchild = rule.shape[0]
    cells = children.reshape(-1, g.dim + 1)
    regions = np.repeat(g.cellregions, nchild)
    # boundary faces
    bn = g.bfacenodes.astype(np.int64)
    if bn.shape[0]:
        if g.dim == 2:
            bf = g.bfacefaces.astype(np.int64)
            m = bf + nn
            bnew = np.stack([np.stack([bn[:, 0], m], 1), np.stack([m, bn[:, 1]], 1)], 1).reshape(-1, 2)
            breg = np.repeat(g.bfaceregions, 2)
        else:
            # midpoint node of an edge given by its two end nodes
            ekeys = np.sort(g.edgenodes.astype(np.int64), axis=1)
            key = ekeys[:, 0] * (nn + 1) + ekeys[:, 1]
            order = np.argsort(key)
            skey = key[order]

            def midnode(a, b):
                lo, hi = np.minimum(a, b), np.maximum(a, b)
                pos = np.searchsorted(skey, lo * (nn + 1) + hi)
                return order[pos] + nn + 1
            a, b, c = bn[:, 0], bn[:, 1], bn[:, 2]
            mab, mbc, mca = midnode(a, b), midnode(b, c), midnode(c, a)
            bnew = np.stack([np.stack([a, mab, mca], 1), np.stack([mab, b, mbc], 1),
                             np.stack([mca, mbc, c], 1), np.stack([mab, mbc, mca], 1)], 1).reshape(-1, 3)
            breg = np.repeat(g.bfaceregions, 4)
    else:
        bnew, breg = None, None
    return ExtendableGrid(coords, cells, regions, bnew, breg)


def perturb_interior_nodes(grid: ExtendableGrid, rel: float = 0.1, seed: int = 20261017) -> ExtendableGrid:
    """Jitter interior nodes by U(-rel*h, rel*h) (SURVEY.md 8d 'perturbed variant')."""
    rng = np.random.default_rng(seed)
    x = grid.coords.copy()
    h = grid.cellvolumes.min() ** (1.0 / grid.dim)
    onb = np.zeros(grid.nnodes, bool)
    onb[grid.bfacenodes.astype(np.int64).ravel() - 1] = True
    x[~onb] += rng.uniform(-rel * h, rel * h, size=(int((~onb).sum()), grid.dim))
    return ExtendableGrid(x, grid.cellnodes, grid.cellregions, grid.bfacenodes, grid.bfaceregions)
