"""CPU-side tests (-m "not gpu"): host mirror logic, oracle cross-checks, and that the C-ABI
library loads and exports every symbol include/grmp.h declares (no compute without a GPU)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import grmp_b200 as G
import oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "grmp.h")).read()
    declared = set(re.findall(r"\b(grmp_[a-z_]+)\s*\(", hdr))
    assert len(declared) >= 20
    G._lib.build()
    L = C.CDLL(G._lib.LIB_PATH)
    for name in sorted(declared):
        assert hasattr(L, name), f"{name} declared in grmp.h but not exported"
    assert declared == set(G._lib.EXPORTS), declared ^ set(G._lib.EXPORTS)


@pytest.mark.skipif(os.path.exists("/dev/nvidia0"), reason="only meaningful on a box without a GPU")
def test_product_path_fails_loudly_without_gpu():
    g = G.uniform_refine(G.grid_unitsquare("Triangle2D"), 1)
    s = G.FESpace(G.H1P1(1), g)
    AP = G.DiscreteSymmetricBilinearForm([G.Gradient, G.Gradient], [s, s])
    with pytest.raises(G._lib.GrmpError, match="no CUDA device"):
        G.assemble_csc(AP)


def test_quadrature_mirror_matches_oracle():
    for edim, geo, maxo in ((2, "Triangle2D", 11), (3, "Tetrahedron3D", 8)):
        for order in range(0, maxo + 1):
            q = G.QuadratureRule(geo, order)
            x, w = O.qrule(edim, order)
            assert x.shape == q.xref.shape
            # hard-coded rules are bitwise equal; eigen-generated Stroud rules agree to rounding
            stroud = edim == 2 and order >= 3 and order != 8
            tol = 2e-15 if stroud else 0.0
            assert np.abs(x - q.xref).max() <= tol and np.abs(w - q.w).max() <= tol, (edim, order)


def test_reference_tables_mirror_bit_identical_to_oracle():
    cases = [(2, "Triangle2D", [G.H1P1(1), G.H1P1(2), G.H1P2(1, 2), G.H1P2(2, 2), G.H1BR(2), G.HDIVRT0(2), G.HDIVBDM1(2), G.L2P0(1)]),
             (3, "Tetrahedron3D", [G.H1P1(1), G.H1P1(3), G.H1P2(1, 3), G.H1BR(3), G.HDIVRT0(3), G.HDIVBDM1(3), G.L2P0(1)])]
    for edim, geo, fes in cases:
        for fe in fes:
            for order in (0, 2, 4, 6):
                q = G.QuadratureRule(geo, order)
                v, d = G.reference_tables(fe, edim, q.xref, True)
                nc_arg = fe.ncomponents
                ov, od = O.reftables(fe.code, nc_arg, edim, q.xref, fe.ndofs_all(edim), fe.ncomponents)
                assert np.array_equal(v, ov) and np.array_equal(d, od), (fe, order)


def _naive_first_encounter(cells, rule):
    seen, items, cellitems = {}, [], []
    for c in cells:
        row = []
        for loc in rule:
            key = tuple(sorted(int(c[i]) for i in loc))
            if key not in seen:
                seen[key] = len(items) + 1
                items.append([int(c[i]) for i in loc])
            row.append(seen[key])
        cellitems.append(row)
    return np.array(items), np.array(cellitems)


@pytest.mark.parametrize("dim", [2, 3])
def test_grid_adjacencies_match_naive_enumeration(dim):
    g = G.uniform_refine(G.grid_unitsquare("Triangle2D") if dim == 2 else G.grid_unitcube("Tetrahedron3D"), 2 if dim == 2 else 1)
    rule = G.grid.TRI_FACENODES if dim == 2 else G.grid.TET_FACENODES
    fn, cf = _naive_first_encounter(g.cellnodes, rule)
    assert np.array_equal(fn, g.facenodes) and np.array_equal(cf, g.cellfaces)
    if dim == 3:
        en, ce = _naive_first_encounter(g.cellnodes, G.grid.TET_EDGENODES)
        assert np.array_equal(en, g.edgenodes) and np.array_equal(ce, g.celledges)
    # signs: +1 for the first cell of a face, -1 for the second; normals outward of the first cell
    for f in range(g.nfaces):
        c0 = g.facecells[f, 0] - 1
        lf = list(g.cellfaces[c0]).index(f + 1)
        assert g.cellfacesigns[c0, lf] == 1
        c1 = g.facecells[f, 1] - 1
        if c1 >= 0:
            lf1 = list(g.cellfaces[c1]).index(f + 1)
            assert g.cellfacesigns[c1, lf1] == -1
    assert abs(g.cellvolumes.sum() - 1) < 1e-14
    # Euler characteristic
    if dim == 2:
        assert g.nnodes - g.nfaces + g.ncells == 1
    else:
        assert g.nnodes - g.nedges + g.nfaces - g.ncells == 1


def test_survey_entity_counts_level3():
    # recurrences V'=V+E, E'=2E+3F+C, F'=4F+8C, C'=8C for red refinement of tets
    g = G.grid_unitcube("Tetrahedron3D")
    V, E, F, Cn = g.nnodes, g.nedges, g.nfaces, g.ncells
    for _ in range(2):
        g = G.uniform_refine(g, 1)
        V, E, F, Cn = V + E, 2 * E + 3 * F + Cn, 4 * F + 8 * Cn, 8 * Cn
        assert (g.nnodes, g.nedges, g.nfaces, g.ncells) == (V, E, F, Cn)


def _naive_celldofs(space):
    """literal transcription of the loop structure of init_dofmap_from_pattern! (dofmaps.jl:280-360)"""
    g = space.xgrid
    nc = space.ncomponents
    out = []
    for cell in range(g.ncells):
        dofs = []
        for c in range(nc):
            offset = c * space.coffset
            for ch, each, q in space.segments:
                if not each:
                    continue
                adj, n = {"N": (g.cellnodes, g.nnodes), "F": (g.cellfaces, g.nfaces), "E": (g.celledges, g.nedges)}[ch]
                for k in range(adj.shape[1]):
                    for m in range(q):
                        dofs.append(adj[cell, k] + offset + m * n)
                offset += n * q
        offset = nc * space.coffset
        for ch, each, q in space.segments:
            if each:
                continue
            adj, n = {"f": (g.cellfaces, g.nfaces), "e": (g.celledges, g.nedges)}[ch]
            for k in range(adj.shape[1]):
                for m in range(q):
                    dofs.append(adj[cell, k] + offset + m * n)
            offset += n * q
        out.append(dofs)
    return np.array(out, dtype=np.int32)


@pytest.mark.parametrize("fe,dim", [(G.H1P1(1), 2), (G.H1P2(2, 2), 2), (G.H1P2(1, 3), 3), (G.H1BR(2), 2), (G.H1BR(3), 3),
                                    (G.HDIVRT0(3), 3), (G.HDIVBDM1(2), 2), (G.HDIVBDM1(3), 3)])
def test_celldofs_match_reference_loop(fe, dim):
    g = G.uniform_refine(G.grid_unitsquare("Triangle2D") if dim == 2 else G.grid_unitcube("Tetrahedron3D"), 1)
    s = G.FESpace(fe, g)
    assert np.array_equal(s.celldofs, _naive_celldofs(s))
    assert s.celldofs.min() == 1 and s.celldofs.max() == s.ndofs
    assert np.unique(s.celldofs).size == s.ndofs


def test_metric_config_dof_counts():
    # P2 on L levels of the unit cube: ndofs = nnodes + nedges
    g = G.uniform_refine(G.grid_unitcube("Tetrahedron3D"), 2)
    s = G.FESpace(G.H1P2(1, 3), g)
    assert s.ndofs == g.nnodes + g.nedges and s.celldofs.shape == (g.ncells, 10)
    p0 = G.FESpace(G.L2P0(1), g)
    assert np.array_equal(p0.celldofs[:, 0], np.arange(1, g.ncells + 1))


def test_fematrix_union_merge_keeps_explicit_zeros():
    g = G.uniform_refine(G.grid_unitsquare("Triangle2D"), 1)
    s = G.FESpace(G.H1P1(1), g)
    A = G.FEMatrix([s])
    cp = np.array([1, 2, 3] + [3] * (s.ndofs - 2), dtype=np.int64)
    A.add_csc(cp, np.array([1, 2], dtype=np.int64), np.array([1.0, 2.0]))
    A.add_csc(cp, np.array([1, 3], dtype=np.int64), np.array([-1.0, 5.0]))
    assert A.nnz == 3
    assert np.array_equal(A.rowval, [1, 2, 3]) and np.array_equal(A.nzval, [0.0, 2.0, 5.0])   # 1 + (-1) stays as explicit zero
    A.fill_block_zero(A[1, 1])
    assert A.nnz == 3 and np.all(A.nzval == 0)


def test_oracle_pattern_is_subset_of_structural_pattern_and_perturbed_is_structural():
    g = G.uniform_refine(G.grid_unitcube("Tetrahedron3D"), 1)
    s = G.FESpace(G.H1P2(1, 3), g)

    def nnz(grid, space):
        A = O.OracleMatrix(space.ndofs, space.ndofs)
        O.blf_assemble(A, grid, space, space, O.OP_GRAD, O.OP_GRAD, apt=O.APT_SYMMETRIC)
        return A.csc()[1].size
    structural = np.unique(np.repeat(s.celldofs.astype(np.int64), 10, axis=1).ravel() * (s.ndofs + 1)
                           + np.tile(s.celldofs.astype(np.int64), (1, 10)).ravel()).size
    n_axis = nnz(g, s)
    gp = G.perturb_interior_nodes(g)
    sp = G.FESpace(G.H1P2(1, 3), gp)
    assert n_axis <= structural
    assert nnz(gp, sp) <= structural


def test_oracle_reproduces_committed_golden_fixtures():
    import glob
    from golden.make_golden import CASES, build_case, oracle_csc
    files = sorted(f for f in glob.glob(os.path.join(ROOT, "tests", "golden", "*.npz")) if not os.path.basename(f).startswith("sg_"))
    assert len(files) == len(CASES)
    for f in files:
        name = os.path.basename(f)[:-4]
        g, s, AP, factor = build_case(G, CASES[name])
        cp, rv, nz = oracle_csc(O, g, s, AP, factor)
        d = np.load(f)
        assert np.array_equal(cp, d["colptr"]) and np.array_equal(rv, d["rowval"]) and np.array_equal(nz, d["nzval"]), name
    # the closed-form anchor of the reference tree (examples/ExampleA01_RationalMassMatrix.jl:17-35)
    d = np.load(os.path.join(ROOT, "tests", "golden", "A01_P1_mass_reference_triangle.npz"))
    M = np.zeros((3, 3))
    for j in range(3):
        for k in range(d["colptr"][j] - 1, d["colptr"][j + 1] - 1):
            M[d["rowval"][k] - 1, j] = d["nzval"][k]
    assert np.abs(M - 0.5 / 12 * np.array([[2, 1, 1], [1, 2, 1], [1, 1, 2]])).max() < 1e-16


@pytest.mark.parametrize("dim", [2, 3])
def test_interpolated_and_homogeneous_boundary_data_host_part(dim):
    """boundarydata.jl:100-258 (no device work): P2 interpolation on boundary faces = nodal values + edge-mean preserving edge dofs,
    exact for a quadratic trace; homogeneous regions are zeroed; fixed dofs come back in first-occurrence order"""
    g = G.perturb_interior_nodes(G.uniform_refine(G.grid_unitsquare() if dim == 2 else G.grid_unitcube(), 1), 0.1)
    s = G.FESpace(G.H1P2(2, dim), g)
    u = lambda x: np.stack([1.0 + x[0] * x[1] - 2.0 * x[dim - 1] ** 2, x[0] ** 2 - 0.5 * x[1]])
    data = G.DataFunction(u, [2, dim], bonus_quadorder=2)
    t = G.FEVector([s])
    t.entries[:] = 9.0
    O = [G.BoundaryData(G.InterpolateDirichletBoundary, data=data, regions=[1, 2]), G.BoundaryData(G.HomogeneousDirichletBoundary, regions=[4])]
    fixed = G.boundarydata(t[1], O)
    assert len(set(fixed)) == fixed.size
    reg12 = np.unique(s.bfacedofs[np.isin(g.bfaceregions, [1, 2])])
    reg4 = np.unique(s.bfacedofs[np.isin(g.bfaceregions, [4])])
    assert set(fixed) == set(reg12) | set(reg4)
    en = (g.facenodes if dim == 2 else g.edgenodes).astype(np.int64) - 1
    xdof = np.concatenate([g.coords, 0.5 * (g.coords[en[:, 0]] + g.coords[en[:, 1]])])
    exact = np.concatenate(list(u(xdof.T)))                  # component-major dof numbering
    only12 = np.setdiff1d(reg12, reg4) - 1
    assert np.abs(t.entries[only12] - exact[only12]).max() < 1e-13
    assert np.all(t.entries[reg4 - 1] == 0.0)
    untouched = np.setdiff1d(np.arange(s.ndofs), fixed - 1)
    assert np.all(t.entries[untouched] == 9.0)


def test_julia_glue_binds_only_declared_entry_points():
    """every `ccall((:grmp_..., lib), ...)` of julia/GRMPCuda.jl and every entry point shown in INTEGRATION.md is declared in include/grmp.h
    (the glue cannot run here: no Julia; this keeps it from drifting away from the ABI)"""
    hdr = open(os.path.join(ROOT, "include", "grmp.h")).read()
    declared = set(re.findall(r"\b(grmp_[a-z_0-9]+)\s*\(", hdr))
    jl = open(os.path.join(ROOT, "julia", "GRMPCuda.jl")).read()
    used = set(re.findall(r":(grmp_[a-z_0-9]+)", jl))
    assert len(used) >= 25
    assert used <= declared, sorted(used - declared)
    doc = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    named = set(re.findall(r"`(grmp_[a-z_0-9]+)`", doc)) - {"grmp_evaltab", "grmp_b200", "grmp_stats", "grmp_ctx", "grmp_grid", "grmp_space", "grmp_blf", "grmp_lf", "grmp_ii"}
    assert named <= declared, sorted(named - declared)
    # operator codes of the glue agree with the header's enum
    ops = dict(re.findall(r"(GRMP_OP_[A-Z0-9_]+) = (\d+)", hdr))
    glue = {"Identity": "GRMP_OP_ID", "Gradient": "GRMP_OP_GRAD", "SymmetricGradient{1}": "GRMP_OP_SYMGRAD", "Divergence": "GRMP_OP_DIV",
            "NormalFlux": "GRMP_OP_NORMALFLUX"}
    for jname, cname in glue.items():
        assert "opcode(::Type{%s}) = %s" % (jname, ops[cname]) in jl, jname
