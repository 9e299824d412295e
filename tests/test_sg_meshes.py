"""Enumeration-free parity cases: the reference's own mesh files (assets/*.sg, SimplexGrid 2.1) carry explicit
NODES / CELLS lists, and for node-based spaces (H1P1) the dof map IS the node numbering -- nothing in these cases depends on this
package's face / edge enumeration or mesh generators.  The parsed lists are committed under tests/golden/sg_*.npz
(make_sg_fixtures.py) because the reference tree does not travel to the GPU box.

CPU: reader vs fixture, oracle vs frozen vectors, analytic checks (area of Cook's membrane, row sums, rigid body modes).
GPU: libgrmp_cuda (generic bit-exact path, column kernels, atomic scatter) vs the oracle on the same lists, incl. Example202's form."""
import glob
import os

import numpy as np
import pytest

import grmp_b200 as G
from grmp_b200.sgfile import parse_sg, simplexgrid
from parity import oracle_blf, oracle_scale, rel_err, tier_report

HERE = os.path.dirname(os.path.abspath(__file__))
FIX = sorted(glob.glob(os.path.join(HERE, "golden", "sg_*.npz")))
ASSETS = "/root/reference/assets"


def grid_of(d):
    return G.ExtendableGrid(d["coords"], d["cellnodes"], d["cellregions"], d["bfacenodes"], d["bfaceregions"])


def forms(g):
    from golden.make_sg_fixtures import forms as f
    return f(G, g)


def test_fixtures_exist():
    assert len(FIX) >= 4


@pytest.mark.parametrize("path", FIX, ids=[os.path.basename(p) for p in FIX])
def test_reader_reproduces_the_committed_lists(path):
    src = os.path.join(ASSETS, os.path.basename(path)[3:-4] + ".sg")
    if not os.path.exists(src):
        pytest.skip("reference assets not present on this machine")
    d = np.load(path)
    p = parse_sg(open(src).read())
    assert np.array_equal(p["coords"], d["coords"]) and np.array_equal(p["cellnodes"], d["cellnodes"])
    assert np.array_equal(p["cellregions"], d["cellregions"])
    assert np.array_equal(p["bfacenodes"], d["bfacenodes"]) and np.array_equal(p["bfaceregions"], d["bfaceregions"])
    g = simplexgrid(src)
    assert g.ncells == d["cellnodes"].shape[0] and g.bfacenodes.shape[0] > 0


def test_sg_parser_rejects_garbage():
    with pytest.raises(ValueError):
        parse_sg("hello")
    with pytest.raises(ValueError):
        parse_sg("SimplexGrid 2.1\nDIMENSION\n2\nNODES\n1 2\n0 0\nEND\n")
    d = parse_sg("SimplexGrid 2.1\n#c\nDIMENSION\n2\nNODES\n3 2\n0 0\n1 0\n0 1\nCELLS\n1\n1 2 3 7\nFACES\n0\nEND\n")
    assert d["cellnodes"].tolist() == [[1, 2, 3]] and d["cellregions"].tolist() == [7]
    e = parse_sg("SimplexGrid 2.1\nDIMENSION\n2\nNODES\n3 2\n0\n0\n1\n0\n0\n1\nCELLS\n1\n1\n2\n3\n7\nFACES\n1\n1 2 4\nEND\n")     # one number per line
    assert np.array_equal(e["coords"], d["coords"]) and e["bfaceregions"].tolist() == [4]


@pytest.mark.parametrize("path", FIX, ids=[os.path.basename(p) for p in FIX])
def test_oracle_on_reference_meshes(path):
    d = np.load(path)
    g = grid_of(d)
    assert np.all(g.cellvolumes > 0)
    if "cookmembrane" in path:           # Cook's membrane: trapezoid (0,0) (48,44) (48,60) (0,44)
        assert abs(g.cellvolumes.sum() - 0.5 * (44 + 16) * 48) < 1e-9
    for name, AP in forms(g).items():
        cp, rv, nz = oracle_blf(AP, 1.0)
        assert np.array_equal(cp, d[name + "_colptr"]) and np.array_equal(rv, d[name + "_rowval"])
        assert np.array_equal(nz, d[name + "_nzval"]), name
    import scipy.sparse as sp_
    n = g.nnodes
    K = sp_.csc_matrix((d["laplace_nzval"], d["laplace_rowval"] - 1, d["laplace_colptr"] - 1), shape=(n, n))
    M = sp_.csc_matrix((d["mass_nzval"], d["mass_rowval"] - 1, d["mass_colptr"] - 1), shape=(n, n))
    assert np.abs(K @ np.ones(n)).max() < 1e-12 * np.abs(K.data).max()           # constants are in the kernel
    assert abs(M.sum() - g.cellvolumes.sum()) < 1e-12 * g.cellvolumes.sum()        # partition of unity
    x = d["coords"][:, 0]
    assert abs(x @ (K @ x) - g.cellvolumes.sum()) < 1e-10 * g.cellvolumes.sum()    # |grad x|^2 = 1
    H = sp_.csc_matrix((d["hooke_nzval"], d["hooke_rowval"] - 1, d["hooke_colptr"] - 1), shape=(2 * n, 2 * n))
    assert abs(H - H.T).max() < 1e-12 * np.abs(H.data).max()
    # boundary mass matrix on the file's explicit FACES list: 1' M 1 = measure of the listed boundary
    B = sp_.csc_matrix((d["bmass_nzval"], d["bmass_rowval"] - 1, d["bmass_colptr"] - 1), shape=(n, n))
    assert abs(B.sum() - g.bfacevolumes.sum()) < 1e-12 * g.bfacevolumes.sum()
    if "cookmembrane" in path:           # perimeter of the trapezoid (0,0) (48,44) (48,60) (0,44)
        assert abs(B.sum() - (np.hypot(48, 44) + 16 + np.hypot(48, 16) + 44)) < 1e-9
    for rigid in (np.concatenate([np.ones(n), np.zeros(n)]), np.concatenate([np.zeros(n), np.ones(n)]),
                  np.concatenate([-d["coords"][:, 1], d["coords"][:, 0]])):      # translations and the rotation carry no strain
        assert np.abs(H @ rigid).max() < 1e-9 * np.abs(H.data).max() * np.abs(rigid).max()


@pytest.mark.gpu
@pytest.mark.parametrize("path", FIX, ids=[os.path.basename(p) for p in FIX])
@pytest.mark.parametrize("backend", ["generic", "columns", "atomic"])
def test_gpu_on_reference_meshes(path, backend):
    d = np.load(path)
    g = grid_of(d)
    code = {"generic": G._lib.PATH_GENERIC, "columns": G._lib.PATH_COLUMNS, "atomic": G._lib.PATH_ATOMIC}[backend]
    for name, AP in forms(g).items():
        if AP.AT == "ON_BFACES":              # boundary forms run on the bit-exact path only
            if backend != "generic":
                continue
        else:
            G.blf_set_path(AP, code)
        cp, rv, nz = G.assemble_csc(AP, 1.0)
        assert np.array_equal(cp, d[name + "_colptr"]) and np.array_equal(rv, d[name + "_rowval"]), name
        if backend == "generic":
            assert np.array_equal(nz, d[name + "_nzval"]), name
        else:
            S = oracle_scale(AP, 1.0)
            assert rel_err(nz, d[name + "_nzval"], S) <= 1e-12, name + ": " + tier_report(nz, d[name + "_nzval"], S)
