"""Bit-level pin against the real reference, where a dump of baseline/run_reference.jl exists (see tests/reference_dump.py);
the loader itself is exercised on a synthetic dump written in the same format."""
import numpy as np
import pytest

import grmp_b200 as G
import oracle as O
import reference_dump as RD
from parity import oracle_blf, rel_err, tier_report


def _form(s):
    return G.DiscreteSymmetricBilinearForm([G.Gradient, G.Gradient], [s, s])


def test_loader_round_trip_on_a_synthetic_dump(tmp_path, monkeypatch):
    g = G.uniform_refine(G.grid_unitcube("Tetrahedron3D"), 1)
    s = G.FESpace(G.H1P2(1, 3), g)
    csc = oracle_blf(_form(s), 1.0)
    RD.write(str(tmp_path), g, s, csc)
    monkeypatch.setenv("GRMP_REFERENCE_DUMP", str(tmp_path))
    d = RD.dump_dir()
    assert d == str(tmp_path)
    g2, s2, ref, same_enum = RD.load(d)
    assert same_enum and g2.ncells == g.ncells and s2.ndofs == s.ndofs
    cp, rv, nz = oracle_blf(_form(s2), 1.0)
    assert np.array_equal(cp, ref[0]) and np.array_equal(rv, ref[1]) and np.array_equal(nz, ref[2])


def test_oracle_bit_identical_to_reference_dump():
    d = RD.dump_dir()
    if d is None:
        pytest.skip("no reference dump (run baseline/run_reference.jl where Julia exists, set GRMP_REFERENCE_DUMP)")
    g, s, ref, same_enum = RD.load(d)
    print("this package's edge enumeration equals the reference's CellDofs:", same_enum)
    cp, rv, nz = oracle_blf(_form(s), 1.0)
    assert np.array_equal(cp, ref[0]), "colptr differs from the reference"
    assert np.array_equal(rv, ref[1]), "rowval differs from the reference"
    assert np.array_equal(nz, ref[2]), f"nzval not bit-identical to the reference (max rel {rel_err(nz, ref[2]):.3e})"


@pytest.mark.gpu
def test_gpu_against_reference_dump():
    d = RD.dump_dir()
    if d is None:
        pytest.skip("no reference dump (run baseline/run_reference.jl where Julia exists, set GRMP_REFERENCE_DUMP)")
    g, s, ref, _ = RD.load(d)
    for path, exact in ((G._lib.PATH_GENERIC, True), (G._lib.PATH_AUTO, False), (G._lib.PATH_COLUMNS, False)):
        AP = _form(s)
        G.blf_set_path(AP, path)
        cp, rv, nz = G.assemble_csc(AP, 1.0)
        assert np.array_equal(cp, ref[0]) and np.array_equal(rv, ref[1])
        if exact:
            assert np.array_equal(nz, ref[2])
        else:
            assert rel_err(nz, ref[2]) <= 1e-12, tier_report(nz, ref[2])
