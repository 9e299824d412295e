"""Generates the golden fixtures tests/golden/*.npz with the CPU oracle (oracle/grmp_oracle.cpp).

The reference ships no stored matrices and Julia is not available here, so these are
outputs of the oracle -- itself pinned by the reference's analytic known-answer tests
(tests/test_oracle_kat.py) -- frozen so that later changes of oracle, host mirror or kernels are
detected.  Regenerate with:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))

# name -> (geometry, level, perturbed, fetype ctor name + args, operator pair, pattern kind, factor)
CASES = {
    "C1_example201_P2tri_laplace": ("square", 4, False, ("H1P2", (1, 2)), ("Gradient", "Gradient"), "sym", 1e-3),
    "C2_P1tet_laplace_L2": ("cube", 2, False, ("H1P1", (1,)), ("Gradient", "Gradient"), "sym", 1.0),
    "C2_P2tet_laplace_L1": ("cube", 1, False, ("H1P2", (1, 3)), ("Gradient", "Gradient"), "sym", 1.0),
    "C2_P2tet_laplace_L1_perturbed": ("cube", 1, True, ("H1P2", (1, 3)), ("Gradient", "Gradient"), "sym", 1.0),
    "C5_RT0tet_mass_L1": ("cube", 1, False, ("HDIVRT0", (3,)), ("Identity", "Identity"), "sym", 1.0),
    "C5_BDM1tet_mass_L0": ("cube", 0, False, ("HDIVBDM1", (3,)), ("Identity", "Identity"), "sym", 1.0),
    "C4_BRtri_laplace_L2": ("square", 2, False, ("H1BR", (2,)), ("Gradient", "Gradient"), "sym", 1.0),
    "A01_P1_mass_reference_triangle": ("reftri", 0, False, ("H1P1", (1,)), ("Identity", "Identity"), "sym", 1.0),
}


def build_case(G, case):
    geo, level, pert, (fe, feargs), ops, kind, factor = case
    if geo == "square":
        g = G.uniform_refine(G.grid_unitsquare("Triangle2D"), level)
    elif geo == "cube":
        g = G.uniform_refine(G.grid_unitcube("Tetrahedron3D"), level)
    else:
        g = G.reference_domain("Triangle2D")
    if pert:
        g = G.perturb_interior_nodes(g)
    s = G.FESpace(getattr(G, fe)(*feargs), g)
    ctor = G.DiscreteSymmetricBilinearForm if kind == "sym" else G.DiscreteBilinearForm
    AP = ctor([getattr(G, ops[0]), getattr(G, ops[1])], [s, s])
    return g, s, AP, factor


def oracle_csc(O, g, s, AP, factor):
    A = O.OracleMatrix(s.ndofs, s.ndofs)
    O.blf_assemble(A, g, s, s, AP.operators[0].code, AP.operators[1].code, apt=O.APT_SYMMETRIC, factor=factor)
    return A.csc()


if __name__ == "__main__":
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import grmp_b200 as G
    import oracle as O
    for name, case in CASES.items():
        g, s, AP, factor = build_case(G, case)
        cp, rv, nz = oracle_csc(O, g, s, AP, factor)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), colptr=cp, rowval=rv, nzval=nz)
        print(name, "ndofs", s.ndofs, "nnz", rv.size)
