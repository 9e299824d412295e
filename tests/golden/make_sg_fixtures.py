"""Fixtures from the reference's own mesh files (/root/reference/assets/*.sg, SimplexGrid 2.1): the parsed NODES / CELLS lists
-- test INPUTS with explicit cell and node numbering, nothing this package enumerates -- plus the oracle's matrices for forms
whose dof map is the node numbering itself:
  * Example202 (examples/Example202_LinearElasticity2D.jl:39-53): H1P1{2}, HookStiffnessOperator2D(mu, lambda) with
    E = 1000, nu = 0.4 on 2d_grid_cookmembrane.sg
  * P1 Laplace stiffness and P1 mass matrix on every 2D mesh file
  * the P1 boundary mass matrix (AT = ON_BFACES, boundarydata.jl:321) on the explicit FACES list of the file (BFaceNodes, BFaceRegions)
The reference tree is not available on the GPU box, so the vectors are committed:  python tests/golden/make_sg_fixtures.py"""
import glob
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
ASSETS = "/root/reference/assets"


def forms(G, g):
    mu = 1000 / (1 + 0.4)
    lam = (0.4 / (1 - 2 * 0.4)) * mu    # Example202_LinearElasticity2D.jl:43-44: mu = (1/(1+nu))*E, lambda = (nu/(1-2nu))*mu
    s1 = G.FESpace(G.H1P1(1), g)
    s2 = G.FESpace(G.H1P1(2), g)
    return {
        "laplace": G.DiscreteSymmetricBilinearForm([G.Gradient, G.Gradient], [s1, s1]),
        "mass": G.DiscreteSymmetricBilinearForm([G.Identity, G.Identity], [s1, s1]),
        "hooke": G.DiscreteBilinearForm([G.SymmetricGradient(1), G.SymmetricGradient(1)], [s2, s2], G.HookeAction(2, mu, lam)),
        "bmass": G.DiscreteSymmetricBilinearForm([G.Identity, G.Identity], [s1, s1], AT="ON_BFACES"),
    }


if __name__ == "__main__":
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import grmp_b200 as G
    from grmp_b200.sgfile import parse_sg, simplexgrid
    from parity import oracle_blf
    for path in sorted(glob.glob(os.path.join(ASSETS, "2d_*.sg"))):
        name = os.path.basename(path)[:-3]
        d = parse_sg(open(path).read())
        g = simplexgrid(d)
        out = {"coords": d["coords"], "cellnodes": d["cellnodes"], "cellregions": d["cellregions"],
               "bfacenodes": d["bfacenodes"], "bfaceregions": d["bfaceregions"]}
        for fname, AP in forms(G, g).items():
            cp, rv, nz = oracle_blf(AP, 1.0)
            out[fname + "_colptr"], out[fname + "_rowval"], out[fname + "_nzval"] = cp, rv, nz
        np.savez_compressed(os.path.join(HERE, "sg_" + name + ".npz"), **out)
        print(name, "nodes", g.nnodes, "cells", g.ncells, "area", g.cellvolumes.sum())
