import sys; sys.path.insert(0,'.')
import numpy as np, grmp_b200 as G
g = G.perturb_interior_nodes(G.uniform_refine(G.grid_unitcube("Tetrahedron3D"), 2))
s = G.FESpace(G.H1P2(1,3), g)
AP = G.DiscreteSymmetricBilinearForm([G.Gradient, G.Gradient], [s, s])
G.blf_set_path(AP, G._lib.PATH_FAST)
cp, rv, nz = G.assemble_csc(AP, 1.0)
_, _, nz2 = G.assemble_csc(AP, 1.0, skip_preps=True)
print("ok", nz.size, np.array_equal(nz, nz2), G.blf_stats(AP).ntiles)
