"""Small end-to-end case for compute-sanitizer (memcheck / racecheck logs are kept under profiles/):
  compute-sanitizer --tool memcheck python tools/sanitize_case.py
covers the ring-walk kernels incl. the device-side record build, one closed-form and one quadrature column kernel, a LinearForm
gather kernel, the bit-exact generic path, matmul / residual and the ItemIntegrator."""
import sys
sys.path.insert(0, '.')
import numpy as np
import grmp_b200 as G

g = G.perturb_interior_nodes(G.uniform_refine(G.grid_unitcube("Tetrahedron3D"), 2))
s = G.FESpace(G.H1P2(1, 3), g)
AP = G.DiscreteSymmetricBilinearForm([G.Gradient, G.Gradient], [s, s])
G.blf_set_path(AP, G._lib.PATH_FAST)
cp, rv, nz = G.assemble_csc(AP, 1.0)
_, _, nz2 = G.assemble_csc(AP, 1.0, skip_preps=True)
print("ring walk ok", nz.size, np.array_equal(nz, nz2), G.blf_stats(AP).ntiles)
x = np.ones(s.ndofs)
print("residual |A 1|^2 =", G.residual(AP, x)[1])
for path, name in ((G._lib.PATH_COLUMNS, "columns"), (G._lib.PATH_GENERIC, "generic")):
    AP2 = G.DiscreteSymmetricBilinearForm([G.Gradient, G.Gradient], [s, s])
    G.blf_set_path(AP2, path)
    _, _, nzc = G.assemble_csc(AP2, 1.0)
    print(name, "max dev vs ring walk / amax", np.abs(nzc - nz).max() / np.abs(nz).max())
g2 = G.perturb_interior_nodes(G.uniform_refine(G.grid_unitsquare("Triangle2D"), 3))
s2 = G.FESpace(G.H1P2(2, 2), g2)
APh = G.DiscreteBilinearForm([G.SymmetricGradient(1), G.SymmetricGradient(1)], [s2, s2], G.HookeAction(2, 2.0, 3.0))
print("hooke path", G.assemble_csc(APh, 1.0)[2].size, G.blf_stats(APh).path)
b = G.FEVector([s2])
G.assemble_operator(b[1], G.LinearForm(G.Identity, G.DataFunction([1.0, 2.0])))
print("lf sum", b.entries.sum())
u = G.FEVector([s])
u.entries[:] = 1.0
print("integral of 1 =", G.evaluate(G.ItemIntegrator([G.Identity]), u[1]))
