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
# several steps back to back: the edge kernel runs as a programmatic dependent of the previous diagonal kernel (multi-step graph)
import ctypes as C
ms = C.c_double(0)
G._lib.check(G._lib.lib().grmp_blf_numeric_steps(AP.AM.h, 1.0, 12, C.byref(ms)))
print("12 chained steps equal the single step:", np.array_equal(G.fetch_values(AP), nz))
# boundary-face items (ON_BFACES), convection form with a coefficient argument, Newton form
for gg, fe in ((g2, G.H1P2(2, 2)), (g, G.H1P2(1, 3))):
    sb = G.FESpace(fe, gg)
    M = G.assemble_csc(G.DiscreteSymmetricBilinearForm([G.Identity, G.Identity], [sb, sb], AT="ON_BFACES"), 1.0)[2]
    bb = G.FEVector([sb])
    G.assemble(bb[1], G.DiscreteLinearForm([G.Identity], [sb], G.fdot_action(G.DataFunction(np.ones(fe.ncomponents))), AT="ON_BFACES"))
    print("boundary measure x ncomp:", M.sum(), bb.entries.sum())
sv = G.FESpace(G.H1P2(2, 2), g2)
a = G.FEVector([sv])
a.entries[:] = np.linspace(0.0, 1.0, sv.ndofs)
Oc = G.ConvectionOperator(1, G.Identity, 2, 2)
A = G.FEMatrix([sv])
G.assemble_operator(A[1, 1], Oc, CurrentSolution=a)
print("convection nnz", A.nnz)
On = G.ConvectionOperator(1, G.Identity, 2, 2, newton=True)
A2 = G.FEMatrix([sv])
rhs = G.FEVector([sv])
G.full_assemble_operator(A2[1, 1], rhs[1], On, CurrentSolution=a)
print("newton nnz", A2.nnz, float(np.abs(rhs.entries).sum()) > 0)
