"""The reference's own solve-level tests (test/runtests.jl) replayed through the device path, under the reference's tolerance (6e-12, runtests.jl:23):
"L2-Bestapproximations" (runtests.jl:355-447) for every FEType of its catalogue that is on the path.  Assembly (AUTO path: column kernels, LinearForm
gather kernels) and the L2ErrorIntegrator run on the device, the solve on the host like in the reference.  The H1 twin with best-approximation boundary data
is tests/test_gpu_bfaces.py::test_reference_kat_h1_bestapproximation_with_bestapprox_boundary; the oracle-only twins are in tests/test_oracle_kat.py."""
import numpy as np
import pytest
import scipy.sparse as sp
import scipy.sparse.linalg as spla

import grmp_b200 as G
from test_oracle_kat import L2_CATALOG, catalog_fetype, exact_function

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("dim,name,order", L2_CATALOG, ids=["%s{%d} order %d" % (n, d, o) for d, n, o in L2_CATALOG])
def test_reference_l2_bestapproximation_on_device(dim, name, order):
    g = G.uniform_refine(G.grid_unitsquare() if dim == 2 else G.grid_unitcube(), 1)
    s = G.FESpace(catalog_fetype(name, dim), g)
    udata = G.DataFunction(exact_function(dim, order), [dim, dim], bonus_quadorder=order)
    A = G.DiscreteBilinearForm([G.Identity, G.Identity], [s, s])                                   # ReactionOperator()
    cp, rv, nz = G.assemble_csc(A, 1.0)
    rhs = G.FEVector([s])
    G.assemble(rhs[1], G.DiscreteLinearForm([G.Identity], [s], G.fdot_action(udata)))              # LinearForm(Identity, u)
    sol = G.FEVector([s])
    sol.entries[:] = spla.spsolve(sp.csc_matrix((nz, rv - 1, cp - 1), shape=(s.ndofs, s.ndofs)), rhs.entries)
    err2 = G.evaluate(G.L2ErrorIntegrator(udata, G.Identity, quadorder=order), sol[1])
    assert np.sqrt(np.abs(np.asarray(err2)).sum()) < 6e-12
