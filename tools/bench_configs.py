#!/usr/bin/env python
"""Throughput of the generic (bit-exact) path on the other BASELINE.json configurations (C3-C5) and of the
generic vs fast path on C2 -- numbers for DESIGN.md 5.  Not the headline bench (that is bench.py).

  python tools/bench_configs.py [--small]
"""
import argparse
import ctypes as C
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import grmp_b200 as G  # noqa: E402


def run(name, AP, factor=1.0, steps=5, path=None, nodes=0):
    L = G._lib.lib()
    G.prepare_assembly(AP)
    h = AP.AM.h
    if path is not None:
        G._lib.check(L.grmp_blf_set_path(h, path))
    nnz = C.c_int64(0)
    t = time.time()
    G._lib.check(L.grmp_blf_symbolic(h, factor, C.byref(nnz)))
    tsym = time.time() - t
    ms = C.c_double(0)
    G._lib.check(L.grmp_blf_numeric_steps(h, factor, 2, C.byref(ms)))
    G._lib.check(L.grmp_blf_numeric_steps(h, factor, steps, C.byref(ms)))
    st = G.blf_stats(AP)
    s1, s2 = AP.FES
    g = s1.xgrid
    per = ms.value / steps
    balg = 8 * nnz.value + g.ncells * 4 * (g.dim + 1 + s1.nd_cell + (s2.nd_cell if s2 is not s1 else 0)) + 8 * g.dim * g.nnodes
    out = {"config": name, "ncells": int(g.ncells), "ndofs": [int(s1.ndofs), int(s2.ndofs)], "nnz": int(nnz.value),
           "path": {1: "generic", 2: "fast"}[int(st.path)], "ms_per_assembly": per, "nnz_per_s": nnz.value / (per * 1e-3),
           "algorithmic_GBs": balg / (per * 1e-3) / 1e9, "frac_of_6538.9": balg / (per * 1e-3) / 1e9 / 6538.9, "symbolic_s": tsym}
    print(json.dumps(out), flush=True)
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--small", action="store_true")
    a = ap.parse_args()
    Lt, Lq = (6, 4) if a.small else (9, 6)
    # C2: P1 / P2 tets
    g3 = G.uniform_refine(G.grid_unitcube("Tetrahedron3D"), Lq - 1)
    s = G.FESpace(G.H1P1(1), g3)
    run("C2 P1 tet Laplace L%d" % (Lq - 1), G.DiscreteSymmetricBilinearForm([G.Gradient, G.Gradient], [s, s]))
    s = G.FESpace(G.H1P2(1, 3), g3)
    run("C2 P2 tet Laplace L%d (generic)" % (Lq - 1), G.DiscreteSymmetricBilinearForm([G.Gradient, G.Gradient], [s, s]), path=1)
    run("C2 P2 tet Laplace L%d (fast)" % (Lq - 1), G.DiscreteSymmetricBilinearForm([G.Gradient, G.Gradient], [s, s]), path=2)
    run("P2 tet mass L%d" % (Lq - 1), G.DiscreteSymmetricBilinearForm([G.Identity, G.Identity], [s, s]))
    # C5: Hdiv mass on tets
    s = G.FESpace(G.HDIVRT0(3), g3)
    run("C5 RT0 tet mass L%d" % (Lq - 1), G.DiscreteSymmetricBilinearForm([G.Identity, G.Identity], [s, s]))
    s = G.FESpace(G.HDIVBDM1(3), g3)
    run("C5 BDM1 tet mass L%d" % (Lq - 1), G.DiscreteSymmetricBilinearForm([G.Identity, G.Identity], [s, s]))
    del g3, s
    # C3: Hooke P2 vector on triangles, C4: BR
    g2 = G.uniform_refine(G.grid_unitsquare("Triangle2D"), Lt)
    s = G.FESpace(G.H1P2(2, 2), g2)
    mu = 1000 / 1.4
    run("C3 Hooke H1P2{2,2} tri L%d" % Lt, G.DiscreteBilinearForm([G.SymmetricGradient(1), G.SymmetricGradient(1)], [s, s],
                                                                G.HookeAction(2, mu, 0.4 * mu / 0.2)))
    sv = G.FESpace(G.H1BR(2), g2)
    sp = G.FESpace(G.L2P0(1), g2)
    run("C4 BR tri Laplace L%d" % Lt, G.DiscreteSymmetricBilinearForm([G.Gradient, G.Gradient], [sv, sv]))
    run("C4 BR x P0 divergence L%d" % Lt, G.DiscreteBilinearForm([G.Divergence, G.Identity], [sv, sp]), factor=-1.0)
    R = G.ReconstructionIdentity(G.HDIVBDM1(2))
    run("C4 BR recon-BDM1 mass L%d" % Lt, G.DiscreteSymmetricBilinearForm([R, R], [sv, sv]))
    # linear forms
    L = G._lib.lib()
    for nm, op, bonus in (("C4 LF recon-BDM1 (tabulated f, 9-pt Stroud)", R, 2), ("LF identity BR", G.Identity, 0)):
        Op = G.LinearForm(op, G.DataFunction(lambda x: np.stack([3 * x[0] ** 2, 3 * x[1] ** 2]), [2, 2], bonus_quadorder=bonus))
        b = G.FEVector([sv])
        AP = G.assemble_operator(b[1], Op)
        t = time.time()
        for _ in range(3):
            G.assemble_operator(b[1], Op, Pattern=AP, skip_preps=True)
        st = G.blf_stats(AP)
        print(json.dumps({"config": nm, "ncells": int(g2.ncells), "ndofs": int(sv.ndofs), "device_ms": st.last_numeric_ms,
                          "host_call_ms": (time.time() - t) / 3 * 1e3}), flush=True)


if __name__ == "__main__":
    main()
