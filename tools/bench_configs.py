#!/usr/bin/env python
"""Throughput of every BASELINE.json configuration (C1-C5) on one GPU, per numeric back end: owner-computes column kernels
(GRMP_PATH_COLUMNS), the ring-walk kernel of the metric form (P2TET), the cell-parallel scatter alternatives (ATOMIC /
COLOURED) and the bit-exact generic path -- the table of profiles/r2_all_configs.md and the scatter-variant measurement of
profiles/r2_scatter_variants.md.  Not the headline bench (that is bench.py).

  python tools/bench_configs.py [--small] [--paths columns,atomic,coloured,generic] [--only substring]
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

PATHS = {"generic": 1, "p2tet": 2, "columns": 3, "atomic": 4, "coloured": 5}
PEAK = 6541.8
try:
    PEAK = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass


def algorithmic_bytes(AP, nnz):
    """8 nnz + 4 ncells (nn + nd_row + nd_col*) + 8 dim nnodes + E_geom"""
    s1, s2 = AP.FES
    g = s1.xgrid
    nn = g.dim + 1
    b = 8 * nnz + g.ncells * 4 * (nn + s1.nd_cell + (s2.nd_cell if s2 is not s1 else 0)) + 8 * g.dim * g.nnodes
    for s in {id(s1): s1, id(s2): s2}.values():
        code = s.fetype.code
        if code == 4:                      # RT0: signs
            b += 4 * nn * g.ncells
        elif code == 5:                    # BDM1: signs (+ orientations in 3D)
            b += (8 if g.dim == 3 else 4) * nn * g.ncells
        elif code == 3:                    # BR: CellFaces + normals/volumes per face
            b += 4 * nn * g.ncells + 8 * (g.dim + 1) * g.nfaces
    return b


def run(name, make_AP, factor=1.0, steps=10, paths=("columns",)):
    L = G._lib.lib()
    out = []
    for pname in paths:
        AP = make_AP()
        G.prepare_assembly(AP)
        h = AP.AM.h
        G._lib.check(L.grmp_blf_set_path(h, PATHS[pname]))
        nnz = C.c_int64(0)
        t = time.time()
        rc = L.grmp_blf_symbolic(h, factor, C.byref(nnz))
        if rc != 0:
            print(json.dumps({"config": name, "path": pname, "skipped": L.grmp_last_error().decode()}), flush=True)
            continue
        tsym = time.time() - t
        ms = C.c_double(0)
        G._lib.check(L.grmp_blf_numeric_steps(h, factor, 3, C.byref(ms)))
        G._lib.check(L.grmp_blf_numeric_steps(h, factor, steps, C.byref(ms)))
        st = G.blf_stats(AP)
        per = ms.value / steps
        balg = algorithmic_bytes(AP, nnz.value)
        g = AP.FES[0].xgrid
        rec = {"config": name, "ncells": int(g.ncells), "nnz": int(nnz.value), "path": G._lib.PATH_NAMES[int(st.path)],
               "launches_per_assembly": int(st.kernel_launches), "ms_per_assembly": round(per, 4), "nnz_per_s": nnz.value / (per * 1e-3),
               "algorithmic_GB": round(balg / 1e9, 4), "algorithmic_GBs": round(balg / (per * 1e-3) / 1e9, 1),
               "frac_of_peak": round(balg / (per * 1e-3) / 1e9 / PEAK, 4), "symbolic_s": round(tsym, 3)}
        print(json.dumps(rec), flush=True)
        out.append(rec)
        del AP
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--small", action="store_true")
    ap.add_argument("--paths", default="columns")
    ap.add_argument("--only", default="")
    ap.add_argument("--lf", action="store_true")
    ap.add_argument("--l2d", type=int, default=10, help="refinement level of the unit square (10: 4.2 M triangles)")
    ap.add_argument("--l3d", type=int, default=5, help="refinement level of the unit cube (5: 0.8 M tetrahedra; P1 / RT0 run one level finer)")
    a = ap.parse_args()
    paths = a.paths.split(",")
    # sizes: SURVEY.md 8(d) -- triangles level 9-10 (1-4.2 M cells), tetrahedra level 5-6 (0.8-6.3 M cells).  The low-order forms run on
    # the larger grid (their matrices are small: level-5 P1 is 2 M non-zeros = 5 us at the roofline, i.e. pure launch latency).
    Lt, Lq = (6, 4) if a.small else (a.l2d, a.l3d)
    Lq_big = Lq if a.small else Lq + 1
    sym, gen = G.DiscreteSymmetricBilinearForm, G.DiscreteBilinearForm
    mu = 1000 / 1.4
    lam = 0.4 * mu / 0.2
    cfgs = []
    g3 = lambda: G.uniform_refine(G.grid_unitcube("Tetrahedron3D"), Lq)       # noqa: E731
    g3b = lambda: G.uniform_refine(G.grid_unitcube("Tetrahedron3D"), Lq_big)  # noqa: E731
    g2 = lambda: G.uniform_refine(G.grid_unitsquare("Triangle2D"), Lt)        # noqa: E731
    cfgs.append(("C1 P2 tri Laplace L%d" % Lt, g2, lambda g: (lambda s: sym([G.Gradient, G.Gradient], [s, s]))(G.FESpace(G.H1P2(1, 2), g)), 1.0))
    cfgs.append(("C3 Hooke H1P2{2,2} tri L%d" % Lt, g2,
                 lambda g: (lambda s: gen([G.SymmetricGradient(1), G.SymmetricGradient(1)], [s, s], G.HookeAction(2, mu, lam)))(G.FESpace(G.H1P2(2, 2), g)), 1.0))
    cfgs.append(("C4 BR tri Laplace L%d" % Lt, g2, lambda g: (lambda s: sym([G.Gradient, G.Gradient], [s, s]))(G.FESpace(G.H1BR(2), g)), 1.0))
    cfgs.append(("C4 BR x P0 divergence L%d" % Lt, g2,
                 lambda g: gen([G.Divergence, G.Identity], [G.FESpace(G.H1BR(2), g), G.FESpace(G.L2P0(1), g)]), -1.0))
    R = G.ReconstructionIdentity(G.HDIVBDM1(2))
    cfgs.append(("C4 BR recon-BDM1 mass L%d" % Lt, g2, lambda g: (lambda s: sym([R, R], [s, s]))(G.FESpace(G.H1BR(2), g)), 1.0))
    cfgs.append(("C2 P2 tet Laplace L%d" % Lq, g3, lambda g: (lambda s: sym([G.Gradient, G.Gradient], [s, s]))(G.FESpace(G.H1P2(1, 3), g)), 1.0))
    cfgs.append(("P2 tet mass L%d" % Lq, g3, lambda g: (lambda s: sym([G.Identity, G.Identity], [s, s]))(G.FESpace(G.H1P2(1, 3), g)), 1.0))
    cfgs.append(("C5 BDM1 tet mass L%d" % Lq, g3, lambda g: (lambda s: sym([G.Identity, G.Identity], [s, s]))(G.FESpace(G.HDIVBDM1(3), g)), 1.0))
    cfgs.append(("C2 P1 tet Laplace L%d" % Lq_big, g3b, lambda g: (lambda s: sym([G.Gradient, G.Gradient], [s, s]))(G.FESpace(G.H1P1(1), g)), 1.0))
    cfgs.append(("C5 RT0 tet mass L%d" % Lq_big, g3b, lambda g: (lambda s: sym([G.Identity, G.Identity], [s, s]))(G.FESpace(G.HDIVRT0(3), g)), 1.0))
    cache = {}
    for name, gf, mk, factor in cfgs:
        if a.only and a.only not in name:
            continue
        if gf not in cache:
            cache.clear()
            cache[gf] = gf()
        g = cache[gf]
        ps = list(paths)
        if "P2 tet Laplace" in name and "columns" in ps and "p2tet" not in ps:
            ps = ["p2tet"] + ps
        run(name, lambda: mk(g), factor=factor, paths=ps)
    if not a.lf:
        return
    cache.clear()
    g = g2()
    sv = G.FESpace(G.H1BR(2), g)
    for nm, op, bonus in (("C4 LF recon-BDM1 (tabulated f, 9-pt Stroud)", R, 2), ("LF identity BR", G.Identity, 0)):
        Op = G.LinearForm(op, G.DataFunction(lambda x: np.stack([3 * x[0] ** 2, 3 * x[1] ** 2]), [2, 2], bonus_quadorder=bonus))
        b = G.FEVector([sv])
        AP = G.assemble_operator(b[1], Op)
        t = time.time()
        for _ in range(3):
            G.assemble_operator(b[1], Op, Pattern=AP, skip_preps=True)
        st = G.blf_stats(AP)
        print(json.dumps({"config": nm, "ncells": int(g.ncells), "ndofs": int(sv.ndofs), "device_ms": st.last_numeric_ms,
                          "host_call_ms": (time.time() - t) / 3 * 1e3}), flush=True)


if __name__ == "__main__":
    main()
