#!/usr/bin/env python
"""Per-rank device times of an N-GPU run measured on ONE GPU: the numeric phase has no exchange (owner-computes columns), so
the time of rank r is the time of its local problem, and the N-GPU step time is the maximum.  Used to tune the partition
(balance, tile shapes) without holding N GPUs.

  python tools/emulate_ranks.py --world 8 [--level 6] [--balance work|cells]
"""
import argparse
import ctypes as C
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import grmp_b200 as G  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--world", type=int, default=8)
    ap.add_argument("--level", type=int, default=6)
    ap.add_argument("--balance", default="halo")
    ap.add_argument("--steps", type=int, default=50)
    a = ap.parse_args()
    L = G._lib.lib()
    g = G.uniform_refine(G.grid_unitcube("Tetrahedron3D"), a.level)
    s = G.FESpace(G.H1P2(1, 3), g)
    out = []
    bounds = G.partition._BALANCERS[a.balance](s, a.world)
    for r in range(a.world):
        lp = G.partition.partition(s, r, a.world, bounds=bounds)
        AP = G.DiscreteSymmetricBilinearForm([G.Gradient, G.Gradient], [lp.space, lp.space])
        G.prepare_assembly(AP)
        h = AP.AM.h
        G._lib.check(L.grmp_blf_set_owned_columns(h, lp.n_owned))
        nnz = C.c_int64(0)
        G._lib.check(L.grmp_blf_symbolic(h, 1.0, C.byref(nnz)))
        cp = np.zeros(lp.space.ndofs + 1, np.int64)
        rv = np.zeros(nnz.value, np.int64)
        G._lib.check(L.grmp_blf_get_pattern(h, G._lib.ptr(cp), G._lib.ptr(rv)))
        ms = C.c_double(0)
        G._lib.check(L.grmp_blf_numeric_steps(h, 1.0, 5, C.byref(ms)))
        G._lib.check(L.grmp_blf_numeric_steps(h, 1.0, a.steps, C.byref(ms)))
        st = G.blf_stats(AP)
        out.append({"rank": r, "ms": ms.value / a.steps, "cells": int(lp.grid.ncells), "nnz_owned": int(cp[lp.n_owned] - 1), "tiles": int(st.ntiles)})
        print(json.dumps(out[-1]), flush=True)
        del AP, lp
    tmax = max(o["ms"] for o in out)
    nnz_total = sum(o["nnz_owned"] for o in out)
    print(json.dumps({"world": a.world, "balance": a.balance, "max_ms": tmax, "min_ms": min(o["ms"] for o in out),
                      "nnz_per_s": nnz_total / (tmax * 1e-3), "cells_all_ranks": sum(o["cells"] for o in out)}))


if __name__ == "__main__":
    main()
