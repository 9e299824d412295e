"""Timing sweep over the tunables of the metric kernel (environment variables read at grmp_blf_symbolic): the grid is built
once, every configuration gets its own pattern handle.  Values of every configuration are compared with the first one.

    python tools/fast_sweep.py --level 6 --steps 20 "" "GRMP_FAST_LAG=2000" "GRMP_FAST_LAG=1000000,GRMP_DEBUG_FLAGS=2"

GRMP_DEBUG_FLAGS is read once per process (static): configurations that set it must come in their own run."""
import argparse
import ctypes as C
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import grmp_b200 as G  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--level", type=int, default=6)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--repeat", type=int, default=3)
    ap.add_argument("configs", nargs="*", default=[""])
    a = ap.parse_args()
    L = G._lib.lib()
    t = time.time()
    g = G.uniform_refine(G.grid_unitcube("Tetrahedron3D"), a.level)
    s = G.FESpace(G.H1P2(1, 3), g)
    s.celldofs
    g.cellvolumes
    print(f"# grid level {a.level}: {g.ncells} cells, {time.time() - t:.1f} s", flush=True)
    ref = None
    for cfg in a.configs:
        env = dict(kv.split("=") for kv in cfg.split(",") if kv)
        for k, v in env.items():
            os.environ[k] = v
        AP = G.DiscreteSymmetricBilinearForm([G.Gradient, G.Gradient], [s, s])
        G.prepare_assembly(AP)
        h = AP.AM.h
        nnz = C.c_int64(0)
        t = time.time()
        G._lib.check(L.grmp_blf_symbolic(h, 1.0, C.byref(nnz)))
        t_sym = time.time() - t
        st = G.blf_stats(AP)
        ms = C.c_double(0)
        G._lib.check(L.grmp_blf_numeric_steps(h, 1.0, 5, C.byref(ms)))
        best = []
        for _ in range(a.repeat):
            G._lib.check(L.grmp_blf_numeric_steps(h, 1.0, a.steps, C.byref(ms)))
            best.append(ms.value / a.steps)
        nz = np.zeros(nnz.value)
        G._lib.check(L.grmp_blf_get_values(h, G._lib.ptr(nz)))
        if ref is None:
            ref = nz
            dev = 0.0
        else:
            dev = float(np.abs(nz - ref).max() / np.abs(ref).max())
        print(json.dumps({"config": cfg, "path": int(st.path), "ntiles": int(st.ntiles), "ms": [round(b, 4) for b in best], "symbolic_s": round(t_sym, 2),
                          "max_abs_dev_vs_first_over_amax": dev}), flush=True)
        del AP
        for k in env:
            del os.environ[k]


if __name__ == "__main__":
    main()
