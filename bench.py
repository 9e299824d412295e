#!/usr/bin/env python
"""bench.py -- assembled nnz/s for the 3D P2 Laplace stiffness matrix (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--level L] [--impl reference]

A "step" is one numeric assembly of LaplaceOperator(1.0) for H1P2{1,3} on
uniform_refine(grid_unitcube(Tetrahedron3D), L) on a frozen sparsity pattern -- the analogue
of fill!(A,0) + assemble!(A, AP; skip_preps = true).  L = 6 (6 291 456 cells,
~2.4e8 non-zeros) is the BASELINE configuration (configs[1]).

  value   device-resident throughput: K steps bracketed by one pair of CUDA events on the library's
          launching stream (grmp_blf_numeric_steps), inputs resident in HBM; the working set
          (nzval 1.9 GB + maps) is far larger than the 126 MB L2, so no explicit flush is needed.
  e2e     the same metric through the reference-facing call with HOST buffers (grmp_blf_assemble_host): per
          step the grid arrays (Coordinates, CellVolumes, CellNodes, CellDofs) are copied from pinned host
          memory, the matrix is assembled and nzval is copied back to pinned host memory; the uploads the
          kernels do not read overlap the download.
  roofline  algorithmic bytes (8 nnz + 4 ncells (nn + nd) + 8 dim nnodes) / kernel time
          against the measured HBM copy bandwidth (MEASURED_PEAKS.json).
  cpu_baseline  the oracle's restatement of the reference loop (1 thread: the reference cell loop is
          serial) on a bounded sample (level 5 unless --cpu-level is given).

N > 1 (torchrun): strong scaling -- the cells of the SAME grid are partitioned into N contiguous
(spatially compact) ranges; a rank owns the dofs whose lowest-numbered cell it holds, assembles the
columns it owns from its cells plus the halo cells touching them (owner-computes, no numeric-phase
exchange), and `value` = global nnz * K / max-over-ranks device time.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402


def _peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def _traffic():
    """DRAM bytes per launch of the dominant kernel from the committed ncu capture, if any"""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(p):
        try:
            return json.load(open(p))
        except Exception:
            return None
    return None


class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i", str(index), "-lms", "100"],
                                      stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        rows = [r.strip().split(",") for r in open(self.f.name) if r.strip()]
        os.unlink(self.f.name)
        sm, reasons, mx = [], set(), None
        for r in rows:
            try:
                sm.append(float(r[0]))
                mx = float(r[1])
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                    if v.strip().lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                continue
        if sm:
            out["sm_mhz"] = float(np.median(sm))
            out["sm_max_mhz"] = mx
            out["reasons"] = sorted(reasons)
            out["samples"] = len(sm)
        return out


def bind_to_gpu_numa(local_rank):
    """N > 1: run this rank (and first-touch its pinned buffers) on the CPUs closest to its GPU (nvmlDeviceGetCpuAffinity);
    returns the CPU list or None.  GRMP_BENCH_NO_BIND=1 disables it."""
    if os.environ.get("GRMP_BENCH_NO_BIND"):
        return None
    try:
        import pynvml
        pynvml.nvmlInit()
        hdl = pynvml.nvmlDeviceGetHandleByIndex(local_rank)
        ncpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(hdl, (ncpu + 63) // 64)
        cpus = [64 * w + b for w, m in enumerate(words) for b in range(64) if (int(m) >> b) & 1]
        allowed = sorted(set(cpus) & set(os.sched_getaffinity(0)))
        if allowed and len(allowed) < len(os.sched_getaffinity(0)):
            os.sched_setaffinity(0, allowed)
            return allowed
    except Exception:
        pass
    return None


def build_problem(level):
    import grmp_b200 as G
    t = time.time()
    g = G.uniform_refine(G.grid_unitcube("Tetrahedron3D"), level)
    s = G.FESpace(G.H1P2(1, 3), g)
    s.celldofs
    g.cellvolumes
    return G, g, s, time.time() - t


def _shm_path(rank):
    tag = os.environ.get("MASTER_PORT", "0")
    base = "/dev/shm" if os.path.isdir("/dev/shm") else tempfile.gettempdir()
    return os.path.join(base, "grmp_bench_%s_rank%d.npz" % (tag, rank))


def local_problem(G, args, rank, world, dist):
    """rank-local grid / space.  Only rank 0 builds the global grid; it partitions it for every rank (cell ranges balanced by
    owner-computes work, partition.py) and hands the pieces over through /dev/shm."""
    if world == 1:
        _, g, s, t_grid = build_problem(args.level)
        return g, s, s.ndofs, {"ncells": int(g.ncells), "ndofs": int(s.ndofs)}, t_grid, None
    t = time.time()
    if rank == 0:
        _, g, s, _ = build_problem(args.level)
        glob = {"ncells": int(g.ncells), "ndofs": int(s.ndofs)}
        bounds = G.partition.halo_balanced_cell_ranges(s, world)
        for r in range(world - 1, -1, -1):
            lp = G.partition.partition(s, r, world, bounds=bounds)
            np.savez(_shm_path(r), coords=lp.grid.coords, cellnodes=lp.grid.cellnodes, vol=lp.grid.cellvolumes, celldofs=lp.space.celldofs,
                     n_owned=lp.n_owned, ndofs=lp.space.ndofs, l2g=lp.local2global, gcells=glob["ncells"], gdofs=glob["ndofs"])
        del g, s, lp
        import gc
        gc.collect()
    dist.barrier()
    d = np.load(_shm_path(rank))
    lg = G.ExtendableGrid(np.ascontiguousarray(d["coords"]), np.ascontiguousarray(d["cellnodes"]))
    lg._cache["vol"] = np.ascontiguousarray(d["vol"])
    ls = G.FESpace(G.H1P2(1, 3), lg)
    ls._celldofs = np.ascontiguousarray(d["celldofs"])
    ls.ndofs = int(d["ndofs"])
    glob = {"ncells": int(d["gcells"]), "ndofs": int(d["gdofs"])}
    l2g = np.ascontiguousarray(d["l2g"])
    n_owned = int(d["n_owned"])
    dist.barrier()
    try:
        os.unlink(_shm_path(rank))
    except OSError:
        pass
    return lg, ls, n_owned, glob, time.time() - t, l2g


def parity_check(G, L, C, lg, ls, n_owned, h_fast, nnz, colptr, world):
    """outside the timed region: the values of the timed kernel against the bit-exact generic path (reference operation and
    summation order, itself bit-equal to the oracle in tests/) on the SAME problem at the SAME size; the pattern must be
    identical.  On uniform_refine grids every entry is well conditioned (S/|ref| <= 14, tests/parity.py), so the bar is the
    pure one: relative 1e-12 above the explicit-zero tier, 1e-15 max|A| absolute inside it."""
    nz_fast = np.zeros(nnz)
    G._lib.check(L.grmp_blf_get_values(h_fast, G._lib.ptr(nz_fast)))
    APg = G.DiscreteSymmetricBilinearForm([G.Gradient, G.Gradient], [ls, ls])
    G.prepare_assembly(APg)
    hg = APg.AM.h
    G._lib.check(L.grmp_blf_set_path(hg, G._lib.PATH_GENERIC))
    nnzg = C.c_int64(0)
    G._lib.check(L.grmp_blf_symbolic(hg, 1.0, C.byref(nnzg)))
    cpg = np.zeros(ls.ndofs + 1, np.int64)
    rvg = np.zeros(nnzg.value, np.int64)
    G._lib.check(L.grmp_blf_get_pattern(hg, G._lib.ptr(cpg), G._lib.ptr(rvg)))
    same_pattern = bool(nnzg.value == nnz and np.array_equal(cpg, colptr))
    nz_gen = np.zeros(nnzg.value)
    G._lib.check(L.grmp_blf_numeric(hg, 1.0, G._lib.ptr(nz_gen)))
    del APg
    end = int(colptr[n_owned] - 1)          # owned columns only (halo columns belong to another rank)
    a, r = nz_fast[:end], nz_gen[:end]
    amax = float(np.abs(r).max())
    err = np.abs(a - r)
    zero = np.abs(r) <= 1e-13 * amax
    rel = float((err[~zero] / np.abs(r[~zero])).max())
    zabs = float(err[zero].max() / amax) if zero.any() else 0.0
    ok = same_pattern and rel <= 1e-12 and zabs <= 1e-15
    return {"parity_checked": bool(ok), "against": "generic bit-exact path, same grid and size, owned columns", "same_pattern": same_pattern,
            "max_rel": rel, "explicit_zero_entries": int(zero.sum()), "explicit_zero_max_abs_over_amax": zabs, "entries": int(end)}, nz_fast


def run_gpu(args):
    import ctypes as C
    import torch
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        args.gpus = world
    torch.cuda.set_device(local_rank)
    dist = None
    bound = bind_to_gpu_numa(local_rank) if world > 1 else None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    import grmp_b200 as G
    L = G._lib.lib()
    lg, ls, n_owned, glob, t_grid, l2g = local_problem(G, args, rank, world, dist)
    AP = G.DiscreteSymmetricBilinearForm([G.Gradient, G.Gradient], [ls, ls])
    G.prepare_assembly(AP)
    h = AP.AM.h
    if args.path != "auto":
        G._lib.check(L.grmp_blf_set_path(h, {"generic": 1, "fast": 2, "columns": 3, "atomic": 4, "coloured": 5}[args.path]))
    if world > 1:
        G._lib.check(L.grmp_blf_set_owned_columns(h, n_owned))
    elif args.owned_cols >= 0:      # experiment: time only the first columns (e.g. the vertex columns)
        n_owned = args.owned_cols
        G._lib.check(L.grmp_blf_set_owned_columns(h, n_owned))
    nnz = C.c_int64(0)
    t0 = time.time()
    G._lib.check(L.grmp_blf_symbolic(h, 1.0, C.byref(nnz)))
    t_sym = time.time() - t0
    colptr = np.zeros(ls.ndofs + 1, np.int64)
    rowval = np.zeros(nnz.value, np.int64)
    G._lib.check(L.grmp_blf_get_pattern(h, G._lib.ptr(colptr), G._lib.ptr(rowval)))
    nnz_owned = int(colptr[n_owned] - 1)
    st = G.blf_stats(AP)
    # clocks are sampled from here (warm-up) until the end of the end-to-end loop: the device-resident timed region
    # alone lasts only ~20 ms, shorter than one nvidia-smi sampling period
    sampler = ClockSampler(local_rank) if rank == 0 else None
    # warm-up
    ms = C.c_double(0)
    G._lib.check(L.grmp_blf_numeric_steps(h, 1.0, max(args.warmup, 3), C.byref(ms)))
    # ---- timed region: K steps, barrier + synchronize on both sides, CUDA events on the launching stream ----
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    w0 = time.time()
    G._lib.check(L.grmp_blf_numeric_steps(h, 1.0, args.steps, C.byref(ms)))
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    wall = time.time() - w0
    dev_ms = ms.value
    launches = int(G.blf_stats(AP).kernel_launches) * args.steps
    # ---- e2e (a): assemble!(A, AP) behind the ABI with HOST buffers: geometry in, the matrix values out, every step ----
    pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory()  # noqa: E731
    h_coords, h_vol = pin(lg.coords), pin(lg.cellvolumes)
    h_nz = torch.empty(nnz.value, dtype=torch.float64).pin_memory()
    e2e_steps = max(2, min(args.steps, 5))

    def e2e_step():
        # one synchronous call, all copies inside (grmp.h: grmp_blf_assemble_host); the topology is frozen with the pattern
        G._lib.check(L.grmp_blf_assemble_host(h, 1.0, h_coords.data_ptr(), h_vol.data_ptr(), None, None, None, h_nz.data_ptr()))

    def timed(fn, n):
        fn()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t = time.time()
        for _ in range(n):
            fn()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        return (time.time() - t) / n
    e2e_s = timed(e2e_step, e2e_steps)
    h2d = sum(t.numel() * t.element_size() for t in (h_coords, h_vol))
    d2h = h_nz.numel() * 8
    # ---- e2e (b): the matrix stays on the device (hand-off to a device solver); per step geometry + a vector go in, the
    #      residual A x - b comes back (solve_direct!'s check, solvers.jl:661-668) ----
    h_x = pin(np.ones(ls.ndofs))
    h_r = torch.empty(ls.ndofs, dtype=torch.float64).pin_memory()
    nrm = C.c_double(0)

    def resident_step():
        G._lib.check(L.grmp_blf_assemble_host(h, 1.0, h_coords.data_ptr(), h_vol.data_ptr(), None, None, None, None))
        G._lib.check(L.grmp_blf_residual(h, h_x.data_ptr(), None, None, 0, h_r.data_ptr(), C.byref(nrm)))
    res_s = timed(resident_step, e2e_steps)
    rowsum_local = h_r.numpy().copy()          # A_local * 1 (local numbering): stiffness rows sum to zero after the merge
    # keep the GPU busy with numeric steps for ~1.5 s more so that the clock record has samples under compute load
    if sampler is not None:
        t_end = time.time() + 1.5
        while time.time() < t_end:
            G._lib.check(L.grmp_blf_numeric_steps(h, 1.0, 50, C.byref(C.c_double(0))))
    clocks = sampler.stop() if sampler else None
    # ---- parity of the timed configuration itself (outside every timed region) ----
    par, nz_fast = ({"parity_checked": False, "skipped": "--no-parity"}, None) if args.no_parity else \
        parity_check(G, L, C, lg, ls, n_owned, h, nnz.value, colptr, world)
    if nz_fast is None:
        nz_fast = h_nz.numpy()
    checksum = float(nz_fast[:nnz_owned].sum())
    checksum_abs = float(np.abs(nz_fast[:nnz_owned]).sum())
    amax = float(np.abs(nz_fast[:nnz_owned]).max())
    # ---- reduce over ranks ----
    per_rank_ms = [dev_ms / args.steps]
    rowsum_inf = None
    if world > 1:
        t = torch.tensor([dev_ms, e2e_s, wall, res_s, amax, par.get("max_rel", 0.0), par.get("explicit_zero_max_abs_over_amax", 0.0)],
                         dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        mine = torch.tensor([dev_ms / args.steps], dtype=torch.float64, device="cuda")
        allms = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(allms, mine)
        per_rank_ms = [float(x.item()) for x in allms]
        dev_ms, e2e_s, wall, res_s, amax, mrel, mzero = [float(x) for x in t.cpu()]
        c = torch.tensor([nnz_owned, h2d, d2h, launches, lg.ncells, checksum, checksum_abs, 1.0 if par.get("parity_checked") else 0.0,
                          ls.ndofs * 16], dtype=torch.float64, device="cuda")
        dist.all_reduce(c, op=dist.ReduceOp.SUM)
        cc = [float(x) for x in c.cpu()]
        nnz_total, h2d, d2h, launches, cells_total = [int(x) for x in cc[:5]]
        checksum, checksum_abs = cc[5], cc[6]
        res_bytes = int(cc[8])
        par = dict(par, parity_checked=bool(cc[7] == world), max_rel=mrel, explicit_zero_max_abs_over_amax=mzero, ranks_checked=world)
        # merged matrix: y = A * 1 assembled from the owned column blocks of all ranks (rows in global numbering)
        y = torch.zeros(glob["ndofs"], dtype=torch.float64, device="cuda")
        y.index_add_(0, torch.from_numpy(l2g).cuda(), torch.from_numpy(rowsum_local).cuda())
        dist.all_reduce(y, op=dist.ReduceOp.SUM)
        rowsum_inf = float(y.abs().max().item())
    else:
        nnz_total, cells_total = nnz_owned, lg.ncells
        res_bytes = ls.ndofs * 16
        rowsum_inf = float(np.abs(rowsum_local).max())
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    par["rowsum_inf_over_amax"] = rowsum_inf / amax       # A * 1 = 0 for a stiffness matrix: checks the merged owned blocks
    par["checksum_abs"] = checksum_abs                    # sum |nzval| over all owned entries: the same number at every N
    ms_step = dev_ms / args.steps
    value = nnz_total / (ms_step * 1e-3)
    peak, peak_src = _peaks()
    # algorithmic bytes of ONE launch on rank 0 (per-launch, like the kernel time it is divided by)
    b_alg = 8 * nnz_owned + lg.ncells * 4 * (4 + 10) + 8 * 3 * lg.nnodes
    achieved = b_alg / (per_rank_ms[0] * 1e-3) / 1e9
    tr = _traffic()
    roofline = {"bound": "hbm", "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s", "frac": round(achieved / peak, 4),
                "traffic": (tr or {}).get("dram_bytes_per_launch") if world == 1 else None, "peak_source": peak_src,
                "kernel": "p2tet_edge_kernel (+ p2tet_vertex_diag_kernel, ~10 % of the step; duration = whole step)" if st.path == 2
                else "col_kernel" if st.path == 3 else "blf_local_kernel+gather_kernel",
                "algorithmic_bytes_per_launch": int(b_alg), "frac_of_nominal_8TBs": round(achieved / 8000.0, 4),
                "note": "rank 0's launch over rank 0's own step time" + ("" if world == 1 else "; traffic (ncu) is captured at N = 1 only")}
    cpu = None
    if world == 1:      # the CPU arm is timed on rank 0 at N = 1 only
        cpu = cpu_baseline(args.cpu_level)
        cpu["all_cores"] = cpu_parallel_baseline(args.cpu_level)
    srt = sorted(per_rank_ms)
    out = {
        "metric": "assembled nnz/s, 3D P2 Laplace stiffness (numeric assembly on a frozen pattern)",
        "value": value, "unit": "nnz/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic (uniform_refine(grid_unitcube(Tetrahedron3D), %d), H1P2{1,3}, LaplaceOperator(1.0))" % args.level,
        "config": {"workload": "Example301 Poisson 3D: H1P2 Laplace stiffness on uniform_refine(grid_unitcube(Tetrahedron3D),%d)" % args.level,
                   "level": args.level, "ncells": glob["ncells"], "ndofs": glob["ndofs"], "nnz": int(nnz_total),
                   "cells_assembled_all_ranks": int(cells_total),
                   "partition": "contiguous cell ranges balanced by walked cells (own + halo), owned columns per rank, no numeric-phase exchange" if world > 1 else "none",
                   "l2": "inputs+outputs (%.2f GB) >> 126 MB L2, no explicit flush" % ((b_alg + 16 * 10 * lg.ncells) / 1e9),
                   "path": G._lib.PATH_NAMES[int(st.path)], "tiles": int(st.ntiles)},
        "e2e": {"value": nnz_total / e2e_s, "unit": "nnz/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                "ms_per_step": e2e_s * 1e3, "steps": e2e_steps,
                "what": "grmp_blf_assemble_host: Coordinates + CellVolumes up (pinned), numeric assembly, nzval down (pinned); PCIe bound"},
        "e2e_resident": {"value": nnz_total / res_s, "unit": "nnz/s", "ms_per_step": res_s * 1e3, "steps": e2e_steps,
                         "h2d_bytes_per_step": int(h2d + res_bytes // 2), "d2h_bytes_per_step": int(res_bytes // 2),
                         "what": "matrix stays on the device (grmp_blf_device_csc hand-off): geometry + x up, assembly, r = A x down "
                                 "(grmp_blf_residual, solvers.jl:661-668)"},
        "gpu_launches": launches,
        "roofline": roofline,
        "host_affinity": ({"rank0_cpus": bound} if bound else None),
        "per_rank_ms": {"min": srt[0], "median": srt[len(srt) // 2], "max": srt[-1], "all": per_rank_ms},
        "parity": par,
        "cpu_baseline": cpu,
        "clocks": clocks,
        "timing": {"wall_s_timed_region": wall, "grid_build_s": t_grid, "symbolic_s": t_sym, "symbolic_device_ms": st.last_symbolic_ms,
                   "first_assembly_nnz_per_s": nnz_owned / max(t_sym, 1e-9)},
        "checksum": checksum,
    }
    print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


def _oracle_problem(level):
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle as O
    G, g, s, _ = build_problem(level)
    A = O.OracleMatrix(s.ndofs, s.ndofs)
    t = time.time()
    O.blf_assemble(A, g, s, s, O.OP_GRAD, O.OP_GRAD, apt=O.APT_SYMMETRIC, factor=1.0)
    A.flush()
    t_first = time.time() - t
    nnz = A.csc()[1].size
    return O, g, s, A, nnz, t_first


def cpu_baseline(level):
    """oracle (port of the reference loop) on a bounded sample: first assembly (pattern + values, LNK
    insertion + flush!) and reassembly on the frozen pattern (CSC binary-search updates)"""
    try:
        O, g, s, A, nnz, t_first = _oracle_problem(level)
        A.fill_zero()
        t = time.time()
        O.blf_assemble(A, g, s, s, O.OP_GRAD, O.OP_GRAD, apt=O.APT_SYMMETRIC, factor=1.0)
        t_re = time.time() - t
        return {"value": nnz / t_re, "unit": "nnz/s", "cores": 1, "kind": "port",
                "sample": "same workload at level %d (%d cells, %d nnz): reassembly on the frozen pattern %.2f s; first assembly "
                          "(pattern+values) %.2f s = %.3g nnz/s" % (level, g.ncells, nnz, t_re, t_first, nnz / t_first),
                "host_cpus": os.cpu_count()}
    except Exception as e:  # the baseline must never take the bench line down
        return {"value": None, "unit": "nnz/s", "cores": 1, "kind": "port", "sample": "failed: %r" % (e,)}


def cpu_parallel_baseline(level, nthreads=None):
    """'best plausible CPU' line: the same oracle loop on all host cores.  The reference's cell loop is serial, so
    this is NOT the reference arm: the cells are split into nthreads contiguous ranges exactly like the multi-GPU partition
    (owner-computes with halo cells, partition.py), every thread reassembles its rank-local matrix on a frozen pattern (the
    ctypes call releases the GIL), value = global nnz / slowest thread."""
    import threading
    try:
        sys.path.insert(0, os.path.join(ROOT, "oracle"))
        import oracle as O
        T = int(nthreads or min(os.cpu_count() or 1, 16))
        G, g, s, _ = build_problem(level)
        lps = [G.partition.partition(s, r, T) for r in range(T)]
        mats, nnz_owned = [], 0
        for lp in lps:                                   # first assembly (pattern), untimed
            A = O.OracleMatrix(lp.space.ndofs, lp.space.ndofs)
            O.blf_assemble(A, lp.grid, lp.space, lp.space, O.OP_GRAD, O.OP_GRAD, apt=O.APT_SYMMETRIC, factor=1.0)
            A.flush()
            cp = A.csc()[0]
            nnz_owned += int(cp[lp.n_owned] - 1)
            A.fill_zero()
            mats.append(A)
        start = threading.Barrier(T + 1)
        ends = [0.0] * T

        def work(r):
            lp = lps[r]
            start.wait()
            O.blf_assemble(mats[r], lp.grid, lp.space, lp.space, O.OP_GRAD, O.OP_GRAD, apt=O.APT_SYMMETRIC, factor=1.0)
            ends[r] = time.time()
        th = [threading.Thread(target=work, args=(r,)) for r in range(T)]
        for t in th:
            t.start()
        start.wait()
        t0 = time.time()
        for t in th:
            t.join()
        dt = max(ends) - t0
        return {"value": nnz_owned / dt, "unit": "nnz/s", "cores": T, "kind": "port, cell ranges on threads (owner-computes with halo)",
                "sample": "level %d: %d threads, slowest %.2f s, cells assembled by all threads %d of %d" %
                          (level, T, dt, sum(lp.grid.ncells for lp in lps), g.ncells)}
    except Exception as e:
        return {"value": None, "sample": "failed: %r" % (e,)}


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path.  The reference is Julia (not
    installable here: no julia binary, no network), so this arm times the oracle's op-for-op port of
    its serial cell loop on the host, 1 thread (the reference loop is serial)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    level = args.cpu_level
    if (args.steps + args.warmup) > 30:
        level = min(level, 4)
    O, g, s, A, nnz, t_first = _oracle_problem(level)
    for _ in range(args.warmup):
        A.fill_zero()
        O.blf_assemble(A, g, s, s, O.OP_GRAD, O.OP_GRAD, apt=O.APT_SYMMETRIC, factor=1.0)
    t = time.time()
    for _ in range(args.steps):
        A.fill_zero()
        O.blf_assemble(A, g, s, s, O.OP_GRAD, O.OP_GRAD, apt=O.APT_SYMMETRIC, factor=1.0)
    dt = (time.time() - t) / args.steps
    value = nnz / dt
    sample = "level %d (%d cells, %d nnz) per step; first assembly %.2f s" % (level, g.ncells, nnz, t_first)
    print(json.dumps({
        "impl": "reference", "metric": "assembled nnz/s, 3D P2 Laplace stiffness (numeric assembly on a frozen pattern)",
        "value": value, "unit": "nnz/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        # the arm's configuration is the GPU arm's; every step assembles a bounded sample of it (one refinement level less =
        # 1/8 of the cells, same element, operator and code path) so that K + W steps end within minutes.  nnz/s is a rate;
        # the per-entry cost of the reference's CSC binary-search update grows with the matrix, so the sample flatters the CPU.
        "config": {"workload": "Example301 Poisson 3D: H1P2 Laplace stiffness on uniform_refine(grid_unitcube(Tetrahedron3D),%d)" % args.level,
                   "level": args.level, "sample_level": level, "sample_ncells": int(g.ncells), "sample_nnz": int(nnz)},
        "cpu_baseline": {"value": value, "unit": "nnz/s", "cores": 1, "kind": "port", "sample": sample,
                         "note": "Julia reference not runnable in this image; oracle port of bilinearform.jl:226-377, serial like the reference "
                                 "(its cell loop uses one thread whatever JULIA_NUM_THREADS is); all_cores = the same loop on cell-range partitions",
                         "all_cores": cpu_parallel_baseline(level)},
        "e2e": {"value": value, "unit": "nnz/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--level", type=int, default=6)
    ap.add_argument("--cpu-level", type=int, default=5)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--owned-cols", type=int, default=-1)
    ap.add_argument("--path", default="auto", choices=["auto", "generic", "fast", "columns", "atomic", "coloured"])
    ap.add_argument("--no-parity", action="store_true", help="skip the post-run comparison with the generic path")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
