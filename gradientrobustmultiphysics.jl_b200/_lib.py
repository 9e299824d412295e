"""ctypes binding of libgrmp_cuda.so (include/grmp.h) -- the in-container stand-in for the
Julia `ccall` glue (julia/GRMPCuda.jl).  There is no CPU fallback: a missing library or a
missing CUDA device raises."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libgrmp_cuda.so")
_lib = None

OK = 0
PATH_AUTO, PATH_GENERIC, PATH_FAST, PATH_P2TET, PATH_COLUMNS, PATH_ATOMIC, PATH_COLOURED = 0, 1, 2, 2, 3, 4, 5
PATH_NAMES = {1: "generic", 2: "p2tet", 3: "columns", 4: "atomic", 5: "coloured"}

EXPORTS = [
    "grmp_last_error", "grmp_init", "grmp_finalize", "grmp_device_synchronize", "grmp_grid_create", "grmp_grid_create_bfaces", "grmp_grid_set_faces",
    "grmp_grid_update_geometry", "grmp_grid_update_cells", "grmp_grid_destroy", "grmp_space_update_dofs",
    "grmp_blf_numeric_steps", "grmp_blf_set_owned_columns", "grmp_space_create", "grmp_space_destroy", "grmp_blf_create",
    "grmp_blf_destroy", "grmp_blf_set_path", "grmp_blf_symbolic", "grmp_blf_get_pattern", "grmp_blf_numeric",
    "grmp_blf_get_values", "grmp_blf_transpose_copy", "grmp_blf_stats", "grmp_blf_device_values", "grmp_lf_create",
    "grmp_lf_destroy", "grmp_lf_assemble", "grmp_lf_stats", "grmp_blf_set_fixed_argument", "grmp_blf_set_newton_argument", "grmp_blf_newton_rhs", "grmp_lf_assemble_feb", "grmp_ii_create", "grmp_ii_destroy", "grmp_ii_resultdim", "grmp_ii_evaluate", "grmp_blf_assemble_host", "grmp_blf_device_csc", "grmp_blf_matmul",
    "grmp_blf_matmul_device", "grmp_blf_residual", "grmp_blf_apply_penalties", "grmp_lf_set_path",
]


class GrmpError(RuntimeError):
    pass


class EvalTab(C.Structure):
    _fields_ = [("nd_all", C.c_int32), ("ncomp", C.c_int32), ("refvals", C.c_void_p), ("refderivs", C.c_void_p)]


class Stats(C.Structure):
    _fields_ = [("last_numeric_ms", C.c_double), ("last_symbolic_ms", C.c_double), ("nnz", C.c_int64), ("ncontrib", C.c_int64),
                ("kernel_launches", C.c_int64), ("path", C.c_int32), ("ntiles", C.c_int32)]


class DeviceCSC(C.Structure):
    _fields_ = [("nrows", C.c_int64), ("ncols", C.c_int64), ("nnz", C.c_int64), ("colptr", C.c_void_p), ("rowval", C.c_void_p),
                ("nzval", C.c_void_p), ("device", C.c_int32), ("reserved", C.c_int32)]


def build(force: bool = False) -> str:
    """compile csrc/ for sm_100a in-tree (nvcc cross-compiles without a GPU)"""
    csrc = os.path.join(_HERE, "csrc")
    srcs = [os.path.join(csrc, f) for f in os.listdir(csrc) if f.endswith((".cu", ".cuh"))] + [
        os.path.join(os.path.dirname(_HERE), "include", "grmp.h")]
    stale = (not os.path.exists(LIB_PATH)) or any(os.path.getmtime(s) > os.path.getmtime(LIB_PATH) for s in srcs)
    if force or stale:
        subprocess.check_call(["make", "-C", csrc, "-j4", "-s"])
    return LIB_PATH


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise GrmpError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                            "(the assembly path has no CPU fallback)")
        L = C.CDLL(LIB_PATH)
        L.grmp_last_error.restype = C.c_char_p
        vp, i32, i64, dbl = C.c_void_p, C.c_int, C.c_int64, C.c_double
        L.grmp_init.argtypes = [i32, C.POINTER(vp)]
        L.grmp_finalize.argtypes = [vp]
        L.grmp_device_synchronize.argtypes = [vp]
        L.grmp_grid_create.argtypes = [vp, i32, i64, vp, i64, vp, vp, vp, C.POINTER(vp)]
        L.grmp_grid_create_bfaces.argtypes = [vp, i32, i64, vp, i64, vp, vp, vp, C.POINTER(vp)]
        L.grmp_grid_set_faces.argtypes = [vp, i64, vp, vp, vp, vp, vp]
        L.grmp_grid_update_geometry.argtypes = [vp, vp, vp]
        L.grmp_grid_destroy.argtypes = [vp]
        L.grmp_grid_update_cells.argtypes = [vp, vp]
        L.grmp_space_update_dofs.argtypes = [vp, vp]
        L.grmp_blf_numeric_steps.argtypes = [vp, dbl, i32, C.POINTER(dbl)]
        L.grmp_blf_set_owned_columns.argtypes = [vp, i64]
        L.grmp_space_create.argtypes = [vp, i32, i32, i64, i32, vp, C.POINTER(vp)]
        L.grmp_space_destroy.argtypes = [vp]
        L.grmp_blf_create.argtypes = [vp, vp, i32, i32, i32, vp, i32, i32, vp, i32, i32, vp, C.POINTER(EvalTab), C.POINTER(EvalTab),
                                      C.POINTER(vp)]
        L.grmp_blf_destroy.argtypes = [vp]
        L.grmp_blf_set_path.argtypes = [vp, i32]
        L.grmp_blf_symbolic.argtypes = [vp, dbl, C.POINTER(i64)]
        L.grmp_blf_get_pattern.argtypes = [vp, vp, vp]
        L.grmp_blf_numeric.argtypes = [vp, dbl, vp]
        L.grmp_blf_assemble_host.argtypes = [vp, dbl, vp, vp, vp, vp, vp, vp]
        L.grmp_blf_get_values.argtypes = [vp, vp]
        L.grmp_blf_transpose_copy.argtypes = [vp, dbl, dbl, vp, vp, vp]
        L.grmp_blf_stats.argtypes = [vp, C.POINTER(Stats)]
        L.grmp_blf_device_values.argtypes = [vp, C.POINTER(vp)]
        L.grmp_lf_create.argtypes = [vp, i32, vp, i32, i32, vp, C.POINTER(EvalTab), C.POINTER(vp)]
        L.grmp_lf_destroy.argtypes = [vp]
        L.grmp_lf_set_path.argtypes = [vp, i32]
        L.grmp_lf_assemble.argtypes = [vp, dbl, i32, vp, vp, i64]
        L.grmp_lf_stats.argtypes = [vp, C.POINTER(Stats)]
        L.grmp_blf_set_newton_argument.argtypes = [vp, i32, C.POINTER(EvalTab), vp, i32]
        L.grmp_blf_newton_rhs.argtypes = [vp, vp, i64]
        L.grmp_lf_assemble_feb.argtypes = [vp, dbl, vp, i32, C.POINTER(EvalTab), vp, vp, i64]
        L.grmp_blf_device_csc.argtypes = [vp, C.POINTER(DeviceCSC)]
        L.grmp_blf_matmul.argtypes = [vp, vp, vp, dbl, i32]
        L.grmp_blf_matmul_device.argtypes = [vp, vp, vp, dbl, i32]
        L.grmp_blf_residual.argtypes = [vp, vp, vp, vp, i64, vp, C.POINTER(dbl)]
        L.grmp_blf_apply_penalties.argtypes = [vp, vp, i64, dbl, C.POINTER(i64)]
        L.grmp_blf_set_fixed_argument.argtypes = [vp, vp, i32, C.POINTER(EvalTab), vp, i32]
        L.grmp_ii_create.argtypes = [vp, i32, i32, vp, i32, i32, vp, C.POINTER(EvalTab), C.POINTER(vp)]
        L.grmp_ii_destroy.argtypes = [vp]
        L.grmp_ii_resultdim.argtypes = [vp, C.POINTER(i32)]
        L.grmp_ii_evaluate.argtypes = [vp, vp, dbl, vp, vp, vp]
        _lib = L
    return _lib


def check(rc: int):
    if rc != OK:
        raise GrmpError(f"libgrmp_cuda error {rc}: {lib().grmp_last_error().decode()}")


def ptr(a):
    return None if a is None else a.ctypes.data


_ctx = {}


def context(device: int | None = None):
    """one grmp_ctx per (process, device); LOCAL_RANK selects the device under torchrun"""
    if device is None:
        device = int(os.environ.get("LOCAL_RANK", "0"))
    if device not in _ctx:
        h = C.c_void_p()
        check(lib().grmp_init(device, C.byref(h)))
        _ctx[device] = h
    return _ctx[device]
