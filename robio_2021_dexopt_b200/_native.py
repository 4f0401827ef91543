"""Loader for the native CUDA library (libtrx_b200.so, built in-tree by build.py).

The product path has no CPU fallback: if the library is missing this raises, and
if it loads but no CUDA device is present every device call returns TRX_ERR_CUDA.
"""
from __future__ import annotations

import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# TRX_B200_LIB: development builds of the same library (e.g. with per-phase clocks); never a different engine
LIB_PATH = os.environ.get("TRX_B200_LIB") or os.path.join(_HERE, "libtrx_b200.so")

_lib = None


class TrxError(RuntimeError):
    def __init__(self, status: int, message: str):
        super().__init__("trx status %d: %s" % (status, message))
        self.status = status


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            "native CUDA engine %s is missing; run `python -c 'import __graft_entry__ as g; g.build()'` "
            "(there is no CPU fallback for the tape engine)" % LIB_PATH)
    L = ctypes.CDLL(LIB_PATH)
    c = ctypes
    L.trx_last_error.restype = c.c_char_p
    L.trx_version.restype = c.c_char_p
    L.trx_device_count.restype = c.c_int
    L.trx_op_count.restype = c.c_int
    L.trx_op_name.restype = c.c_char_p
    L.trx_op_name.argtypes = [c.c_int]
    L.trx_op_signature.restype = c.c_char_p
    L.trx_op_signature.argtypes = [c.c_int]
    L.trx_kernel_launch_count.restype = c.c_uint64
    L.trx_program_create_from_blob.argtypes = [c.c_void_p, c.c_size_t, c.c_int, c.POINTER(c.c_void_p)]
    L.trx_program_destroy.argtypes = [c.c_void_p]
    L.trx_program_destroy.restype = None
    L.trx_program_buffer_bytes.argtypes = [c.c_void_p, c.c_int, c.c_uint64, c.POINTER(c.c_uint64)]
    L.trx_program_stat.argtypes = [c.c_void_p, c.c_char_p, c.POINTER(c.c_uint64)]
    L.trx_memory_create.argtypes = [c.c_uint64, c.c_int, c.POINTER(c.c_void_p)]
    L.trx_memory_destroy.argtypes = [c.c_void_p]
    L.trx_memory_destroy.restype = None
    L.trx_memory_bytes.argtypes = [c.c_void_p, c.POINTER(c.c_uint64)]
    for name in ("trx_input", "trx_parameterize", "trx_output", "trx_input_device", "trx_parameterize_device",
                 "trx_output_device"):
        getattr(L, name).argtypes = [c.c_void_p, c.c_void_p, c.c_void_p, c.c_uint64, c.c_void_p]
    L.trx_execute.argtypes = [c.c_void_p, c.c_void_p, c.c_void_p]
    L.trx_memory_peek.argtypes = [c.c_void_p, c.c_uint64, c.c_uint64, c.c_void_p, c.c_uint64]
    L.trx_collision_shape_create.argtypes = [c.c_void_p, c.POINTER(c.c_uint64)]
    L.trx_collision_pair_create.argtypes = [c.c_uint64, c.c_uint64, c.POINTER(c.c_uint64)]
    L.trx_collision_shape_center.argtypes = [c.c_uint64, c.POINTER(c.c_double)]
    L.trx_collision_clear.restype = None
    L.trx_collision_table.argtypes = [c.c_void_p, c.c_uint64, c.POINTER(c.c_uint64)]
    L.trx_collision_query_pairs.argtypes = [c.c_uint64, c.c_void_p, c.c_uint64, c.c_void_p, c.c_int]
    L.trx_collision_project_points.argtypes = [c.c_uint64, c.c_void_p, c.c_uint64, c.c_void_p, c.c_int]
    L.trx_objective_create.argtypes = [c.c_void_p, c.c_size_t, c.c_void_p, c.c_size_t, c.c_void_p, c.c_size_t,
                                       c.c_uint64, c.c_uint64, c.c_int, c.c_int, c.POINTER(c.c_void_p)]
    L.trx_objective_destroy.argtypes = [c.c_void_p]
    L.trx_objective_destroy.restype = None
    L.trx_objective_info.argtypes = [c.c_void_p, c.c_char_p, c.POINTER(c.c_uint64)]
    L.trx_objective_parameterize.argtypes = [c.c_void_p, c.c_void_p, c.c_uint64, c.c_void_p]
    L.trx_objective_evaluate.argtypes = [c.c_void_p, c.c_void_p, c.POINTER(c.c_double), c.c_void_p, c.c_void_p]
    L.trx_objective_evaluate_device.argtypes = [c.c_void_p, c.c_void_p, c.c_void_p, c.c_void_p]
    L.trx_objective_gd_steps.argtypes = [c.c_void_p, c.c_void_p, c.c_void_p, c.c_double, c.c_double, c.c_uint32,
                                         c.c_void_p, c.c_void_p]
    L.trx_objective_last_times.argtypes = [c.c_void_p, c.POINTER(c.c_float)]
    L.trx_objective_set_profiling.argtypes = [c.c_void_p, c.c_int]
    L.trx_objective_profile_collect.argtypes = [c.c_void_p, c.POINTER(c.c_double), c.POINTER(c.c_uint64)]
    L.trx_objective_attach_fprop.argtypes = [c.c_void_p, c.c_void_p, c.c_size_t]
    L.trx_objective_hessian_product.argtypes = [c.c_void_p, c.c_void_p, c.c_void_p, c.c_void_p]
    L.trx_objective_sq_steps.argtypes = [c.c_void_p, c.c_void_p, c.c_double, c.c_uint32, c.c_double, c.c_double, c.c_uint32,
                                         c.c_void_p, c.c_void_p, c.c_void_p]
    L.trx_comm_unique_id.argtypes = [c.c_void_p]
    L.trx_comm_create.argtypes = [c.c_void_p, c.c_int, c.c_int, c.c_int, c.POINTER(c.c_void_p)]
    L.trx_comm_destroy.argtypes = [c.c_void_p]
    L.trx_comm_destroy.restype = None
    L.trx_comm_allreduce.argtypes = [c.c_void_p, c.c_void_p, c.c_uint64, c.c_void_p]
    L.trx_objective_set_comm.argtypes = [c.c_void_p, c.c_void_p]
    L.trx_dexsyn_build.argtypes = [c.c_char_p, c.POINTER(c.POINTER(c.c_uint8)), c.POINTER(c.c_uint64)]
    L.trx_builder_free.argtypes = [c.POINTER(c.c_uint8)]
    L.trx_builder_free.restype = None
    L.trx_builder_last_error.restype = c.c_char_p
    _lib = L
    return L


def check(status: int) -> None:
    if status != 0:
        raise TrxError(status, lib().trx_last_error().decode(errors="replace"))
