"""ctypes binding of libchrom_cuda.so (include/chrom_cuda.h).

The library is built in-tree by ``__graft_entry__.build()`` / ``make -C chromatography_b200/csrc``.
There is no fallback of any kind: if the shared library is missing this module raises, and if no
B200 is visible ``chrom_cuda_create`` fails with CHROM_ERR_NO_DEVICE.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# CHROM_CUDA_LIB: another build of the same library (kernel A/B experiments); default: the in-tree build
LIB_PATH = os.environ.get("CHROM_CUDA_LIB") or os.path.join(_HERE, "libchrom_cuda.so")

CHROM_OK = 0
CHROM_ERR_INVALID, CHROM_ERR_CUDA, CHROM_ERR_UNSUPPORTED, CHROM_ERR_NO_DEVICE = -1, -2, -3, -4
EULER, RK4 = 0, 1
ARITH_FAST, ARITH_EXACT, ARITH_FAST_STRUCTURED, ARITH_FAST_DENSE = 0, 1, 3, 4
INJ_NONE, INJ_DIRAC, INJ_GAUSSIAN, INJ_RECTANGLE = 0, 1, 2, 3
WANT_OUTLET, WANT_SNAPSHOTS, WANT_FINAL, WANT_SUMMARY, WANT_STATUS = 1, 2, 4, 8, 16
SUMMARY_LEN = 6  # area, first moment, max, time of max, sigma, min (CHROM_SUMMARY_LEN)


class SingleCol(C.Structure):
    _fields_ = [("lambda_", C.c_double), ("langmuir_k", C.c_double), ("port_number", C.c_double),
                ("fe", C.c_double), ("ue", C.c_double), ("dz", C.c_double),
                ("inj_table", C.c_uint32), ("detector_cell", C.c_uint32)]


class MultiCol(C.Structure):
    _fields_ = [("fe", C.c_double), ("ue", C.c_double), ("dz", C.c_double), ("stationary_fraction", C.c_double),
                ("species_off", C.c_uint32), ("detector_cell", C.c_uint32)]


class Species(C.Structure):
    _fields_ = [("lambda_", C.c_double), ("langmuir_k", C.c_double), ("port_number", C.c_uint32),
                ("inj_table", C.c_uint32)]


class RunCfg(C.Structure):
    _fields_ = [("method", C.c_int32), ("arith", C.c_int32), ("total_time", C.c_double),
                ("time_steps", C.c_uint64), ("n_points", C.c_uint32), ("n_species", C.c_uint32),
                ("outlet_stride", C.c_uint64), ("snapshot_stride", C.c_uint64),
                ("outlet_steps", C.POINTER(C.c_uint64)), ("n_outlet_steps", C.c_uint64),
                ("snapshot_steps", C.POINTER(C.c_uint64)), ("n_snapshot_steps", C.c_uint64)]


class Timing(C.Structure):
    _fields_ = [("h2d_ms", C.c_double), ("kernel_ms", C.c_double), ("d2h_ms", C.c_double),
                ("total_ms", C.c_double), ("kernel_launches", C.c_uint64), ("h2d_bytes", C.c_uint64),
                ("d2h_bytes", C.c_uint64)]


# every symbol include/chrom_cuda.h declares: name -> (restype, argtypes)
_dp = C.POINTER(C.c_double)
_vp = C.c_void_p
SYMBOLS = {
    "chrom_cuda_device_count": (C.c_int32, []),
    "chrom_cuda_create": (C.c_int32, [C.POINTER(C.c_int32), C.c_int32, C.POINTER(_vp)]),
    "chrom_cuda_destroy": (None, [_vp]),
    "chrom_cuda_last_error": (C.c_char_p, [_vp]),
    "chrom_cuda_version": (C.c_char_p, []),
    "chrom_cuda_last_timing": (C.c_int32, [_vp, C.POINTER(Timing)]),
    "chrom_cuda_host_alloc": (C.c_int32, [C.c_size_t, C.POINTER(_vp)]),
    "chrom_cuda_host_free": (None, [_vp]),
    "chrom_cuda_bind_host_thread": (C.c_int32, [_vp, C.c_int32, C.POINTER(C.c_int32), C.POINTER(C.c_int32)]),
    "chrom_cuda_validate_cfg": (C.c_char_p, [C.POINTER(RunCfg)]),
    "chrom_cuda_n_samples": (C.c_uint64, [C.c_uint64, C.c_uint64]),
    "chrom_cuda_cfg_n_out": (C.c_uint64, [C.POINTER(RunCfg)]),
    "chrom_cuda_cfg_n_snap": (C.c_uint64, [C.POINTER(RunCfg)]),
    "chrom_cuda_sample_indices": (C.c_uint64, [C.c_uint64, C.c_uint64, C.POINTER(C.c_uint64), C.c_uint64]),
    "chrom_cuda_node_index": (C.c_uint32, [C.c_double, C.c_double, C.c_uint32]),
    "chrom_cuda_time_points": (None, [C.c_double, C.c_uint64, _dp]),
    "chrom_cuda_table_stages": (C.c_uint32, [C.c_int32]),
    "chrom_cuda_tabulate_injection": (C.c_int32, [C.c_int32, C.c_double, C.c_double, C.c_double, C.c_int32,
                                                  C.c_double, C.c_uint64, _dp]),
    "chrom_cuda_shard_range": (None, [C.c_uint64, C.c_uint32, C.c_uint32, C.POINTER(C.c_uint64),
                                      C.POINTER(C.c_uint64)]),
    "chrom_cuda_status_message": (C.c_size_t, [C.c_int64, C.c_char_p, C.c_size_t]),
    "chrom_cuda_solve_single_batch": (C.c_int32, [_vp, C.POINTER(RunCfg), C.POINTER(SingleCol), C.c_uint64, _dp,
                                                  C.c_uint32, _dp, _dp, _dp, _dp, _dp, C.POINTER(C.c_int64)]),
    "chrom_cuda_solve_multi_batch": (C.c_int32, [_vp, C.POINTER(RunCfg), C.POINTER(MultiCol), C.c_uint64,
                                                 C.POINTER(Species), C.c_uint64, _dp, C.c_uint32, _dp, _dp, _dp,
                                                 _dp, _dp, C.POINTER(C.c_int64)]),
    "chrom_cuda_plan_single": (C.c_int32, [_vp, C.POINTER(RunCfg), C.POINTER(SingleCol), C.c_uint64, _dp,
                                           C.c_uint32, _dp, C.c_uint32, C.POINTER(_vp)]),
    "chrom_cuda_plan_multi": (C.c_int32, [_vp, C.POINTER(RunCfg), C.POINTER(MultiCol), C.c_uint64,
                                          C.POINTER(Species), C.c_uint64, _dp, C.c_uint32, _dp, C.c_uint32,
                                          C.POINTER(_vp)]),
    "chrom_cuda_plan_execute": (C.c_int32, [_vp]),
    "chrom_cuda_plan_download": (C.c_int32, [_vp, _dp, _dp, _dp, _dp, C.POINTER(C.c_int64)]),
    "chrom_cuda_plan_rsf": (C.c_int32, [_vp, _vp, _dp]),
    "chrom_cuda_plan_describe": (C.c_int32, [_vp, C.c_char_p, C.c_size_t]),
    "chrom_cuda_plan_destroy": (None, [_vp]),
    "chrom_cuda_measure_fp64_peak": (C.c_int32, [_vp, C.c_int32, _dp, _dp]),
    "chrom_cuda_flush_l2": (C.c_int32, [_vp]),
    "chrom_cuda_measure_copy_bw": (C.c_int32, [_vp, C.c_int32, C.c_size_t, _dp]),
    "chrom_cuda_probe_rcp": (C.c_int32, [_vp, C.c_double, C.c_double, C.c_uint32, _dp]),
}

_lib = None


def lib():
    """The loaded library; raises if it has not been built (no fallback)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "or `make -C chromatography_b200/csrc`. chromatography_b200 has no CPU fallback.")
        handle = C.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(handle, name)  # AttributeError if the header and the library disagree
            fn.restype = res
            fn.argtypes = args
        _lib = handle
    return _lib
