"""Where the time of one C3 one-shot solve (host buffers in, host buffers out) goes: the library's own timing fields
next to the wall clock of the call."""
import ctypes as C
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import bench  # noqa: E402
import chromatography_b200 as cb  # noqa: E402
from chromatography_b200 import _ffi  # noqa: E402

lib = _ffi.lib()
ctx = cb.Context([0])
spec = bench.workload_spec("c3")
inp = bench.build_inputs(spec, spec["n_cols"], 0)
n_cols = spec["n_cols"]
cfg = cb.make_cfg(cb.RK4, inp["total_time"], spec["steps"], spec["nz"], 1, 1, 0, cb.ARITH_FAST)
outlet, p1 = bench.pinned_array(lib, (n_cols, spec["steps"] + 1))
final, p2 = bench.pinned_array(lib, (n_cols, spec["nz"]))
status, p3 = bench.pinned_array(lib, (n_cols,), np.int64)
dp = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))  # noqa: E731
tables = inp["tables"]
for it in range(4):
    t0 = time.perf_counter()
    rc = lib.chrom_cuda_solve_single_batch(ctx._h, C.byref(cfg), inp["cols"], n_cols, dp(tables), tables.shape[0], None,
                                           dp(outlet), None, dp(final), None, status.ctypes.data_as(C.POINTER(C.c_int64)))
    wall = (time.perf_counter() - t0) * 1e3
    ctx.check(rc)
    print("call %d: wall %.1f ms" % (it, wall), {k: round(v, 2) if isinstance(v, float) else v for k, v in ctx.timing().items()})
