"""Where the time of a one-context multi-device solve goes (run with gpurun --gpus N): resident plan vs one-shot solve,
all devices vs one device with the same share of columns.  Prints one JSON line per case."""
import ctypes as C
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import bench  # noqa: E402
import chromatography_b200 as cb  # noqa: E402
from chromatography_b200 import _ffi  # noqa: E402

lib = _ffi.lib()
n_dev = lib.chrom_cuda_device_count()
spec = bench.workload_spec("c3")
inp = bench.build_inputs(spec, 0, 1, "strong", 0)
n_all = len(inp["cols"])
tables = inp["tables"]
dp = lambda a: None if a is None else a.ctypes.data_as(C.POINTER(C.c_double))  # noqa: E731
slen = cb.SUMMARY_LEN


def run(devs, n_cols, mode):
    ctx = cb.Context(devs)
    cols = (_ffi.SingleCol * n_cols).from_buffer_copy(bytes(inp["cols"])[: n_cols * C.sizeof(_ffi.SingleCol)])
    stride = {"sweep": bench.SWEEP_OUTLET_STRIDE, "full": 1, "final": 0}[mode]
    cfg = cb.make_cfg(cb.RK4, inp["total_time"], spec["steps"], spec["nz"], 1, stride, 0, cb.ARITH_FAST)
    out = {"devices": devs, "n_cols": n_cols, "mode": mode}
    # resident plan
    want = cb.WANT_FINAL | cb.WANT_STATUS | (cb.WANT_OUTLET if stride else 0) | (cb.WANT_SUMMARY if mode == "sweep" else 0)
    plan = cb.Plan(ctx, cfg, cols, tables, want=want)
    out["kernel"] = plan.describe()
    plan.execute()
    t0 = time.perf_counter()
    t = plan.execute()
    out["plan_execute"] = {"wall_ms": (time.perf_counter() - t0) * 1e3, "kernel_ms": t["kernel_ms"]}
    plan.close()
    # one-shot
    final, p1 = bench.pinned_array(lib, (n_cols, spec["nz"]))
    status, p2 = bench.pinned_array(lib, (n_cols,), np.int64)
    summ, p3 = bench.pinned_array(lib, (n_cols, slen)) if mode == "sweep" else (None, None)
    n_out = spec["steps"] // stride + 1 if stride else 0
    outlet, p4 = bench.pinned_array(lib, (n_cols, n_out)) if stride else (None, None)
    for it in range(3):
        t0 = time.perf_counter()
        ctx.check(lib.chrom_cuda_solve_single_batch(ctx._h, C.byref(cfg), cols, n_cols, dp(tables), tables.shape[0], None,
                                                    dp(outlet), None, dp(final), dp(summ),
                                                    status.ctypes.data_as(C.POINTER(C.c_int64))))
        wall = (time.perf_counter() - t0) * 1e3
        tm = ctx.timing()
    out["one_shot"] = {"wall_ms": wall, **{k: tm[k] for k in ("kernel_ms", "h2d_ms", "d2h_ms", "total_ms", "kernel_launches")}}
    for p in (p1, p2, p3, p4):
        if p:
            lib.chrom_cuda_host_free(p)
    ctx.close()
    print(json.dumps(out), flush=True)


share = n_all // n_dev
for mode in ("final", "sweep", "full"):
    run([0], share, mode)
    run(list(range(n_dev)), n_all, mode)
