"""Scratch measurements on a B200: peaks, reciprocal accuracy, C3 timing per chunk size M."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import chromatography_b200 as cb
from chromatography_b200 import workloads as W, _ffi

ctx = cb.Context([0])
out = {}
out["fp64_peak_tflops"] = [ctx.fp64_peak_tflops() for _ in range(3)]
out["copy_bw_gbs"] = ctx.copy_bw_gbs()
out["rcp_err_seed_cubic_cubic+newton"] = ctx.probe_rcp(0.5, 64.0, 1 << 22)
out["rcp_err_wide"] = ctx.probe_rcp(1e-3, 1e6, 1 << 22)
print(json.dumps(out), flush=True)

n_cols = int(os.environ.get("COLS", 65536))
steps = int(os.environ.get("STEPS", 10000))
a = W.tfa_sweep_arrays(n_cols)
cols = (_ffi.SingleCol * n_cols)()
arr = np.frombuffer(cols, dtype=np.float64).reshape(n_cols, 7)
for j, k in enumerate(["lambda_", "langmuir_k", "port_number", "fe", "ue", "dz"]):
    arr[:, j] = a[k]
arr[:, 6] = 0.0
tables = cb.tabulate_injection(cb.TemporalInjection.gaussian(10.0, 3.0, 0.1), cb.RK4, 600.0, steps)[None]
units = n_cols * 100 * steps * 4
for arith in (cb.ARITH_FAST,) + ((cb.ARITH_EXACT,) if os.environ.get("EXACT") else ()):
    for M in os.environ.get("MS", "25,20,10,5").split(","):
        os.environ["CHROM_FORCE_M"] = M
        for want in (cb.WANT_OUTLET | cb.WANT_FINAL,):
            cfg = cb.make_cfg(cb.RK4, 600.0, steps, 100, 1, 1, 0, arith)
            plan = cb.Plan(ctx, cfg, cols, tables, want=want | cb.WANT_STATUS)
            plan.execute()
            ts = [plan.execute()["kernel_ms"] for _ in range(2)]
            ms = min(ts)
            print(json.dumps({"arith": arith, "desc": plan.describe(), "want": want, "kernel_ms": ts,
                              "cell_stage_per_s": units / (ms * 1e-3),
                              "alg_tflops": units * 15.25 / (ms * 1e-3) / 1e12}), flush=True)
            plan.close()
        if arith == cb.ARITH_EXACT:
            break
