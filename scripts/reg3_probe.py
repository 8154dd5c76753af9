"""Runs ONE mixed LangmuirMulti batch through one kernel variant and prints OK (used under `timeout` to find which
variant / column kind stalls).  python scripts/reg3_probe.py n nz method(rk4|euler) gen sync kind(mild|strong|late|unstable|all) [steps]"""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
n, nz, method, gen, sync, kind = int(sys.argv[1]), int(sys.argv[2]), sys.argv[3], sys.argv[4], sys.argv[5], sys.argv[6]
steps = int(sys.argv[7]) if len(sys.argv) > 7 else 150
os.environ["CHROM_REG2"] = gen
os.environ["CHROM_REG3_SYNC"] = sync
import numpy as np  # noqa: E402

import chromatography_b200 as cb  # noqa: E402
import test_gpu_multi_reg3 as T  # noqa: E402

dt = 0.2 * (0.25 / nz) / (1e-3 / 0.4)
models, y0 = T.mixed_batch(n, nz, 100 * n + nz, late_start=100.5 * dt)
pick = {"mild": [0, 4], "strong": [1], "late": [2], "unstable": [3], "all": [0, 1, 2, 3, 4]}[kind]
models = [models[i] for i in pick]
y0 = np.ascontiguousarray(y0[pick])
ctx = cb.Context([0])
m = cb.RK4 if method == "rk4" else cb.EULER
r = T.run(ctx, models, y0, m, steps * dt, steps, snapshot_stride=50)
print("OK", n, nz, method, gen, sync, kind, "status", r.status.tolist(), "kernel_ms", round(r.timing["kernel_ms"], 3), flush=True)
ctx.close()
