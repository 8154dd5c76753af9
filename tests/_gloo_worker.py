"""Worker of tests/test_dist_gloo.py: one rank of the bench's N>1 path on CPU (gloo).
The device solve is replaced by the oracle (this is a test of the HOST side of the multi-GPU path:
per-rank input generation, shard ranges of the strong- and weak-scaling modes, barrier, max/sum reductions).
GLOO_TEST_SCALING selects the mode; GLOO_TEST_COLS the batch size (strong: in total, weak: per rank)."""
import hashlib
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

import bench  # noqa: E402
from oracle import oracle as O  # noqa: E402

ranks = bench.Ranks("gloo")
spec = dict(bench.workload_spec("c3"))
spec["steps"] = 200          # shortened horizon for the CPU stand-in
spec["total_time"] = 12.0
scaling = os.environ.get("GLOO_TEST_SCALING", "strong")
n_total = int(os.environ.get("GLOO_TEST_COLS", "11"))
inp = bench.build_inputs(spec, ranks.rank, ranks.world, scaling, n_total)
n_cols = len(inp["cols"])
cols = np.frombuffer(inp["cols"], dtype=np.float64).reshape(n_cols, 7).copy()
ranks.barrier()
t0 = time.perf_counter()
models = [O.LangmuirSingle(c[0], c[1], c[2], c[3], c[4], c[5], spec["nz"], O.TemporalInjection.gaussian(10.0, 3.0, 0.1))
          for c in cols]
_, final, status = O.solve_single_batch(models, O.RK4, spec["total_time"], spec["steps"], n_threads=1)
ms = (time.perf_counter() - t0) * 1e3 + 10.0 * ranks.rank   # make the max-over-ranks visible
ranks.barrier()
units = bench.units_of(spec, n_cols)
value = bench.aggregate_value(ranks, units, ms)
out = json.dumps({"rank": ranks.rank, "world": ranks.world, "units": units, "n_cols": n_cols, "col0": inp["col0"], "ms": ms, "value": value,
                  "max_ms": ranks.max(ms), "sum_units": ranks.sum(units),
                  "params_sha": hashlib.sha256(cols.tobytes()).hexdigest(),
                  "final_sha": hashlib.sha256(final.tobytes()).hexdigest(), "bad": int(np.count_nonzero(status))})
with open(os.path.join(os.environ["GLOO_TEST_OUT"], f"rank{ranks.rank}.json"), "w") as f:
    f.write(out)
ranks.close()
