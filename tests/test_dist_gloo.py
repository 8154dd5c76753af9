"""world_size = 2 on CPU (gloo): the host side of the multi-GPU path.  Columns are independent, so the
N>1 design is: every rank owns its own columns (weak scaling, seed 42 + rank), no data-path collective,
one barrier and two scalar reductions for the timing (bench.Ranks / bench.aggregate_value)."""
import hashlib
import json
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_two_ranks_gloo(tmp_path):
    env = dict(os.environ, OMP_NUM_THREADS="1", GLOO_TEST_OUT=str(tmp_path))
    p = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", "29517",
                        os.path.join(ROOT, "tests", "_gloo_worker.py")], capture_output=True, text=True, env=env,
                       timeout=600)
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [json.load(open(tmp_path / f"rank{r}.json")) for r in (0, 1)]
    assert sorted(r["rank"] for r in lines) == [0, 1] and all(r["world"] == 2 for r in lines)
    r0, r1 = sorted(lines, key=lambda r: r["rank"])
    # weak scaling: same unit count per rank, different columns per rank
    assert r0["units"] == r1["units"] and r0["params_sha"] != r1["params_sha"] and r0["final_sha"] != r1["final_sha"]
    assert r0["bad"] == 0 and r1["bad"] == 0
    # reductions: both ranks agree on max time, summed units and the aggregate value
    assert r0["max_ms"] == r1["max_ms"] == max(r0["ms"], r1["ms"])
    assert r0["sum_units"] == r1["sum_units"] == r0["units"] + r1["units"]
    want = (r0["units"] + r1["units"]) / (max(r0["ms"], r1["ms"]) * 1e-3)
    assert abs(r0["value"] - want) <= 1e-9 * want and r0["value"] == r1["value"]
    # rank r's inputs are exactly what a single process generates for that rank (deterministic shards)
    sys.path.insert(0, ROOT)
    import bench
    spec = dict(bench.workload_spec("c3"))
    for r in (r0, r1):
        cols = np.frombuffer(bench.build_inputs(spec, 6, r["rank"])["cols"], dtype=np.float64)
        assert hashlib.sha256(cols.tobytes()).hexdigest() == r["params_sha"]


def test_single_rank_needs_no_process_group():
    sys.path.insert(0, ROOT)
    import bench
    os.environ.pop("WORLD_SIZE", None)
    r = bench.Ranks("gloo")
    assert not r.distributed and r.max(3.0) == 3.0 and r.sum(4.0) == 4.0
    r.barrier()
    assert bench.aggregate_value(r, 100.0, 50.0) == 2000.0
