"""world_size = 2 on CPU (gloo): the host side of the multi-GPU path.  Columns are independent, so the N>1 design
is: ONE batch cut into contiguous column ranges (chrom_cuda_shard_range), rank r solves its range, no data-path
collective, one barrier and two scalar reductions for the timing (bench.Ranks / bench.aggregate_value).  The weak
mode (every rank its own batch, seed 42 + rank) is kept as an option and tested too."""
import hashlib
import json
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def run_two_ranks(tmp_path, scaling, cols, port):
    env = dict(os.environ, OMP_NUM_THREADS="1", GLOO_TEST_OUT=str(tmp_path), GLOO_TEST_SCALING=scaling,
               GLOO_TEST_COLS=str(cols))
    p = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", str(port),
                        os.path.join(ROOT, "tests", "_gloo_worker.py")], capture_output=True, text=True, env=env,
                       timeout=600)
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [json.load(open(tmp_path / f"rank{r}.json")) for r in (0, 1)]
    assert sorted(r["rank"] for r in lines) == [0, 1] and all(r["world"] == 2 for r in lines)
    return sorted(lines, key=lambda r: r["rank"])


def check_reductions(r0, r1):
    assert r0["bad"] == 0 and r1["bad"] == 0
    # both ranks agree on max time, summed units and the aggregate value
    assert r0["max_ms"] == r1["max_ms"] == max(r0["ms"], r1["ms"])
    assert r0["sum_units"] == r1["sum_units"] == r0["units"] + r1["units"]
    want = (r0["units"] + r1["units"]) / (max(r0["ms"], r1["ms"]) * 1e-3)
    assert abs(r0["value"] - want) <= 1e-9 * want and r0["value"] == r1["value"]


def test_two_ranks_strong_scaling(tmp_path):
    """An odd batch of 11 columns over 2 ranks: 6 + 5, contiguous, and together exactly the single-process batch."""
    import bench
    r0, r1 = run_two_ranks(tmp_path, "strong", 11, 29517)
    check_reductions(r0, r1)
    assert (r0["col0"], r0["n_cols"], r1["col0"], r1["n_cols"]) == (0, 6, 6, 5)
    spec = dict(bench.workload_spec("c3"))
    whole = np.frombuffer(bench.build_inputs(spec, 0, 1, "strong", 11)["cols"], dtype=np.float64).reshape(11, 7)
    for r in (r0, r1):
        part = whole[r["col0"]:r["col0"] + r["n_cols"]]
        assert hashlib.sha256(np.ascontiguousarray(part).tobytes()).hexdigest() == r["params_sha"]
        # and what a single process generates for that rank
        again = np.frombuffer(bench.build_inputs(spec, r["rank"], 2, "strong", 11)["cols"], dtype=np.float64)
        assert hashlib.sha256(again.tobytes()).hexdigest() == r["params_sha"]


def test_two_ranks_weak_scaling(tmp_path):
    import bench
    r0, r1 = run_two_ranks(tmp_path, "weak", 6, 29518)
    check_reductions(r0, r1)
    # same unit count per rank, different columns per rank
    assert r0["units"] == r1["units"] and r0["params_sha"] != r1["params_sha"] and r0["final_sha"] != r1["final_sha"]
    spec = dict(bench.workload_spec("c3"))
    for r in (r0, r1):
        cols = np.frombuffer(bench.build_inputs(spec, r["rank"], 2, "weak", 6)["cols"], dtype=np.float64)
        assert hashlib.sha256(cols.tobytes()).hexdigest() == r["params_sha"]


@pytest.mark.parametrize("n_cols,world", [(65536, 8), (4096, 8), (11, 2), (5, 8), (1, 4), (100, 3)])
def test_shard_range_covers_the_batch(n_cols, world):
    """bench.shard_range restates chrom_cuda_shard_range (api.cu); the C function itself is checked against the
    same values in tests/test_cabi_host.py."""
    import bench
    pieces = [bench.shard_range(n_cols, world, r) for r in range(world)]
    assert pieces[0][0] == 0 and sum(c for _, c in pieces) == n_cols
    for (a0, ac), (b0, _) in zip(pieces, pieces[1:]):
        assert b0 == min(a0 + ac, n_cols)
    per = -(-n_cols // world)
    assert all(c <= per for _, c in pieces)


def test_single_rank_needs_no_process_group():
    import bench
    os.environ.pop("WORLD_SIZE", None)
    r = bench.Ranks("gloo")
    assert not r.distributed and r.max(3.0) == 3.0 and r.sum(4.0) == 4.0
    r.barrier()
    assert bench.aggregate_value(r, 100.0, 50.0) == 2000.0
