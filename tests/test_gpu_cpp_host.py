"""GPU: the C++ host layer end to end.  (1) the reference's validation suite and solver tests through
`RK4Solver().solve(scenario, config)` (tests/cpp/test_validation_gpu.cpp); (2) bit parity of what
`Solver::solve` returns (time points, outlet, trajectory states, final state) with the CPU oracle."""
import os
import subprocess

import numpy as np
import pytest

import chromatography_b200 as cb
from chromatography_b200 import workloads as W
from oracle import oracle as O
from helpers import PROFILE_RTOL, assert_bit_equal, max_rel_err, to_oracle

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "tests", "cpp", "_bin")


def test_reference_validation_suite_through_cpp_solver():
    r = subprocess.run([os.path.join(BIN, "test_validation_gpu")], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    assert " 0 failed" in r.stdout


def _ramp(multi, nz=100):
    i = np.arange(nz)
    a = 0.01 * (i % 7).astype(np.float64) / 7.0
    if not multi:
        return a
    return np.stack([a, 0.02 * (i % 5).astype(np.float64) / 5.0], axis=1)  # [nz, 2]


@pytest.mark.parametrize("ramp", [0, 1], ids=["zero_ic", "user_ic"])
@pytest.mark.parametrize("arith", ["exact", "fast"])
@pytest.mark.parametrize("method", ["rk4", "euler"])
@pytest.mark.parametrize("kind", ["single", "multi"])
def test_cpp_solve_matches_oracle(tmp_path, kind, method, arith, ramp):
    T, steps, nz = 120.0, 2000, 100
    out = tmp_path / "dump.bin"
    r = subprocess.run([os.path.join(BIN, "chrom_solve_dump"), kind, method, arith, repr(T), str(steps), str(ramp), str(out)],
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "steps=2000 dt=0.06" in r.stdout
    data = np.fromfile(out, dtype=np.float64)
    om = O.RK4 if method == "rk4" else O.EULER
    multi = kind == "multi"
    n = 2 if multi else 1
    y0 = _ramp(multi) if ramp else None
    if multi:
        ref = O.solve_multi(to_oracle(W.acids_multi()), om, T, steps, y0=y0, want_trajectory=True)
        ref_outlet = ref.outlet                                                          # [n][steps+1]
        ref_traj = [np.asarray(s).T.ravel() for s in ref.trajectory[::50]]               # [nz, n] -> column-major
        ref_final = np.asarray(ref.final_state).T.ravel()
    else:
        ref = O.solve_single(to_oracle(W.tfa_single()), om, T, steps, y0=y0, want_trajectory=True)
        ref_outlet = ref.outlet[None, :]
        ref_traj = [np.asarray(s) for s in ref.trajectory[::50]]
        ref_final = np.asarray(ref.final_state)
    assert ref.status == 0
    n_traj = steps // 50 + 1
    assert data.size == (steps + 1) * (1 + n) + (n_traj + 1) * nz * n
    tp, rest = data[:steps + 1], data[steps + 1:]
    outlet, rest = rest[:n * (steps + 1)].reshape(n, steps + 1), rest[n * (steps + 1):]
    traj, final = rest[:n_traj * nz * n].reshape(n_traj, nz * n), rest[n_traj * nz * n:]
    assert_bit_equal(tp, ref.time_points, "time_points")
    if arith == "exact":
        assert_bit_equal(outlet, ref_outlet, "outlet")
        assert_bit_equal(traj, np.stack(ref_traj), "state_trajectory[::50]")
        assert_bit_equal(final, ref_final, "final_state")
    else:
        assert max_rel_err(outlet, ref_outlet) <= PROFILE_RTOL
        assert max_rel_err(traj, np.stack(ref_traj)) <= PROFILE_RTOL
        assert max_rel_err(final, ref_final) <= PROFILE_RTOL
