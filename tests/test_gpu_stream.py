"""GPU parity: the streaming LangmuirSingle path (nz > 8192, one launch per time step, 4-cell halo
recomputation per warp tile) against the CPU oracle, plus a size-independent property at the full
BASELINE config-4 size (nz = 2^22): the upwind scheme has no downstream dependence, so the first
8192 cells of the long column must equal — bit for bit in EXACT mode — a resident-path solve of
an 8192-cell column with the same dz."""
import numpy as np
import pytest

import chromatography_b200 as cb
from chromatography_b200 import workloads as W
from oracle import oracle as O
from helpers import PROFILE_RTOL, assert_bit_equal, max_rel_err, to_oracle

pytestmark = pytest.mark.gpu


def long_column(nz, cfl=0.5):
    """C1 parameters on nz cells of width 0.25/100 (so the front moves), dt from the CFL rule."""
    m = W.tfa_single(nz=nz)
    m.dz = 0.25 / 100.0
    n_bar = m.fe / (1.0 + m.fe) * m.port_number
    u_eff0 = m.ue / (1.0 + m.fe * (m.lambda_ + n_bar * m.langmuir_k))
    return m, cfl * m.dz / u_eff0


def run_gpu(ctx, models, method, total_time, steps, arith, y0=None, snapshot_stride=0, outlet_stride=1,
            want_summary=False):
    cfg = cb.make_cfg(method, total_time, steps, models[0].n_points, 1, outlet_stride, snapshot_stride, arith)
    cols, tables = cb.pack_single(models, method, total_time, steps)
    plan = cb.Plan(ctx, cfg, cols, tables, y0=y0, want=cb.WANT_OUTLET | cb.WANT_FINAL | cb.WANT_STATUS |
                   (cb.WANT_SNAPSHOTS if snapshot_stride else 0) | (cb.WANT_SUMMARY if want_summary else 0))
    desc = plan.describe()
    plan.execute()
    r = plan.download()
    plan.close()
    return r, desc


@pytest.mark.parametrize("nz", [8193, 8200, 8448, 10001, 65539, 1 << 18])
@pytest.mark.parametrize("method", [cb.RK4, cb.EULER], ids=["rk4", "euler"])
def test_stream_matches_oracle(ctx, nz, method):
    m, dt = long_column(nz)
    steps = 24
    T = dt * steps
    m.set_injection(cb.TemporalInjection.gaussian(0.3 * T, 0.1 * T, 0.1))
    rng = np.random.default_rng(nz)
    y0 = rng.uniform(0.0, 0.02, size=(1, nz))
    ref = O.solve_single(to_oracle(m), method, T, steps, y0=y0[0], want_trajectory=True)
    r, desc = run_gpu(ctx, [m], method, T, steps, cb.ARITH_EXACT, y0=y0, snapshot_stride=8)
    assert "single_stream" in desc
    assert r.status[0] == 0
    assert_bit_equal(r.snapshots[0], ref.trajectory[::8], "snapshots")
    assert_bit_equal(r.outlet[0], ref.outlet, "outlet")
    assert_bit_equal(r.final_state[0], ref.final_state, "final")
    f, _ = run_gpu(ctx, [m], method, T, steps, cb.ARITH_FAST, y0=y0, want_summary=True)
    assert max_rel_err(f.final_state[0], ref.final_state) <= PROFILE_RTOL
    assert max_rel_err(f.outlet[0], ref.outlet) <= PROFILE_RTOL
    area = O.trapezoid(ref.time_points, ref.outlet)
    assert abs(f.summary[0][0] - area) <= 1e-7 * abs(area)


def test_stream_batch_and_strides(ctx):
    nz, steps = 9000, 21
    models = []
    for k in (0.2, 0.4, 0.7):
        m, dt = long_column(nz)
        m.langmuir_k = k
        models.append(m)
    T = dt * steps
    y0 = np.random.default_rng(1).uniform(0.0, 0.05, size=(3, nz))
    oo, of, ost = O.solve_single_batch([to_oracle(m) for m in models], O.RK4, T, steps, y0=y0)
    r, desc = run_gpu(ctx, models, cb.RK4, T, steps, cb.ARITH_EXACT, y0=y0, outlet_stride=5, snapshot_stride=7)
    assert "single_stream" in desc and r.outlet.shape == (3, steps // 5 + 1) and r.snapshots.shape == (3, 4, nz)
    assert_bit_equal(r.outlet, oo[:, ::5], "strided outlet")
    assert_bit_equal(r.final_state, of, "final")
    assert_bit_equal(r.snapshots[:, 0], y0, "snapshot 0 = initial state")


@pytest.mark.parametrize("steps,snap", [(23, 0), (22, 6), (9, 0), (3, 0), (26, 13)])
def test_stream_fused_steps_tail_and_snapshot_strides(ctx, steps, snap):
    """Temporal blocking (up to 4 RK4 steps per launch): a step count that is not a multiple of the fusion
    depth, fewer steps than the depth, and snapshot strides that force depth 2 or 1 — all bit-identical."""
    nz = 8200 + 37
    m, dt = long_column(nz)
    T = dt * steps
    m.set_injection(cb.TemporalInjection.gaussian(0.3 * T, 0.1 * T, 0.1))
    y0 = np.random.default_rng(steps).uniform(0.0, 0.02, size=(1, nz))
    ref = O.solve_single(to_oracle(m), O.RK4, T, steps, y0=y0[0], want_trajectory=True)
    r, desc = run_gpu(ctx, [m], cb.RK4, T, steps, cb.ARITH_EXACT, y0=y0, snapshot_stride=snap)
    assert "single_stream" in desc
    assert_bit_equal(r.outlet[0], ref.outlet, "outlet (every step, sampled inside fused launches)")
    assert_bit_equal(r.final_state[0], ref.final_state, "final")
    if snap:
        assert_bit_equal(r.snapshots[0], ref.trajectory[::snap], "snapshots")


def test_stream_status_matches_oracle(ctx):
    """Unstable step on a long column: first non-finite step / class equal to the oracle's."""
    nz, steps = 9000, 300
    m, dt = long_column(nz, cfl=6.0)
    T = dt * steps
    m.set_injection(cb.TemporalInjection.rectangle(0.0, T, 1.0))
    ref = O.solve_single(to_oracle(m), O.RK4, T, steps)
    assert ref.status != 0
    r, _ = run_gpu(ctx, [m], cb.RK4, T, steps, cb.ARITH_EXACT)
    assert int(r.status[0]) == ref.status


@pytest.mark.parametrize("arith", [cb.ARITH_EXACT, cb.ARITH_FAST], ids=["exact", "fast"])
def test_full_size_prefix_property(ctx, arith):
    """BASELINE config 4 at full size (nz = 2^22): cells [0, 8192) of the streamed column equal the
    resident-path solve of the 8192-cell column (EXACT: bitwise; FAST: 1e-9), and the oracle on a
    shorter prefix."""
    nz, steps = 1 << 22, 40
    m, total_time, _ = W.fine_grid_single(nz, 1000)
    T = total_time * steps / 1000
    m.set_injection(cb.TemporalInjection.gaussian(0.4 * T, 0.15 * T, 0.1))
    big, desc = run_gpu(ctx, [m], cb.RK4, T, steps, arith)
    assert "single_stream" in desc and big.status[0] == 0
    small_model = cb.LangmuirSingle(m.lambda_, m.langmuir_k, m.port_number, m.column_length, 8192, m.dz, m.fe, m.ue,
                                    m.injection)
    small, sdesc = run_gpu(ctx, [small_model], cb.RK4, T, steps, arith)
    assert "single_resident" in sdesc
    ref = O.solve_single(to_oracle(small_model), O.RK4, T, steps)
    if arith == cb.ARITH_EXACT:
        assert_bit_equal(big.final_state[0][:8192], small.final_state[0], "prefix vs resident path")
        assert_bit_equal(big.final_state[0][:8192], ref.final_state, "prefix vs oracle")
    else:
        assert max_rel_err(big.final_state[0][:8192], ref.final_state) <= PROFILE_RTOL
    assert np.isfinite(big.final_state).all() and big.final_state[0].max() > 0.0
