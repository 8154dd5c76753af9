"""GPU parity: LangmuirSingle RK4 / Euler through the C ABI against the CPU oracle.

EXACT arithmetic must be bit-identical to the oracle (integer-exact comparison of the f64 bit
patterns); FAST arithmetic must be within north_star's 1e-9 (profiles) / 1e-7 (retention time, area).
The reference's own tests this mirrors: validation/main.rs:292-440 (linear/non-linear TFA, Rsf),
injection.rs:470-546 (Dirac / Rectangle / Gaussian semantics through the solver), rk4.rs / euler.rs
trajectory-length and time-grid tests, solver/mod.rs validate_state.
"""
import numpy as np
import pytest

import chromatography_b200 as cb
from chromatography_b200 import workloads as W
from oracle import oracle as O
from helpers import MOMENT_RTOL, PROFILE_RTOL, assert_bit_equal, max_rel_err, to_oracle

pytestmark = pytest.mark.gpu

METHODS = {"rk4": (cb.RK4, O.RK4), "euler": (cb.EULER, O.EULER)}


def run_gpu(ctx, models, method, total_time, steps, arith, y0=None, snapshot_stride=0, outlet_stride=1,
            want_summary=False):
    cfg = cb.make_cfg(method, total_time, steps, models[0].n_points, 1, outlet_stride, snapshot_stride, arith)
    cols, tables = cb.pack_single(models, method, total_time, steps)
    return cb.solve_single_batch(ctx, cfg, cols, tables, y0, True, snapshot_stride > 0, True, want_summary)


@pytest.mark.parametrize("method", ["rk4", "euler"])
def test_c1_exact_is_bit_identical_full_trajectory(ctx, method):
    """Config 1 (README quick start): every state of the trajectory, bit for bit."""
    gm, om = METHODS[method]
    m = W.tfa_single()
    T, steps = W.C1["total_time"], W.C1["steps"]
    ref = O.solve_single(to_oracle(m), om, T, steps, want_trajectory=True)
    r = run_gpu(ctx, [m], gm, T, steps, cb.ARITH_EXACT, snapshot_stride=1)
    assert r.status[0] == 0 and ref.status == 0
    assert_bit_equal(r.snapshots[0], ref.trajectory, "trajectory")
    assert_bit_equal(r.outlet[0], ref.outlet, "outlet")
    assert_bit_equal(r.final_state[0], ref.final_state, "final_state")
    # the survey-session sanity values of the restated RK4
    if method == "rk4":
        i = int(np.argmax(r.outlet[0]))
        assert abs(r.outlet[0][i] - 8.531523813e-3) < 1e-11 and abs(cb.time_points(T, steps)[i] - 358.08) < 1e-9


@pytest.mark.parametrize("method", ["rk4", "euler"])
def test_c1_fast_within_tolerance(ctx, method):
    gm, om = METHODS[method]
    m = W.tfa_single()
    T, steps = W.C1["total_time"], W.C1["steps"]
    ref = O.solve_single(to_oracle(m), om, T, steps, want_trajectory=True)
    r = run_gpu(ctx, [m], gm, T, steps, cb.ARITH_FAST, snapshot_stride=100, want_summary=True)
    assert r.status[0] == 0
    err_traj = max_rel_err(r.snapshots[0], ref.trajectory[::100])
    err_out = max_rel_err(r.outlet[0], ref.outlet)
    assert err_traj <= PROFILE_RTOL and err_out <= PROFILE_RTOL, (err_traj, err_out)
    tp = ref.time_points
    area_ref, tr_ref = O.trapezoid(tp, ref.outlet), O.first_moment(tp, ref.outlet)
    assert abs(O.trapezoid(tp, r.outlet[0]) - area_ref) <= MOMENT_RTOL * abs(area_ref)
    assert abs(O.first_moment(tp, r.outlet[0]) - tr_ref) <= MOMENT_RTOL * abs(tr_ref)
    # on-device summary: area, first moment, max, time of max
    s = r.summary[0]
    assert abs(s[0] - area_ref) <= MOMENT_RTOL * abs(area_ref)
    assert abs(s[1] - tr_ref) <= MOMENT_RTOL * abs(tr_ref)
    i = int(np.argmax(ref.outlet))
    assert abs(s[2] - ref.outlet[i]) <= PROFILE_RTOL * ref.outlet[i] and abs(s[3] - tp[i]) < 1e-9


@pytest.mark.parametrize("arith", [cb.ARITH_EXACT, cb.ARITH_FAST])
def test_validation_linear_tfa_known_answers(ctx, arith):
    """validation/main.rs:292-440 through the CUDA path: t_R 624 s +-1 %, sigma 61.39 s +-10 %,
    mass recovery >= 0.90, Rsf(Euler, RK4) < 0.05 (published 0.016)."""
    T, steps = 900.0, 4500
    t_inj = 2.0 * T / steps
    m = cb.LangmuirSingle.new(1.0, 0.5, 6.0, 0.4, 1e-3, 0.3, 100, cb.TemporalInjection.rectangle(0.0, t_inj, 1e-3))
    rk = run_gpu(ctx, [m], cb.RK4, T, steps, arith)
    eu = run_gpu(ctx, [m], cb.EULER, T, steps, arith)
    tp = cb.time_points(T, steps)
    t_r = O.first_moment(tp, rk.outlet[0])
    assert abs(t_r - 624.0) / 624.0 < 0.01
    assert abs(O.peak_sigma(tp, rk.outlet[0]) - 61.39) / 61.39 < 0.10
    assert O.trapezoid(tp, rk.outlet[0]) / (1e-3 * t_inj) >= 0.90
    rsf = O.rsf(tp, eu.outlet[0], tp, rk.outlet[0])
    assert rsf < 0.05 and abs(rsf - 0.016) < 1e-3
    if arith == cb.ARITH_EXACT:
        ref = O.solve_single(to_oracle(m), O.RK4, T, steps)
        assert_bit_equal(rk.outlet[0], ref.outlet, "linear TFA outlet")


def test_nonlinear_tfa_peak_compression(ctx):
    """validation/main.rs:353-428: overloaded column elutes earlier, mass recovery >= 0.85."""
    T, steps = 900.0, 4500
    t_inj = 2.0 * T / steps
    tp = cb.time_points(T, steps)
    modes = []
    for c0 in (1e-3, 1.0):
        m = cb.LangmuirSingle.new(1.0, 0.5, 6.0, 0.4, 1e-3, 0.3, 100, cb.TemporalInjection.rectangle(0.0, t_inj, c0))
        r = run_gpu(ctx, [m], cb.RK4, T, steps, cb.ARITH_FAST)
        ref = O.solve_single(to_oracle(m), O.RK4, T, steps)
        assert max_rel_err(r.outlet[0], ref.outlet) <= PROFILE_RTOL
        modes.append(tp[int(np.argmax(r.outlet[0]))])
        if c0 == 1.0:
            assert O.trapezoid(tp, r.outlet[0]) / (c0 * t_inj) >= 0.85
    assert modes[1] < modes[0]


@pytest.mark.parametrize("inj", [cb.TemporalInjection.dirac(0.0, 1e-3), cb.TemporalInjection.dirac(5.0, 0.1),
                                 cb.TemporalInjection.dirac(0.03, 0.1),  # hit only at a mid-stage time t + dt/2.0
                                 cb.TemporalInjection.rectangle(5.0, 15.0, 0.05), cb.TemporalInjection.none()],
                         ids=["dirac0", "dirac5", "dirac_midstage", "rectangle", "none"])
@pytest.mark.parametrize("method", ["rk4", "euler"])
def test_injection_semantics_exact(ctx, inj, method):
    """Dirac uses bit-exact time equality on the stage times t, t+dt/2.0, t+dt (injection.rs:418)."""
    gm, om = METHODS[method]
    m = W.tfa_single(injection=inj)
    T, steps = 60.0, 1000  # dt = 0.06: step 0 stage 2 is t = 0.03
    ref = O.solve_single(to_oracle(m), om, T, steps, want_trajectory=True)
    r = run_gpu(ctx, [m], gm, T, steps, cb.ARITH_EXACT, snapshot_stride=1)
    assert_bit_equal(r.snapshots[0], ref.trajectory, "trajectory")
    if inj.kind == 0:
        assert not r.final_state.any()


@pytest.mark.parametrize("nz", [2, 3, 5, 31, 32, 33, 64, 100, 101, 257, 640, 1000, 1024, 1025, 1500, 4100, 8192])
def test_grid_sizes_exact_and_fast(ctx, nz):
    """Every lane mapping: sub-warp columns, one warp, multi-warp CTA; ragged last chunk."""
    steps, T = 300, 30.0
    rng = np.random.default_rng(nz)
    y0 = rng.uniform(0.0, 0.02, size=(1, nz))
    m = W.tfa_single(nz=nz)
    m.dz = 0.25 / 100.0  # keep CFL fixed as nz grows
    ref = O.solve_single(to_oracle(m), O.RK4, T, steps, y0=y0[0], want_trajectory=True)
    r = run_gpu(ctx, [m], cb.RK4, T, steps, cb.ARITH_EXACT, y0=y0, snapshot_stride=50)
    assert r.status[0] == 0
    assert_bit_equal(r.snapshots[0], ref.trajectory[::50], f"nz={nz} snapshots")
    assert_bit_equal(r.outlet[0], ref.outlet, f"nz={nz} outlet")
    assert_bit_equal(r.final_state[0], ref.final_state, f"nz={nz} final")
    f = run_gpu(ctx, [m], cb.RK4, T, steps, cb.ARITH_FAST, y0=y0)
    assert max_rel_err(f.final_state[0], ref.final_state) <= PROFILE_RTOL
    assert max_rel_err(f.outlet[0], ref.outlet) <= PROFILE_RTOL


@pytest.mark.parametrize("n_cols", [1, 7, 8, 9, 257])
def test_sweep_batch_matches_oracle(ctx, n_cols):
    """C3-shaped sweep (varied K, N, u, porosity), ragged warp occupancy, shorter horizon."""
    models = W.tfa_sweep_models(n_cols, seed=42)
    T, steps = 120.0, 2000
    oo, of, ost = O.solve_single_batch([to_oracle(m) for m in models], O.RK4, T, steps)
    r = run_gpu(ctx, models, cb.RK4, T, steps, cb.ARITH_EXACT)
    assert not r.status.any() and not ost.any()
    assert_bit_equal(r.outlet, oo, "outlet")
    assert_bit_equal(r.final_state, of, "final")
    f = run_gpu(ctx, models, cb.RK4, T, steps, cb.ARITH_FAST)
    assert max_rel_err(f.final_state, of) <= PROFILE_RTOL


def test_strides_and_sample_counts(ctx):
    m = W.tfa_single()
    T, steps = 60.0, 1003
    ref = O.solve_single(to_oracle(m), O.RK4, T, steps, want_trajectory=True)
    r = run_gpu(ctx, [m], cb.RK4, T, steps, cb.ARITH_EXACT, snapshot_stride=100, outlet_stride=7)
    assert r.outlet.shape == (1, steps // 7 + 1) and r.snapshots.shape == (1, steps // 100 + 1, 100)
    assert_bit_equal(r.outlet[0], ref.outlet[::7], "strided outlet")
    assert_bit_equal(r.snapshots[0], ref.trajectory[::100], "strided snapshots")


@pytest.mark.parametrize("T", [6000.0, 12000.0])  # RK4: Inf first at 6000 (-147), NaN first at 12000 (+84)
@pytest.mark.parametrize("method", ["rk4", "euler"])
def test_validate_state_status_matches_oracle(ctx, method, T):
    """CFL-violating steps blow up: the first step holding Inf (no NaN) / NaN must match the oracle
    (solver/mod.rs:514-553), per column, and healthy columns in the same batch are untouched."""
    gm, om = METHODS[method]
    steps = 400  # dt >= 15 s on dz = 2.5 mm: unstable
    bad = W.tfa_single(injection=cb.TemporalInjection.rectangle(0.0, 100.0, 1.0))
    good = cb.LangmuirSingle.new(1.2, 0.4, 2.0, 0.4, 1e-6, 0.25, 100, cb.TemporalInjection.rectangle(0.0, 100.0, 1.0))
    models = [bad, good, bad, good, good, bad, good, good, good]
    oo, of, ost = O.solve_single_batch([to_oracle(m) for m in models], om, T, steps)
    assert ost[0] != 0 and ost[1] == 0
    r = run_gpu(ctx, models, gm, T, steps, cb.ARITH_EXACT)
    np.testing.assert_array_equal(r.status, ost)
    assert_bit_equal(r.final_state[1], of[1], "healthy column")
    msg = cb.status_message(int(r.status[0]))
    assert msg == O.validate_state_message(int(ost[0]))
    f = run_gpu(ctx, models, gm, T, steps, cb.ARITH_FAST)
    assert (f.status != 0).tolist() == (ost != 0).tolist()


def test_solver_trait_mirror(ctx):
    """The reference-facing API: Scenario + SolverConfiguration -> SimulationResult (rk4.rs:205-350)."""
    m = W.tfa_single()
    sc = cb.Scenario.new(m, cb.DomainBoundaries.temporal(m.setup_initial_state()))
    cfgs = cb.SolverConfiguration.time_evolution(60.0, 1000)
    res = cb.CudaRK4Solver(ctx, arith=cb.ARITH_EXACT).solve(sc, cfgs)
    ref = O.solve_single(to_oracle(m), O.RK4, 60.0, 1000, want_trajectory=True)
    assert len(res) == 1001 and res.state_trajectory.shape == (1001, 100)
    assert_bit_equal(res.time_points, ref.time_points, "time_points")
    assert_bit_equal(res.state_trajectory, ref.trajectory, "state_trajectory")
    assert_bit_equal(res.final_state, ref.final_state, "final_state")
    assert res.metadata == {"solver": "Runge-Kutta 4", "time steps": "1000", "dt": "0.06", "total time": "60",
                            "function evaluations": "4000"}
    eu = cb.CudaEulerSolver(ctx).solve(sc, cfgs)
    assert eu.metadata == {"solver": "Forward Euler", "time steps": "1000", "dt": "0.06", "total time": "60"}
    assert cb.CudaRK4Solver(ctx).name() == "Runge Kutta (RK4)" and cb.CudaEulerSolver(ctx).name() == "Forward Euler"
    with pytest.raises(cb.SolverError, match="Total time must be positive"):
        cb.CudaRK4Solver(ctx).solve(sc, cb.SolverConfiguration.time_evolution(-1.0, 10))
    with pytest.raises(cb.SolverError, match="TimeSteps must be greater than 0"):
        cb.CudaEulerSolver(ctx).solve(sc, cb.SolverConfiguration.time_evolution(1.0, 0))
    with pytest.raises(cb.SolverError, match="EulerSolver only supports TimeEvolution configuration, got Iterative"):
        cb.CudaRK4Solver(ctx).solve(sc, cb.SolverConfiguration.iterative(1e-6, 10))
    with pytest.raises(cb.SolverError, match="No initial condition found in domain boundaries"):
        cb.CudaRK4Solver(ctx).solve(cb.Scenario.new(m, cb.DomainBoundaries.temporal(None)), cfgs)


def test_plan_reexecution_is_deterministic(ctx):
    models = W.tfa_sweep_models(64)
    cfg = cb.make_cfg(cb.RK4, 60.0, 500, 100)
    cols, tables = cb.pack_single(models, cb.RK4, 60.0, 500)
    plan = cb.Plan(ctx, cfg, cols, tables)
    assert "single_resident" in plan.describe()
    t1 = plan.execute()
    a = plan.download()
    plan.execute()
    b = plan.download()
    assert t1["kernel_ms"] > 0 and t1["kernel_launches"] == 1
    assert_bit_equal(a.final_state, b.final_state, "re-execution")
    plan.close()


def test_pipelined_one_shot_equals_plan_and_oracle(ctx):
    """A batch large enough for the chunked compute / copy-back pipeline of the one-shot solve
    (>= 2 waves of columns, >= 64 MB of outputs) must give the same bits as the plan path (one launch)
    and match the oracle on a sample of columns."""
    n_cols, T, steps = 20_000, 30.0, 500
    a = W.tfa_sweep_arrays(n_cols, seed=7)
    from chromatography_b200 import _ffi
    cols = (_ffi.SingleCol * n_cols)()
    arr = np.frombuffer(cols, dtype=np.float64).reshape(n_cols, 7)
    for j, k in enumerate(["lambda_", "langmuir_k", "port_number", "fe", "ue", "dz"]):
        arr[:, j] = a[k]
    arr[:, 6] = 0.0
    tables = cb.tabulate_injection(cb.TemporalInjection.gaussian(10.0, 3.0, 0.1), cb.RK4, T, steps)[None]
    cfg = cb.make_cfg(cb.RK4, T, steps, 100, arith=cb.ARITH_EXACT)
    one = cb.solve_single_batch(ctx, cfg, cols, tables, want_summary=True)
    assert one.timing["kernel_launches"] > 1, "expected the chunked pipeline"
    plan = cb.Plan(ctx, cfg, cols, tables, want=cb.WANT_OUTLET | cb.WANT_FINAL | cb.WANT_STATUS | cb.WANT_SUMMARY)
    plan.execute()
    ref = plan.download()
    plan.close()
    assert_bit_equal(one.outlet, ref.outlet, "outlet")
    assert_bit_equal(one.final_state, ref.final_state, "final")
    assert_bit_equal(one.summary, ref.summary, "summary")
    assert not one.status.any()
    pick = [0, 1, 9471, 9472, 9473, 18943, 18944, 19_999]
    models = W.tfa_sweep_models(n_cols, seed=7)
    oo, of, _ = O.solve_single_batch([to_oracle(models[i]) for i in pick], O.RK4, T, steps)
    assert_bit_equal(one.outlet[pick], oo, "outlet vs oracle")
    assert_bit_equal(one.final_state[pick], of, "final vs oracle")


def test_solve_scan_cfl_sweep(ctx):
    """The reference's CFL scan (examples/cfl_stability_scan.rs:378-394: one solve per step count, sequential loop)
    and spatial-convergence scan (spatial_temporal_convergence.rs:261-277: one per grid) as ONE solve_scan call:
    groups run concurrently on their own contexts; every entry equals its own oracle solve bit for bit, unstable
    entries carry the reference's validate_state error."""
    T = 6000.0
    scenarios, configs, expect = [], [], []
    for steps in (100, 1500, 4000, 10000, 1500, 4000):  # CFL from unstable (NaN at step 62) to comfortable, two repeated
        m = W.tfa_single()
        scenarios.append(cb.Scenario.new(m, cb.DomainBoundaries.temporal(m.setup_initial_state())))
        configs.append(cb.SolverConfiguration.time_evolution(T, steps))
        expect.append((m, steps))
    for nz in (50, 200):
        m = W.tfa_single(nz=nz)
        scenarios.append(cb.Scenario.new(m, cb.DomainBoundaries.temporal(m.setup_initial_state())))
        configs.append(cb.SolverConfiguration.time_evolution(T, 20000))
        expect.append((m, 20000))
    solver = cb.CudaRK4Solver(ctx, arith=cb.ARITH_EXACT, full_trajectory=False)
    out = solver.solve_scan(scenarios, configs, max_concurrent=3)
    assert len(out) == len(scenarios)
    n_err = 0
    for r, (m, steps) in zip(out, expect):
        ref = O.solve_single(to_oracle(m), O.RK4, T, steps)
        if ref.status != 0:
            n_err += 1
            assert isinstance(r, cb.SolverError) and str(r) == cb.status_message(ref.status)
            continue
        assert_bit_equal(r.final_state, ref.final_state, f"steps={steps} nz={m.points()} final")
        assert_bit_equal(r.outlet, ref.outlet, f"steps={steps} nz={m.points()} outlet")
        assert r.metadata["time steps"] == str(steps)
    assert 0 < n_err < len(out)  # the scan crosses the stability limit
    with pytest.raises(cb.SolverError, match="one SolverConfiguration per scenario"):
        solver.solve_scan(scenarios, configs[:-1])


@pytest.mark.parametrize("method", ["rk4", "euler"])
@pytest.mark.parametrize("stride", [1, 3])
def test_time_sliced_last_chunk(ctx, method, stride):
    """One-shot solves without an on-device summary advance their LAST chunk of columns in time slices (state
    carried through final_state) so that its outlet series goes back slice by slice.  Same bits as the single
    launch of the plan path, including the slot arithmetic of a non-unit outlet stride and a validate_state
    failure that happens in an early slice of a column of the last chunk (later slices must not overwrite it)."""
    gm, om = METHODS[method]
    n_cols, T, steps = (20_000 if stride == 1 else 48_000), 30.0, 500  # >= 64 MB of outputs: the chunked pipeline
    a = W.tfa_sweep_arrays(n_cols, seed=11)
    from chromatography_b200 import _ffi
    cols = (_ffi.SingleCol * n_cols)()
    arr = np.frombuffer(cols, dtype=np.float64).reshape(n_cols, 7)
    for j, k in enumerate(["lambda_", "langmuir_k", "port_number", "fe", "ue", "dz"]):
        arr[:, j] = a[k]
    arr[:, 6] = 0.0
    arr[n_cols - 2, 4] *= 400.0  # CFL >> 1 in a column of the last chunk: non-finite before the last slice
    tables = cb.tabulate_injection(cb.TemporalInjection.gaussian(10.0, 3.0, 0.1), gm, T, steps)[None]
    cfg = cb.make_cfg(gm, T, steps, 100, 1, stride, 0, cb.ARITH_EXACT)
    one = cb.solve_single_batch(ctx, cfg, cols, tables)
    plan = cb.Plan(ctx, cfg, cols, tables, want=cb.WANT_OUTLET | cb.WANT_FINAL | cb.WANT_STATUS)
    plan.execute()
    ref = plan.download()
    plan.close()
    assert one.timing["kernel_launches"] > ref.timing["kernel_launches"] + 3, "expected chunks and time slices"
    assert ref.status[n_cols - 2] != 0 and abs(ref.status[n_cols - 2]) < steps - steps // 4  # before the last slice
    np.testing.assert_array_equal(one.status, ref.status)
    ok = ref.status == 0
    assert ok.sum() == n_cols - 1
    assert_bit_equal(one.outlet[ok], ref.outlet[ok], "outlet")
    assert_bit_equal(one.final_state[ok], ref.final_state[ok], "final")
    models = W.tfa_sweep_models(n_cols, seed=11)
    pick = [n_cols - 1, n_cols - 3, n_cols - 4000]
    oo, of, _ = O.solve_single_batch([to_oracle(models[i]) for i in pick], om, T, steps)
    assert_bit_equal(one.outlet[pick], oo[:, ::stride], "outlet vs oracle")
    assert_bit_equal(one.final_state[pick], of, "final vs oracle")


@pytest.mark.parametrize("method", ["rk4", "euler"])
def test_randomised_isotherms_including_extremely_steep(ctx, monkeypatch, method):
    """Random LangmuirSingle columns over wide parameter ranges, loads up to K c >> 1, and isotherms so steep
    (fe nbar K / (1 + fe lambda) > 1e3) that the host must drop the polynomial FAST form (which subtracts two nearly
    equal numbers there) for the cancellation-free one; a batch mixing both kinds takes the safe form as a whole.
    The plan description says which form ran; both stay within 1e-9 of the oracle, EXACT is bit-identical."""
    gm, om = METHODS[method]
    rng = np.random.default_rng(77)
    nz, steps = 120, 60
    inj = cb.TemporalInjection.rectangle(0.0, 0.3, 0.5)

    def model(k, n_ports, lam):
        return cb.LangmuirSingle.new(lam, k, n_ports, 0.4, 1e-3, 0.25, nz, inj)

    mild = [model(float(np.exp(rng.uniform(np.log(0.05), np.log(20.0)))), float(rng.uniform(0.5, 3.0)),
                  float(rng.uniform(0.0, 1.5))) for _ in range(12)]
    steep = [model(float(rng.uniform(2e3, 2e4)), float(rng.uniform(1.0, 3.0)), 0.0) for _ in range(4)]
    y0 = rng.uniform(0.0, 1.0, size=(16, nz))
    T = steps * 0.2 * mild[0].dz / mild[0].ue
    for name, models in (("mild", mild), ("steep", steep), ("mixed", mild[:4] + steep)):
        cfg = cb.make_cfg(gm, T, steps, nz, 1, 1, 0, cb.ARITH_FAST)
        cols, tables = cb.pack_single(models, gm, T, steps)
        plan = cb.Plan(ctx, cfg, cols, tables, y0=y0[:len(models)])
        desc = plan.describe()
        plan.close()
        assert ("fast-poly" in desc) == (name == "mild"), (name, desc)
        oo, of, ost = O.solve_single_batch([to_oracle(m) for m in models], om, T, steps, y0=y0[:len(models)])
        f = run_gpu(ctx, models, gm, T, steps, cb.ARITH_FAST, y0=y0[:len(models)])
        np.testing.assert_array_equal(f.status, ost)
        ok = ost == 0
        assert ok.any()
        assert max_rel_err(f.final_state[ok], of[ok]) <= PROFILE_RTOL, name
        assert max_rel_err(f.outlet[ok], oo[ok]) <= PROFILE_RTOL, name
        e = run_gpu(ctx, models, gm, T, steps, cb.ARITH_EXACT, y0=y0[:len(models)])
        np.testing.assert_array_equal(e.status, ost)
        assert_bit_equal(e.final_state[ok], of[ok], f"{name} exact")
