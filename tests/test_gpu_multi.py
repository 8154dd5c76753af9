"""GPU parity: LangmuirMulti RK4 / Euler through the C ABI against the CPU oracle.

EXACT arithmetic (nalgebra closed-form inverses n<=4, nalgebra LU n>=5, gemv order) must be
bit-identical to the oracle; FAST arithmetic (register LU solve with partial pivoting) within 1e-9.
Reference tests mirrored: validation/main.rs:451-528 (ascorbic/erythorbic, glucose/fructose
retention times), langmuir_multi.rs:1592-1613 (interior dC/dt = 0 on an empty column).
"""
import numpy as np
import pytest

import chromatography_b200 as cb
from chromatography_b200 import workloads as W
from oracle import oracle as O
from helpers import MOMENT_RTOL, PROFILE_RTOL, assert_bit_equal, max_rel_err, to_oracle

pytestmark = pytest.mark.gpu


def run_gpu(ctx, models, method, total_time, steps, arith, y0=None, snapshot_stride=0, outlet_stride=1,
            want_summary=False):
    n = models[0].n_species()
    cfg = cb.make_cfg(method, total_time, steps, models[0].n_points, n, outlet_stride, snapshot_stride, arith)
    cols, species, tables = cb.pack_multi(models, method, total_time, steps)
    return cb.solve_multi_batch(ctx, cfg, cols, species, tables, y0, True, snapshot_stride > 0, True, want_summary)


def random_species(n, seed, k_spread=1.0):
    rng = np.random.default_rng(seed)
    inj = cb.TemporalInjection.gaussian(3.0, 1.0, 0.1)
    return [cb.SpeciesParams.new(f"S{k}", rng.uniform(0.8, 1.5), rng.uniform(0.1, 0.8) * (k_spread ** rng.uniform(-1, 1)),
                                 int(rng.integers(1, 4)), inj) for k in range(n)]


@pytest.mark.parametrize("method", [cb.RK4, cb.EULER], ids=["rk4", "euler"])
def test_c2_acids_exact_bit_identical(ctx, method):
    """Config 2 (examples/acids_multi.rs) at a shortened horizon, whole trajectory bit for bit."""
    m = W.acids_multi()
    T, steps = 120.0, 2000
    ref = O.solve_multi(to_oracle(m), method, T, steps, want_trajectory=True)
    r = run_gpu(ctx, [m], method, T, steps, cb.ARITH_EXACT, snapshot_stride=1)
    assert r.status[0] == 0 and ref.status == 0
    assert_bit_equal(r.snapshots[0].transpose(0, 2, 1), ref.trajectory, "trajectory")
    assert_bit_equal(r.outlet[0], ref.outlet, "outlet")
    assert_bit_equal(r.final_state[0].T, ref.final_state, "final_state")


def test_c2_acids_full_config_fast(ctx):
    """Config 2 in full (RK4, 1200 s / 20 000 steps): profiles 1e-9, retention time / area 1e-7,
    and the survey's sanity peaks 6.8385735e-3 @ 449.28 s, 5.3867524e-3 @ 553.56 s."""
    m = W.acids_multi()
    T, steps = W.C2["total_time"], W.C2["steps"]
    ref = O.solve_multi(to_oracle(m), O.RK4, T, steps)
    r = run_gpu(ctx, [m], cb.RK4, T, steps, cb.ARITH_FAST, want_summary=True)
    assert r.status[0] == 0
    assert max_rel_err(r.outlet[0], ref.outlet) <= PROFILE_RTOL
    assert max_rel_err(r.final_state[0].T, ref.final_state) <= PROFILE_RTOL
    tp = ref.time_points
    for k, (peak, t_peak) in enumerate([(6.8385735e-3, 449.28), (5.3867524e-3, 553.56)]):
        i = int(np.argmax(r.outlet[0][k]))
        assert abs(r.outlet[0][k][i] - peak) < 1e-9 and abs(tp[i] - t_peak) < 1e-9
        area, tr = O.trapezoid(tp, ref.outlet[k]), O.first_moment(tp, ref.outlet[k])
        assert abs(O.trapezoid(tp, r.outlet[0][k]) - area) <= MOMENT_RTOL * area
        assert abs(O.first_moment(tp, r.outlet[0][k]) - tr) <= MOMENT_RTOL * tr
        assert abs(r.summary[0][k][0] - area) <= MOMENT_RTOL * area
        assert abs(r.summary[0][k][1] - tr) <= MOMENT_RTOL * tr


@pytest.mark.parametrize("arith", [cb.ARITH_EXACT, cb.ARITH_FAST], ids=["exact", "fast"])
def test_validation_multi_known_answers(ctx, arith):
    """validation/main.rs:451-528: ascorbic/erythorbic t_R 448 / 556 s +-1 %, gap 108 s +-1 %;
    glucose/fructose (linear, lambda = 0) 168.6 / 202.8 s +-1 %."""
    cases = [
        (0.25, [("Ascorbic", 1.0, 1.1, 2), ("Erythorbic", 1.0, 1.7, 2)], 800.0, 4000, [448.0, 556.0]),
        (0.3, [("Glucose", 0.0, 0.27 / 0.6, 1), ("Fructose", 0.0, 0.46 / 0.6, 1)], 400.0, 2000, [168.6, 202.8]),
    ]
    for length, sp, T, steps, t_ref in cases:
        inj = cb.TemporalInjection.rectangle(0.0, 2.0 * T / steps, 1e-3)
        m = cb.LangmuirMulti.new([cb.SpeciesParams.new(n, l, k, p, inj) for n, l, k, p in sp], 100, 0.4, 1e-3, length)
        r = run_gpu(ctx, [m], cb.RK4, T, steps, arith)
        tp = cb.time_points(T, steps)
        t_r = [O.first_moment(tp, r.outlet[0][k]) for k in range(2)]
        for a, b in zip(t_r, t_ref):
            assert abs(a - b) / b < 0.01
        gap = t_ref[1] - t_ref[0]
        assert abs((t_r[1] - t_r[0]) - gap) / gap < 0.01 or sp[0][0] == "Glucose"
        if arith == cb.ARITH_EXACT:
            ref = O.solve_multi(to_oracle(m), O.RK4, T, steps)
            assert_bit_equal(r.outlet[0], ref.outlet, "outlet")


@pytest.mark.parametrize("n", [1, 2, 3, 4, 5, 6, 7, 8])
@pytest.mark.parametrize("nz", [33, 100, 300])
def test_species_counts_exact_and_fast(ctx, n, nz):
    """Every inverse path of nalgebra (n=1, 2, 3, 4 closed forms; n>=5 LU) and every register-LU size."""
    steps, T = 120, 12.0
    rng = np.random.default_rng(100 * n + nz)
    y0 = rng.uniform(0.0, 0.05, size=(nz, n))
    m = cb.LangmuirMulti.new(random_species(n, seed=n), nz, 0.4, 1e-3, 0.25 * nz / 100.0)
    ref = O.solve_multi(to_oracle(m), O.RK4, T, steps, y0=y0, want_trajectory=True)
    y0_dev = np.ascontiguousarray(y0.T)[None]
    r = run_gpu(ctx, [m], cb.RK4, T, steps, cb.ARITH_EXACT, y0=y0_dev, snapshot_stride=40)
    assert r.status[0] == 0 and ref.n_singular == 0
    assert_bit_equal(r.snapshots[0].transpose(0, 2, 1), ref.trajectory[::40], f"n={n} nz={nz} snapshots")
    assert_bit_equal(r.outlet[0], ref.outlet, f"n={n} nz={nz} outlet")
    for arith in (cb.ARITH_FAST, cb.ARITH_FAST_DENSE):
        f = run_gpu(ctx, [m], cb.RK4, T, steps, arith, y0=y0_dev)
        assert max_rel_err(f.final_state[0].T, ref.final_state) <= PROFILE_RTOL
        assert max_rel_err(f.outlet[0], ref.outlet) <= PROFILE_RTOL


@pytest.mark.parametrize("n", [1, 2, 3, 4, 5, 6, 7, 8])
@pytest.mark.parametrize("nz", [33, 100, 512])
def test_species_counts_euler(ctx, n, nz):
    """EulerSolver (euler.rs:227-269) through every kernel family that serves n <= 8: EXACT bit-identical, FAST and
    FAST_DENSE within 1e-9, three columns per batch (the persistent-CTA column loop), several warps per column."""
    steps = 90
    T = steps * 0.2 * (0.25 / 100.0) / (1e-3 / 0.4)
    rng = np.random.default_rng(7 * n + nz)
    y0 = rng.uniform(0.0, 0.05, size=(nz, n))
    m = cb.LangmuirMulti.new(random_species(n, seed=n), nz, 0.4, 1e-3, 0.25 * nz / 100.0)
    ref = O.solve_multi(to_oracle(m), O.EULER, T, steps, y0=y0, want_trajectory=True)
    y0_dev = np.repeat(np.ascontiguousarray(y0.T)[None], 3, axis=0)
    r = run_gpu(ctx, [m] * 3, cb.EULER, T, steps, cb.ARITH_EXACT, y0=y0_dev, snapshot_stride=30)
    assert not r.status.any() and ref.status == 0
    for c in range(3):
        assert_bit_equal(r.snapshots[c].transpose(0, 2, 1), ref.trajectory[::30], f"n={n} nz={nz} snapshots")
        assert_bit_equal(r.outlet[c], ref.outlet, f"n={n} nz={nz} outlet")
    for arith in (cb.ARITH_FAST, cb.ARITH_FAST_DENSE):
        f = run_gpu(ctx, [m] * 3, cb.EULER, T, steps, arith, y0=y0_dev, snapshot_stride=30, want_summary=True)
        assert not f.status.any()
        for c in range(3):
            assert max_rel_err(f.snapshots[c].transpose(0, 2, 1), ref.trajectory[::30]) <= PROFILE_RTOL
            assert max_rel_err(f.final_state[c].T, ref.final_state) <= PROFILE_RTOL
            assert max_rel_err(f.outlet[c], ref.outlet) <= PROFILE_RTOL


@pytest.mark.parametrize("n", [9, 12, 16, 24, 40])
@pytest.mark.parametrize("method", [cb.RK4, cb.EULER], ids=["rk4", "euler"])
def test_many_species_runtime_n(ctx, n, method):
    """n > 8: one warp per cell (multi_tiled.cu).  EXACT = nalgebra LU + gemv restated with lanes over rows /
    columns (bit-identical), FAST = the closed-form structured elimination (dense row-exchanging LU on a scratch
    matrix when a row exchange is needed)."""
    nz, steps, T = 40, 60, 6.0
    rng = np.random.default_rng(1000 + n)
    y0 = rng.uniform(0.0, 0.05, size=(nz, n))
    m = cb.LangmuirMulti.new(random_species(n, seed=n), nz, 0.4, 1e-3, 0.1)
    ref = O.solve_multi(to_oracle(m), method, T, steps, y0=y0, want_trajectory=True)
    y0_dev = np.ascontiguousarray(y0.T)[None]
    r = run_gpu(ctx, [m, m], method, T, steps, cb.ARITH_EXACT, y0=np.concatenate([y0_dev, y0_dev]), snapshot_stride=20)
    assert not r.status.any() and ref.n_singular == 0
    for c in range(2):
        assert_bit_equal(r.snapshots[c].transpose(0, 2, 1), ref.trajectory[::20], f"n={n} snapshots col {c}")
        assert_bit_equal(r.outlet[c], ref.outlet, f"n={n} outlet col {c}")
    f = run_gpu(ctx, [m], method, T, steps, cb.ARITH_FAST, y0=y0_dev)
    assert max_rel_err(f.final_state[0].T, ref.final_state) <= PROFILE_RTOL
    assert max_rel_err(f.outlet[0], ref.outlet) <= PROFILE_RTOL


def pivoting_case(n):
    """A column whose cell matrices need row exchanges.  |a[r][i]| / a[i][i] is bounded by
    nbar_r / nbar_i * K_r c_r / (1 + K_r c_r), so pivoting needs a weakly-sited (N=1), strongly binding,
    nearly absent species ahead of heavily loaded N=3 species, and lambda = 0."""
    nz = 64
    rng = np.random.default_rng(n)
    inj = cb.TemporalInjection.rectangle(0.0, 1.0, 2.0)
    half = max(1, n // 2)
    sp = [cb.SpeciesParams.new(f"S{k}", 0.0, 50.0 * (1 + k), 1, inj) if k < half else
          cb.SpeciesParams.new(f"S{k}", 0.0, 1.0, 3, inj) for k in range(n)]
    m = cb.LangmuirMulti.new(sp, nz, 0.4, 1e-3, 0.16)
    y0 = np.zeros((nz, n))
    y0[:, :half] = rng.uniform(0.0, 1e-4, (nz, half))
    y0[:, half:] = rng.uniform(0.0, 1e-3, (nz, n - half))
    y0[:, -1] = rng.uniform(1.5, 3.0, nz)  # one dominant loaded species
    return m, y0


@pytest.mark.parametrize("n", [2, 3, 5, 8, 12])
def test_partial_pivoting_is_exercised(ctx, n):
    """The LU must pivot (premise asserted on the first cell's matrix).  EXACT stays bit-identical
    (nalgebra's row exchanges restated), FAST (compare-and-swap pivoting) within tolerance."""
    steps, T = 60, 0.6
    m, y0 = pivoting_case(n)
    om = to_oracle(m)
    A = np.eye(n) + m.fe * om.jacobian(y0[0])
    assert any(abs(A[r, i]) > abs(A[i, i]) for i in range(n) for r in range(i + 1, n))
    ref = O.solve_multi(om, O.RK4, T, steps, y0=y0)
    y0_dev = np.ascontiguousarray(y0.T)[None]
    r = run_gpu(ctx, [m], cb.RK4, T, steps, cb.ARITH_EXACT, y0=y0_dev)
    assert r.status[0] == 0 and ref.status == 0
    assert_bit_equal(r.final_state[0].T, ref.final_state, f"n={n} pivoting, exact")
    for arith in (cb.ARITH_FAST, cb.ARITH_FAST_DENSE):  # structured: pivot test fires -> dense redo
        f = run_gpu(ctx, [m], cb.RK4, T, steps, arith, y0=y0_dev)
        assert max_rel_err(f.final_state[0].T, ref.final_state) <= PROFILE_RTOL


def test_c5_shaped_batch(ctx):
    """Config 5 shape (n=8, nz=512) on a few columns and a short horizon, against the oracle batch."""
    models = W.multi_batch_models(5, 8, 512, seed=42)
    T, steps = 600.0 * 40 / 4096, 40
    oo, of, ost = O.solve_multi_batch([to_oracle(m) for m in models], O.RK4, T, steps)
    r = run_gpu(ctx, models, cb.RK4, T, steps, cb.ARITH_EXACT)
    assert not r.status.any() and not ost.any()
    assert_bit_equal(r.final_state, of, "final")
    assert_bit_equal(r.outlet, oo, "outlet")
    for arith in (cb.ARITH_FAST, cb.ARITH_FAST_DENSE):
        f = run_gpu(ctx, models, cb.RK4, T, steps, arith)
        assert max_rel_err(f.final_state, of) <= PROFILE_RTOL
    many = W.multi_batch_models(700, 8, 64, seed=3)  # more columns than persistent CTAs: the column loop
    T2 = 600.0 * 10 / 4096
    oo, of, ost = O.solve_multi_batch([to_oracle(m) for m in many], O.RK4, T2, 10)
    r = run_gpu(ctx, many, cb.RK4, T2, 10, cb.ARITH_EXACT)
    assert_bit_equal(r.final_state, of, "700 columns, persistent CTAs")


def test_interior_zero_on_empty_column(ctx):
    """langmuir_multi.rs:1592-1613: one Euler step on an empty column moves only the inlet cell."""
    inj = cb.TemporalInjection.dirac(0.0, 1e-3)
    m = cb.LangmuirMulti.new([cb.SpeciesParams.new("A", 1.0, 0.5, 1, inj), cb.SpeciesParams.new("B", 1.0, 2.0, 1, inj)],
                             100, 0.4, 1e-3, 0.25)
    r = run_gpu(ctx, [m], cb.EULER, 0.1, 1, cb.ARITH_EXACT)
    fs = r.final_state[0]  # [n, nz]
    assert (fs[:, 0] > 0).all() and not fs[:, 1:].any()


def test_status_matches_oracle(ctx):
    """Unstable step size: first non-finite step and its NaN/Inf class must equal the oracle's."""
    inj = cb.TemporalInjection.rectangle(0.0, 100.0, 1.0)
    sp = [cb.SpeciesParams.new("A", 1.0, 1.1, 2, inj), cb.SpeciesParams.new("B", 1.0, 1.7, 2, inj)]
    bad = cb.LangmuirMulti.new(sp, 100, 0.4, 1e-3, 0.25)
    good = cb.LangmuirMulti.new(sp, 100, 0.4, 1e-6, 0.25)
    for T in (6000.0, 12000.0, 60000.0):
        oo, of, ost = O.solve_multi_batch([to_oracle(m) for m in (bad, good)], O.RK4, T, 400)
        r = run_gpu(ctx, [bad, good], cb.RK4, T, 400, cb.ARITH_EXACT)
        assert ost[0] != 0 and ost[1] == 0
        np.testing.assert_array_equal(r.status, ost)
        assert_bit_equal(r.final_state[1], of[1], "healthy column")


def test_solver_trait_mirror_multi(ctx):
    m = W.acids_multi()
    sc = cb.Scenario.new(m, cb.DomainBoundaries.temporal(m.setup_initial_state()))
    res = cb.CudaRK4Solver(ctx, arith=cb.ARITH_EXACT).solve(sc, cb.SolverConfiguration.time_evolution(30.0, 500))
    ref = O.solve_multi(to_oracle(m), O.RK4, 30.0, 500, want_trajectory=True)
    assert res.state_trajectory.shape == (501, 100, 2) and res.final_state.shape == (100, 2)
    assert_bit_equal(res.state_trajectory, ref.trajectory, "state_trajectory")
    assert_bit_equal(res.final_state, ref.final_state, "final_state")
    assert res.metadata["function evaluations"] == "2000"


@pytest.mark.parametrize("nz", [2, 3, 31, 32, 33, 64, 65, 257])
@pytest.mark.parametrize("n", [1, 3, 8])
def test_ragged_grid_sizes(ctx, n, nz):
    """Grid sizes around the warp / thread-tile boundaries (the smallest legal column has 2 cells): the register-
    resident FAST kernel (1, 2 or 4 cells per thread), the dense-LU and the exact shared-memory kernels."""
    steps, T = 30, 3.0
    rng = np.random.default_rng(17 * n + nz)
    y0 = rng.uniform(0.0, 0.05, size=(nz, n))
    m = cb.LangmuirMulti.new(random_species(n, seed=n), nz, 0.4, 1e-3, 0.25 * nz / 100.0)
    ref = O.solve_multi(to_oracle(m), O.RK4, T, steps, y0=y0, want_trajectory=True)
    y0_dev = np.ascontiguousarray(y0.T)[None]
    batch = [m] * 3  # three columns: the persistent-CTA column loop with a non-zero initial state
    r = run_gpu(ctx, batch, cb.RK4, T, steps, cb.ARITH_EXACT, y0=np.repeat(y0_dev, 3, axis=0), snapshot_stride=10)
    assert not r.status.any()
    for c in range(3):
        assert_bit_equal(r.snapshots[c].transpose(0, 2, 1), ref.trajectory[::10], f"n={n} nz={nz} snapshots")
        assert_bit_equal(r.outlet[c], ref.outlet, f"n={n} nz={nz} outlet")
    for arith in (cb.ARITH_FAST, cb.ARITH_FAST_DENSE):
        f = run_gpu(ctx, batch, cb.RK4, T, steps, arith, y0=np.repeat(y0_dev, 3, axis=0), snapshot_stride=10,
                    want_summary=True)
        for c in range(3):
            assert max_rel_err(f.snapshots[c].transpose(0, 2, 1), ref.trajectory[::10]) <= PROFILE_RTOL
            assert max_rel_err(f.outlet[c], ref.outlet) <= PROFILE_RTOL
            assert max_rel_err(f.final_state[c].T, ref.final_state) <= PROFILE_RTOL
            for k in range(n):
                area = O.trapezoid(ref.time_points, ref.outlet[k])
                assert abs(f.summary[c][k][0] - area) <= MOMENT_RTOL * abs(area)


@pytest.mark.parametrize("seed", range(12))
def test_randomised_strongly_nonlinear_cases(ctx, seed):
    """Random species counts, grids, affinities over three decades and loads up to strongly non-linear (K c >> 1):
    whichever path a cell takes in FAST mode — sufficient bound, per-step pivot test, dense row-exchanging LU —
    the result must stay within 1e-9 of the oracle and validate_state must agree; EXACT stays bit-identical."""
    rng = np.random.default_rng(9000 + seed)
    n = int(rng.choice([2, 3, 5, 8, 11, 16, 20, 33]))
    nz = int(rng.integers(10, 300))
    lam = rng.uniform(0.0, 1.5, n)
    K = np.exp(rng.uniform(np.log(0.05), np.log(50.0), n))
    ports = rng.integers(1, 4, n)
    conc = rng.uniform(0.01, 2.0)
    inj = cb.TemporalInjection.rectangle(0.0, 0.5, conc)
    sp = [cb.SpeciesParams.new(f"S{k}", float(lam[k]), float(K[k]), int(ports[k]), inj) for k in range(n)]
    m = cb.LangmuirMulti.new(sp, nz, 0.4, 1e-3, 0.25 * nz / 100.0)
    y0 = rng.uniform(0.0, 1.0, size=(nz, n)) * (rng.uniform(size=(1, n)) < 0.7)  # some species absent at first
    steps, T = 40, 40 * 0.2 * m.dz / m.ue  # CFL 0.2 on the unretained velocity
    om = to_oracle(m)
    ref = O.solve_multi(om, O.RK4, T, steps, y0=y0)
    y0_dev = np.ascontiguousarray(y0.T)[None]
    e = run_gpu(ctx, [m], cb.RK4, T, steps, cb.ARITH_EXACT, y0=y0_dev)
    assert e.status[0] == ref.status
    if ref.status == 0:
        assert_bit_equal(e.final_state[0].T, ref.final_state, f"seed {seed}: n={n} nz={nz} exact")
    for arith in (cb.ARITH_FAST,) + ((cb.ARITH_FAST_DENSE,) if n <= 8 else ()):
        f = run_gpu(ctx, [m], cb.RK4, T, steps, arith, y0=y0_dev)
        assert f.status[0] == ref.status
        if ref.status == 0:
            assert max_rel_err(f.final_state[0].T, ref.final_state) <= PROFILE_RTOL, (seed, n, nz, arith)
            assert max_rel_err(f.outlet[0], ref.outlet) <= PROFILE_RTOL, (seed, n, nz, arith)


@pytest.mark.parametrize("n", [11, 16])
@pytest.mark.parametrize("nz", [2, 17, 32])
def test_short_columns_many_species_inlet_rows(ctx, n, nz):
    """nz <= 32 with n >= 11: the register-resident kernel runs 32 threads per column while a step's inlet rows
    hold n * 3 = 33..48 values (RK4) — every row must be loaded (strided load), with a non-zero injection that
    differs per species so a stale or missing entry cannot go unnoticed."""
    steps, T = 90, 9.0
    rng = np.random.default_rng(7000 + 100 * n + nz)
    species = [cb.SpeciesParams.new(f"S{k}", rng.uniform(0.8, 1.5), rng.uniform(0.1, 0.8), int(rng.integers(1, 4)),
                                    cb.TemporalInjection.gaussian(2.0 + 0.3 * k, 0.8 + 0.05 * k, 0.02 * (k + 1)))
               for k in range(n)]
    m = cb.LangmuirMulti.new(species, nz, 0.4, 1e-3, 0.25 * nz / 100.0)
    ref = O.solve_multi(to_oracle(m), O.RK4, T, steps)
    for arith in (cb.ARITH_FAST, cb.ARITH_EXACT):
        r = run_gpu(ctx, [m], cb.RK4, T, steps, arith)
        assert r.status[0] == 0
        if arith == cb.ARITH_EXACT:
            assert_bit_equal(r.outlet[0], ref.outlet, f"n={n} nz={nz} outlet")
        else:
            assert max_rel_err(r.outlet[0], ref.outlet) <= PROFILE_RTOL
            assert max_rel_err(r.final_state[0].T, ref.final_state) <= PROFILE_RTOL
