"""Pins the CPU oracle (oracle/chrom_oracle.cpp) against every known-answer test the reference holds
for the hot path (tests/golden/reference_known_answers.json, transcribed with file:line), plus
structure-independent cross-checks of the nalgebra restatement (numpy.linalg, Sherman-Morrison).
Runs on CPU; no GPU, no product code."""
import json
import os

import numpy as np
import pytest

from oracle import oracle as O

KA = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "reference_known_answers.json")))


# ---- injection semantics: src/models/injection.rs:470-546 ------------------------------------------------
def test_dirac_exact_equality():
    d = KA["injection"]["dirac"]
    inj = O.TemporalInjection.dirac(d["time"], d["amount"])
    for t, v in d["evaluate"]:
        assert inj.evaluate(t) == v
    assert O.TemporalInjection.dirac(10.0, 1.0).evaluate(10.0 + 1e-9) == 0.0


def test_gaussian_values():
    g = KA["injection"]["gaussian"]
    inj = O.TemporalInjection.gaussian(g["center"], g["width"], g["peak"])
    assert abs(inj.evaluate(g["at_center"]["t"]) - g["at_center"]["value"]) < g["at_center"]["abs_tol"]
    for t in g["at_sigma"]["t"]:
        assert abs(inj.evaluate(t) - g["at_sigma"]["value"]) < g["at_sigma"]["abs_tol"]
    for t in g["far"]["t"]:
        assert inj.evaluate(t) < g["far"]["less_than"]
    # the formula itself: peak * exp(((-d)*d)/2.0), d = (t-center)/width
    d = (7.3 - 10.0) / 2.0
    assert inj.evaluate(7.3) == 0.1 * np.exp(((-d) * d) / 2.0)


def test_rectangle_half_open_and_none():
    r = KA["injection"]["rectangle"]
    inj = O.TemporalInjection.rectangle(r["start"], r["end"], r["concentration"])
    for t, v in r["evaluate"]:
        assert inj.evaluate(t) == v
    for t, v in KA["injection"]["none"]["evaluate"]:
        assert O.TemporalInjection.none().evaluate(t) == v


# ---- model quantities ----------------------------------------------------------------------------------------
def test_derived_quantities():
    q = KA["derived_quantities"]
    m = O.LangmuirMulti.new([O.SpeciesParams.new("A", 1.0, 0.5, 1, O.TemporalInjection.dirac(0.0, 1e-3))],
                            q["n_points"], q["porosity"], q["velocity"], q["column_length"])
    for k in ("fe", "ue", "dz", "stationary_fraction"):
        assert abs(getattr(m, k) - q[k]) <= q["epsilon"] * abs(q[k])
    s = O.LangmuirSingle.new(1.2, 0.4, 2.0, q["porosity"], q["velocity"], q["column_length"], q["n_points"],
                             O.TemporalInjection.none())
    assert (s.fe, s.ue, s.dz) == (m.fe, m.ue, m.dz)


def test_jacobian_values_at_zero():
    for c in KA["jacobian_at_zero"]["cases"]:
        m = O.LangmuirMulti.new([O.SpeciesParams.new("T", c["lambda"], c["K"], c["N"], O.TemporalInjection.none())],
                                50, c["porosity"], 0.001, 0.25)
        assert abs(m.jacobian([0.0])[0, 0] - c["M00"]) <= KA["jacobian_at_zero"]["epsilon"] * c["M00"]


def two_species_model():
    inj = O.TemporalInjection.dirac(0.0, 1e-3)
    sp = [O.SpeciesParams.new(n, l, k, p, inj) for n, l, k, p in KA["inverse_times_original"]["species"]]
    return O.LangmuirMulti.new(sp, 100, 0.4, 0.001, 0.25)


def test_jacobian_sign_structure_and_finiteness():
    """langmuir_multi.rs:1391-1428: diagonal > 0, off-diagonal <= 0, all finite, shapes."""
    m = two_species_model()
    j = m.jacobian([1e-4, 1e-4])
    assert j.shape == (2, 2) and j[0, 0] > 0 and j[1, 1] > 0 and j[0, 1] <= 0 and j[1, 0] <= 0
    assert np.isfinite(m.jacobian([1e-6, 2e-6])).all()


def test_inverse_times_original_is_identity():
    k = KA["inverse_times_original"]
    m = two_species_model()
    mat = np.eye(2) + m.fe * m.jacobian(k["c"])
    inv = m.inverse_propagation(k["c"])
    assert inv is not None and np.abs(mat @ inv - np.eye(2)).max() < k["abs_tol"]


def test_single_rhs_inlet_effect():
    k = KA["single_rhs"]
    m = O.LangmuirSingle.new(*k["model"], O.TemporalInjection.gaussian(*k["gaussian"]))
    z = np.zeros(100)
    d10, d0 = m.compute_physics(z, 10.0), m.compute_physics(z, 0.0)
    assert abs(d10[0]) > 0 and abs(d0[0]) < k["inlet_abs_below_at_t0"] and abs(d10[0]) > abs(d0[0])
    assert not d10[1:].any()


def test_multi_rhs_interior_zero_on_empty_column():
    """langmuir_multi.rs:1592-1613."""
    m = two_species_model()
    d = m.compute_physics(np.zeros((100, 2)), 0.0)
    assert np.isfinite(d).all() and np.abs(d[1:]).max() < 1e-15 and (d[0] > 0).all()
    # the rayon path (cells across threads) gives the same bits as the serial path
    big = O.LangmuirMulti.new(two_species_model().species, 600, 0.4, 0.001, 0.25)
    c = np.random.default_rng(0).uniform(0, 0.01, (600, 2))
    assert np.array_equal(big.compute_physics(c, 1.0), big.compute_physics(c, 1.0, cells_parallel=True))


# ---- nalgebra restatement -------------------------------------------------------------------------------------
@pytest.mark.parametrize("n", [1, 2, 3, 4, 5, 6, 8, 12, 20])
def test_try_inverse_against_numpy(n):
    rng = np.random.default_rng(n)
    for _ in range(20):
        a = rng.standard_normal((n, n)) + (2.0 if _ % 2 else 0.0) * np.eye(n)
        inv = O.try_inverse(a)
        assert inv is not None
        assert np.abs(inv - np.linalg.inv(a)).max() <= 1e-10 * np.linalg.cond(a)
        assert np.abs(a @ inv - np.eye(n)).max() <= 1e-12 * np.linalg.cond(a)


@pytest.mark.parametrize("n", [1, 2, 3, 4, 5, 7])
def test_try_inverse_singular_is_none(n):
    assert O.try_inverse(np.zeros((n, n))) is None
    if n >= 2:
        a = np.arange(1.0, n * n + 1).reshape(n, n)
        a[1] = 2.0 * a[0]  # exactly dependent rows
        inv = O.try_inverse(a)
        # nalgebra returns None only on an exact zero determinant / pivot; rounding may hide it for n >= 3
        assert inv is None or not np.isfinite(inv).all() or np.abs(inv).max() > 1e12


def test_lu_pivots_rows():
    a = np.array([[1e-20, 1.0, 0, 0, 0], [1.0, 1.0, 0, 0, 0], [0, 0, 1.0, 0, 0], [0, 0, 0, 0.0, 2.0], [0, 0, 0, 3.0, 1.0]])
    inv = O.try_inverse(a)
    assert np.abs(a @ inv - np.eye(5)).max() < 1e-15


def test_gemv_column_order():
    a = np.array([[1e16, 1.0, -1e16], [3.0, 4.0, 5.0], [0.5, 0.25, 0.125]])
    x = np.array([1.0, 1.0, 1.0])
    y = O.gemv(a, x)
    assert y[0] == (1e16 * 1.0 + 1.0) + (-1e16)  # accumulated j = 0, 1, 2 in that order (= 0.0, not 1.0)
    assert y[1] == 12.0


@pytest.mark.parametrize("n", [2, 3, 4, 5, 8])
def test_multi_row_against_sherman_morrison(n):
    """I + Fe*M = diag(a) - p q^T, so dv follows in O(n) without any factorisation."""
    rng = np.random.default_rng(n)
    inj = O.TemporalInjection.none()
    sp = [O.SpeciesParams.new(f"S{k}", rng.uniform(0.8, 1.5), rng.uniform(0.1, 0.8), int(rng.integers(1, 4)), inj)
          for k in range(n)]
    m = O.LangmuirMulti.new(sp, 64, 0.4, 1e-3, 0.25)
    c = rng.uniform(0.0, 0.5, (64, n))
    got = m.compute_physics(c, 0.0)
    K = np.array([s.langmuir_k for s in sp])
    lam = np.array([s.lambda_ for s in sp])
    nbar = m.stationary_fraction * np.array([float(s.port_number) for s in sp])
    up = np.vstack([np.zeros((1, n)), c[:-1]])
    grad = (c - up) / m.dz
    D = 1.0 + c @ K
    a = 1.0 + m.fe * (lam[None, :] + nbar * K / D[:, None])
    p = m.fe * nbar * K * c / (D ** 2)[:, None]
    ainv_g = grad / a
    ainv_p = p / a
    dv = ainv_g + ainv_p * ((ainv_g @ K) / (1.0 - ainv_p @ K))[:, None]
    want = -m.ue * dv
    assert np.abs(got - want).max() <= 1e-12 * np.abs(want).max()


# ---- integrators on mock ODEs: rk4.rs:361-1103, euler.rs:299-977 ----------------------------------------
def test_time_grid_and_trajectory_length():
    tp = O.time_points(10.0, 100)
    assert len(tp) == 101 and tp[0] == 0.0 and tp[-1] == 10.0
    assert np.array_equal(tp[1:], (np.arange(100) + 1.0) * (10.0 / 100.0))
    traj, final, st = O.solve_linear_ode(O.RK4, -0.1, 0.0, 10.0, 100, [1.0, 2.0])
    assert traj.shape == (101, 2) and st == 0 and np.array_equal(traj[-1], final) and np.array_equal(traj[0], [1.0, 2.0])


def test_exponential_decay_and_constant_growth():
    _, f, _ = O.solve_linear_ode(O.RK4, -0.1, 0.0, 10.0, 100, [1.0])
    assert abs(f[0] - np.exp(-1.0)) < 1e-8
    _, f, _ = O.solve_linear_ode(O.EULER, -0.1, 0.0, 10.0, 1000, [1.0])
    assert abs(f[0] - np.exp(-1.0)) < 1e-3
    for method in (O.RK4, O.EULER):  # dy/dt = 2: both exact
        _, f, _ = O.solve_linear_ode(method, 0.0, 2.0, 5.0, 50, [1.0])
        assert abs(f[0] - 11.0) < 1e-12


def test_convergence_order():
    k = KA["integrator_order"]
    def err(method, steps):
        _, f, _ = O.solve_linear_ode(method, -0.1, 0.0, 10.0, steps, [1.0])
        return abs(f[0] - np.exp(-1.0))
    r = err(O.RK4, 10) / err(O.RK4, 20)
    assert k["rk4_error_ratio_when_halving_dt"][0] <= r <= k["rk4_error_ratio_when_halving_dt"][1]
    r = err(O.EULER, 100) / err(O.EULER, 200)
    assert k["euler_error_ratio_when_halving_dt"][0] <= r <= k["euler_error_ratio_when_halving_dt"][1]


def test_validate_state_nan_and_inf():
    """solver/mod.rs:514-553 through a mock that overflows: dy/dt = 1e300*y."""
    _, _, st = O.solve_linear_ode(O.EULER, 1e300, 0.0, 10.0, 10, [1.0])
    assert st == -2  # 1e300 after step 1, Inf after step 2 (no NaN yet)
    assert O.validate_state_message(st).startswith("Infinity detected in Concentration at step 2.")
    # alternating overflow: the stage slopes are +-Inf and cancel to NaN inside the first step
    _, _, st = O.solve_linear_ode(O.RK4, -1e200, 0.0, 10.0, 10, [1.0])
    assert st == 1 and O.validate_state_message(st).startswith("NaN detected in Concentration at step 1.")


# ---- the validation suite: validation/main.rs ----------------------------------------------------------------
def linear_tfa(c0):
    k = KA["linear_tfa"]
    t_inj = 2.0 * k["total_time"] / k["steps"]
    return O.LangmuirSingle.new(k["lambda"], k["K"], k["N"], k["porosity"], k["velocity"], k["column_length"],
                                k["n_points"], O.TemporalInjection.rectangle(0.0, t_inj, c0)), t_inj


def test_linear_tfa_validation():
    k = KA["linear_tfa"]
    m, t_inj = linear_tfa(k["c0"])
    dead = k["column_length"] / m.ue
    assert abs(dead - k["dead_time"]) < 1e-6 and abs(dead * (1.0 + m.fe * k["linear_ka"]) - k["t_retention"]) < 1e-6
    r = O.solve_single(m, O.RK4, k["total_time"], k["steps"])
    e = O.solve_single(m, O.EULER, k["total_time"], k["steps"])
    tp = r.time_points
    assert abs(O.first_moment(tp, r.outlet) - k["t_retention"]) / k["t_retention"] < k["t_retention_rel_tol"]
    assert abs(O.peak_sigma(tp, r.outlet) - k["sigma"]) / k["sigma"] < k["sigma_rel_tol"]
    assert O.trapezoid(tp, r.outlet) / (k["c0"] * t_inj) >= k["mass_recovery_min"]
    rsf = O.rsf(e.time_points, e.outlet, tp, r.outlet)
    assert rsf < k["rsf_euler_rk4_max"] and round(rsf, 3) == k["rsf_published"]
    s = KA["survey_sanity_values"]["linear_tfa"]
    assert abs(O.first_moment(tp, r.outlet) - s["t_retention"]) < 1e-3 and abs(rsf - s["rsf"]) < 1e-5
    assert abs(O.peak_sigma(tp, r.outlet) - s["sigma"]) < 1e-3
    assert abs(O.trapezoid(tp, r.outlet) / (k["c0"] * t_inj) - s["recovery"]) < 1e-5


def test_nonlinear_tfa_validation():
    k, nl = KA["linear_tfa"], KA["nonlinear_tfa"]
    modes = []
    for c0 in (k["c0"], nl["c0"]):
        m, t_inj = linear_tfa(c0)
        r = O.solve_single(m, O.RK4, k["total_time"], k["steps"])
        modes.append(r.time_points[int(np.argmax(r.outlet))])
        if c0 == nl["c0"]:
            assert O.trapezoid(r.time_points, r.outlet) / (c0 * t_inj) >= nl["mass_recovery_min"]
    assert modes[1] < modes[0]


@pytest.mark.parametrize("case", ["ascorbic_erythorbic", "glucose_fructose_linear"])
def test_multi_species_validation(case):
    k = KA[case]
    inj = O.TemporalInjection.rectangle(0.0, 2.0 * k["total_time"] / k["steps"], k["c0"])
    sp = [O.SpeciesParams.new(n, l, kk, p, inj) for n, l, kk, p, _ in k["species"]]
    m = O.LangmuirMulti.new(sp, k["n_points"], k["porosity"], k["velocity"], k["column_length"])
    r = O.solve_multi(m, O.RK4, k["total_time"], k["steps"])
    assert r.status == 0 and r.n_singular == 0
    t_r = [O.first_moment(r.time_points, r.outlet[i]) for i in range(2)]
    for got, (_, _, _, _, want) in zip(t_r, k["species"]):
        assert abs(got - want) / want < k["rel_tol"]
    if "gap" in k:
        assert abs((t_r[1] - t_r[0]) - k["gap"]) / k["gap"] < k["rel_tol"]
        s = KA["survey_sanity_values"]["ascorbic_erythorbic_t_retention"]
        assert abs(t_r[0] - s[0]) < 1e-3 and abs(t_r[1] - s[1]) < 1e-3


def test_survey_sanity_values_c1_c2():
    s = KA["survey_sanity_values"]
    m = O.LangmuirSingle.new(1.2, 0.4, 2.0, 0.4, 1e-3, 0.25, 100, O.TemporalInjection.gaussian(10.0, 3.0, 0.1))
    r = O.solve_single(m, O.RK4, 600.0, 10_000)
    i = int(np.argmax(r.outlet))
    assert abs(r.outlet[i] - s["c1_rk4"]["peak"]) < 1e-11 and abs(r.time_points[i] - s["c1_rk4"]["t_peak"]) < 1e-9
    assert abs(r.outlet[-1] - s["c1_rk4"]["final_outlet"]) < 1e-12
    inj = O.TemporalInjection.gaussian(10.0, 3.0, 0.1)
    mm = O.LangmuirMulti.new([O.SpeciesParams.new("Ascorbic", 1.0, 1.1, 2, inj),
                              O.SpeciesParams.new("Erythorbic", 1.0, 1.7, 2, inj)], 100, 0.4, 1e-3, 0.25)
    rm = O.solve_multi(mm, O.RK4, 1200.0, 20_000)
    for k, (peak, t_peak) in enumerate(s["c2_rk4_peaks"]):
        i = int(np.argmax(rm.outlet[k]))
        assert abs(rm.outlet[k][i] - peak) < 1e-9 and abs(rm.time_points[i] - t_peak) < 1e-9


def test_oracle_regression_vectors():
    """Oracle-generated vectors (tests/golden/make_oracle_vectors.py): guards the oracle against drift.
    They are NOT reference outputs (the Rust reference cannot run here)."""
    path = os.path.join(os.path.dirname(__file__), "golden", "oracle_vectors.npz")
    g = np.load(path)
    from golden.make_oracle_vectors import cases
    for name, arr in cases().items():
        assert np.array_equal(arr.view(np.uint64), g[name].view(np.uint64)), name


def test_mock_ode_values():
    """rk4.rs:589-717 / euler.rs:475-548: the integrators on the reference's mock models (dy/dt = a y + b)."""
    k = KA["mock_ode_values"]
    g = k["constant_growth"]
    for method in (O.RK4, O.EULER):
        _, f, st = O.solve_linear_ode(method, 0.0, g["rate"], g["total_time"], g["steps"], [0.0])
        assert st == 0 and abs(f[0] - g["expected_final"]) < g["abs_tol"]
    for key, method in (("exponential_decay_rk4", O.RK4), ("exponential_decay_euler", O.EULER)):
        d = k[key]
        _, f, st = O.solve_linear_ode(method, d["rate"], 0.0, d["total_time"], d["steps"], [1.0])
        assert st == 0 and abs(f[0] - np.exp(d["rate"] * d["total_time"])) < d["abs_tol"]
    c = k["rk4_convergence"]
    errs = [abs(O.solve_linear_ode(O.RK4, -0.1, 0.0, c["total_time"], s, [1.0])[1][0] - np.exp(-0.1 * c["total_time"]))
            for s in c["steps"]]
    for a, b in zip(errs, errs[1:]):
        assert c["ratio_min"] < a / b < c["ratio_max"]
