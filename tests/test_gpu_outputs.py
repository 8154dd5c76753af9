"""GPU parity of the OUTPUT side of the C ABI in every kernel family:

* the on-device summary [area, first moment, max, time of max, sigma, min] against the reference's own analytics
  (validation/dissimilarity.rs:81-88 trapezoid, validation/main.rs:190-206 first_moment / peak_sigma) evaluated on
  the oracle's full-rate series — 1e-7 (north_star's tolerance for retention time and peak area);
* explicit sample lists — the reference's sample_indices (physics/traits.rs:1010-1023) — for the outlet series and
  the profile snapshots: bit-identical to the corresponding rows of the oracle trajectory in EXACT mode;
* a detector that is not at the outlet (Detector::node_index, domain/detector.rs:285);
* Rsf between two device-resident series (validation/dissimilarity.rs:173-203; the Euler-vs-RK4 criterion of
  validation/main.rs:430-440 with its published value 0.016);
* a column at rest (ue = 0) in FAST mode; a batch that cannot fill the machine (split launch)."""
import numpy as np
import pytest

import chromatography_b200 as cb
from chromatography_b200 import _ffi, workloads as W
from oracle import oracle as O
from helpers import MOMENT_RTOL, PROFILE_RTOL, assert_bit_equal, max_rel_err, to_oracle

pytestmark = pytest.mark.gpu

WANT_ALL = cb.WANT_OUTLET | cb.WANT_SNAPSHOTS | cb.WANT_FINAL | cb.WANT_SUMMARY | cb.WANT_STATUS


def reference_summary(tp, series):
    """What the reference's analytics give for one detector signal."""
    area = O.trapezoid(tp, series)
    i = int(np.argmax(series))
    return np.array([area, O.first_moment(tp, series), series[i], tp[i], O.peak_sigma(tp, series), np.min(series)])


def check_summary(got, tp, series, what):
    want = reference_summary(tp, series)
    for k, name in enumerate(["area", "first moment", "max", "time of max", "sigma", "min"]):
        tol = MOMENT_RTOL * max(abs(want[k]), 1e-300) if k != 5 else 1e-9 * np.max(np.abs(series)) + 1e-300
        assert abs(got[k] - want[k]) <= tol, (what, name, got[k], want[k])


def detector_series(ref, m):
    """The oracle trajectory at the model's detector cell (ABI: detector_cell 0 = outlet, k = cell k - 1)."""
    d = getattr(m, "detector_cell", 0)
    return ref.trajectory[:, d - 1 if d else m.n_points - 1]


def long_column(nz, cfl=0.5):
    m = W.tfa_single(nz=nz)
    m.dz = 0.25 / 100.0
    n_bar = m.fe / (1.0 + m.fe) * m.port_number
    u_eff0 = m.ue / (1.0 + m.fe * (m.lambda_ + n_bar * m.langmuir_k))
    return m, cfl * m.dz / u_eff0


def single_case(kind):
    """(models, total_time, steps, expected kernel family) of a LangmuirSingle case per kernel family."""
    if kind == "resident":
        return W.tfa_sweep_models(5), 600.0, 3000, "single_resident<"
    if kind == "resident_mw":
        ms = W.tfa_sweep_models(3, nz=1500)   # the peak must elute: full horizon, CFL ~ 0.4 on the finer grid
        return ms, 600.0, 6000, "single_resident_mw<"
    m, dt = long_column(9000)
    steps = 160
    m.set_injection(cb.TemporalInjection.gaussian(0.3 * dt * steps, 0.1 * dt * steps, 0.1))
    m.detector_cell = 40 + 1  # 160 steps at CFL 0.5 move the front ~80 cells: the detector sits where the peak passes
    return [m], dt * steps, steps, "single_stream<"


def run_single(ctx, models, T, steps, arith=cb.ARITH_FAST, method=cb.RK4, want=WANT_ALL, **cfg_kw):
    kw = dict(outlet_stride=1, snapshot_stride=0)
    kw.update(cfg_kw)
    cfg = cb.make_cfg(method, T, steps, models[0].n_points, 1, arith=arith, **kw)
    if not (kw["snapshot_stride"] or kw.get("snapshot_steps") is not None):
        want &= ~cb.WANT_SNAPSHOTS
    cols, tables = cb.pack_single(models, method, T, steps)
    plan = cb.Plan(ctx, cfg, cols, tables, want=want)
    desc = plan.describe()
    plan.execute()
    r = plan.download()
    plan.close()
    return r, desc


def multi_case(kind):
    if kind == "reg":      # first-generation register kernel (n <= 4)
        return [W.acids_multi()], 600.0, 3000, "multi_reg<"
    if kind == "reg2":     # second generation (forced with CHROM_REG2 = 2 by the tests: see force_generation)
        return W.multi_batch_models(2, 6, 64, seed=5), 300.0, 500, "multi_reg2<"
    if kind == "reg3":     # third generation: the default for n >= 5
        return W.multi_batch_models(2, 6, 64, seed=5), 300.0, 500, "multi_reg3<"
    if kind == "resident":  # shared-memory resident (EXACT arithmetic)
        return W.multi_batch_models(2, 3, 96, seed=6), 300.0, 500, "multi_resident<"
    return W.multi_batch_models(1, 20, 40, seed=9), 100.0, 120, "multi_tiled<"  # warp per cell


def force_generation(monkeypatch, kind):
    """The register-resident kernel generation is chosen when the plan is built (csrc/multi_reg.cu)."""
    if kind == "reg2":
        monkeypatch.setenv("CHROM_REG2", "2")


def run_multi(ctx, models, T, steps, arith=cb.ARITH_FAST, want=WANT_ALL, **cfg_kw):
    kw = dict(outlet_stride=1, snapshot_stride=0)
    kw.update(cfg_kw)
    n = models[0].n_species()
    cfg = cb.make_cfg(cb.RK4, T, steps, models[0].n_points, n, arith=arith, **kw)
    if not (kw["snapshot_stride"] or kw.get("snapshot_steps") is not None):
        want &= ~cb.WANT_SNAPSHOTS
    cols, species, tables = cb.pack_multi(models, cb.RK4, T, steps)
    plan = cb.Plan(ctx, cfg, cols, tables, species=species, want=want)
    desc = plan.describe()
    plan.execute()
    r = plan.download()
    plan.close()
    return r, desc


# ---- summary ----------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("kind", ["resident", "resident_mw", "stream"])
def test_summary_single(ctx, kind):
    models, T, steps, family = single_case(kind)
    # the summary must not depend on how (or whether) the outlet series itself is sampled
    r, desc = run_single(ctx, models, T, steps, outlet_stride=7)
    assert family in desc and not r.status.any()
    tp = cb.time_points(T, steps)
    for c, m in enumerate(models):
        ref = O.solve_single(to_oracle(m), O.RK4, T, steps, want_trajectory=True)
        series = detector_series(ref, m)
        check_summary(r.summary[c], tp, series, f"{kind} column {c}")
        assert max_rel_err(r.outlet[c], series[::7]) <= PROFILE_RTOL


@pytest.mark.parametrize("kind", ["reg", "reg2", "reg3", "resident", "tiled"])
def test_summary_multi(ctx, kind, monkeypatch):
    force_generation(monkeypatch, kind)
    models, T, steps, family = multi_case(kind)
    arith = cb.ARITH_EXACT if kind == "resident" else cb.ARITH_FAST
    r, desc = run_multi(ctx, models, T, steps, arith=arith)
    assert family in desc and not r.status.any()
    tp = cb.time_points(T, steps)
    for c, m in enumerate(models):
        ref = O.solve_multi(to_oracle(m), O.RK4, T, steps)
        for k in range(m.n_species()):
            check_summary(r.summary[c][k], tp, ref.outlet[k], f"{kind} column {c} species {k}")


def test_summary_through_the_one_shot_call(ctx):
    """chrom_cuda_solve_single_batch with summary + strided outlet: the sweep-mode call of bench.py."""
    models = W.tfa_sweep_models(40)
    T, steps = 600.0, 2000
    cfg = cb.make_cfg(cb.RK4, T, steps, 100, 1, outlet_stride=50)
    cols, tables = cb.pack_single(models, cb.RK4, T, steps)
    r = cb.solve_single_batch(ctx, cfg, cols, tables, want_summary=True)
    assert r.summary.shape == (40, cb.SUMMARY_LEN) and r.outlet.shape == (40, steps // 50 + 1)
    tp = cb.time_points(T, steps)
    for c in (0, 17, 39):
        ref = O.solve_single(to_oracle(models[c]), O.RK4, T, steps)
        check_summary(r.summary[c], tp, ref.outlet, f"column {c}")
        assert max_rel_err(r.outlet[c], ref.outlet[::50]) <= PROFILE_RTOL


# ---- explicit sample lists ------------------------------------------------------------------------------------------
def test_sample_indices_is_the_reference_rule():
    assert cb.sample_indices(100, 5).tolist() == [0, 24, 49, 74, 99]   # physics/traits.rs:1003-1006 (doc test)
    assert cb.sample_indices(4, None).tolist() == [0, 1, 2, 3]


@pytest.mark.parametrize("kind", ["resident", "resident_mw", "stream"])
def test_sample_lists_single(ctx, kind):
    models, T, steps, family = single_case(kind)
    out_steps = cb.sample_indices(steps + 1, 37)          # includes step 0 and the last step
    snap_steps = np.array([8, 16, steps // 2 // 8 * 8, steps], dtype=np.uint64)  # no initial state, ends on the last step
    r, desc = run_single(ctx, models, T, steps, arith=cb.ARITH_EXACT, outlet_steps=out_steps, snapshot_steps=snap_steps)
    assert family in desc and not r.status.any()
    assert r.outlet.shape == (len(models), 37) and r.snapshots.shape == (len(models), 4, models[0].n_points)
    for c, m in enumerate(models):
        ref = O.solve_single(to_oracle(m), O.RK4, T, steps, want_trajectory=True)
        assert_bit_equal(r.outlet[c], detector_series(ref, m)[out_steps.astype(int)], f"{kind} outlet list")
        assert_bit_equal(r.snapshots[c], ref.trajectory[snap_steps.astype(int)], f"{kind} snapshot list")


@pytest.mark.parametrize("kind", ["reg", "reg2", "reg3", "resident", "tiled"])
def test_sample_lists_multi(ctx, kind, monkeypatch):
    force_generation(monkeypatch, kind)
    models, T, steps, family = multi_case(kind)
    out_steps = cb.sample_indices(steps + 1, 23)
    snap_steps = np.array([0, 3, steps - 1], dtype=np.uint64)
    arith = cb.ARITH_FAST if kind in ("reg", "reg2", "reg3") else cb.ARITH_EXACT
    r, desc = run_multi(ctx, models, T, steps, arith=arith, outlet_steps=out_steps, snapshot_steps=snap_steps)
    if arith == cb.ARITH_FAST:
        assert family in desc
    assert not r.status.any()
    for c, m in enumerate(models):
        ref = O.solve_multi(to_oracle(m), O.RK4, T, steps, want_trajectory=True)
        want_out = ref.outlet[:, out_steps.astype(int)]
        want_snap = ref.trajectory[snap_steps.astype(int)].transpose(0, 2, 1)
        if arith == cb.ARITH_EXACT:
            assert_bit_equal(r.outlet[c], want_out, f"{kind} outlet list")
            assert_bit_equal(r.snapshots[c], want_snap, f"{kind} snapshot list")
        else:
            assert max_rel_err(r.outlet[c], want_out) <= PROFILE_RTOL
            assert max_rel_err(r.snapshots[c], want_snap) <= PROFILE_RTOL


def test_bad_sample_lists_are_rejected(ctx):
    m = W.tfa_single()
    cols, tables = cb.pack_single([m], cb.RK4, 6.0, 100)
    for bad in ([5, 5, 9], [9, 5], [0, 101]):
        cfg = cb.make_cfg(cb.RK4, 6.0, 100, 100, outlet_steps=np.array(bad, dtype=np.uint64))
        with pytest.raises(cb.SolverError, match="strictly increasing"):
            cb.solve_single_batch(ctx, cfg, cols, tables)


# ---- detector position ----------------------------------------------------------------------------------------------
def test_node_index_is_the_reference_rule():
    assert cb.node_index(0.25, 0.25, 100) == 99      # Detector::outlet(): domain/detector.rs:278 (doc test)
    assert cb.node_index(0.125, 0.25, 100) == 50     # Relative(0.5): domain/detector.rs:281-282 (doc test)
    assert cb.node_index(0.0, 0.25, 100) == 0 and cb.node_index(1.0, 0.25, 100) == 99


@pytest.mark.parametrize("kind", ["resident", "resident_mw", "stream"])
def test_detector_cell_single(ctx, kind):
    models, T, steps, family = single_case(kind)
    nz = models[0].n_points
    # mid-column by the reference's rule (the long streaming column is only 160 steps old: stay where the peak is)
    cells = [min(cb.node_index(0.5 * nz * models[0].dz, nz * models[0].dz, nz), 64), 0, nz - 2, 7, 33][:len(models)]
    for m, cell in zip(models, cells):
        m.detector_cell = cell + 1    # ABI: 0 = outlet, k = cell k - 1
    r, desc = run_single(ctx, models, T, steps, arith=cb.ARITH_EXACT)
    assert family in desc
    tp = cb.time_points(T, steps)
    for c, (m, cell) in enumerate(zip(models, cells)):
        ref = O.solve_single(to_oracle(m), O.RK4, T, steps, want_trajectory=True)
        series = ref.trajectory[:, cell]
        assert_bit_equal(r.outlet[c], series, f"{kind} detector at cell {cell}")
        check_summary(r.summary[c], tp, series, f"{kind} detector at cell {cell}")


@pytest.mark.parametrize("kind", ["reg", "reg2", "reg3", "resident", "tiled"])
def test_detector_cell_multi(ctx, kind, monkeypatch):
    force_generation(monkeypatch, kind)
    models, T, steps, family = multi_case(kind)
    nz = models[0].n_points
    cells = [nz // 2, 1][:len(models)]
    for m, cell in zip(models, cells):
        m.detector_cell = cell + 1
    arith = cb.ARITH_FAST if kind in ("reg", "reg2", "reg3") else cb.ARITH_EXACT
    r, desc = run_multi(ctx, models, T, steps, arith=arith)
    tp = cb.time_points(T, steps)
    for c, (m, cell) in enumerate(zip(models, cells)):
        ref = O.solve_multi(to_oracle(m), O.RK4, T, steps, want_trajectory=True)
        series = ref.trajectory[:, cell, :].T    # [species][step]
        if arith == cb.ARITH_EXACT:
            assert_bit_equal(r.outlet[c], series, f"{kind} detector at cell {cell}")
        else:
            assert max_rel_err(r.outlet[c], series) <= PROFILE_RTOL
        for k in range(m.n_species()):
            check_summary(r.summary[c][k], tp, series[k], f"{kind} species {k}")


def test_detector_beyond_the_column_is_rejected(ctx):
    m = W.tfa_single()
    m.detector_cell = 101 + 1
    cols, tables = cb.pack_single([m], cb.RK4, 6.0, 100)
    with pytest.raises(cb.SolverError, match="detector_cell"):
        cb.solve_single_batch(ctx, cb.make_cfg(cb.RK4, 6.0, 100, 100), cols, tables)


# ---- Rsf on the device -----------------------------------------------------------------------------------------------
def test_rsf_euler_vs_rk4_on_device(ctx):
    """validation/main.rs:430-440: Rsf(Euler, RK4) < 0.05 on the linear TFA case (CHANGELOG 0.4.0 publishes 0.016);
    here for a batch, without either series leaving the device."""
    T, steps = 900.0, 4500                                                   # validation/reference.rs:34-70
    inj = cb.TemporalInjection.rectangle(0.0, 2.0 * T / steps, 1e-3)
    base = cb.LangmuirSingle.new(1.0, 0.5, 6.0, 0.4, 1e-3, 0.3, 100, inj)    # the linear TFA reference case
    models = [base] + W.tfa_sweep_models(6, injection=inj)
    plans = []
    for method in (cb.EULER, cb.RK4):
        cfg = cb.make_cfg(method, T, steps, 100, 1, outlet_stride=2)
        cols, tables = cb.pack_single(models, method, T, steps)
        p = cb.Plan(ctx, cfg, cols, tables, want=cb.WANT_OUTLET | cb.WANT_STATUS)
        p.execute()
        plans.append(p)
    got = plans[0].rsf(plans[1])
    tp = cb.time_points(T, steps)[::2]
    for c, m in enumerate(models):
        e = O.solve_single(to_oracle(m), O.EULER, T, steps).outlet[::2]
        r = O.solve_single(to_oracle(m), O.RK4, T, steps).outlet[::2]
        want = O.rsf(tp, e, tp, r)
        assert abs(got[c] - want) <= MOMENT_RTOL * want, (c, got[c], want)
    assert got[0] < 0.05 and abs(got[0] - 0.016) < 0.002
    for p in plans:
        p.close()


def test_rsf_multi_and_mismatched_plans(ctx):
    models = W.multi_batch_models(3, 2, 64, seed=3)
    T, steps = 300.0, 600
    cols, species, tables = cb.pack_multi(models, cb.RK4, T, steps)
    a = cb.Plan(ctx, cb.make_cfg(cb.RK4, T, steps, 64, 2, arith=cb.ARITH_FAST), cols, tables, species=species)
    b = cb.Plan(ctx, cb.make_cfg(cb.RK4, T, steps, 64, 2, arith=cb.ARITH_EXACT), cols, tables, species=species)
    a.execute()
    b.execute()
    same = a.rsf(b)
    assert same.shape == (3, 2) and np.all(same < 1e-9)      # FAST vs EXACT: the same chromatogram
    c = cb.Plan(ctx, cb.make_cfg(cb.RK4, T, steps, 64, 2, outlet_stride=2), cols, tables, species=species)
    c.execute()
    with pytest.raises(cb.SolverError, match="same columns, species and outlet sampling"):
        a.rsf(c)
    for p in (a, b, c):
        p.close()


# ---- FAST arithmetic on a column at rest ---------------------------------------------------------------------------
def test_column_at_rest_fast_equals_reference(ctx):
    """ue = 0: the reference computes -sigma * 0 * dc/dz = 0 and the state stays as it is; the FAST forms fold
    -ue/dz into a denominator, so such a batch must run the reference's operation order."""
    models = W.tfa_sweep_models(3)
    models[1].ue = 0.0
    y0 = np.random.default_rng(1).uniform(0.0, 0.05, size=(3, 100))
    cfg = cb.make_cfg(cb.RK4, 6.0, 100, 100, arith=cb.ARITH_FAST)
    cols, tables = cb.pack_single(models, cb.RK4, 6.0, 100)
    r = cb.solve_single_batch(ctx, cfg, cols, tables, y0=y0)
    assert not r.status.any()
    assert_bit_equal(r.final_state[1], y0[1], "column at rest")
    for c in (0, 2):
        ref = O.solve_single(to_oracle(models[c]), O.RK4, 6.0, 100, y0=y0[c])
        assert max_rel_err(r.final_state[c], ref.final_state) <= PROFILE_RTOL


# ---- a batch that cannot fill the machine: split launch --------------------------------------------------------------
def test_split_launch_is_bit_identical_to_the_single_launch(ctx, monkeypatch):
    """Between one and two warps per SM sub-partition the batch runs as two launches with different cells per lane
    (single_resident_choose).  Chunking never enters the per-cell arithmetic, so every output must be bit-identical
    to the unsplit launch, and both must match the oracle."""
    n_cols, T, steps = 8192, 9.0, 150     # 1.73 warps per sub-partition of a 148-SM part at 25 cells per lane
    models = W.tfa_sweep_models(n_cols)
    r, desc = run_single(ctx, models, T, steps, outlet_stride=3)
    monkeypatch.setenv("CHROM_NO_SPLIT", "1")
    r1, desc1 = run_single(ctx, models, T, steps, outlet_stride=3)
    assert "split launch" in desc and "split launch" not in desc1, (desc, desc1)
    seam = int(desc.split("split launch: ")[1].split(" columns")[0])
    assert 0 < seam < n_cols
    assert not r.status.any()
    assert_bit_equal(r.outlet, r1.outlet, "outlet")
    assert_bit_equal(r.final_state, r1.final_state, "final state")
    assert_bit_equal(r.summary, r1.summary, "summary")
    for c in (0, seam - 1, seam, seam + 1, n_cols - 1):     # both sides of the seam between the two launches
        ref = O.solve_single(to_oracle(models[c]), O.RK4, T, steps)
        assert max_rel_err(r.final_state[c], ref.final_state) <= PROFILE_RTOL
        assert max_rel_err(r.outlet[c], ref.outlet[::3]) <= PROFILE_RTOL
