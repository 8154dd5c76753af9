"""GPU parity of the tiled LangmuirMulti path (multi_tiled.cu): columns whose 4*n*nz doubles exceed one
CTA's shared memory are cut into z-tiles with a recomputed upstream halo, and n > 64 runs one warp per cell.

Same bar as test_gpu_multi.py: EXACT arithmetic bit-identical to the oracle (nalgebra inverse + gemv order),
FAST within 1e-9.  Small shapes are pushed onto the tiled kernels with the library's environment knobs
(CHROM_MULTI_TILED / CHROM_TILED_TW / CHROM_TILED_FUSE / CHROM_TILED_WARP) so that the oracle finishes in seconds;
the natural triggers (nz too long, n > 64) are covered at the smallest sizes that reach them.
"""
import numpy as np
import pytest

import chromatography_b200 as cb
from oracle import oracle as O
from helpers import MOMENT_RTOL, PROFILE_RTOL, assert_bit_equal, max_rel_err, to_oracle
from test_gpu_multi import pivoting_case, random_species, run_gpu

pytestmark = pytest.mark.gpu


def describe(ctx, models, method, T, steps, arith):
    n = models[0].n_species()
    cfg = cb.make_cfg(method, T, steps, models[0].n_points, n, 1, 0, arith)
    cols, species, tables = cb.pack_multi(models, method, T, steps)
    plan = cb.Plan(ctx, cfg, cols, tables, species=species, want=cb.WANT_FINAL)
    try:
        return plan.describe()
    finally:
        plan.close()


def force(monkeypatch, tw=None, fuse=None, warp=False):
    monkeypatch.setenv("CHROM_MULTI_TILED", "1")
    if tw:
        monkeypatch.setenv("CHROM_TILED_TW", str(tw))
    if fuse:
        monkeypatch.setenv("CHROM_TILED_FUSE", str(fuse))
    if warp:
        monkeypatch.setenv("CHROM_TILED_WARP", "1")


def case(n, nz, seed, length=None):
    rng = np.random.default_rng(seed)
    y0 = rng.uniform(0.0, 0.05, size=(nz, n))
    m = cb.LangmuirMulti.new(random_species(n, seed=n), nz, 0.4, 1e-3, length or 0.25 * nz / 100.0)
    return m, y0, np.ascontiguousarray(y0.T)[None]


@pytest.mark.parametrize("n", [1, 2, 5, 8, 12])
@pytest.mark.parametrize("method", [cb.RK4, cb.EULER], ids=["rk4", "euler"])
@pytest.mark.parametrize("fuse", [1, 3])
def test_forced_tiles_thread_per_cell(ctx, monkeypatch, n, method, fuse):
    """nz=100 cut into tiles of <= 24 cells, 1 or 3 steps per launch (halo 4 / 12 for RK4): trajectory,
    outlet and final state bit for bit; snapshot stride 7 is no multiple of the fused steps."""
    nz, steps, T = 100, 45, 4.5
    m, y0, y0_dev = case(n, nz, 7 * n + fuse)
    ref = O.solve_multi(to_oracle(m), method, T, steps, y0=y0, want_trajectory=True)
    force(monkeypatch, tw=24 if fuse == 1 else 40, fuse=fuse)
    assert "multi_tiled" in describe(ctx, [m], method, T, steps, cb.ARITH_EXACT)
    assert "tiles/col=1 " not in describe(ctx, [m], method, T, steps, cb.ARITH_EXACT)
    r = run_gpu(ctx, [m, m], method, T, steps, cb.ARITH_EXACT, y0=np.concatenate([y0_dev, y0_dev]), snapshot_stride=7)
    assert not r.status.any() and ref.n_singular == 0
    for c in range(2):
        assert_bit_equal(r.snapshots[c].transpose(0, 2, 1), ref.trajectory[::7], f"n={n} snapshots col {c}")
        assert_bit_equal(r.outlet[c], ref.outlet, f"n={n} outlet col {c}")
        assert_bit_equal(r.final_state[c].T, ref.final_state, f"n={n} final col {c}")
    for arith in (cb.ARITH_FAST, cb.ARITH_FAST_DENSE):
        f = run_gpu(ctx, [m], method, T, steps, arith, y0=y0_dev, want_summary=True)
        assert max_rel_err(f.final_state[0].T, ref.final_state) <= PROFILE_RTOL
        assert max_rel_err(f.outlet[0], ref.outlet) <= PROFILE_RTOL
        tp = ref.time_points
        for k in range(n):
            area = O.trapezoid(tp, ref.outlet[k])
            assert abs(f.summary[0][k][0] - area) <= MOMENT_RTOL * area
            assert abs(f.summary[0][k][1] - O.first_moment(tp, ref.outlet[k])) <= MOMENT_RTOL * tp[-1]


@pytest.mark.parametrize("n", [5, 12, 33, 40])
@pytest.mark.parametrize("tiles", ["one", "many"])
def test_forced_warp_per_cell(ctx, monkeypatch, n, tiles):
    """One warp per cell (1 or 2 species per lane) on shapes the thread-per-cell kernels also cover."""
    nz, steps, T = 40, 24, 2.4
    m, y0, y0_dev = case(n, nz, 11 * n, length=0.1)
    force(monkeypatch, tw=16 if tiles == "many" else None, fuse=2 if tiles == "many" else None, warp=True)
    d = describe(ctx, [m], cb.RK4, T, steps, cb.ARITH_EXACT)
    assert "warp/cell" in d and (("tiles/col=1 " in d) == (tiles == "one")), d
    for method in (cb.RK4, cb.EULER):
        ref = O.solve_multi(to_oracle(m), method, T, steps, y0=y0, want_trajectory=True)
        r = run_gpu(ctx, [m], method, T, steps, cb.ARITH_EXACT, y0=y0_dev, snapshot_stride=8)
        assert r.status[0] == 0 and ref.n_singular == 0
        assert_bit_equal(r.snapshots[0].transpose(0, 2, 1), ref.trajectory[::8], f"n={n} snapshots")
        assert_bit_equal(r.outlet[0], ref.outlet, f"n={n} outlet")
        f = run_gpu(ctx, [m], method, T, steps, cb.ARITH_FAST, y0=y0_dev)
        assert max_rel_err(f.final_state[0].T, ref.final_state) <= PROFILE_RTOL
        assert max_rel_err(f.outlet[0], ref.outlet) <= PROFILE_RTOL


@pytest.mark.parametrize("n,nz,steps", [(70, 30, 10), (100, 100, 6), (130, 24, 6)])
def test_more_than_64_species(ctx, n, nz, steps):
    """n > 64 (the reference benches up to n = 100, nz = 100: benches/langmuir_performance.rs:1367): one warp per
    cell, 4 or 8 species per lane; (100, 100) does not fit one tile either.  No environment knobs."""
    T = 0.1 * steps
    m, y0, y0_dev = case(n, nz, n, length=0.1)
    d = describe(ctx, [m], cb.RK4, T, steps, cb.ARITH_FAST)
    assert "multi_tiled" in d and "warp/cell" in d, d
    ref = O.solve_multi(to_oracle(m), O.RK4, T, steps, y0=y0)
    r = run_gpu(ctx, [m, m], cb.RK4, T, steps, cb.ARITH_EXACT, y0=np.concatenate([y0_dev, y0_dev]))
    assert not r.status.any() and ref.n_singular == 0
    for c in range(2):
        assert_bit_equal(r.final_state[c].T, ref.final_state, f"n={n} final col {c}")
        assert_bit_equal(r.outlet[c], ref.outlet, f"n={n} outlet col {c}")
    f = run_gpu(ctx, [m], cb.RK4, T, steps, cb.ARITH_FAST, y0=y0_dev)
    assert max_rel_err(f.final_state[0].T, ref.final_state) <= PROFILE_RTOL
    assert max_rel_err(f.outlet[0], ref.outlet) <= PROFILE_RTOL


@pytest.mark.parametrize("n", [5, 12, 40])
def test_warp_per_cell_pivoting(ctx, monkeypatch, n):
    """Cells whose LU must exchange rows: the scan form's pivot test fires and the warp redoes the cell with the
    dense row-exchanging LU (FAST); the EXACT warp LU restates nalgebra's exchanges bit for bit."""
    steps, T = 40, 0.4
    m, y0 = pivoting_case(n)
    om = to_oracle(m)
    A = np.eye(n) + m.fe * om.jacobian(y0[0])
    assert any(abs(A[r, i]) > abs(A[i, i]) for i in range(n) for r in range(i + 1, n))
    ref = O.solve_multi(om, O.RK4, T, steps, y0=y0)
    y0_dev = np.ascontiguousarray(y0.T)[None]
    force(monkeypatch, tw=32, fuse=2, warp=True)
    r = run_gpu(ctx, [m], cb.RK4, T, steps, cb.ARITH_EXACT, y0=y0_dev)
    assert r.status[0] == 0 and ref.status == 0
    assert_bit_equal(r.final_state[0].T, ref.final_state, f"n={n} pivoting, exact")
    f = run_gpu(ctx, [m], cb.RK4, T, steps, cb.ARITH_FAST, y0=y0_dev)
    assert max_rel_err(f.final_state[0].T, ref.final_state) <= PROFILE_RTOL


def test_long_column_natural_trigger(ctx):
    """n = 2, nz = 6000: 4*n*nz doubles = 384 KB > 227 KB, no knobs.  Also a batch of unequal columns."""
    nz, steps, T = 6000, 24, 0.6
    rng = np.random.default_rng(5)
    models, y0s = [], []
    for c in range(3):
        inj = cb.TemporalInjection.gaussian(0.2, 0.1, 0.1)
        sp = [cb.SpeciesParams.new("A", 1.0, 1.1 + 0.1 * c, 2, inj), cb.SpeciesParams.new("B", 1.0, 1.7, 2, inj)]
        models.append(cb.LangmuirMulti.new(sp, nz, 0.4, 1e-3, 0.25))
        y0s.append(rng.uniform(0.0, 0.05, size=(nz, 2)))
    d = describe(ctx, models, cb.RK4, T, steps, cb.ARITH_FAST)
    assert "multi_tiled" in d and "thread/cell" in d and "tiles/col=1 " not in d, d
    y0_dev = np.stack([np.ascontiguousarray(y.T) for y in y0s])
    r = run_gpu(ctx, models, cb.RK4, T, steps, cb.ARITH_EXACT, y0=y0_dev, snapshot_stride=12)
    f = run_gpu(ctx, models, cb.RK4, T, steps, cb.ARITH_FAST, y0=y0_dev)
    for c, (m, y0) in enumerate(zip(models, y0s)):
        ref = O.solve_multi(to_oracle(m), O.RK4, T, steps, y0=y0, want_trajectory=True)
        assert r.status[c] == 0
        assert_bit_equal(r.snapshots[c].transpose(0, 2, 1), ref.trajectory[::12], f"col {c} snapshots")
        assert_bit_equal(r.outlet[c], ref.outlet, f"col {c} outlet")
        assert_bit_equal(r.final_state[c].T, ref.final_state, f"col {c} final")
        assert max_rel_err(f.final_state[c].T, ref.final_state) <= PROFILE_RTOL


@pytest.mark.parametrize("warp", [False, True], ids=["thread", "warp"])
def test_tiled_status_matches_oracle(ctx, monkeypatch, warp):
    """validate_state across tiles and launches: first non-finite step and its NaN/Inf class as the oracle's."""
    inj = cb.TemporalInjection.rectangle(0.0, 100.0, 1.0)
    sp = [cb.SpeciesParams.new("A", 1.0, 1.1, 2, inj), cb.SpeciesParams.new("B", 1.0, 1.7, 2, inj)]
    bad = cb.LangmuirMulti.new(sp, 100, 0.4, 1e-3, 0.25)
    good = cb.LangmuirMulti.new(sp, 100, 0.4, 1e-6, 0.25)
    force(monkeypatch, tw=32, fuse=2, warp=warp)
    for T in (6000.0, 12000.0, 60000.0):
        oo, of, ost = O.solve_multi_batch([to_oracle(m) for m in (bad, good)], O.RK4, T, 400)
        r = run_gpu(ctx, [bad, good], cb.RK4, T, 400, cb.ARITH_EXACT)
        assert ost[0] != 0 and ost[1] == 0
        np.testing.assert_array_equal(r.status, ost)
        assert_bit_equal(r.final_state[1], of[1], "healthy column")


def test_tiled_plan_reexecute(ctx, monkeypatch):
    """A multi-launch tiled plan replays as a CUDA graph: executing it twice gives the same bits."""
    nz, steps, T = 100, 30, 3.0
    m, y0, y0_dev = case(4, nz, 3)
    force(monkeypatch, tw=24, fuse=1)
    cfg = cb.make_cfg(cb.RK4, T, steps, nz, 4, 1, 0, cb.ARITH_FAST)
    cols, species, tables = cb.pack_multi([m], cb.RK4, T, steps)
    plan = cb.Plan(ctx, cfg, cols, tables, species=species, y0=y0_dev, want=cb.WANT_FINAL | cb.WANT_OUTLET)
    try:
        plan.execute()
        a = plan.download()
        plan.execute()
        b = plan.download()
    finally:
        plan.close()
    assert_bit_equal(a.final_state, b.final_state, "re-executed plan")
    assert_bit_equal(a.outlet, b.outlet, "re-executed plan outlet")
    ref = O.solve_multi(to_oracle(m), O.RK4, T, steps, y0=y0)
    assert max_rel_err(a.final_state[0].T, ref.final_state) <= PROFILE_RTOL


@pytest.mark.parametrize("n,nz", [(12, 600), (16, 400), (10, 1500)])
def test_nine_to_sixteen_species_long_columns(ctx, n, nz):
    """n = 9..16 on columns beyond the register-resident capacity: FAST runs the closed-form elimination one thread
    per cell on the shared-memory-resident or tiled kernels (EXACT: one warp per cell).  No environment knobs."""
    steps = 16
    m, y0, y0_dev = case(n, nz, 31 * n + nz)
    T = steps * 0.2 * m.dz / m.ue
    d = describe(ctx, [m], cb.RK4, T, steps, cb.ARITH_FAST)
    assert ("multi_resident" in d or "thread/cell" in d) and "fast-structured-LU" in d, d
    ref = O.solve_multi(to_oracle(m), O.RK4, T, steps, y0=y0)
    f = run_gpu(ctx, [m, m], cb.RK4, T, steps, cb.ARITH_FAST, y0=np.concatenate([y0_dev, y0_dev]))
    assert not f.status.any()
    for c in range(2):
        assert max_rel_err(f.final_state[c].T, ref.final_state) <= PROFILE_RTOL
        assert max_rel_err(f.outlet[c], ref.outlet) <= PROFILE_RTOL
    e = run_gpu(ctx, [m], cb.RK4, T, steps, cb.ARITH_EXACT, y0=y0_dev)
    assert_bit_equal(e.final_state[0].T, ref.final_state, f"n={n} nz={nz} exact")
