"""GPU parity of the third-generation register-resident LangmuirMulti kernel (csrc/multi_reg3.cuh).

k_multi_reg3 decides the partial-pivoting guard per column RUN: the hot run uses the closed-form elimination on every
cell and a column in which any cell fails the sufficient no-row-exchange bound is integrated again by the guarded
code.  The contract tested here: whatever the seam protocol (CHROM_REG3_SYNC = 0, 3, 4: the three the build keeps) and whenever the bound fails
(never, at step 0, in the middle of a run, after the last poll), outlet / snapshots / final state / summary / status
are BIT-IDENTICAL to the second-generation kernel (CHROM_REG2 = 2), and within 1e-9 of the oracle
(langmuir_multi.rs:717-899 through rk4.rs:266-332 / euler.rs:227-269).
"""
import os

import numpy as np
import pytest

import chromatography_b200 as cb
from oracle import oracle as O
from helpers import PROFILE_RTOL, assert_bit_equal, max_rel_err, to_oracle

pytestmark = pytest.mark.gpu


class kernel_env:
    """Kernel-selection knobs of csrc/multi_reg.cu, read when a plan is built."""

    def __init__(self, gen, sync=None):
        self.kv = {"CHROM_REG2": str(gen)}
        if sync is not None:
            self.kv["CHROM_REG3_SYNC"] = str(sync)

    def __enter__(self):
        self.old = {k: os.environ.get(k) for k in ("CHROM_REG2", "CHROM_REG3_SYNC")}
        for k in self.old:
            os.environ.pop(k, None)
        os.environ.update(self.kv)

    def __exit__(self, *a):
        for k, v in self.old.items():
            os.environ.pop(k, None)
            if v is not None:
                os.environ[k] = v


def mixed_batch(n, nz, seed, late_start):
    """Five columns of one shape: mild (never suspect), overloaded from step 0 (needs row exchanges), overloaded from
    `late_start` on (a strong, concentrated feed arrives in the middle of the run), unstable (non-finite states:
    the bound fails on NaN as well), mild again."""
    rng = np.random.default_rng(seed)
    mild_inj = cb.TemporalInjection.gaussian(0.05, 0.02, 0.1)

    def mild():
        return cb.LangmuirMulti.new([cb.SpeciesParams.new(f"S{k}", rng.uniform(0.8, 1.5), rng.uniform(0.1, 0.8),
                                                          int(rng.integers(1, 4)), mild_inj) for k in range(n)],
                                    nz, 0.4, 1e-3, 0.25)

    half = max(1, n // 2)

    def strong(inj):
        return cb.LangmuirMulti.new(
            [cb.SpeciesParams.new(f"S{k}", 0.0, 50.0 * (1 + k), 1, inj) if k < half else
             cb.SpeciesParams.new(f"S{k}", 0.0, 1.0, 3, inj) for k in range(n)], nz, 0.4, 1e-3, 0.25)

    y0 = np.zeros((5, n, nz))
    y0[1, :half] = rng.uniform(0.0, 1e-4, (half, nz))
    y0[1, half:] = rng.uniform(0.0, 1e-3, (n - half, nz))
    y0[1, -1] = rng.uniform(1.5, 3.0, nz)
    unstable = cb.LangmuirMulti.new([cb.SpeciesParams.new(f"S{k}", 1.0, 1.0 + 0.1 * k, 2,
                                                          cb.TemporalInjection.rectangle(0.0, 1e9, 1.0))
                                     for k in range(n)], nz, 0.4, 5.0, 0.25)
    models = [mild(), strong(cb.TemporalInjection.rectangle(0.0, 1.0, 2.0)),
              strong(cb.TemporalInjection.rectangle(late_start, 1e9, 2.0)), unstable, mild()]
    return models, y0


def run(ctx, models, y0, method, T, steps, snapshot_stride=0):
    n = models[0].n_species()
    cfg = cb.make_cfg(method, T, steps, models[0].n_points, n, 1, snapshot_stride, cb.ARITH_FAST)
    cols, species, tables = cb.pack_multi(models, method, T, steps)
    return cb.solve_multi_batch(ctx, cfg, cols, species, tables, y0, True, snapshot_stride > 0, True, True)


def assert_same(a, b, what):
    np.testing.assert_array_equal(a.status, b.status, err_msg=what + " status")
    ok = a.status == 0  # columns that went non-finite stop being compared at the first bad step by contract
    assert_bit_equal(a.final_state[ok], b.final_state[ok], what + " final")
    assert_bit_equal(a.outlet[ok], b.outlet[ok], what + " outlet")
    assert_bit_equal(a.summary[ok], b.summary[ok], what + " summary")
    if a.snapshots is not None:
        assert_bit_equal(a.snapshots[ok], b.snapshots[ok], what + " snapshots")


@pytest.mark.parametrize("method", [cb.RK4, cb.EULER], ids=["rk4", "euler"])
@pytest.mark.parametrize("n,nz", [(8, 512), (8, 100), (7, 257), (6, 64), (5, 33), (3, 130), (1, 40), (2, 2)])
def test_bit_identical_to_second_generation(ctx, n, nz, method):
    steps = 150
    dt = 0.2 * (0.25 / nz) / (1e-3 / 0.4)  # CFL 0.2 on the unretained velocity
    T = steps * dt
    models, y0 = mixed_batch(n, nz, 100 * n + nz, late_start=100.5 * dt)
    with kernel_env(2):
        ref = run(ctx, models, y0, method, T, steps, snapshot_stride=50)
    assert ref.status[0] == 0 and ref.status[1] == 0 and ref.status[2] == 0 and ref.status[4] == 0
    assert ref.status[3] != 0
    for sync in (0, 3, 4):
        with kernel_env(3, sync):
            r = run(ctx, models, y0, method, T, steps, snapshot_stride=50)
        assert_same(r, ref, f"n={n} nz={nz} sync={sync}")
    # and the oracle, on the three kinds of healthy column
    om = [to_oracle(m) for m in models]
    for c in (0, 1, 2):
        o = O.solve_multi(om[c], method, T, steps, y0=np.ascontiguousarray(y0[c].T))
        assert o.status == 0
        assert max_rel_err(ref.final_state[c].T, o.final_state) <= PROFILE_RTOL, (n, nz, c)
        assert max_rel_err(ref.outlet[c], o.outlet) <= PROFILE_RTOL, (n, nz, c)


def test_late_suspect_after_the_last_poll(ctx):
    """The strong feed arrives 3 steps before the end (after the last 64-step poll of the split-barrier protocol): the
    end-of-run check must still send the column to the guarded code."""
    n, nz, steps = 8, 96, 200
    dt = 0.2 * (0.25 / nz) / (1e-3 / 0.4)
    models, y0 = mixed_batch(n, nz, 7, late_start=196.5 * dt)
    with kernel_env(2):
        ref = run(ctx, models, y0, cb.RK4, steps * dt, steps)
    for sync in (0, 3, 4):
        with kernel_env(3, sync):
            assert_same(run(ctx, models, y0, cb.RK4, steps * dt, steps), ref, f"late suspect sync={sync}")


def test_many_columns_persistent_loop_and_default_choice(ctx):
    """More columns than persistent CTAs, every third one overloaded: runs of both kinds alternate inside one CTA
    (barriers re-initialised between runs).  No knob set: the default kernel for n >= 5 must be the third
    generation."""
    n, nz, steps = 5, 64, 70
    dt = 0.2 * (0.25 / nz) / (1e-3 / 0.4)
    models, y0 = [], []
    for k in range(120):
        m, y = mixed_batch(n, nz, 1000 + k, late_start=(20.5 + k % 40) * dt)
        for c in (0, 2, 1):
            models.append(m[c])
            y0.append(y[c])
    y0 = np.stack(y0)
    with kernel_env(2):
        ref = run(ctx, models, y0, cb.RK4, steps * dt, steps)
    cfg = cb.make_cfg(cb.RK4, steps * dt, steps, nz, n, 1, 0, cb.ARITH_FAST)
    cols, species, tables = cb.pack_multi(models, cb.RK4, steps * dt, steps)
    plan = cb.Plan(ctx, cfg, cols, tables, species, y0)
    assert "multi_reg3" in plan.describe(), plan.describe()
    plan.close()
    for sync in (None, 0, 3, 4):
        with kernel_env(3, sync):
            assert_same(run(ctx, models, y0, cb.RK4, steps * dt, steps), ref, f"360 columns sync={sync}")
