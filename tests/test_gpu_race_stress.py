"""GPU: a stand-in for `compute-sanitizer --tool racecheck`, which is closed on this pool.

The register- and shared-memory-resident kernels synchronise by hand: double-buffered inlet rows, one barrier per
stage or per step, a seam ring between adjacent warps guarded by release/acquire counters
(multi_reg2.cuh, multi_reg3.cuh: counters or mbarriers), warp-boundary cells through shared memory (single_resident.cu, multi-warp columns), ping-pong stage
buffers (multi_kernels.cuh), tile halos (multi_tiled.cuh, single_stream.cu).  A race in any of them shows up as
timing-dependent bits.  So: a batch of IDENTICAL columns, several per persistent CTA and more CTAs than the machine
holds at once, is solved repeatedly; every column of every run must carry exactly the bits of column 0 of run 0, and
that column must agree with the oracle.  Timing differs from CTA to CTA, warp to warp and run to run; the bits may not.
"""
import os

import numpy as np
import pytest

import chromatography_b200 as cb
from chromatography_b200 import workloads as W
from oracle import oracle as O
from helpers import PROFILE_RTOL, assert_bit_equal, max_rel_err, to_oracle

pytestmark = pytest.mark.gpu

RUNS = 4


def _all_columns_equal(arr, what):
    first = np.ascontiguousarray(arr[0])
    for c in range(1, arr.shape[0]):
        if not np.array_equal(first.view(np.uint64), np.ascontiguousarray(arr[c]).view(np.uint64)):
            raise AssertionError(f"{what}: column {c} differs from column 0 (identical inputs)")


def _multi_model(n, nz, seed):
    rng = np.random.default_rng(seed)
    inj = cb.TemporalInjection.gaussian(2.0, 0.7, 0.2)
    sp = [cb.SpeciesParams.new(f"S{k}", rng.uniform(0.8, 1.5), rng.uniform(0.1, 0.8), int(rng.integers(1, 4)), inj)
          for k in range(n)]
    return cb.LangmuirMulti.new(sp, nz, 0.4, 1e-3, 0.25 * nz / 100.0)


# (n_species, nz, arithmetic, environment): every hand-synchronised multi-species kernel family
MULTI_CASES = [
    (8, 512, cb.ARITH_FAST, {}),                       # k_multi_reg3 (default protocol): 8 warps, seam ring, split step barrier
    (6, 200, cb.ARITH_FAST, {}),                       # k_multi_reg3, ragged last warp
    (5, 97, cb.ARITH_FAST, {}),                        # k_multi_reg3, odd n, nz not a multiple of anything
    (8, 512, cb.ARITH_FAST, {"CHROM_REG3_SYNC": "0"}),  # k_multi_reg3, release/acquire counters instead of mbarriers
    (8, 512, cb.ARITH_FAST, {"CHROM_REG3_SYNC": "3"}),  # k_multi_reg3, predicated seam code, CTA barrier per step
    (3, 300, cb.ARITH_FAST, {}),                        # k_multi_reg3 at n = 3 (machine-filling batch: 600 columns)
    (8, 512, cb.ARITH_FAST, {"CHROM_REG2": "2"}),      # k_multi_reg2 (second generation, guard per cell)
    (2, 300, cb.ARITH_FAST, {}),                       # k_multi_reg: shared-memory exchange, one barrier per stage
    (12, 256, cb.ARITH_FAST, {}),                      # k_multi_reg, acc in shared slots
    (8, 512, cb.ARITH_FAST_DENSE, {}),                 # k_multi_resident: ping-pong stage buffers
    (3, 400, cb.ARITH_EXACT, {}),                      # k_multi_resident, reference order
    (4, 300, cb.ARITH_FAST, {"CHROM_MULTI_TILED": "1", "CHROM_TILED_TW": "64", "CHROM_TILED_FUSE": "2"}),  # tiles + halos
    (20, 64, cb.ARITH_FAST, {}),                       # warp per cell
]


@pytest.mark.parametrize("n,nz,arith,env", MULTI_CASES,
                         ids=[f"n{n}-nz{nz}-a{a}" + "".join("-" + k.split("_")[-1].lower() + v for k, v in e.items() if len(v) < 3)
                              for n, nz, a, e in MULTI_CASES])
def test_identical_multi_columns_agree_bit_for_bit(ctx, monkeypatch, n, nz, arith, env):
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    m = _multi_model(n, nz, seed=100 * n + nz)
    n_cols, T, steps = 600, 12.0, 160  # > 148 SMs x resident CTAs: persistent CTAs loop over several columns
    cfg = cb.make_cfg(cb.RK4, T, steps, nz, n, 1, 40, arith)
    cols, species, tables = cb.pack_multi([m] * n_cols, cb.RK4, T, steps)
    plan = cb.Plan(ctx, cfg, cols, tables, species=species,
                   want=cb.WANT_OUTLET | cb.WANT_FINAL | cb.WANT_STATUS | cb.WANT_SNAPSHOTS | cb.WANT_SUMMARY)
    first = None
    for run in range(RUNS):
        plan.execute()
        r = plan.download()
        assert not r.status.any()
        for name in ("outlet", "final_state", "snapshots", "summary"):
            _all_columns_equal(getattr(r, name), f"{plan.describe()} run {run} {name}")
        if first is None:
            first = r
        else:
            assert_bit_equal(r.outlet[0], first.outlet[0], "outlet, run to run")
            assert_bit_equal(r.final_state[0], first.final_state[0], "final, run to run")
    plan.close()
    ref = O.solve_multi(to_oracle(m), O.RK4, T, steps)
    if arith == cb.ARITH_EXACT:
        assert_bit_equal(first.final_state[0].T, ref.final_state, "vs oracle")
    else:
        assert max_rel_err(first.final_state[0].T, ref.final_state) <= PROFILE_RTOL
        assert max_rel_err(first.outlet[0], ref.outlet) <= PROFILE_RTOL


# single species: one warp per several columns (shuffles only), several warps per column (shared-memory warp seam,
# one barrier per stage), and the streaming kernel (tile halos, ping-pong through L2)
@pytest.mark.parametrize("nz,n_cols,steps", [(100, 40000, 200), (1500, 700, 120), (5000, 400, 60), (40000, 6, 40)],
                         ids=["one-warp", "multi-warp-1500", "multi-warp-5000", "stream"])
def test_identical_single_columns_agree_bit_for_bit(ctx, nz, n_cols, steps):
    probe = W.tfa_single(nz=nz)
    T = steps * min(0.05, 0.4 * probe.dz / probe.ue)  # CFL <= 0.4 whatever the grid
    m = W.tfa_single(injection=cb.TemporalInjection.gaussian(0.3 * T, 0.1 * T, 0.1), nz=nz)
    cfg = cb.make_cfg(cb.RK4, T, steps, nz, 1, 1, max(1, steps // 4), cb.ARITH_FAST)
    cols, tables = cb.pack_single([m] * n_cols, cb.RK4, T, steps)
    plan = cb.Plan(ctx, cfg, cols, tables,
                   want=cb.WANT_OUTLET | cb.WANT_FINAL | cb.WANT_STATUS | cb.WANT_SNAPSHOTS | cb.WANT_SUMMARY)
    first = None
    for run in range(RUNS):
        plan.execute()
        r = plan.download()
        assert not r.status.any()
        for name in ("outlet", "final_state", "snapshots", "summary"):
            _all_columns_equal(getattr(r, name), f"{plan.describe()} run {run} {name}")
        if first is None:
            first = r
        else:
            assert_bit_equal(r.final_state[0], first.final_state[0], "final, run to run")
    plan.close()
    ref = O.solve_single(to_oracle(m), O.RK4, T, steps)
    assert max_rel_err(first.final_state[0], ref.final_state) <= PROFILE_RTOL
    assert max_rel_err(first.outlet[0], ref.outlet) <= PROFILE_RTOL


def test_environment_is_restored():
    assert "CHROM_MULTI_TILED" not in os.environ
