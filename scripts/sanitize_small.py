"""Small solves through every kernel family, for compute-sanitizer (memcheck / racecheck) runs:
   compute-sanitizer --tool racecheck python scripts/sanitize_small.py
Shapes are tiny (sanitizer tools slow kernels down 10-100x); results are still checked against the oracle."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
import chromatography_b200 as cb  # noqa: E402
from chromatography_b200 import workloads as W  # noqa: E402
from oracle import oracle as O  # noqa: E402
from helpers import max_rel_err, to_oracle  # noqa: E402


def multi_case(ctx, n, nz, steps, arith, env=None, method=cb.RK4):
    for k, v in (env or {}).items():
        os.environ[k] = v
    try:
        rng = np.random.default_rng(n * 1000 + nz)
        inj = cb.TemporalInjection.gaussian(0.3, 0.1, 0.1)
        sp = [cb.SpeciesParams.new(f"S{k}", rng.uniform(0.8, 1.5), rng.uniform(0.1, 0.8), int(rng.integers(1, 4)), inj)
              for k in range(n)]
        m = cb.LangmuirMulti.new(sp, nz, 0.4, 1e-3, 0.25 * nz / 100.0)
        T = 0.1 * steps
        y0 = rng.uniform(0.0, 0.05, size=(nz, n))
        cfg = cb.make_cfg(method, T, steps, nz, n, 1, max(1, steps // 2), arith)
        cols, species, tables = cb.pack_multi([m, m], method, T, steps)
        y0d = np.ascontiguousarray(y0.T)[None].repeat(2, axis=0)
        r = cb.solve_multi_batch(ctx, cfg, cols, species, tables, y0d, True, True, True, True)
        ref = O.solve_multi(to_oracle(m), O.RK4 if method == cb.RK4 else O.EULER, T, steps, y0=y0)
        err = max_rel_err(r.final_state[0].T, ref.final_state)
        assert err <= 1e-9 and not r.status.any(), (n, nz, arith, env, err)
        return err
    finally:
        for k in (env or {}):
            os.environ.pop(k, None)


def main():
    ctx = cb.Context([0])
    out = []
    # single-species: register-resident (one warp / multi warp) and streaming
    for nz, steps in ((100, 20), (1500, 6), (9000, 4)):
        m = W.tfa_single(nz=nz)
        T = 0.5 * steps
        cfg = cb.make_cfg(cb.RK4, T, steps, nz, 1, 1, max(1, steps // 2), cb.ARITH_FAST)
        cols, tables = cb.pack_single([m, m, m], cb.RK4, T, steps)
        r = cb.solve_single_batch(ctx, cfg, cols, tables, None, True, True, True, True)
        ref = O.solve_single(to_oracle(m), O.RK4, T, steps)
        err = max_rel_err(r.final_state[0], ref.final_state)
        assert err <= 1e-9 and not r.status.any(), (nz, err)
        out.append(("single", nz, err))
    # multi-species: register-resident (M = 1, 2, 4), shared-memory (exact, dense), tiled, warp per cell
    for n, nz, arith, env in (
            (2, 100, cb.ARITH_FAST, {"CHROM_REG_M": "1"}), (2, 100, cb.ARITH_FAST, {"CHROM_REG_M": "4"}),
            (8, 70, cb.ARITH_FAST, {"CHROM_REG_M": "2"}), (8, 70, cb.ARITH_FAST, None),
            (3, 90, cb.ARITH_EXACT, None), (8, 64, cb.ARITH_FAST_DENSE, None), (5, 64, cb.ARITH_EXACT, None),
            (4, 100, cb.ARITH_FAST, {"CHROM_MULTI_TILED": "1", "CHROM_TILED_TW": "24", "CHROM_TILED_FUSE": "2"}),
            (4, 100, cb.ARITH_EXACT, {"CHROM_MULTI_TILED": "1", "CHROM_TILED_TW": "24", "CHROM_TILED_FUSE": "1"}),
            (12, 40, cb.ARITH_FAST, None), (12, 40, cb.ARITH_EXACT, None),
            (40, 24, cb.ARITH_FAST, {"CHROM_MULTI_TILED": "1", "CHROM_TILED_TW": "12", "CHROM_TILED_FUSE": "1"}),
            (70, 16, cb.ARITH_FAST, None), (70, 12, cb.ARITH_EXACT, None)):
        out.append(("multi", n, nz, arith, multi_case(ctx, n, nz, 6, arith, env)))
    out.append(("multi-euler", multi_case(ctx, 8, 70, 6, cb.ARITH_FAST, None, cb.EULER)))
    ctx.close()
    for o in out:
        print(o)
    print("sanitize_small OK")


if __name__ == "__main__":
    main()
