"""Prints the measured FAST-vs-oracle error levels (max relative error, entry by entry) for a few workloads:
the margin under north_star's 1e-9.  python scripts/fast_error_levels.py"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
import chromatography_b200 as cb  # noqa: E402
from chromatography_b200 import workloads as W  # noqa: E402
from oracle import oracle as O  # noqa: E402
from helpers import max_rel_err, to_oracle  # noqa: E402

ctx = cb.Context([0])
# C1 / C3 column, full horizon
m = W.tfa_single()
ref = O.solve_single(to_oracle(m), O.RK4, 600.0, 10000)
cfg = cb.make_cfg(cb.RK4, 600.0, 10000, 100, 1, 1, 0, cb.ARITH_FAST)
cols, tables = cb.pack_single([m], cb.RK4, 600.0, 10000)
r = cb.solve_single_batch(ctx, cfg, cols, tables, None, True, False, True, False)
print("C1 single fast: outlet %.2e final %.2e" % (max_rel_err(r.outlet[0], ref.outlet), max_rel_err(r.final_state[0], ref.final_state)))
# C2 full
m = W.acids_multi()
ref = O.solve_multi(to_oracle(m), O.RK4, 1200.0, 20000)
for name, arith in (("fast", cb.ARITH_FAST), ("dense", cb.ARITH_FAST_DENSE)):
    cfg = cb.make_cfg(cb.RK4, 1200.0, 20000, 100, 2, 1, 0, arith)
    c, sp, tab = cb.pack_multi([m], cb.RK4, 1200.0, 20000)
    r = cb.solve_multi_batch(ctx, cfg, c, sp, tab, None, True, False, True, False)
    print("C2 multi n=2 %s: outlet %.2e final %.2e" % (name, max_rel_err(r.outlet[0], ref.outlet), max_rel_err(r.final_state[0].T, ref.final_state)))
# C5 columns, full horizon
models = W.multi_batch_models(4, 8, 512)
oo, of, ost = O.solve_multi_batch([to_oracle(x) for x in models], O.RK4, 600.0, 4096)
for name, arith in (("fast", cb.ARITH_FAST), ("dense", cb.ARITH_FAST_DENSE)):
    cfg = cb.make_cfg(cb.RK4, 600.0, 4096, 512, 8, 1, 0, arith)
    c, sp, tab = cb.pack_multi(models, cb.RK4, 600.0, 4096)
    r = cb.solve_multi_batch(ctx, cfg, c, sp, tab, None, True, False, True, False)
    print("C5 multi n=8 %s: outlet %.2e final %.2e" % (name, max_rel_err(r.outlet, oo), max_rel_err(r.final_state, of)))
# n = 32 (warp per cell)
rng = np.random.default_rng(1)
inj = cb.TemporalInjection.gaussian(10.0, 3.0, 0.1)
sp = [cb.SpeciesParams.new(f"S{k}", rng.uniform(0.8, 1.5), rng.uniform(0.1, 0.8), int(rng.integers(1, 4)), inj) for k in range(32)]
m = cb.LangmuirMulti.new(sp, 64, 0.4, 1e-3, 0.25)
ref = O.solve_multi(to_oracle(m), O.RK4, 300.0, 600)
cfg = cb.make_cfg(cb.RK4, 300.0, 600, 64, 32, 1, 0, cb.ARITH_FAST)
c, sp2, tab = cb.pack_multi([m], cb.RK4, 300.0, 600)
r = cb.solve_multi_batch(ctx, cfg, c, sp2, tab, None, True, False, True, False)
print("n=32 warp per cell fast: outlet %.2e final %.2e" % (max_rel_err(r.outlet[0], ref.outlet), max_rel_err(r.final_state[0].T, ref.final_state)))
ctx.close()
