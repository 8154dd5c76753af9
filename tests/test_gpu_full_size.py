"""GPU parity at BASELINE.json's FULL sizes (configs 3 and 5), through the C ABI.

The oracle cannot run 65 536 x 10 000 steps in seconds, so the full-size runs are checked by
  * a seeded sample of columns against the oracle (profiles 1e-9, retention time / area 1e-7),
  * a size-independent property of the domain: columns are independent scenarios, so a column's result must
    not depend on where in the batch (which warp / CTA / wave) it was solved — the same columns re-solved as
    a small batch must agree with the full-size run bit for bit,
  * EVERY column, not a sample: the full batch is solved a second time in EXACT arithmetic (bit-identical to the oracle
    on the sampled columns, checked here) and the FAST run must agree with it within 1e-9 on every entry of every column,
  * validate_state: every status 0, every summary finite, mass recovered at the outlet within physical bounds.
Config 4 at full size: tests/test_gpu_stream.py::test_full_size_prefix_property.
"""
import numpy as np
import pytest

import chromatography_b200 as cb
from chromatography_b200 import _ffi, workloads as W
from oracle import oracle as O
from helpers import MOMENT_RTOL, PROFILE_RTOL, assert_bit_equal, max_rel_err, to_oracle

pytestmark = pytest.mark.gpu


def test_c3_full_size(ctx):
    """Config 3: 65 536 TFA-like columns, nz = 100, RK4 600 s / 10 000 steps (the bench workload)."""
    n_cols, nz, T, steps, stride = 65536, 100, 600.0, 10_000, 50
    models = W.tfa_sweep_models(n_cols)
    cfg = cb.make_cfg(cb.RK4, T, steps, nz, 1, stride, 0, cb.ARITH_FAST)
    cols, tables = cb.pack_single(models, cb.RK4, T, steps)
    r = cb.solve_single_batch(ctx, cfg, cols, tables, None, True, False, True, True)
    assert not r.status.any()
    assert np.isfinite(r.summary).all() and np.isfinite(r.final_state).all()
    # the Gaussian(10, 3, 0.1) injection carries 0.1*3*sqrt(2 pi) mol s / L; what has left by t = 600 s is between 0 and all of it
    injected = 0.1 * 3.0 * np.sqrt(2.0 * np.pi)
    assert (r.summary[:, 0] >= 0.0).all() and (r.summary[:, 0] <= injected * (1.0 + 1e-6)).all()
    rng = np.random.default_rng(42)
    sample = np.sort(rng.choice(n_cols, size=48, replace=False))
    sample[0], sample[-1] = 0, n_cols - 1
    oo, of, ost = O.solve_single_batch([to_oracle(models[i]) for i in sample], O.RK4, T, steps)
    assert not ost.any()
    tp = cb.time_points(T, steps)
    for k, i in enumerate(sample):
        assert max_rel_err(r.outlet[i], oo[k][::stride]) <= PROFILE_RTOL, i
        assert max_rel_err(r.final_state[i], of[k]) <= PROFILE_RTOL, i
        area, tr = O.trapezoid(tp, oo[k]), O.first_moment(tp, oo[k])
        assert abs(r.summary[i][0] - area) <= MOMENT_RTOL * max(area, 1e-300), i
        if area > 1e-12:
            assert abs(r.summary[i][1] - tr) <= MOMENT_RTOL * tr, i
    # every column: the EXACT run of the whole batch carries the oracle's bits on the sample, and FAST agrees with it
    # within the contract on all 65 536 columns
    cfg_x = cb.make_cfg(cb.RK4, T, steps, nz, 1, stride, 0, cb.ARITH_EXACT)
    x = cb.solve_single_batch(ctx, cfg_x, cols, tables, None, True, False, True, False)
    assert not x.status.any()
    for k, i in enumerate(sample):
        assert_bit_equal(x.outlet[i], oo[k][::stride], f"EXACT outlet of column {i} vs oracle")
        assert_bit_equal(x.final_state[i], of[k], f"EXACT final state of column {i} vs oracle")
    assert max_rel_err(r.outlet, x.outlet) <= PROFILE_RTOL
    assert max_rel_err(r.final_state, x.final_state) <= PROFILE_RTOL
    # batch-position independence, bit for bit
    sub = (_ffi.SingleCol * len(sample))(*[cols[int(i)] for i in sample])
    s = cb.solve_single_batch(ctx, cfg, sub, tables, None, True, False, True, True)
    assert_bit_equal(s.outlet, r.outlet[sample], "outlet of the sampled columns")
    assert_bit_equal(s.final_state, r.final_state[sample], "final state of the sampled columns")
    assert_bit_equal(s.summary, r.summary[sample], "summary of the sampled columns")


def test_c5_full_size(ctx):
    """Config 5: 4 096 LangmuirMulti columns, n = 8, nz = 512, RK4 600 s / 4 096 steps."""
    n_cols, n, nz, T, steps, stride = 4096, 8, 512, 600.0, 4096, 64
    models = W.multi_batch_models(n_cols, n, nz)
    cfg = cb.make_cfg(cb.RK4, T, steps, nz, n, stride, 0, cb.ARITH_FAST)
    cols, species, tables = cb.pack_multi(models, cb.RK4, T, steps)
    r = cb.solve_multi_batch(ctx, cfg, cols, species, tables, None, True, False, True, True)
    assert not r.status.any()
    assert np.isfinite(r.summary).all() and np.isfinite(r.final_state).all()
    sample = [0, 1777, n_cols - 1]
    oo, of, ost = O.solve_multi_batch([to_oracle(models[i]) for i in sample], O.RK4, T, steps)
    assert not ost.any()
    for k, i in enumerate(sample):
        assert max_rel_err(r.outlet[i], oo[k][:, ::stride]) <= PROFILE_RTOL, i
        assert max_rel_err(r.final_state[i], of[k]) <= PROFILE_RTOL, i
    # every column: EXACT (nalgebra LU order; the oracle's bits on the sample) against FAST on all 4 096 columns
    cfg_x = cb.make_cfg(cb.RK4, T, steps, nz, n, stride, 0, cb.ARITH_EXACT)
    x = cb.solve_multi_batch(ctx, cfg_x, cols, species, tables, None, True, False, True, False)
    assert not x.status.any()
    for k, i in enumerate(sample):
        assert_bit_equal(x.outlet[i], oo[k][:, ::stride], f"EXACT outlet of column {i} vs oracle")
        assert_bit_equal(x.final_state[i], of[k], f"EXACT final state of column {i} vs oracle")
    assert max_rel_err(r.outlet, x.outlet) <= PROFILE_RTOL
    assert max_rel_err(r.final_state, x.final_state) <= PROFILE_RTOL
    # batch-position independence, bit for bit: 20 columns re-solved on their own
    rng = np.random.default_rng(5)
    pick = np.sort(rng.choice(n_cols, size=20, replace=False))
    sub_models = [models[int(i)] for i in pick]
    c2, s2, t2 = cb.pack_multi(sub_models, cb.RK4, T, steps)
    s = cb.solve_multi_batch(ctx, cfg, c2, s2, t2, None, True, False, True, True)
    assert_bit_equal(s.outlet, r.outlet[pick], "outlet of the re-solved columns")
    assert_bit_equal(s.final_state, r.final_state[pick], "final state of the re-solved columns")
    assert_bit_equal(s.summary, r.summary[pick], "summary of the re-solved columns")
