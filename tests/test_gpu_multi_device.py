"""GPU, >= 2 devices: one libchrom_cuda context over several GPUs cuts a batch into contiguous column
ranges (chrom_cuda_shard_range, SURVEY §8e: no collective, host-side concatenation only).  The result
must be bit-identical to the single-device solve of the same batch — for the one-shot call (chunked,
copy-back overlapped) and for a resident plan — and to the oracle."""
import numpy as np
import pytest

import chromatography_b200 as cb
from chromatography_b200 import _ffi, workloads as W
from oracle import oracle as O
from helpers import assert_bit_equal, to_oracle

pytestmark = pytest.mark.gpu


def _need_two():
    n = _ffi.lib().chrom_cuda_device_count()
    if n < 2:
        pytest.skip(f"{n} visible device(s): the multi-device path needs 2 (run with gpurun --gpus 2)")
    return n


@pytest.mark.parametrize("n_cols", [2, 37, 2048])
def test_single_batch_sharded_over_devices(ctx, n_cols):
    n_dev = _need_two()
    T, steps = 30.0, 500
    models = W.tfa_sweep_models(n_cols)
    cols, tables = cb.pack_single(models, cb.RK4, T, steps)
    cfg = cb.make_cfg(cb.RK4, T, steps, 100, arith=cb.ARITH_EXACT)
    one = cb.solve_single_batch(ctx, cfg, cols, tables, want_summary=True)
    with cb.Context(list(range(n_dev))) as many:
        r = cb.solve_single_batch(many, cfg, cols, tables, want_summary=True)
        plan = cb.Plan(many, cfg, cols, tables)
        plan.execute()
        p = plan.download()
        plan.close()
    for got in (r, p):
        assert_bit_equal(got.outlet, one.outlet, "outlet")
        assert_bit_equal(got.final_state, one.final_state, "final")
        assert (got.status == one.status).all()
    assert_bit_equal(r.summary, one.summary, "summary")
    k = min(n_cols, 8)
    oo, of, _ = O.solve_single_batch([to_oracle(m) for m in models[:k]], O.RK4, T, steps)
    assert_bit_equal(r.outlet[:k], oo, "outlet vs oracle")
    assert_bit_equal(r.final_state[:k], of, "final vs oracle")


def test_large_one_shot_takes_the_pipelined_path_on_every_device(ctx):
    """Enough output (> 64 MB per shard) for the chunked compute / copy-back pipeline on each device."""
    n_dev = _need_two()
    n_cols, T, steps = 16384, 120.0, 2000
    a = W.tfa_sweep_arrays(n_cols, seed=7)
    cols = (_ffi.SingleCol * n_cols)()
    arr = np.frombuffer(cols, dtype=np.float64).reshape(n_cols, 7)
    for j, k in enumerate(["lambda_", "langmuir_k", "port_number", "fe", "ue", "dz"]):
        arr[:, j] = a[k]
    arr[:, 6] = 0.0
    tables = cb.tabulate_injection(cb.TemporalInjection.gaussian(10.0, 3.0, 0.1), cb.RK4, T, steps)[None]
    cfg = cb.make_cfg(cb.RK4, T, steps, 100)
    one = cb.solve_single_batch(ctx, cfg, cols, np.ascontiguousarray(tables))
    with cb.Context(list(range(n_dev))) as many:
        r = cb.solve_single_batch(many, cfg, cols, np.ascontiguousarray(tables))
    assert_bit_equal(r.outlet, one.outlet, "outlet")
    assert_bit_equal(r.final_state, one.final_state, "final")
    assert not r.status.any()


def test_multi_species_batch_sharded_over_devices(ctx):
    n_dev = _need_two()
    T, steps = 20.0, 200
    models = W.multi_batch_models(9, 3, 64)
    cols, species, tables = cb.pack_multi(models, cb.RK4, T, steps)
    cfg = cb.make_cfg(cb.RK4, T, steps, 64, 3, arith=cb.ARITH_EXACT)
    one = cb.solve_multi_batch(ctx, cfg, cols, species, tables)
    with cb.Context(list(range(n_dev))) as many:
        r = cb.solve_multi_batch(many, cfg, cols, species, tables)
    assert_bit_equal(r.outlet, one.outlet, "outlet")
    assert_bit_equal(r.final_state, one.final_state, "final")
