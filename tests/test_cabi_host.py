"""CPU-side checks of the drop-in boundary: libchrom_cuda.so loads and exports every symbol
include/chrom_cuda.h declares; the host-only helpers of the C ABI restate the reference's boundary
semantics bit for bit (checked against the oracle); the Python mirror reproduces the reference's
validation and error texts; without a GPU the library refuses to create a context (no CPU fallback).
No compute call is made here."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import chromatography_b200 as cb
from chromatography_b200 import _ffi, workloads as W
from oracle import oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions():
    src = open(os.path.join(ROOT, "include", "chrom_cuda.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(chrom_cuda_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    names = header_functions()
    assert len(names) >= 25
    lib = C.CDLL(_ffi.LIB_PATH)
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/chrom_cuda.h but not exported"
    assert sorted(_ffi.SYMBOLS) == names, "ctypes table and header disagree"
    assert b"sm_100a" in _ffi.lib().chrom_cuda_version()


def test_constants_match_header():
    """Every CHROM_* constant of include/chrom_cuda.h equals its mirror in the ctypes table and the C++ host enum."""
    src = open(os.path.join(ROOT, "include", "chrom_cuda.h")).read()
    defs = {k: int(v.strip("()u")) for k, v in re.findall(r"#define (CHROM_[A-Z0-9_]+) (\(?-?\d+u?\)?)", src)}
    mirror = {"CHROM_OK": _ffi.CHROM_OK, "CHROM_EULER": _ffi.EULER, "CHROM_RK4": _ffi.RK4,
              "CHROM_ARITH_FAST": _ffi.ARITH_FAST, "CHROM_ARITH_EXACT": _ffi.ARITH_EXACT,
              "CHROM_ARITH_FAST_STRUCTURED": _ffi.ARITH_FAST_STRUCTURED, "CHROM_ARITH_FAST_DENSE": _ffi.ARITH_FAST_DENSE,
              "CHROM_WANT_OUTLET": _ffi.WANT_OUTLET, "CHROM_WANT_SNAPSHOTS": _ffi.WANT_SNAPSHOTS,
              "CHROM_WANT_FINAL": _ffi.WANT_FINAL, "CHROM_WANT_SUMMARY": _ffi.WANT_SUMMARY,
              "CHROM_WANT_STATUS": _ffi.WANT_STATUS}
    for k, v in mirror.items():
        assert defs[k] == v, (k, defs[k], v)
    hpp = open(os.path.join(ROOT, "include", "chrom", "solver.hpp")).read()
    for name, key in (("Fast", "CHROM_ARITH_FAST"), ("Exact", "CHROM_ARITH_EXACT"),
                      ("FastStructured", "CHROM_ARITH_FAST_STRUCTURED"), ("FastDense", "CHROM_ARITH_FAST_DENSE")):
        assert re.search(rf"\b{name} = {defs[key]},", hpp), name
    rs = open(os.path.join(ROOT, "bindings", "rust", "chrom-cuda", "src", "ffi.rs")).read()
    for key in ("CHROM_ARITH_FAST", "CHROM_ARITH_EXACT", "CHROM_ARITH_FAST_STRUCTURED", "CHROM_ARITH_FAST_DENSE"):
        assert re.search(rf"pub const {key}: i32 = {defs[key]};", rs), key


def test_struct_layouts_match_header():
    assert C.sizeof(_ffi.SingleCol) == 56 and C.sizeof(_ffi.MultiCol) == 40 and C.sizeof(_ffi.Species) == 24
    assert C.sizeof(_ffi.RunCfg) == 80 and C.sizeof(_ffi.Timing) == 56  # 48 + two (pointer, count) sample lists


def test_no_cpu_fallback():
    lib = _ffi.lib()
    if lib.chrom_cuda_device_count() > 0:
        pytest.skip("a GPU is visible")
    h = C.c_void_p()
    assert lib.chrom_cuda_create(None, 0, C.byref(h)) == _ffi.CHROM_ERR_NO_DEVICE
    assert b"no CPU fallback" in lib.chrom_cuda_last_error(None)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        cb.Context()
    m = W.tfa_single()
    sc = cb.Scenario.new(m, cb.DomainBoundaries.temporal(m.setup_initial_state()))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        cb.CudaRK4Solver().solve(sc, cb.SolverConfiguration.time_evolution(1.0, 10))


def test_validate_cfg_texts():
    """SolverType::validate — src/solver/traits.rs:172-185."""
    lib = _ffi.lib()
    assert lib.chrom_cuda_validate_cfg(C.byref(cb.make_cfg(cb.RK4, 10.0, 100, 100))) is None
    assert lib.chrom_cuda_validate_cfg(C.byref(cb.make_cfg(cb.RK4, -1.0, 100, 100))) == b"Total time must be positive"
    assert lib.chrom_cuda_validate_cfg(C.byref(cb.make_cfg(cb.RK4, 0.0, 100, 100))) == b"Total time must be positive"
    assert lib.chrom_cuda_validate_cfg(C.byref(cb.make_cfg(cb.RK4, 1.0, 0, 100))) == b"TimeSteps must be greater than 0"


@pytest.mark.parametrize("T,steps", [(600.0, 10_000), (10.0, 100), (1.0, 3), (0.7, 7)])
def test_time_points_bitwise(T, steps):
    tp = cb.time_points(T, steps)
    assert np.array_equal(tp.view(np.uint64), O.time_points(T, steps).view(np.uint64))
    assert tp[0] == 0.0 and len(tp) == steps + 1


def test_n_samples_and_shard_ranges():
    lib = _ffi.lib()
    assert lib.chrom_cuda_n_samples(1000, 0) == 0 and lib.chrom_cuda_n_samples(1000, 1) == 1001
    assert lib.chrom_cuda_n_samples(1003, 7) == 144 and lib.chrom_cuda_n_samples(5, 10) == 1
    a, b = C.c_uint64(), C.c_uint64()
    for n_cols, g in [(65536, 8), (10, 4), (3, 8), (1, 1), (4097, 2)]:
        covered = []
        for d in range(g):
            lib.chrom_cuda_shard_range(n_cols, g, d, C.byref(a), C.byref(b))
            covered += list(range(a.value, a.value + b.value))
            assert b.value <= -(-n_cols // g)
        assert covered == list(range(n_cols)), "contiguous, disjoint, complete"


INJECTIONS = [cb.TemporalInjection.none(), cb.TemporalInjection.dirac(0.0, 1e-3), cb.TemporalInjection.dirac(5.0, 0.1),
              cb.TemporalInjection.dirac(0.03, 0.1), cb.TemporalInjection.dirac(0.09, 0.1),
              cb.TemporalInjection.gaussian(10.0, 3.0, 0.1), cb.TemporalInjection.rectangle(5.0, 15.0, 0.05),
              cb.TemporalInjection.rectangle(0.0, 0.12, 1e-3)]


@pytest.mark.parametrize("inj", INJECTIONS)
def test_injection_tables_bitwise(inj):
    """Stage times t, t + dt/2.0, t + dt with t = (double)step*dt (rk4.rs:267-295); Euler t = dt*step."""
    T, steps = 60.0, 1000
    dt = T / float(steps)
    oi = O.TemporalInjection(inj.kind, inj.p0, inj.p1, inj.p2)
    rk = cb.tabulate_injection(inj, cb.RK4, T, steps)
    eu = cb.tabulate_injection(inj, cb.EULER, T, steps)
    assert rk.shape == (steps, 3) and eu.shape == (steps, 1)
    for step in list(range(0, 200)) + [steps - 1]:
        t = float(step) * dt
        want = [oi.evaluate(t), oi.evaluate(t + dt / 2.0), oi.evaluate(t + dt)]
        assert [float(x) for x in rk[step]] == want
        assert float(eu[step, 0]) == oi.evaluate(dt * float(step))
    if inj == cb.TemporalInjection.dirac(0.03, 0.1):
        assert rk[0, 1] == 0.1 and rk[:, 0].sum() == 0.0  # hit only at a mid-stage time
        assert eu.sum() == 0.0                            # Euler never evaluates t = 0.03


def test_status_messages_match_reference_text():
    assert cb.status_message(0) is None
    for s in (1, 42, -7, -10_000):
        assert cb.status_message(s) == O.validate_state_message(s)
    assert cb.status_message(3) == ("NaN detected in Concentration at step 3. This indicates numerical instability. "
                                    "Try reducing time step (increase time_steps parameter).")


def test_model_validation_texts():
    """langmuir_single.rs:245-259 (panics) and langmuir_multi.rs:178-204, 350-392 (Err strings)."""
    inj = cb.TemporalInjection.none()
    with pytest.raises(ValueError, match=r"Porosity must be in \]0,1\[, got 1.5"):
        cb.LangmuirSingle.new(1.2, 0.4, 2.0, 1.5, 1e-3, 0.25, 100, inj)
    with pytest.raises(ValueError, match="Column length must be positive, got 0"):
        cb.LangmuirSingle.new(1.2, 0.4, 2.0, 0.4, 1e-3, 0.0, 100, inj)
    with pytest.raises(ValueError, match="Need at least 2 spatial points, got 1"):
        cb.LangmuirSingle.new(1.2, 0.4, 2.0, 0.4, 1e-3, 0.25, 1, inj)
    with pytest.raises(ValueError, match="Rectangle end must be > start"):
        cb.TemporalInjection.rectangle(5.0, 5.0, 1.0)
    sp = cb.SpeciesParams.new
    with pytest.raises(ValueError, match="must have at least one species"):
        cb.LangmuirMulti.new([], 100, 0.4, 1e-3, 0.25)
    with pytest.raises(ValueError, match="Species 'X' : lambda must be >= 0, got -1"):
        cb.LangmuirMulti.new([sp("X", -1.0, 1.0, 1, inj)], 100, 0.4, 1e-3, 0.25)
    with pytest.raises(ValueError, match="Langmuir K̃ must be > 0, got 0"):
        cb.LangmuirMulti.new([sp("X", 1.0, 0.0, 1, inj)], 100, 0.4, 1e-3, 0.25)
    with pytest.raises(ValueError, match="Port number must be > 0"):
        cb.LangmuirMulti.new([sp("X", 1.0, 1.0, 0, inj)], 100, 0.4, 1e-3, 0.25)
    with pytest.raises(ValueError, match="Duplicate speciees name 'A'"):
        cb.LangmuirMulti.new([sp("A", 1.0, 1.0, 1, inj), sp("A", 1.0, 2.0, 1, inj)], 100, 0.4, 1e-3, 0.25)
    with pytest.raises(ValueError, match="column length must be strictly positive"):
        cb.LangmuirMulti.new([sp("A", 1.0, 1.0, 1, inj)], 100, 0.4, 1e-3, 0.0)
    m = cb.LangmuirMulti.new([sp("A", 1.0, 0.5, 1, inj)], 100, 0.4, 1e-3, 0.25)
    m.add_species(sp("B", 1.0, 2.0, 1, inj))
    with pytest.raises(ValueError, match="Species 'A' already exists in this model"):
        m.add_species(sp("A", 1.0, 2.0, 1, inj))
    with pytest.raises(ValueError):
        m.add_species(sp("C", -1.0, 2.0, 1, inj))
    assert m.n_species() == 2 and m.species_names() == ["A", "B"] and m.points() == 100
    assert (m.fe, m.ue, m.dz, m.stationary_fraction) == ((1.0 - 0.4) / 0.4, 1e-3 / 0.4, 0.25 / 100.0, 1.0 - 0.4)


def test_serde_shapes_keep_derived_fields_verbatim():
    """src/config/model.rs:111-148: fe / ue / dz come from the file, not from the constructor:
    `fe: 1.5` is not the double (1-0.4)/0.4."""
    m = cb.LangmuirSingle.from_map({"LangmuirSingle": {
        "lambda": 1.2, "langmuir_k": 0.4, "port_number": 2.0, "column_length": 0.25, "n_points": 100,
        "dz": 0.0025, "fe": 1.5, "ue": 0.0025, "injection": {"type": "Gaussian", "center": 10.0, "width": 3.0,
                                                            "peak_concentration": 0.1}}})
    assert m.fe == 1.5 and m.fe != W.tfa_single().fe
    assert cb.LangmuirSingle.from_map(m.to_map()) == m
    mm = W.acids_multi()
    assert cb.LangmuirMulti.from_map(mm.to_map()).to_map() == mm.to_map()
    with pytest.raises(ValueError, match="Custom"):
        cb.TemporalInjection.from_map({"type": "Custom"})


def test_rust_float_formatting():
    """metadata values are `f64::to_string()` (rk4.rs:345-346)."""
    for x, s in [(600.0, "600"), (0.06, "0.06"), (1e-7, "0.0000001"), (1200.0, "1200"), (0.1 + 0.2, "0.30000000000000004"),
                 (1e21, "1000000000000000000000"), (2.5e-3, "0.0025")]:
        assert cb.rust_f64(x) == s


def test_solver_errors_before_any_device_work():
    """The reference's early returns (rk4.rs:213-245) need no device: raised before a context exists."""
    m = W.tfa_single()
    sc = cb.Scenario.new(m, cb.DomainBoundaries.temporal(m.setup_initial_state()))
    s = cb.CudaRK4Solver()
    with pytest.raises(cb.SolverError, match="Total time must be positive"):
        s.solve(sc, cb.SolverConfiguration.time_evolution(0.0, 10))
    with pytest.raises(cb.SolverError, match="TimeSteps must be greater than 0"):
        s.solve(sc, cb.SolverConfiguration.time_evolution(1.0, 0))
    with pytest.raises(cb.SolverError, match="EulerSolver only supports TimeEvolution configuration, got Analytical"):
        cb.CudaEulerSolver().solve(sc, cb.SolverConfiguration.analytical(1.0))
    with pytest.raises(cb.SolverError, match="Dimension boundaries cannot be empty."):
        s.solve(cb.Scenario.new(m, cb.DomainBoundaries.default()), cb.SolverConfiguration.time_evolution(1.0, 10))
    with pytest.raises(cb.SolverError, match="No initial condition found in domain boundaries"):
        s.solve(cb.Scenario.new(m, cb.DomainBoundaries.temporal(None)), cb.SolverConfiguration.time_evolution(1.0, 10))

    class FisherKPP:  # examples/diffusion.rs custom model: not covered, and no fallback
        def points(self):
            return 100

        def name(self):
            return "FisherKPP"
    with pytest.raises(cb.SolverError, match="not supported by the CUDA backend"):
        s.solve(cb.Scenario.new(FisherKPP(), cb.DomainBoundaries.temporal(np.zeros(100))),
                cb.SolverConfiguration.time_evolution(1.0, 10))


def test_packing_deduplicates_injection_tables():
    models = W.tfa_sweep_models(64)
    cols, tables = cb.pack_single(models, cb.RK4, 60.0, 100)
    assert tables.shape == (1, 100, 3) and all(c.inj_table == 0 for c in cols)
    models[3].set_injection(cb.TemporalInjection.dirac(0.0, 1.0))
    cols, tables = cb.pack_single(models, cb.RK4, 60.0, 100)
    assert tables.shape == (2, 100, 3) and cols[3].inj_table == 1 and cols[4].inj_table == 0
    mm = W.multi_batch_models(3, 4, 64)
    mcols, species, mt = cb.pack_multi(mm, cb.RK4, 60.0, 100)
    assert len(species) == 12 and [c.species_off for c in mcols] == [0, 4, 8] and mt.shape[0] == 1


def test_splitmix64_reference_stream():
    """Vigna's splitmix64.c with seed 1234567 starts 6457827717110365317, 3203168211198807973."""
    r = W.SplitMix64(1234567)
    assert [r.next_u64() for _ in range(2)] == [6457827717110365317, 3203168211198807973]
    a = W.SplitMix64(42)
    scalar = [a.uniform(0.0, 1.0) for _ in range(100)]
    assert np.array_equal(W.splitmix_uniform(100, 42), np.array(scalar))
    p = W.tfa_sweep_arrays(1000)
    assert 0.1 <= p["langmuir_k"].min() and p["langmuir_k"].max() <= 0.8 and 1.0 <= p["port_number"].min()
    assert p["port_number"].max() <= 3.0 and p["fe"].min() >= 1.0 - 1e-12 and p["fe"].max() <= 0.7 / 0.3 + 1e-12


def test_sample_indices_and_node_index_known_answers():
    """chrom_cuda_sample_indices / chrom_cuda_node_index against the reference's own doc and unit tests
    (physics/traits.rs:1003-1023, 1466-1482; domain/detector.rs:278-290, 391-406)."""
    import json
    ka = json.load(open(os.path.join(ROOT, "tests", "golden", "reference_known_answers.json")))
    for c in ka["sample_indices"]["cases"]:
        assert cb.sample_indices(c["total"], c["n"]).tolist() == c["expected"], c
    for c in ka["detector_node_index"]["cases"]:
        assert cb.node_index(c["position"], c["column_length"], c["n_points"]) == c["expected"], c
    # the rule itself: n evenly spaced by integer division, last forced to total - 1, for a sweep of sizes
    for total in (2, 7, 100, 10001):
        for n in (2, 3, 5, 50, total - 1):
            if n < 2 or n >= total:
                continue
            want = [(i * (total - 1)) // (n - 1) for i in range(n)]
            want[-1] = total - 1
            assert cb.sample_indices(total, n).tolist() == want
    # a sample list turns into the buffer length the ABI expects
    cfg = cb.make_cfg(cb.RK4, 600.0, 10000, 100, outlet_steps=cb.sample_indices(10001, 250), snapshot_stride=2500)
    lib = _ffi.lib()
    assert lib.chrom_cuda_cfg_n_out(C.byref(cfg)) == 250 and lib.chrom_cuda_cfg_n_snap(C.byref(cfg)) == 5
    plain = cb.make_cfg(cb.RK4, 600.0, 10000, 100, outlet_stride=50)
    assert lib.chrom_cuda_cfg_n_out(C.byref(plain)) == 201 and lib.chrom_cuda_cfg_n_snap(C.byref(plain)) == 0
