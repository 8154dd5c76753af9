"""Host-side mirror of the reference's solver boundary, calling libchrom_cuda.so through ctypes.

Mirrors (paths in chrom-rs):
  SolverType / SolverConfiguration / SimulationResult / Solver   src/solver/traits.rs:62-121, 258-280, 449-474, 606-642
  Scenario                                                       src/solver/scenario.rs:50-89
  DomainBoundaries::temporal / initial_condition                 src/solver/boundary.rs:98-100, 273-276
  RK4Solver::solve / EulerSolver::solve                          src/solver/methods/rk4.rs:205-350, euler.rs:166-288
`Err(String)` becomes `SolverError(str)` with the same text.  Anything the CUDA backend does not
cover (unknown model types, Custom injections) raises: there is no CPU fallback.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field

import numpy as np

from . import _ffi
from .models import LangmuirMulti, LangmuirSingle, TemporalInjection, rust_f64


class SolverError(Exception):
    """The reference's `Err(String)`."""


# ---- configuration types ---------------------------------------------------------------------
@dataclass
class SolverType:
    kind: str = "TimeEvolution"
    total_time: float = 0.0
    time_steps: int = 0

    @staticmethod
    def TimeEvolution(total_time, time_steps) -> "SolverType":
        return SolverType("TimeEvolution", float(total_time), int(time_steps))

    def name(self) -> str:
        return self.kind

    def validate(self) -> None:
        if self.kind == "TimeEvolution":
            if self.total_time <= 0.0:
                raise SolverError("Total time must be positive")
            if self.time_steps == 0:
                raise SolverError("TimeSteps must be greater than 0")


@dataclass
class SolverConfiguration:
    solver_type: SolverType
    step: int | None = None

    @staticmethod
    def time_evolution(total_time, time_steps) -> "SolverConfiguration":
        return SolverConfiguration(SolverType.TimeEvolution(total_time, time_steps))

    @staticmethod
    def iterative(tolerance, max_iterations) -> "SolverConfiguration":
        return SolverConfiguration(SolverType("Iterative"))

    @staticmethod
    def analytical(evaluation_time) -> "SolverConfiguration":
        return SolverConfiguration(SolverType("Analytical"))

    def with_step(self, n: int) -> "SolverConfiguration":
        self.step = int(n)
        return self

    def validate(self) -> None:
        self.solver_type.validate()


@dataclass
class DomainBoundaries:
    """Only the temporal form the hot path uses: one 't' dimension holding the initial state."""
    initial: np.ndarray | None = None
    n_dimensions: int = 1

    @staticmethod
    def temporal(initial) -> "DomainBoundaries":
        return DomainBoundaries(None if initial is None else np.asarray(initial, dtype=np.float64))

    @staticmethod
    def default() -> "DomainBoundaries":
        return DomainBoundaries(None, 0)

    def initial_condition(self):
        return self.initial

    def validate(self) -> None:
        if self.n_dimensions == 0:
            raise SolverError("Dimension boundaries cannot be empty.")


@dataclass
class Scenario:
    model: object
    conditions: DomainBoundaries

    @staticmethod
    def new(model, conditions) -> "Scenario":
        return Scenario(model, conditions)

    def validate(self) -> None:
        self.conditions.validate()

    def get_model_name(self) -> str:
        return self.model.name()


@dataclass
class SimulationResult:
    time_points: np.ndarray
    state_trajectory: np.ndarray | None   # [n_snap, nz] or [n_snap, nz, n_species]; all steps when snapshot_stride == 1
    final_state: np.ndarray               # [nz] or [nz, n_species]
    metadata: dict = field(default_factory=dict)
    outlet: np.ndarray | None = None      # [n_out] or [n_species, n_out]
    outlet_time_points: np.ndarray | None = None
    summary: np.ndarray | None = None     # [4] or [n_species, 4]: area, first moment, max, time of max

    def __len__(self) -> int:
        return len(self.time_points)

    def is_empty(self) -> bool:
        return len(self.time_points) == 0


# ---- context ------------------------------------------------------------------------------------
class Context:
    """One libchrom_cuda context (a set of devices + streams).  Not thread-safe."""

    def __init__(self, device_ids=None):
        lib = _ffi.lib()
        h = C.c_void_p()
        if device_ids is None:
            rc = lib.chrom_cuda_create(None, 0, C.byref(h))
        else:
            ids = (C.c_int32 * len(device_ids))(*device_ids)
            rc = lib.chrom_cuda_create(ids, len(device_ids), C.byref(h))
        if rc != _ffi.CHROM_OK:
            raise RuntimeError(f"chrom_cuda_create failed ({rc}): {lib.chrom_cuda_last_error(None).decode()}")
        self._h = h
        self._lib = lib

    def close(self):
        if getattr(self, "_h", None):
            self._lib.chrom_cuda_destroy(self._h)
            self._h = None

    __del__ = close

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def check(self, rc: int):
        if rc != _ffi.CHROM_OK:
            raise SolverError(self._lib.chrom_cuda_last_error(self._h).decode())

    def timing(self) -> dict:
        t = _ffi.Timing()
        self._lib.chrom_cuda_last_timing(self._h, C.byref(t))
        return {n: getattr(t, n) for n, _ in _ffi.Timing._fields_}

    def bind_host_thread(self, device_index=0) -> tuple:
        """Pin the calling thread (and its later allocations) to the NUMA node of the device: (node, n_cpus);
        node == -1 when the host exposes no topology (nothing changed)."""
        node, ncpu = C.c_int32(-1), C.c_int32(0)
        self.check(self._lib.chrom_cuda_bind_host_thread(self._h, device_index, C.byref(node), C.byref(ncpu)))
        return node.value, ncpu.value

    def flush_l2(self) -> None:
        self.check(self._lib.chrom_cuda_flush_l2(self._h))

    def fp64_peak_tflops(self, device_index=0) -> float:
        tf, ms = C.c_double(), C.c_double()
        self.check(self._lib.chrom_cuda_measure_fp64_peak(self._h, device_index, C.byref(tf), C.byref(ms)))
        return tf.value

    def copy_bw_gbs(self, nbytes=1 << 30, device_index=0) -> float:
        g = C.c_double()
        self.check(self._lib.chrom_cuda_measure_copy_bw(self._h, device_index, nbytes, C.byref(g)))
        return g.value

    def probe_rcp(self, lo=0.5, hi=64.0, n=1 << 20):
        out = (C.c_double * 3)()
        self.check(self._lib.chrom_cuda_probe_rcp(self._h, lo, hi, n, out))
        return list(out)


_default_ctx = None


def default_context() -> Context:
    global _default_ctx
    if _default_ctx is None:
        _default_ctx = Context()
    return _default_ctx


# ---- low-level batch calls (thin, array-in / array-out) ------------------------------------------
def _dp(a):
    return None if a is None else a.ctypes.data_as(C.POINTER(C.c_double))


def time_points(total_time: float, time_steps: int) -> np.ndarray:
    out = np.empty(time_steps + 1)
    _ffi.lib().chrom_cuda_time_points(float(total_time), time_steps, _dp(out))
    return out


def tabulate_injection(inj: TemporalInjection, method: int, total_time: float, time_steps: int) -> np.ndarray:
    stages = _ffi.lib().chrom_cuda_table_stages(method)
    out = np.empty((time_steps, stages))
    rc = _ffi.lib().chrom_cuda_tabulate_injection(inj.kind, inj.p0, inj.p1, inj.p2, method, float(total_time),
                                                  time_steps, _dp(out))
    if rc != _ffi.CHROM_OK:
        raise SolverError("chrom_cuda_tabulate_injection: invalid arguments")
    return out


def status_message(status: int):
    if status == 0:
        return None
    buf = C.create_string_buffer(512)
    _ffi.lib().chrom_cuda_status_message(int(status), buf, 512)
    return buf.value.decode()


def make_cfg(method, total_time, time_steps, n_points, n_species=1, outlet_stride=1, snapshot_stride=0,
             arith=_ffi.ARITH_FAST, outlet_steps=None, snapshot_steps=None) -> _ffi.RunCfg:
    """outlet_steps / snapshot_steps: explicit strictly increasing step lists (what sample_indices(time_steps + 1, n)
    returns, physics/traits.rs:1010-1023); a list replaces the corresponding stride."""
    cfg = _ffi.RunCfg(method, arith, float(total_time), int(time_steps), int(n_points), int(n_species),
                      int(outlet_stride), int(snapshot_stride))
    keep = []
    for name, steps in (("outlet", outlet_steps), ("snapshot", snapshot_steps)):
        if steps is None:
            continue
        arr = np.ascontiguousarray(steps, dtype=np.uint64)
        keep.append(arr)  # the struct only borrows the memory: the array lives as long as the cfg object
        setattr(cfg, f"{name}_steps", arr.ctypes.data_as(C.POINTER(C.c_uint64)))
        setattr(cfg, f"n_{name}_steps", arr.size)
    cfg._keepalive = keep
    return cfg


def sample_indices(total: int, n: int | None) -> np.ndarray:
    """sample_indices(total, n) of the reference (physics/traits.rs:1010-1023); n None / 0 = every index."""
    lib = _ffi.lib()
    n = int(n or 0)
    count = lib.chrom_cuda_sample_indices(int(total), n, None, 0)
    out = np.zeros(count, dtype=np.uint64)
    lib.chrom_cuda_sample_indices(int(total), n, out.ctypes.data_as(C.POINTER(C.c_uint64)), count)
    return out


def node_index(position: float, column_length: float, n_points: int) -> int:
    """Detector::node_index (domain/detector.rs:285-290): the grid node closest to a position in metres."""
    return int(_ffi.lib().chrom_cuda_node_index(float(position), float(column_length), int(n_points)))


class _TableSet:
    """De-duplicates injection profiles of a batch into [n_tables][steps][stages]."""

    def __init__(self, method, total_time, time_steps):
        self.key = (method, float(total_time), int(time_steps))
        self.index = {}
        self.tables = []

    def add(self, inj: TemporalInjection) -> int:
        if not isinstance(inj, TemporalInjection):
            raise SolverError("Temporal Injection::Custom cannot be serialized")  # injection.rs:289-293
        i = self.index.get(inj)
        if i is None:
            i = len(self.tables)
            self.index[inj] = i
            self.tables.append(tabulate_injection(inj, *self.key))
        return i

    def array(self) -> np.ndarray:
        return np.ascontiguousarray(np.stack(self.tables))


def pack_single(models, method, total_time, time_steps):
    ts = _TableSet(method, total_time, time_steps)
    cols = (_ffi.SingleCol * len(models))()
    for i, m in enumerate(models):
        cols[i] = _ffi.SingleCol(m.lambda_, m.langmuir_k, m.port_number, m.fe, m.ue, m.dz, ts.add(m.injection),
                                 int(getattr(m, "detector_cell", 0)))  # 0 = outlet, k = cell k - 1
    return cols, ts.array()


def pack_multi(models, method, total_time, time_steps):
    ts = _TableSet(method, total_time, time_steps)
    n = models[0].n_species()
    cols = (_ffi.MultiCol * len(models))()
    species = (_ffi.Species * (len(models) * n))()
    for i, m in enumerate(models):
        if m.n_species() != n:
            raise SolverError("all columns of a batch must have the same number of species")
        cols[i] = _ffi.MultiCol(m.fe, m.ue, m.dz, m.stationary_fraction, i * n, int(getattr(m, "detector_cell", 0)))
        for k, s in enumerate(m.species):
            species[i * n + k] = _ffi.Species(s.lambda_, s.langmuir_k, s.port_number, ts.add(s.injection))
    return cols, species, ts.array()


@dataclass
class BatchResult:
    outlet: np.ndarray | None        # [n_cols, n_out] / [n_cols, n_species, n_out]
    snapshots: np.ndarray | None     # [n_cols, n_snap, nz] / [n_cols, n_snap, n_species, nz]
    final_state: np.ndarray | None   # [n_cols, nz] / [n_cols, n_species, nz]
    summary: np.ndarray | None       # [n_cols, 6] / [n_cols, n_species, 6]: area, first moment, max, t_max, sigma, min
    status: np.ndarray               # [n_cols] int64
    timing: dict


def solve_single_batch(ctx: Context, cfg: _ffi.RunCfg, cols, tables: np.ndarray, y0=None, want_outlet=True,
                       want_snapshots=False, want_final=True, want_summary=False) -> BatchResult:
    lib = _ffi.lib()
    n_cols, nz = len(cols), cfg.n_points
    n_out = lib.chrom_cuda_cfg_n_out(C.byref(cfg))
    n_snap = lib.chrom_cuda_cfg_n_snap(C.byref(cfg))
    outlet = np.empty((n_cols, n_out)) if (want_outlet and n_out) else None
    snaps = np.empty((n_cols, n_snap, nz)) if (want_snapshots and n_snap) else None
    final = np.empty((n_cols, nz)) if want_final else None
    summ = np.empty((n_cols, _ffi.SUMMARY_LEN)) if want_summary else None
    status = np.zeros(n_cols, dtype=np.int64)
    y0c = None if y0 is None else np.ascontiguousarray(y0, dtype=np.float64).reshape(n_cols, nz)
    tables = np.ascontiguousarray(tables, dtype=np.float64)
    ctx.check(lib.chrom_cuda_solve_single_batch(ctx._h, C.byref(cfg), cols, n_cols, _dp(tables), tables.shape[0],
                                                _dp(y0c), _dp(outlet), _dp(snaps), _dp(final), _dp(summ),
                                                status.ctypes.data_as(C.POINTER(C.c_int64))))
    return BatchResult(outlet, snaps, final, summ, status, ctx.timing())


def solve_multi_batch(ctx: Context, cfg: _ffi.RunCfg, cols, species, tables: np.ndarray, y0=None, want_outlet=True,
                      want_snapshots=False, want_final=True, want_summary=False) -> BatchResult:
    lib = _ffi.lib()
    n_cols, nz, n = len(cols), cfg.n_points, cfg.n_species
    n_out = lib.chrom_cuda_cfg_n_out(C.byref(cfg))
    n_snap = lib.chrom_cuda_cfg_n_snap(C.byref(cfg))
    outlet = np.empty((n_cols, n, n_out)) if (want_outlet and n_out) else None
    snaps = np.empty((n_cols, n_snap, n, nz)) if (want_snapshots and n_snap) else None
    final = np.empty((n_cols, n, nz)) if want_final else None
    summ = np.empty((n_cols, n, _ffi.SUMMARY_LEN)) if want_summary else None
    status = np.zeros(n_cols, dtype=np.int64)
    y0c = None if y0 is None else np.ascontiguousarray(y0, dtype=np.float64).reshape(n_cols, n, nz)
    tables = np.ascontiguousarray(tables, dtype=np.float64)
    ctx.check(lib.chrom_cuda_solve_multi_batch(ctx._h, C.byref(cfg), cols, n_cols, species, len(species),
                                               _dp(tables), tables.shape[0], _dp(y0c), _dp(outlet), _dp(snaps),
                                               _dp(final), _dp(summ), status.ctypes.data_as(C.POINTER(C.c_int64))))
    return BatchResult(outlet, snaps, final, summ, status, ctx.timing())


class Plan:
    """A batch kept resident in HBM: upload once, execute many times, download on demand."""

    def __init__(self, ctx: Context, cfg: _ffi.RunCfg, cols, tables, species=None, y0=None,
                 want=_ffi.WANT_OUTLET | _ffi.WANT_FINAL | _ffi.WANT_STATUS):
        lib = _ffi.lib()
        self.ctx, self.cfg, self.n_cols, self.want = ctx, cfg, len(cols), want
        self.multi = species is not None
        tables = np.ascontiguousarray(tables, dtype=np.float64)
        y0c = None if y0 is None else np.ascontiguousarray(y0, dtype=np.float64)
        h = C.c_void_p()
        if self.multi:
            rc = lib.chrom_cuda_plan_multi(ctx._h, C.byref(cfg), cols, len(cols), species, len(species), _dp(tables),
                                           tables.shape[0], _dp(y0c), want, C.byref(h))
        else:
            rc = lib.chrom_cuda_plan_single(ctx._h, C.byref(cfg), cols, len(cols), _dp(tables), tables.shape[0],
                                            _dp(y0c), want, C.byref(h))
        ctx.check(rc)
        self._h = h

    def describe(self) -> str:
        buf = C.create_string_buffer(512)
        _ffi.lib().chrom_cuda_plan_describe(self._h, buf, 512)
        return buf.value.decode()

    def execute(self) -> dict:
        self.ctx.check(_ffi.lib().chrom_cuda_plan_execute(self._h))
        return self.ctx.timing()

    def download(self) -> BatchResult:
        lib = _ffi.lib()
        cfg, n_cols, nz, n = self.cfg, self.n_cols, self.cfg.n_points, self.cfg.n_species
        n_out = lib.chrom_cuda_cfg_n_out(C.byref(cfg))
        n_snap = lib.chrom_cuda_cfg_n_snap(C.byref(cfg))
        sp = (n,) if self.multi else ()
        outlet = np.empty((n_cols,) + sp + (n_out,)) if (self.want & _ffi.WANT_OUTLET and n_out) else None
        snaps = np.empty((n_cols, n_snap) + sp + (nz,)) if (self.want & _ffi.WANT_SNAPSHOTS and n_snap) else None
        final = np.empty((n_cols,) + sp + (nz,)) if self.want & _ffi.WANT_FINAL else None
        summ = np.empty((n_cols,) + sp + (_ffi.SUMMARY_LEN,)) if self.want & _ffi.WANT_SUMMARY else None
        status = np.zeros(n_cols, dtype=np.int64)
        self.ctx.check(lib.chrom_cuda_plan_download(self._h, _dp(outlet), _dp(snaps), _dp(final), _dp(summ),
                                                    status.ctypes.data_as(C.POINTER(C.c_int64))))
        return BatchResult(outlet, snaps, final, summ, status, self.ctx.timing())

    def rsf(self, other: "Plan") -> np.ndarray:
        """Surface resolution Rsf (validation/dissimilarity.rs:173-203) between this plan's outlet series and
        `other`'s, per column (and species), computed on the device: neither series moves to the host."""
        sp = (self.cfg.n_species,) if self.multi else ()
        out = np.empty((self.n_cols,) + sp)
        self.ctx.check(_ffi.lib().chrom_cuda_plan_rsf(self._h, other._h, _dp(out)))
        return out

    def close(self):
        if getattr(self, "_h", None):
            _ffi.lib().chrom_cuda_plan_destroy(self._h)
            self._h = None

    __del__ = close


# ---- the reference's Solver trait, CUDA-backed ------------------------------------------------------
class _CudaSolver:
    METHOD = _ffi.RK4
    NAME = ""
    META_SOLVER = ""

    def __init__(self, ctx: Context | None = None, arith=_ffi.ARITH_FAST, full_trajectory=True):
        self._ctx = ctx
        self.arith = arith
        self.full_trajectory = full_trajectory  # the reference keeps every step's state (rk4.rs:315)

    @classmethod
    def new(cls, *a, **k):
        return cls(*a, **k)

    def name(self) -> str:
        return self.NAME

    @property
    def ctx(self) -> Context:
        if self._ctx is None:
            self._ctx = default_context()
        return self._ctx

    def solve(self, scenario: Scenario, config: SolverConfiguration) -> SimulationResult:
        r = self.solve_batch([scenario], config)[0]
        if isinstance(r, SolverError):
            raise r
        return r

    def solve_batch(self, scenarios, config: SolverConfiguration, snapshot_stride=None, outlet_stride=1,
                    want_summary=True) -> list:
        """`solve` over independent scenarios of one shape in a single device batch.
        Each entry of the returned list is a SimulationResult, or the SolverError that scenario's
        `solve` would have returned."""
        config.validate()                                   # rk4.rs:213
        for sc in scenarios:
            sc.validate()                                   # rk4.rs:217
        st = config.solver_type
        if st.kind != "TimeEvolution":                      # rk4.rs:226-231 (the RK4 text says Euler too)
            raise SolverError(f"EulerSolver only supports TimeEvolution configuration, got {st.name()}")
        total_time, steps = st.total_time, st.time_steps
        models = [sc.model for sc in scenarios]
        y0s = []
        for sc in scenarios:
            ic = sc.conditions.initial_condition()
            if ic is None:                                  # rk4.rs:242-245
                raise SolverError("No initial condition found in domain boundaries")
            y0s.append(ic)
        first = models[0]
        if not all(type(m) is type(first) and m.points() == first.points() for m in models):
            raise SolverError("solve_batch needs scenarios of one model type and grid size")
        if snapshot_stride is None:
            snapshot_stride = 1 if self.full_trajectory else 0
        nz = first.points()
        if isinstance(first, LangmuirSingle):
            cfg = make_cfg(self.METHOD, total_time, steps, nz, 1, outlet_stride, snapshot_stride, self.arith)
            cols, tables = pack_single(models, self.METHOD, total_time, steps)
            y0 = np.stack([np.asarray(y, dtype=np.float64).reshape(nz) for y in y0s])
            y0 = None if not y0.any() else y0
            r = solve_single_batch(self.ctx, cfg, cols, tables, y0, True, snapshot_stride > 0, True, want_summary)
            multi = False
        elif isinstance(first, LangmuirMulti):
            n = first.n_species()
            cfg = make_cfg(self.METHOD, total_time, steps, nz, n, outlet_stride, snapshot_stride, self.arith)
            cols, species, tables = pack_multi(models, self.METHOD, total_time, steps)
            # reference state is [nz, n]; the device layout is [n][nz] (nalgebra column-major)
            y0 = np.stack([np.asarray(y, dtype=np.float64).reshape(nz, n).T for y in y0s])
            y0 = None if not y0.any() else y0
            r = solve_multi_batch(self.ctx, cfg, cols, species, tables, y0, True, snapshot_stride > 0, True,
                                  want_summary)
            multi = True
        else:
            raise SolverError(f"model '{type(first).__name__}' is not supported by the CUDA backend "
                              "(LangmuirSingle and LangmuirMulti only; no CPU fallback)")
        tp = time_points(total_time, steps)
        dt = total_time / float(steps)
        results = []
        for i in range(len(scenarios)):
            if r.status[i] != 0:                            # validate_state: rk4.rs:331 -> that scenario's Err
                results.append(SolverError(status_message(int(r.status[i]))))
                continue
            meta = {"solver": self.META_SOLVER, "time steps": str(steps), "dt": rust_f64(dt),
                    "total time": rust_f64(total_time)}
            if self.METHOD == _ffi.RK4:
                meta["function evaluations"] = str(4 * steps)
            traj = None if r.snapshots is None else (r.snapshots[i].transpose(0, 2, 1) if multi else r.snapshots[i])
            final = r.final_state[i].T if multi else r.final_state[i]
            results.append(SimulationResult(tp, traj, final, meta, None if r.outlet is None else r.outlet[i],
                                            tp[::outlet_stride] if outlet_stride else None,
                                            None if r.summary is None else r.summary[i]))
        return results


    def solve_scan(self, scenarios, configs, max_concurrent=4, **batch_options) -> list:
        """Heterogeneous sweeps (a different time grid / spatial grid / model type per scenario: the reference's
        sequential loops in examples/cfl_stability_scan.rs:378-394, stiffness_convergence.rs:360,
        spatial_temporal_convergence.rs:261-277).  Entries are grouped by (model type, n_points, species count,
        configuration); each group is one device batch and up to `max_concurrent` groups run at the same time,
        each on its own context and host thread (distinct contexts are thread-safe).  Entry i is what
        solve(scenarios[i], configs[i]) would have returned, or its SolverError."""
        from concurrent.futures import ThreadPoolExecutor
        import threading
        if len(configs) != len(scenarios):
            raise SolverError("solve_scan needs one SolverConfiguration per scenario")
        groups: dict = {}
        for i, (sc, cf) in enumerate(zip(scenarios, configs)):
            m, st = sc.model, cf.solver_type
            key = (type(m).__name__, m.points(), m.n_species() if isinstance(m, LangmuirMulti) else 1, st.kind,
                   getattr(st, "total_time", None), getattr(st, "time_steps", None))
            groups.setdefault(key, []).append(i)
        results: list = [None] * len(scenarios)
        local = threading.local()
        first_ctx = self.ctx

        def run(idx):
            if getattr(local, "solver", None) is None:
                with lock:
                    own = first_ctx if not taken else Context()
                    taken.append(own)
                local.solver = type(self)(own, self.arith, self.full_trajectory)
            try:
                out = local.solver.solve_batch([scenarios[i] for i in idx], configs[idx[0]], **batch_options)
            except SolverError as e:      # a failure of the whole group: every member reports it
                out = [e] * len(idx)
            for i, r in zip(idx, out):
                results[i] = r

        lock, taken = threading.Lock(), []
        with ThreadPoolExecutor(max_workers=max(1, min(max_concurrent, len(groups)))) as pool:
            list(pool.map(run, groups.values()))
        for c in taken:
            if c is not first_ctx:
                c.close()
        return results


class CudaRK4Solver(_CudaSolver):
    METHOD = _ffi.RK4
    NAME = "Runge Kutta (RK4)"        # rk4.rs:352-354
    META_SOLVER = "Runge-Kutta 4"     # rk4.rs:343


class CudaEulerSolver(_CudaSolver):
    METHOD = _ffi.EULER
    NAME = "Forward Euler"            # euler.rs:290-292
    META_SOLVER = "Forward Euler"     # euler.rs:282
