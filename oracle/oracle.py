"""ctypes front-end of the CPU oracle (oracle/chrom_oracle.cpp).

TEST INFRASTRUCTURE ONLY — see the header of chrom_oracle.cpp.  Only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this.
The product package (chromatography_b200) never does.

Names mirror the reference (chrom-rs) so tests read like its own:
`TemporalInjection.dirac/gaussian/rectangle/none` (src/models/injection.rs:349-399),
`LangmuirSingle.new(lambda, langmuir_k, port_number, porosity, velocity, column_length,
spatial_points, injection)` (src/models/langmuir_single.rs:235-276),
`SpeciesParams.new` / `LangmuirMulti.new` (src/models/langmuir_multi.rs:161, 342-410).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass, field

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "liboracle.so")


def build(force: bool = False) -> str:
    """Compile the oracle with its Makefile (g++ -O2 -ffp-contract=off -fopenmp)."""
    src = os.path.join(_HERE, "chrom_oracle.cpp")
    stale = (not os.path.exists(_LIB_PATH)) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src)
    if force or stale:
        subprocess.run(["make", "-C", _HERE, "-s"], check=True)
    return _LIB_PATH


class _Inj(C.Structure):
    _fields_ = [("kind", C.c_int32), ("_pad", C.c_int32), ("p0", C.c_double), ("p1", C.c_double), ("p2", C.c_double)]


class _Single(C.Structure):
    _fields_ = [(n, C.c_double) for n in ("lambda_", "langmuir_k", "port_number", "fe", "ue", "dz")] + [("inj", _Inj)]


class _Species(C.Structure):
    _fields_ = [("lambda_", C.c_double), ("langmuir_k", C.c_double), ("port_number", C.c_uint32),
                ("_pad", C.c_uint32), ("inj", _Inj)]


class _Multi(C.Structure):
    _fields_ = [("fe", C.c_double), ("ue", C.c_double), ("dz", C.c_double), ("stationary_fraction", C.c_double),
                ("n_species", C.c_uint32), ("_pad", C.c_uint32), ("species", C.POINTER(_Species))]


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
        dp = C.POINTER(C.c_double)
        _lib.oracle_injection_evaluate.restype = C.c_double
        _lib.oracle_injection_evaluate.argtypes = [C.POINTER(_Inj), C.c_double]
        _lib.oracle_single_rhs.argtypes = [C.POINTER(_Single), C.c_uint32, dp, C.c_double, dp]
        _lib.oracle_multi_jacobian.argtypes = [C.POINTER(_Multi), dp, dp]
        _lib.oracle_try_inverse.restype = C.c_int32
        _lib.oracle_try_inverse.argtypes = [C.c_uint32, dp, dp]
        _lib.oracle_gemv.argtypes = [C.c_uint32, dp, dp, dp]
        _lib.oracle_multi_rhs.argtypes = [C.POINTER(_Multi), C.c_uint32, dp, C.c_double, dp, C.c_int32]
        _lib.oracle_solve_linear_ode.restype = C.c_int64
        _lib.oracle_solve_linear_ode.argtypes = [C.c_int32, C.c_double, C.c_double, C.c_double, C.c_uint64,
                                                 C.c_uint32, dp, dp, dp]
        _lib.oracle_solve_single.restype = C.c_int64
        _lib.oracle_solve_single.argtypes = [C.POINTER(_Single), C.c_int32, C.c_double, C.c_uint64, C.c_uint32,
                                             dp, dp, dp, dp]
        _lib.oracle_solve_multi.restype = C.c_int64
        _lib.oracle_solve_multi.argtypes = [C.POINTER(_Multi), C.c_int32, C.c_double, C.c_uint64, C.c_uint32,
                                            dp, dp, dp, dp, C.c_int32, C.POINTER(C.c_uint64)]
        _lib.oracle_solve_single_batch.argtypes = [C.POINTER(_Single), C.c_uint64, C.c_int32, C.c_double,
                                                   C.c_uint64, C.c_uint32, dp, dp, dp, C.POINTER(C.c_int64),
                                                   C.c_int32]
        _lib.oracle_solve_multi_batch.argtypes = [C.POINTER(_Multi), C.c_uint64, C.c_int32, C.c_double, C.c_uint64,
                                                  C.c_uint32, dp, dp, dp, C.POINTER(C.c_int64), C.c_int32,
                                                  C.c_int32]
        _lib.oracle_max_threads.restype = C.c_int32
    return _lib


def _dp(a):
    return None if a is None else a.ctypes.data_as(C.POINTER(C.c_double))


EULER, RK4 = 0, 1


@dataclass(frozen=True)
class TemporalInjection:
    kind: int = 0
    p0: float = 0.0
    p1: float = 0.0
    p2: float = 0.0

    @staticmethod
    def none():
        return TemporalInjection(0)

    @staticmethod
    def dirac(time, amount):
        return TemporalInjection(1, time, amount)

    @staticmethod
    def gaussian(center, width, peak_concentration):
        return TemporalInjection(2, center, width, peak_concentration)

    @staticmethod
    def rectangle(start, end, concentration):
        assert end > start, "Rectangle end must be > start"
        return TemporalInjection(3, start, end, concentration)

    def _c(self):
        return _Inj(self.kind, 0, self.p0, self.p1, self.p2)

    def evaluate(self, t: float) -> float:
        return lib().oracle_injection_evaluate(C.byref(self._c()), float(t))


@dataclass
class LangmuirSingle:
    lambda_: float
    langmuir_k: float
    port_number: float
    fe: float
    ue: float
    dz: float
    n_points: int
    injection: TemporalInjection

    @staticmethod
    def new(lambda_, langmuir_k, port_number, porosity, velocity, column_length, spatial_points, injection):
        # langmuir_single.rs:261-263
        fe = (1.0 - porosity) / porosity
        ue = velocity / porosity
        dz = column_length / float(spatial_points)
        return LangmuirSingle(lambda_, langmuir_k, port_number, fe, ue, dz, spatial_points, injection)

    def _c(self):
        return _Single(self.lambda_, self.langmuir_k, self.port_number, self.fe, self.ue, self.dz,
                       self.injection._c())

    def compute_physics(self, c: np.ndarray, t: float) -> np.ndarray:
        c = np.ascontiguousarray(c, dtype=np.float64)
        out = np.empty_like(c)
        lib().oracle_single_rhs(C.byref(self._c()), self.n_points, _dp(c), float(t), _dp(out))
        return out


@dataclass
class SpeciesParams:
    name: str
    lambda_: float
    langmuir_k: float
    port_number: int
    injection: TemporalInjection

    @staticmethod
    def new(name, lambda_, langmuir_k, port_number, injection):
        return SpeciesParams(name, lambda_, langmuir_k, int(port_number), injection)


@dataclass
class LangmuirMulti:
    species: list
    n_points: int
    fe: float
    ue: float
    dz: float
    stationary_fraction: float
    _keep: list = field(default_factory=list, repr=False)

    @staticmethod
    def new(species, n_points, porosity, velocity, column_length):
        # langmuir_multi.rs:394-397
        dz = column_length / float(n_points)
        fe = (1.0 - porosity) / porosity
        ue = velocity / porosity
        sf = 1.0 - porosity
        return LangmuirMulti(list(species), n_points, fe, ue, dz, sf)

    def n_species(self):
        return len(self.species)

    def _c(self):
        arr = (_Species * len(self.species))(*[
            _Species(s.lambda_, s.langmuir_k, s.port_number, 0, s.injection._c()) for s in self.species])
        self._keep.append(arr)
        return _Multi(self.fe, self.ue, self.dz, self.stationary_fraction, len(self.species), 0,
                      C.cast(arr, C.POINTER(_Species)))

    def jacobian(self, c) -> np.ndarray:
        """n x n Jacobian M (M[i, j] = dCbar_i/dC_j)."""
        n = self.n_species()
        c = np.ascontiguousarray(c, dtype=np.float64)
        out = np.empty(n * n)
        lib().oracle_multi_jacobian(C.byref(self._c()), _dp(c), _dp(out))
        return out.reshape(n, n).T.copy()  # oracle is column-major

    def inverse_propagation(self, c):
        n = self.n_species()
        mat = np.eye(n) + self.fe * self.jacobian(c)
        return try_inverse(mat)

    def compute_physics(self, c: np.ndarray, t: float, cells_parallel=False) -> np.ndarray:
        """c: [n_points, n_species] (reference matrix shape); returns the same shape."""
        cf = np.asfortranarray(c, dtype=np.float64)
        out = np.empty_like(cf, order="F")
        lib().oracle_multi_rhs(C.byref(self._c()), self.n_points, _dp(cf), float(t), _dp(out),
                               int(cells_parallel))
        return np.array(out, order="C")


def from_model(model):
    """Oracle twin of a host-side model object (duck-typed on the reference's field names, so this
    module never imports the product package)."""
    def inj(i):
        return TemporalInjection(i.kind, i.p0, i.p1, i.p2)
    if hasattr(model, "species"):
        sp = [SpeciesParams(s.name, s.lambda_, s.langmuir_k, s.port_number, inj(s.injection)) for s in model.species]
        return LangmuirMulti(sp, model.n_points, model.fe, model.ue, model.dz, model.stationary_fraction)
    return LangmuirSingle(model.lambda_, model.langmuir_k, model.port_number, model.fe, model.ue, model.dz,
                          model.n_points, inj(model.injection))


def try_inverse(mat: np.ndarray):
    """nalgebra `try_inverse` restatement.  Returns None when nalgebra would."""
    n = mat.shape[0]
    a = np.asfortranarray(mat, dtype=np.float64)
    out = np.empty_like(a, order="F")
    ok = lib().oracle_try_inverse(n, _dp(a), _dp(out))
    return np.array(out, order="C") if ok else None


def gemv(mat: np.ndarray, x: np.ndarray) -> np.ndarray:
    n = mat.shape[0]
    a = np.asfortranarray(mat, dtype=np.float64)
    x = np.ascontiguousarray(x, dtype=np.float64)
    y = np.empty(n)
    lib().oracle_gemv(n, _dp(a), _dp(x), _dp(y))
    return y


@dataclass
class OracleResult:
    time_points: np.ndarray
    outlet: np.ndarray            # [steps+1] (single) or [n_species, steps+1] (multi)
    final_state: np.ndarray       # [nz] or [nz, n_species]
    trajectory: np.ndarray | None  # [steps+1, nz] or [steps+1, nz, n_species]
    status: int                   # 0 ok, +s NaN at step s, -s Inf at step s
    n_singular: int = 0

    def error(self):
        return validate_state_message(self.status)


def validate_state_message(status: int):
    """Error strings of src/solver/mod.rs:527-548 for a nonzero status."""
    if status > 0:
        return (f"NaN detected in Concentration at step {status}. This indicates numerical instability. "
                "Try reducing time step (increase time_steps parameter).")
    if status < 0:
        return (f"Infinity detected in Concentration at step {-status}. This indicates numerical overflow. "
                "Try reducing time step or check physics model for division by zero.")
    return None


def time_points(total_time: float, steps: int) -> np.ndarray:
    """[0.0] + [(s as f64 + 1.0) * dt] — rk4.rs:254, 327."""
    dt = total_time / float(steps)
    return np.concatenate([[0.0], (np.arange(steps, dtype=np.float64) + 1.0) * dt])


def solve_single(model: LangmuirSingle, method: int, total_time: float, steps: int, y0=None,
                 want_trajectory=False) -> OracleResult:
    nz = model.n_points
    outlet = np.zeros(steps + 1)
    final = np.zeros(nz)
    traj = np.zeros((steps + 1, nz)) if want_trajectory else None
    y0c = None if y0 is None else np.ascontiguousarray(y0, dtype=np.float64)
    st = lib().oracle_solve_single(C.byref(model._c()), method, float(total_time), steps, nz, _dp(y0c), _dp(traj),
                                   _dp(outlet), _dp(final))
    return OracleResult(time_points(total_time, steps), outlet, final, traj, int(st))


def solve_multi(model: LangmuirMulti, method: int, total_time: float, steps: int, y0=None, want_trajectory=False,
                cells_parallel=False) -> OracleResult:
    nz, n = model.n_points, model.n_species()
    outlet = np.zeros((n, steps + 1))
    final = np.zeros((n, nz))
    traj = np.zeros((steps + 1, n, nz)) if want_trajectory else None
    y0c = None if y0 is None else np.ascontiguousarray(np.asarray(y0, dtype=np.float64).T)  # -> [n][nz]
    ns = C.c_uint64(0)
    st = lib().oracle_solve_multi(C.byref(model._c()), method, float(total_time), steps, nz, _dp(y0c), _dp(traj),
                                  _dp(outlet), _dp(final), int(cells_parallel), C.byref(ns))
    return OracleResult(time_points(total_time, steps), outlet, final.T.copy(),
                        None if traj is None else traj.transpose(0, 2, 1).copy(), int(st), int(ns.value))


def solve_linear_ode(method: int, a: float, b: float, total_time: float, steps: int, y0) -> tuple:
    """Mock-model integrator check: dy/dt = a*y + b (ExponentialDecay / ConstantGrowth mocks)."""
    y0c = np.ascontiguousarray(y0, dtype=np.float64)
    n = y0c.size
    traj = np.zeros((steps + 1, n))
    final = np.zeros(n)
    st = lib().oracle_solve_linear_ode(method, a, b, float(total_time), steps, n, _dp(y0c), _dp(traj), _dp(final))
    return traj, final, int(st)


def solve_single_batch(models, method, total_time, steps, n_threads=0, want_outlet=True, y0=None):
    """One column per OpenMP thread.  Returns (outlet [n_cols, steps+1] | None, final [n_cols, nz], status)."""
    n_cols = len(models)
    nz = models[0].n_points
    arr = (_Single * n_cols)(*[m._c() for m in models])
    outlet = np.zeros((n_cols, steps + 1)) if want_outlet else None
    final = np.zeros((n_cols, nz))
    status = np.zeros(n_cols, dtype=np.int64)
    y0c = None if y0 is None else np.ascontiguousarray(y0, dtype=np.float64)
    lib().oracle_solve_single_batch(arr, n_cols, method, float(total_time), steps, nz, _dp(y0c), _dp(outlet),
                                    _dp(final), status.ctypes.data_as(C.POINTER(C.c_int64)), int(n_threads))
    return outlet, final, status


def solve_multi_batch(models, method, total_time, steps, n_threads=0, want_outlet=True, cells_parallel=False):
    """Returns (outlet [n_cols, n, steps+1] | None, final [n_cols, n, nz], status)."""
    n_cols = len(models)
    nz, n = models[0].n_points, models[0].n_species()
    arr = (_Multi * n_cols)(*[m._c() for m in models])
    outlet = np.zeros((n_cols, n, steps + 1)) if want_outlet else None
    final = np.zeros((n_cols, n, nz))
    status = np.zeros(n_cols, dtype=np.int64)
    lib().oracle_solve_multi_batch(arr, n_cols, method, float(total_time), steps, nz, None, _dp(outlet), _dp(final),
                                   status.ctypes.data_as(C.POINTER(C.c_int64)), int(n_threads),
                                   int(cells_parallel))
    return outlet, final, status


def max_threads() -> int:
    return int(lib().oracle_max_threads())


class inverse_variant:
    """Context manager: run the LangmuirMulti oracle with a sensitivity variant of its n x n step
    (chrom_oracle.cpp g_inverse_variant: 1 LU inverse with true divisions for every n, 2 the same in long double,
    3 direct Gaussian-elimination solve without an inverse).  Used by tests/test_oracle_sensitivity.py only."""

    def __init__(self, v: int):
        self.v = int(v)

    def __enter__(self):
        lib().oracle_set_inverse_variant(self.v)
        return self

    def __exit__(self, *a):
        lib().oracle_set_inverse_variant(0)


# --- chromatogram analytics of the reference's validation suite (validation/dissimilarity.rs,
# validation/main.rs:190-206): test-side helpers, numpy ---------------------------------------
def trapezoid(times, values):
    t = np.asarray(times)
    v = np.asarray(values)
    return float(np.sum(0.5 * (v[:-1] + v[1:]) * (t[1:] - t[:-1])))


def first_moment(times, concs):
    t = np.asarray(times)
    c = np.asarray(concs)
    return trapezoid(t, t * c) / trapezoid(t, c)


def peak_sigma(times, concs):
    t = np.asarray(times)
    c = np.asarray(concs)
    tm = first_moment(t, c)
    return float(np.sqrt(trapezoid(t, (t - tm) ** 2 * c) / trapezoid(t, c)))


def rsf(t1, c1, t2, c2):
    """validation/dissimilarity.rs `rsf`: L1 distance of area-normalised signals on the merged grid."""
    t1, c1, t2, c2 = map(np.asarray, (t1, c1, t2, c2))
    y1 = c1 / trapezoid(t1, c1)
    y2 = c2 / trapezoid(t2, c2)
    min_dt = min(np.min(np.abs(np.diff(t1))), np.min(np.abs(np.diff(t2))))
    tol = min_dt * 0.01
    merged = np.sort(np.concatenate([t1, t2]))
    keep = np.concatenate([[True], np.diff(merged) >= tol])
    grid = merged[keep]
    return trapezoid(grid, np.abs(np.interp(grid, t1, y1) - np.interp(grid, t2, y2)))
