"""chromatography_b200 — B200-native (sm_100a) backend for chrom-rs's method-of-lines time integrator.

The product is `libchrom_cuda.so` (hand-written f64 CUDA kernels behind the C ABI of
`include/chrom_cuda.h`).  This package is the Python host-side mirror of the reference's solver
boundary over that ABI; it contains no numerics and no fallback path.
"""
from ._ffi import (ARITH_EXACT, ARITH_FAST, ARITH_FAST_DENSE, ARITH_FAST_STRUCTURED, EULER, RK4, SUMMARY_LEN, WANT_FINAL,
                   WANT_OUTLET, WANT_SNAPSHOTS, WANT_STATUS, WANT_SUMMARY)
from .models import LangmuirMulti, LangmuirSingle, SpeciesParams, TemporalInjection, rust_f64
from .solver import (BatchResult, Context, CudaEulerSolver, CudaRK4Solver, DomainBoundaries, Plan, Scenario,
                     SimulationResult, SolverConfiguration, SolverError, SolverType, default_context, make_cfg,
                     node_index, pack_multi, pack_single, sample_indices, solve_multi_batch, solve_single_batch,
                     status_message, tabulate_injection, time_points)

__all__ = [n for n in dir() if not n.startswith("_")]
