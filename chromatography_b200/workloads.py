"""Synthetic workloads of BASELINE.json (`configs`) — shapes and parameter ranges per SURVEY.md §8(d).

Self-contained SplitMix64 (seed 42) because the reference's bench generator uses Rust's SmallRng,
which cannot be reproduced without Rust.  Parameter ranges: K in U[0.1, 0.8], N in U[1, 3]
(benches/langmuir_performance.rs:312-319); u in U[5e-4, 2e-3], porosity in U[0.3, 0.5] (survey's choice).
Used by bench.py and by tests; carries no numerics.
"""
from __future__ import annotations

import numpy as np

from .models import LangmuirMulti, LangmuirSingle, SpeciesParams, TemporalInjection

_MASK = (1 << 64) - 1


class SplitMix64:
    def __init__(self, seed: int = 42):
        self.s = seed & _MASK

    def next_u64(self) -> int:
        self.s = (self.s + 0x9E3779B97F4A7C15) & _MASK
        z = self.s
        z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & _MASK
        z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & _MASK
        return z ^ (z >> 31)

    def uniform(self, lo: float, hi: float) -> float:
        return lo + (hi - lo) * ((self.next_u64() >> 11) * (1.0 / (1 << 53)))


def splitmix_uniform(n: int, seed: int = 42) -> np.ndarray:
    """n uniforms in [0,1) from SplitMix64, vectorised (same stream as SplitMix64.uniform(0,1))."""
    idx = np.arange(1, n + 1, dtype=np.uint64)
    with np.errstate(over="ignore"):
        z = np.uint64(seed) + idx * np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
    return (z >> np.uint64(11)).astype(np.float64) * (1.0 / (1 << 53))


# ---- C1: README quick start (README.md:44-61, examples/tfa.rs:44-64, 86) ---------------------------
C1 = dict(lambda_=1.2, langmuir_k=0.4, port_number=2.0, porosity=0.4, velocity=1e-3, column_length=0.25,
          nz=100, total_time=600.0, steps=10_000)


def tfa_single(injection=None, nz=None) -> LangmuirSingle:
    inj = injection or TemporalInjection.gaussian(10.0, 3.0, 0.1)
    return LangmuirSingle.new(C1["lambda_"], C1["langmuir_k"], C1["port_number"], C1["porosity"], C1["velocity"],
                              C1["column_length"], nz or C1["nz"], inj)


# ---- C2: acids_multi example (examples/acids_multi.rs:130-174) -------------------------------------
C2 = dict(nz=100, porosity=0.4, velocity=1e-3, column_length=0.25, total_time=1200.0, steps=20_000)


def acids_multi(injection=None) -> LangmuirMulti:
    inj = injection or TemporalInjection.gaussian(10.0, 3.0, 0.1)
    sp = [SpeciesParams.new("Ascorbic", 1.0, 1.1, 2, inj), SpeciesParams.new("Erythorbic", 1.0, 1.7, 2, inj)]
    return LangmuirMulti.new(sp, C2["nz"], C2["porosity"], C2["velocity"], C2["column_length"])


# ---- C3: parameter sweep of TFA-like columns ----------------------------------------------------------
def tfa_sweep_arrays(n_cols: int, seed: int = 42, nz: int = 100) -> dict:
    """Per-column parameter arrays of the C3 sweep (K, N, u, porosity varied; lambda, L fixed)."""
    u = splitmix_uniform(4 * n_cols, seed).reshape(n_cols, 4)
    k = 0.1 + 0.7 * u[:, 0]
    n = 1.0 + 2.0 * u[:, 1]
    vel = 5e-4 + 1.5e-3 * u[:, 2]
    eps = 0.3 + 0.2 * u[:, 3]
    return dict(lambda_=np.full(n_cols, 1.2), langmuir_k=k, port_number=n, fe=(1.0 - eps) / eps, ue=vel / eps,
                dz=np.full(n_cols, 0.25 / float(nz)))


def tfa_sweep_models(n_cols: int, seed: int = 42, nz: int = 100, injection=None) -> list:
    a = tfa_sweep_arrays(n_cols, seed, nz)
    inj = injection or TemporalInjection.gaussian(10.0, 3.0, 0.1)
    return [LangmuirSingle(float(a["lambda_"][i]), float(a["langmuir_k"][i]), float(a["port_number"][i]), 0.25, nz,
                           float(a["dz"][i]), float(a["fe"][i]), float(a["ue"][i]), inj) for i in range(n_cols)]


# ---- C4: one fine-grid TFA column -----------------------------------------------------------------------
def fine_grid_single(nz: int = 1 << 22, steps: int = 1000):
    """C1 parameters on nz cells; dt = 0.5*dz/u_eff(0) (benches/langmuir_performance.rs:253-271)."""
    m = tfa_single(nz=nz)
    n_bar = m.fe / (1.0 + m.fe) * m.port_number
    u_eff0 = m.ue / (1.0 + m.fe * (m.lambda_ + n_bar * m.langmuir_k))
    dt = 0.5 * m.dz / u_eff0
    total_time = dt * steps
    # Gaussian inlet scaled to the simulated window
    m.set_injection(TemporalInjection.gaussian(0.3 * total_time, 0.05 * total_time, 0.1))
    return m, total_time, steps


# ---- C5: LangmuirMulti n = 8 batch ---------------------------------------------------------------------
def multi_batch_models(n_cols: int, n_species: int = 8, nz: int = 512, seed: int = 42, injection=None) -> list:
    """lambda in U[0.8,1.5], K in U[0.1,0.8], N in {1,2,3} (benches/langmuir_performance.rs:312-319, 514-537)."""
    u = splitmix_uniform(3 * n_cols * n_species, seed).reshape(n_cols, n_species, 3)
    inj = injection or TemporalInjection.gaussian(10.0, 3.0, 0.1)
    out = []
    for c in range(n_cols):
        sp = [SpeciesParams.new(f"S{k}", 0.8 + 0.7 * u[c, k, 0], 0.1 + 0.7 * u[c, k, 1],
                                int(round(1.0 + 2.0 * u[c, k, 2])), inj) for k in range(n_species)]
        out.append(LangmuirMulti.new(sp, nz, 0.4, 1e-3, 0.25))
    return out
