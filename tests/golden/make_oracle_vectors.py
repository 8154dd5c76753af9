"""Generates tests/golden/oracle_vectors.npz from the CPU oracle.

These are ORACLE outputs (regression guard + a travelling fixture for the GPU box), not outputs of
the Rust reference, which cannot be built in this environment (no cargo/rustc, no network).
Run: python tests/golden/make_oracle_vectors.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
from oracle import oracle as O  # noqa: E402


def cases() -> dict:
    out = {}
    g = O.TemporalInjection.gaussian(10.0, 3.0, 0.1)
    m = O.LangmuirSingle.new(1.2, 0.4, 2.0, 0.4, 1e-3, 0.25, 100, g)
    for name, method in (("rk4", O.RK4), ("euler", O.EULER)):
        r = O.solve_single(m, method, 600.0, 10_000)
        out[f"c1_{name}_outlet"] = r.outlet[::50].copy()
        out[f"c1_{name}_final"] = r.final_state
    d = O.LangmuirSingle.new(1.2, 0.4, 2.0, 0.4, 1e-3, 0.25, 100, O.TemporalInjection.dirac(0.03, 0.1))
    out["dirac_midstage_rk4_final"] = O.solve_single(d, O.RK4, 60.0, 1000).final_state
    inj = O.TemporalInjection.gaussian(10.0, 3.0, 0.1)
    mm = O.LangmuirMulti.new([O.SpeciesParams.new("Ascorbic", 1.0, 1.1, 2, inj),
                              O.SpeciesParams.new("Erythorbic", 1.0, 1.7, 2, inj)], 100, 0.4, 1e-3, 0.25)
    r = O.solve_multi(mm, O.RK4, 120.0, 2000)
    out["c2_rk4_final"] = r.final_state
    rng = np.random.default_rng(7)
    for n in (3, 4, 5, 8):
        sp = [O.SpeciesParams.new(f"S{k}", 0.8 + 0.1 * k, 0.2 + 0.07 * k, 1 + k % 3, inj) for k in range(n)]
        mn = O.LangmuirMulti.new(sp, 48, 0.4, 1e-3, 0.12)
        y0 = rng.uniform(0.0, 0.05, (48, n))
        out[f"multi_n{n}_rk4_final"] = O.solve_multi(mn, O.RK4, 6.0, 60, y0=y0).final_state
        out[f"multi_n{n}_y0"] = y0
        a = np.eye(n) + mn.fe * mn.jacobian(y0[0])
        out[f"inverse_n{n}"] = O.try_inverse(a)
    return out


if __name__ == "__main__":
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "oracle_vectors.npz")
    np.savez_compressed(path, **cases())
    print("wrote", path, os.path.getsize(path), "bytes")
