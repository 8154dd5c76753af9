"""Times LangmuirMulti plans over a list of shapes (kernel time from the library's CUDA events) and prints one
JSON line per shape: which kernel family ran, cell-stage updates/s, algorithmic TFLOP/s.  Used to place the
tiled / warp-per-cell kernels (multi_tiled.cu) against the resident ones; environment knobs select the path.

  python scripts/multi_shapes.py n,nz,cols,steps[,ENV=VAL...] ...
"""
import json
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import chromatography_b200 as cb  # noqa: E402
from chromatography_b200 import workloads as W  # noqa: E402


def flops_per_cell_stage(n):
    # SURVEY §8d accounting, LU-solve formulation: jacobian/matrix build, LU, two triangular solves, stage algebra
    lu = 2.0 * n ** 3 / 3.0 - n ** 2 / 2.0 - n / 6.0 + n
    return 3 * n + 5 + 3 * n * n + lu + 2 * n * n + 2 * n + 3.25 * n


def main():
    ctx = cb.Context([0])
    for spec in sys.argv[1:]:
        parts = spec.split(",")
        n, nz, cols, steps = (int(x) for x in parts[:4])
        env = dict(p.split("=") for p in parts[4:])
        method = cb.EULER if env.pop("METHOD", "rk4") == "euler" else cb.RK4  # cell-stage updates: 1 per step for Euler
        arith = {"exact": cb.ARITH_EXACT, "fast": cb.ARITH_FAST, "dense": cb.ARITH_FAST_DENSE}[env.pop("ARITH", "fast")]
        for k, v in env.items():
            os.environ[k] = v
        try:
            T = 600.0 * steps / 4096 * (512.0 / nz)  # the C5 time step scaled with dz: same CFL
            models = W.multi_batch_models(cols, n, nz)
            cfg = cb.make_cfg(method, T, steps, nz, n, 1, 0, arith)
            c, sp, tab = cb.pack_multi(models, method, T, steps)
            plan = cb.Plan(ctx, cfg, c, tab, species=sp, want=cb.WANT_OUTLET | cb.WANT_FINAL)
            desc = plan.describe()
            ms = []
            for it in range(4):
                t = plan.execute()
                ms.append(t["kernel_ms"])
            r = plan.download()
            plan.close()
            best = min(ms[1:])
            units = cols * nz * steps * (4.0 if method == cb.RK4 else 1.0)
            print(json.dumps({"n": n, "nz": nz, "cols": cols, "steps": steps, "env": env, "kernel": desc,
                              "kernel_ms": best, "cell_stage_per_s": units / (best * 1e-3),
                              "tflops_alg": units * flops_per_cell_stage(n) / (best * 1e-3) / 1e12,
                              "bad_status": int((r.status != 0).sum())}), flush=True)
        finally:
            for k in env:
                os.environ.pop(k, None)
    ctx.close()


if __name__ == "__main__":
    main()
