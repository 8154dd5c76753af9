"""Times LangmuirSingle plans over a list of shapes (kernel time from the library's CUDA events): which kernel family
ran, cell-stage updates/s, fraction of the FP64 peak measured in the same process (15.25 flop per cell-stage).
  python scripts/single_shapes.py nz,cols,steps[,ENV=VAL...] ..."""
import json
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import chromatography_b200 as cb  # noqa: E402
from chromatography_b200 import workloads as W  # noqa: E402


def main():
    ctx = cb.Context([0])
    peak = ctx.fp64_peak_tflops()
    for spec in sys.argv[1:]:
        parts = spec.split(",")
        nz, cols, steps = (int(x) for x in parts[:3])
        env = dict(p.split("=") for p in parts[3:])
        method = cb.EULER if env.pop("METHOD", "rk4") == "euler" else cb.RK4  # cell-stage updates: 1 per step for Euler
        os.environ.update(env)
        try:
            T = 600.0 * steps / 10000 * (100.0 / nz)  # the C1 time step scaled with dz: same CFL
            models = W.tfa_sweep_models(cols, nz=nz)
            cfg = cb.make_cfg(method, T, steps, nz, 1, 1, 0, cb.ARITH_FAST)
            c, tab = cb.pack_single(models, method, T, steps)
            plan = cb.Plan(ctx, cfg, c, tab, want=cb.WANT_OUTLET | cb.WANT_FINAL)
            desc = plan.describe()
            ms = [plan.execute()["kernel_ms"] for _ in range(4)]
            r = plan.download()
            plan.close()
            best = min(ms[1:])
            units = cols * nz * steps * (4.0 if method == cb.RK4 else 1.0)
            print(json.dumps({"nz": nz, "cols": cols, "steps": steps, "env": env, "kernel": desc, "kernel_ms": best,
                              "cell_stage_per_s": units / (best * 1e-3),
                              "frac_fp64_peak": units * 15.25 / (best * 1e-3) / 1e12 / peak,
                              "bad_status": int((r.status != 0).sum())}), flush=True)
        finally:
            for k in env:
                os.environ.pop(k, None)
    ctx.close()


if __name__ == "__main__":
    main()
