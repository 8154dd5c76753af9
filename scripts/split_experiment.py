"""Experiment: a batch that cannot fill the machine (8192 C3 columns on one B200: 1.73 warps per SM sub-partition at
25 cells per lane) split into two sub-batches with different cells per lane, launched concurrently from two host
threads (two contexts = two streams), so that every sub-partition gets one warp of each kind.
  python scripts/split_experiment.py colsA,MA colsB,MB [steps]"""
import json
import os
import sys
import threading
import time

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import chromatography_b200 as cb  # noqa: E402
from chromatography_b200 import workloads as W  # noqa: E402


def make_plan(ctx, cols, m, steps):
    os.environ["CHROM_FORCE_M"] = str(m)
    models = W.tfa_sweep_models(cols, nz=100)
    T = 600.0 * steps / 10000
    cfg = cb.make_cfg(cb.RK4, T, steps, 100, 1, 1, 0, cb.ARITH_FAST)
    c, tab = cb.pack_single(models, cb.RK4, T, steps)
    p = cb.Plan(ctx, cfg, c, tab, want=cb.WANT_OUTLET | cb.WANT_FINAL)
    os.environ.pop("CHROM_FORCE_M")
    return p


def main():
    parts = [tuple(int(x) for x in a.split(",")) for a in sys.argv[1:3]]
    steps = int(sys.argv[3]) if len(sys.argv) > 3 else 10000
    ctxs = [cb.Context([0]) for _ in parts]
    plans = [make_plan(c, cols, m, steps) for c, (cols, m) in zip(ctxs, parts)]
    descs = [p.describe() for p in plans]
    for p in plans:
        p.execute()
    alone = [min(p.execute()["kernel_ms"] for _ in range(2)) for p in plans]
    walls = []
    for _ in range(4):
        bar = threading.Barrier(len(plans) + 1)
        res = [None] * len(plans)

        def run(i):
            bar.wait()
            res[i] = plans[i].execute()["kernel_ms"]

        ths = [threading.Thread(target=run, args=(i,)) for i in range(len(plans))]
        for t in ths:
            t.start()
        bar.wait()
        t0 = time.perf_counter()
        for t in ths:
            t.join()
        walls.append(((time.perf_counter() - t0) * 1e3, list(res)))
    print(json.dumps({"parts": parts, "steps": steps, "kernels": descs, "alone_ms": alone,
                      "concurrent_wall_ms_and_kernel_ms": walls}), flush=True)


if __name__ == "__main__":
    main()
