"""bench.py — the headline metric of BASELINE.json on B200:
RK4 cell-stage updates/sec on a batch of Langmuir columns.

A "step" is one pass of the hot path over one batch: the complete solve (all time steps) of
every column of the workload.  Default workload (N=1): config 3 of BASELINE.json, the one the
metric is quoted on — 65 536 TFA-like LangmuirSingle columns, nz=100, RK4, 10 000 steps per
GPU (weak scaling: every rank gets its own 65 536 columns; columns are independent, no collective).

  value   : whole-job cell-stage updates/s with inputs resident in HBM (plan executed K times,
            kernel time from CUDA events on the launch stream, max over ranks)
  e2e     : the same through chrom_cuda_solve_single_batch with pinned HOST buffers: parameter +
            injection-table H2D, kernel, outlet chromatograms + final state + status D2H, every step
  roofline: algorithmic flops (15.25 per cell-stage) / kernel time vs the FP64 (DFMA)
            peak measured in this run (MEASURED_PEAKS.json has no FP64 entry)
  cpu_baseline: the C++ oracle port of the reference arithmetic on the box's host cores, bounded sample

`--impl reference` times the reference's CPU path (the oracle port: the Rust reference cannot be
built here — no cargo/rustc) with all host threads on a bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "RK4 cell-stage updates/sec, batched Langmuir columns"
UNIT = "cell-stage updates/s"

# algorithmic flops per cell-stage (a divide counts as one flop)
FLOPS_PER_CELL_STAGE = {"single": 15.25, "multi2": 53.5, "multi8": 776.0}


def workload_spec(name: str) -> dict:
    if name == "c3":
        return dict(kind="single", n_cols=65536, nz=100, n_species=1, total_time=600.0, steps=10_000,
                    label="C3: 65,536 TFA-like LangmuirSingle columns (K,N,u,eps varied), nz=100, RK4 600 s / 10,000 steps",
                    flops=FLOPS_PER_CELL_STAGE["single"])
    if name == "c1":
        return dict(kind="single", n_cols=1, nz=100, n_species=1, total_time=600.0, steps=10_000,
                    label="C1: one TFA LangmuirSingle column, nz=100, RK4 600 s / 10,000 steps",
                    flops=FLOPS_PER_CELL_STAGE["single"])
    if name == "c5":
        return dict(kind="multi", n_cols=4096, nz=512, n_species=8, total_time=600.0, steps=4096,
                    label="C5: 4,096 LangmuirMulti columns, n=8, nz=512, RK4 600 s / 4,096 steps",
                    flops=FLOPS_PER_CELL_STAGE["multi8"])
    if name == "c2":
        return dict(kind="multi", n_cols=1, nz=100, n_species=2, total_time=1200.0, steps=20_000,
                    label="C2: ascorbic/erythorbic LangmuirMulti n=2, nz=100, RK4 1200 s / 20,000 steps",
                    flops=FLOPS_PER_CELL_STAGE["multi2"])
    if name == "c4":
        return dict(kind="single", n_cols=1, nz=1 << 22, n_species=1, total_time=None, steps=1000,
                    label="C4: one fine-grid TFA column nz=2^22, RK4 1,000 steps (streaming path)",
                    flops=FLOPS_PER_CELL_STAGE["single"])
    raise SystemExit(f"unknown workload {name}")


def units_of(spec: dict, n_cols: int) -> float:
    return float(n_cols) * spec["nz"] * spec["steps"] * 4.0


def measured_peaks() -> dict:
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return json.load(f)
    except (OSError, ValueError):
        return {}


def ncu_traffic(family: str, workload: str):
    """dram__bytes_read.sum + dram__bytes_write.sum of the dominant kernel, per launch, from the committed
    `ncu --set full` capture of this workload (profiles/ncu_traffic.json, written by scripts/ncu_traffic.py).
    Not measured in this run: a number taken under a profiler is never a bench value, this is a counter."""
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
            t = json.load(f)
        e = t[workload]
        if e.get("kernel_family") != family:
            return None, f"profiles/ncu_traffic.json has {e.get('kernel_family')} for {workload}, this run used {family}"
        return e["dram_bytes_per_launch"], e["source"]
    except (OSError, ValueError, KeyError):
        return None, "no ncu capture committed for this workload"


# ---- host cores actually usable --------------------------------------------------------------------
def host_cores() -> int:
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


# ---- CPU arm (oracle port of the reference arithmetic) ------------------------------------------
def cpu_sample(spec: dict, n_cols: int, threads: int, steps: int | None = None):
    """Time the oracle on `n_cols` columns of the workload (one column per OpenMP thread; a lone
    column runs on one thread: LangmuirSingle has no parallel path in the reference).
    Returns (cell-stage/s, seconds)."""
    from chromatography_b200 import workloads as W
    from oracle import oracle as O
    to_oracle = O.from_model
    steps = steps or spec["steps"]
    if spec["kind"] == "single":
        if spec["nz"] > 100000:
            m, total_time, _ = W.fine_grid_single(spec["nz"], spec["steps"])
            models = [to_oracle(m)]
            total_time = total_time * steps / spec["steps"]
        else:
            models = [to_oracle(m) for m in W.tfa_sweep_models(n_cols, nz=spec["nz"])]
            total_time = spec["total_time"] * steps / spec["steps"]
        t0 = time.perf_counter()
        O.solve_single_batch(models, O.RK4, total_time, steps, n_threads=threads, want_outlet=True)
        dt = time.perf_counter() - t0
    else:
        models = [to_oracle(m) for m in W.multi_batch_models(n_cols, spec["n_species"], spec["nz"])]
        total_time = spec["total_time"] * steps / spec["steps"]
        t0 = time.perf_counter()
        O.solve_multi_batch(models, O.RK4, total_time, steps, n_threads=threads, want_outlet=True)
        dt = time.perf_counter() - t0
    return len(models) * spec["nz"] * steps * 4.0 / dt, dt


def cpu_sample_size(spec: dict, cores: int) -> tuple:
    """(columns, steps) of a bounded sample: roughly 10-20 s of CPU work spread over all cores."""
    per_core_rate = 1.0e8 if spec["kind"] == "single" else 2.0e8 / (spec["n_species"] ** 2.2)  # cell-stage/s, rough
    budget = 4.0 * per_core_rate * cores  # ~4 s wall
    steps = spec["steps"]
    per_col = spec["nz"] * steps * 4.0
    if per_col > budget:  # very long column (C4): cut steps instead
        return 1, max(8, int(budget / (spec["nz"] * 4.0)))
    n = max(cores, int(budget / per_col))
    n = (n // cores) * cores
    return min(n, spec["n_cols"]), steps


def run_reference_arm(args, spec, rank):
    """--impl reference: the reference's CPU implementation of the path, all host threads."""
    if rank != 0:
        return
    cores = host_cores()
    n, steps = cpu_sample_size(spec, cores)
    for _ in range(max(1, min(args.warmup, 1))):
        cpu_sample(spec, max(cores, n // 4), cores, steps)
    vals, secs = [], []
    for _ in range(args.steps):
        v, s = cpu_sample(spec, n, cores, steps)
        vals.append(v)
        secs.append(s)
    units_sum = n * spec["nz"] * steps * 4.0 * len(vals)
    value = units_sum / sum(secs)
    sample = f"{n} of {spec['n_cols']} columns x {steps} of {spec['steps']} time steps per step, one column per OpenMP thread"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * sum(secs) / len(secs), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": spec["label"], "note": "CPU arm: C++ oracle port of the reference arithmetic "
                   "(Rust reference cannot be built: no cargo/rustc in the image)"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ---- clocks sampler ------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--id={gpu_index}", f"--query-gpu={self.Q}",
                                       "--format=csv,noheader,nounits", "-lms", "100"], stdout=self.f,
                                      stderr=subprocess.DEVNULL)
        except OSError:
            self.p = None

    def stop(self) -> dict:
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, mx, pw, reasons = [], [], [], set()
        for ln in self.f.read().splitlines():
            c = [x.strip() for x in ln.split(",")]
            if len(c) < 9:
                continue
            try:
                sm.append(float(c[1]))
                mx.append(float(c[2]))
                pw.append(float(c[3]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), c[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        self.f.close()
        os.unlink(self.f.name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        hi = max(pw)
        load = [s for s, p in zip(sm, pw) if p >= 0.5 * hi] or sm
        return {"sm_mhz": statistics.median(load), "sm_max_mhz": max(mx), "reasons": sorted(reasons),
                "power_w_max": hi, "samples": len(sm)}


# ---- one process per GPU: the only cross-rank traffic is the barrier and two scalar reductions ----------
class Ranks:
    """torch.distributed plumbing of the bench (no data-path collective: columns are independent).
    backend "nccl" on the GPU box, "gloo" in the CPU tests (tests/test_dist_gloo.py)."""

    def __init__(self, backend: str, local_rank: int = 0):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.distributed = self.world > 1
        self.cuda = backend == "nccl"
        if self.distributed:
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            if self.cuda:
                dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
            else:
                dist.init_process_group(backend)

    def barrier(self):
        if self.distributed:
            self.dist.barrier()
        if self.cuda:
            self.torch.cuda.synchronize()

    def _reduce(self, x: float, op) -> float:
        if not self.distributed:
            return x
        t = self.torch.tensor([x], dtype=self.torch.float64, device="cuda" if self.cuda else "cpu")
        self.dist.all_reduce(t, op=op)
        return float(t.item())

    def max(self, x: float) -> float:
        return self._reduce(x, self.dist.ReduceOp.MAX)

    def sum(self, x: float) -> float:
        return self._reduce(x, self.dist.ReduceOp.SUM)

    def close(self):
        if self.distributed:
            self.dist.destroy_process_group()


def aggregate_value(ranks: "Ranks", units_this_rank: float, device_ms_this_rank: float) -> float:
    """Whole-job throughput: units all ranks processed / max-over-ranks of the device time."""
    return ranks.sum(units_this_rank) / (ranks.max(device_ms_this_rank) * 1e-3)


# ---- GPU arm ------------------------------------------------------------------------------------------
def pinned_array(lib, shape, dtype=np.float64):
    n = int(np.prod(shape))
    p = C.c_void_p()
    rc = lib.chrom_cuda_host_alloc(n * np.dtype(dtype).itemsize, C.byref(p))
    if rc != 0:
        raise RuntimeError("chrom_cuda_host_alloc failed")
    buf = (C.c_byte * (n * np.dtype(dtype).itemsize)).from_address(p.value)
    return np.frombuffer(buf, dtype=dtype).reshape(shape), p


def build_inputs(spec, n_cols, rank):
    """Synthetic inputs of the workload for this rank (seed 42 + rank), packed for the C ABI."""
    import chromatography_b200 as cb
    from chromatography_b200 import _ffi, workloads as W
    steps = spec["steps"]
    if spec["kind"] == "single":
        if spec["nz"] > 100000:
            m, total_time, _ = W.fine_grid_single(spec["nz"], steps)
            cols, tables = cb.pack_single([m], cb.RK4, total_time, steps)
            return dict(cols=cols, tables=tables, species=None, total_time=total_time)
        a = W.tfa_sweep_arrays(n_cols, seed=42 + rank, nz=spec["nz"])
        cols = (_ffi.SingleCol * n_cols)()
        arr = np.frombuffer(cols, dtype=np.float64).reshape(n_cols, 7)
        for j, k in enumerate(["lambda_", "langmuir_k", "port_number", "fe", "ue", "dz"]):
            arr[:, j] = a[k]
        arr[:, 6] = 0.0  # inj_table = 0 (shared Gaussian(10, 3, 0.1) table)
        tables = cb.tabulate_injection(cb.TemporalInjection.gaussian(10.0, 3.0, 0.1), cb.RK4, spec["total_time"],
                                       steps)[None]
        return dict(cols=cols, tables=np.ascontiguousarray(tables), species=None, total_time=spec["total_time"])
    models = (W.multi_batch_models(n_cols, spec["n_species"], spec["nz"], seed=42 + rank) if spec["n_cols"] > 1
              else [W.acids_multi()])
    cols, species, tables = cb.pack_multi(models, cb.RK4, spec["total_time"], steps)
    return dict(cols=cols, tables=tables, species=species, total_time=spec["total_time"])


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c3", choices=["c1", "c2", "c3", "c4", "c5"])
    ap.add_argument("--cols", type=int, default=0, help="override columns per GPU (debug only; invalidates the metric)")
    ap.add_argument("--arith", default="fast", choices=["fast", "dense"],
                    help="multi-species workloads: closed-form structured LU (fast, the library default) or the "
                         "dense register LU with partial pivoting (dense)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-numa", action="store_true", help="do not bind the rank to its GPU's NUMA node")
    ap.add_argument("--pin-cpus", action="store_true", help="experiment: pin each local rank to its share of the CPUs")
    args = ap.parse_args()

    spec = workload_spec(args.workload)
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        run_reference_arm(args, spec, rank)
        return

    W_ = max(args.warmup, 3)  # timing rule: at least 3 warm-up steps
    import torch
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: chromatography_b200 has no CPU fallback")
    torch.cuda.set_device(local_rank)
    ranks = Ranks("nccl", local_rank)
    distributed, barrier, max_over_ranks, sum_over_ranks = ranks.distributed, ranks.barrier, ranks.max, ranks.sum

    import chromatography_b200 as cb
    from chromatography_b200 import _ffi
    lib = _ffi.lib()
    ctx = cb.Context([local_rank])
    # multi-GPU hosts: keep this rank's thread and its page-locked buffers on the NUMA node of its GPU
    # (restored before the CPU baseline, which uses every host core)
    affinity0 = os.sched_getaffinity(0)
    numa = ctx.bind_host_thread(0) if not args.no_numa else (-1, 0)
    if args.pin_cpus and world > 1:  # experiment: an equal share of the allowed CPUs per local rank
        cpus = sorted(affinity0)
        share = max(1, len(cpus) // world)
        os.sched_setaffinity(0, set(cpus[local_rank * share:(local_rank + 1) * share]))

    n_cols = args.cols or spec["n_cols"]
    inp = build_inputs(spec, n_cols, rank)
    n_cols = len(inp["cols"])
    units_rank = units_of(spec, n_cols)
    arith = cb.ARITH_FAST
    if spec["kind"] == "multi":
        if args.arith == "dense":
            arith = cb.ARITH_FAST_DENSE  # the dense register LU: SURVEY 8d's 776 flops per cell-stage at n = 8
        else:
            # FAST = the closed-form structured elimination: per species b 2, S 2, p 1, alpha 2, 1/alpha 1, z 1, v 2,
            # q.v 2, q.z 2, x 2, stage algebra 3.25 = 20.25 flops; per cell 1+S, 1/D, 1/D^2, 1-q.v, f 2, pivot bound 4 = 10
            spec = dict(spec, flops=20.25 * spec["n_species"] + 10.0)
    cfg = cb.make_cfg(cb.RK4, inp["total_time"], spec["steps"], spec["nz"], spec["n_species"], 1, 0, arith)
    want = cb.WANT_OUTLET | cb.WANT_FINAL | cb.WANT_STATUS

    # ---- value: inputs resident in HBM --------------------------------------------------------------
    plan = cb.Plan(ctx, cfg, inp["cols"], inp["tables"], species=inp["species"], want=want)
    desc = plan.describe()
    for _ in range(W_):
        plan.execute()
    peak_tflops = ctx.fp64_peak_tflops() if rank == 0 else 0.0
    barrier()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    kernel_ms, launches = [], 0
    wall0 = time.perf_counter()
    for _ in range(args.steps):
        ctx.flush_l2()  # cold L2 between timed iterations (outside the event-timed kernel region)
        t = plan.execute()
        kernel_ms.append(t["kernel_ms"])
        launches += int(t["kernel_launches"])
    barrier()
    wall_ms = (time.perf_counter() - wall0) * 1e3
    clocks = sampler.stop() if sampler else None
    t_dev_ms = max_over_ranks(sum(kernel_ms))
    value = aggregate_value(ranks, units_rank * args.steps, sum(kernel_ms))
    launches_value = launches
    st = plan.download().status
    n_bad = int(np.count_nonzero(st))
    plan.close()

    # ---- e2e: host buffers through the C-ABI solve call ----------------------------------------------
    e2e = None
    if not args.no_e2e:
        n_out = spec["steps"] + 1
        sp = (spec["n_species"],) if spec["kind"] == "multi" else ()
        outlet, p1 = pinned_array(lib, (n_cols,) + sp + (n_out,))
        final, p2 = pinned_array(lib, (n_cols,) + sp + (spec["nz"],))
        status, p3 = pinned_array(lib, (n_cols,), np.int64)
        tables = inp["tables"]
        dp = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))  # noqa: E731

        def call():
            if spec["kind"] == "single":
                rc = lib.chrom_cuda_solve_single_batch(ctx._h, C.byref(cfg), inp["cols"], n_cols, dp(tables),
                                                       tables.shape[0], None, dp(outlet), None, dp(final), None,
                                                       status.ctypes.data_as(C.POINTER(C.c_int64)))
            else:
                rc = lib.chrom_cuda_solve_multi_batch(ctx._h, C.byref(cfg), inp["cols"], n_cols, inp["species"],
                                                      len(inp["species"]), dp(tables), tables.shape[0], None,
                                                      dp(outlet), None, dp(final), None,
                                                      status.ctypes.data_as(C.POINTER(C.c_int64)))
            ctx.check(rc)
            return ctx.timing()

        call()  # warm-up (first touch of the pinned buffers)
        barrier()
        e_steps = max(1, min(args.steps, 3))
        t0 = time.perf_counter()
        h2d = d2h = 0
        for _ in range(e_steps):
            tm = call()
            h2d, d2h = int(tm["h2d_bytes"]), int(tm["d2h_bytes"])
            launches += int(tm["kernel_launches"])
        barrier()
        e_ms = max_over_ranks((time.perf_counter() - t0) * 1e3)
        e2e = {"value": sum_over_ranks(units_rank) * e_steps / (e_ms * 1e-3), "unit": UNIT,
               "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "ms_per_step": e_ms / e_steps,
               "steps": e_steps, "api": "chrom_cuda_solve_%s_batch, pinned host buffers" % spec["kind"],
               "host_numa_node": numa[0], "host_cpus_bound": numa[1]}
        # The same call in the sweep mode the ABI offers: per-column on-device summaries
        # (area, first moment, max, time of max) + final state + status instead of the full-rate outlet
        # series.  Reported beside e2e, never instead of it: the D2H volume is what differs.
        if spec["kind"] == "single" and spec["nz"] <= 8192:
            summ, p4 = pinned_array(lib, (n_cols, 4))

            def call_summary():
                rc = lib.chrom_cuda_solve_single_batch(ctx._h, C.byref(cfg), inp["cols"], n_cols, dp(tables),
                                                       tables.shape[0], None, None, None, dp(final), dp(summ),
                                                       status.ctypes.data_as(C.POINTER(C.c_int64)))
                ctx.check(rc)
                return ctx.timing()

            call_summary()
            barrier()
            t0 = time.perf_counter()
            for _ in range(e_steps):
                tm = call_summary()
                launches += int(tm["kernel_launches"])
            barrier()
            s_ms = max_over_ranks((time.perf_counter() - t0) * 1e3)
            e2e["summary_mode"] = {"value": sum_over_ranks(units_rank) * e_steps / (s_ms * 1e-3), "unit": UNIT,
                                   "h2d_bytes_per_step": int(tm["h2d_bytes"]), "d2h_bytes_per_step": int(tm["d2h_bytes"]),
                                   "ms_per_step": s_ms / e_steps,
                                   "returns": "summary[n_cols][4] + final_state + status (no outlet series)"}
            lib.chrom_cuda_host_free(p4)
        for p in (p1, p2, p3):
            lib.chrom_cuda_host_free(p)

    # ---- CPU baseline (rank 0, N = 1 only) -------------------------------------------------------------
    cpu = None
    os.sched_setaffinity(0, affinity0)
    if rank == 0 and world == 1 and not args.no_cpu:
        cores = host_cores()
        n, csteps = cpu_sample_size(spec, cores)
        v, secs = cpu_sample(spec, n, cores, csteps)
        cpu = {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
               "sample": f"{n} of {spec['n_cols']} columns x {csteps} of {spec['steps']} time steps, one column per "
                         f"OpenMP thread, {secs:.1f} s wall"}

    if rank == 0:
        per_rank_rate = units_rank * args.steps / (sum(kernel_ms) * 1e-3)
        achieved = per_rank_rate * spec["flops"] / 1e12
        # dominant kernel: launches of the solver kernel per execute and its average duration (CUDA events)
        solver_launches = max(1, int(launches_value / args.steps) - (3 if "single_stream" in desc else 0))
        launch_ms = (sum(kernel_ms) / args.steps) / solver_launches
        family = desc.split("<")[0]
        traffic, traffic_src = ncu_traffic(family, args.workload)
        # algorithmic HBM bytes per solver launch: streaming path 16 B per cell-step;
        # resident paths only write their sampled outputs (outlet series + final state)
        if "single_stream" in desc:
            fuse = int(desc.split(" fused")[0].split()[-1])
            alg_bytes = 16.0 * spec["nz"] * n_cols * fuse
        else:
            alg_bytes = 8.0 * n_cols * spec["n_species"] * (spec["steps"] + 1 + spec["nz"])
        hbm_peak = measured_peaks().get("hbm_gbs")
        hbm_gbs = alg_bytes / (launch_ms * 1e-3) / 1e9
        out_gb = 8.0 * n_cols * spec["n_species"] * (spec["steps"] + 1) / 1e9
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": W_,
            "ms_per_step": t_dev_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": spec["label"], "columns_per_gpu": n_cols,
                       "arith": "fast (FMA contraction, one refined MUFU reciprocal per cell-stage / pivot; <=1e-9 "
                                "relative to the reference's operation order)",
                       "kernel": desc,
                       "l2": f"256 MiB buffer written between timed iterations (L2 flush); outputs {out_gb:.2f} GB/step",
                       "columns_with_nonfinite_status": n_bad},
            "roofline": {"bound": "fp64", "achieved": achieved, "peak": peak_tflops, "unit": "TFLOP/s",
                         "frac": achieved / peak_tflops if peak_tflops else None,
                         "traffic": traffic, "traffic_source": traffic_src,
                         "peak_source": "DFMA micro-benchmark in this run (chrom_cuda_measure_fp64_peak); "
                                        "MEASURED_PEAKS.json holds HBM and bf16 peaks only, no FP64 entry",
                         "flops_per_cell_stage": spec["flops"],
                         "kernel_launch_ms": launch_ms, "kernel_launches_per_step": solver_launches,
                         "algorithmic_bytes": alg_bytes,
                         "hbm": {"achieved": hbm_gbs, "peak": hbm_peak, "unit": "GB/s",
                                 "frac": hbm_gbs / hbm_peak if hbm_peak else None,
                                 "peak_source": "MEASURED_PEAKS.json hbm_gbs (burst copy)" if hbm_peak else "absent"}},
            "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": launches, "clocks": clocks,
            "wall_ms_per_step_incl_flush": wall_ms / args.steps,
        }
        print(json.dumps(line), flush=True)
    ctx.close()
    ranks.close()


if __name__ == "__main__":
    main()
