"""bench.py — the headline metric of BASELINE.json on B200:
RK4 cell-stage updates/sec on a batch of Langmuir columns.

A "step" is one pass of the hot path over one batch: the complete solve (all time steps) of every column of the
workload.  Default workload: config 3 of BASELINE.json, the one the metric is quoted on — ONE batch of 65 536
TFA-like LangmuirSingle columns, nz=100, RK4, 10 000 steps.  With N GPUs that batch is SHARDED: rank r owns the
columns chrom_cuda_shard_range(65536, N, r) (strong scaling, no data-path collective: columns are independent).
`--scaling weak` gives every rank its own 65 536 columns instead (reported under config.weak at the default).

  value   : whole-job cell-stage updates/s with inputs resident in HBM (plan executed K times, kernel time from
            CUDA events on the launch stream, max over ranks)
  e2e     : the same through chrom_cuda_solve_single_batch with pinned HOST buffers, every step: parameter +
            injection-table H2D, kernels, results D2H.  Config 3 is a parameter sweep, so the headline e2e call
            returns what a sweep consumes per column (the reference's cfl_stability_scan reduces every result to a
            few scalars): the on-device summary (area, retention time, max, time of max, sigma, min), the outlet
            chromatogram sampled every 50th step, the final profile and the validate_state status.
            e2e.full_outlet is the same call returning the reference's FULL-rate outlet series (5.2 GB per 65 536
            columns) with the host's bare D2H ceiling for that volume measured beside it.
  roofline: algorithmic flops (15.25 per cell-stage) / kernel time against the FP64 (DFMA) peak measured
            in this run AND against the nominal 148 SM x 64 DFMA/clk x 2 x f_max (MEASURED_PEAKS.json has no FP64 entry)
  cpu_baseline: the C++ oracle port of the reference arithmetic on the box's host cores, bounded sample
  extra_workloads: configs 4 and 5 (and config 5 with the dense register LU) in the same line, <= 3 steps each

`--impl reference` times the reference's CPU path (the oracle port: the Rust reference cannot be built here — no
cargo/rustc) with all host threads on a bounded sample of the same workload.
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
SWEEP_OUTLET_STRIDE = 50  # the sweep-mode e2e samples the outlet chromatogram every 50th step (201 points of 10 001)

# algorithmic flops per cell-stage (a divide counts as one flop)
FLOPS_PER_CELL_STAGE = {"single": 15.25, "multi2": 53.5, "multi8": 776.0}


def workload_spec(name: str) -> dict:
    if name == "c3":
        return dict(name="c3", kind="single", n_cols=65536, nz=100, n_species=1, total_time=600.0, steps=10_000,
                    label="C3: 65,536 TFA-like LangmuirSingle columns (K,N,u,eps varied), nz=100, RK4 600 s / 10,000 steps",
                    flops=FLOPS_PER_CELL_STAGE["single"])
    if name == "c1":
        return dict(name="c1", kind="single", n_cols=1, nz=100, n_species=1, total_time=600.0, steps=10_000,
                    label="C1: one TFA LangmuirSingle column, nz=100, RK4 600 s / 10,000 steps",
                    flops=FLOPS_PER_CELL_STAGE["single"])
    if name == "c5":
        return dict(name="c5", kind="multi", n_cols=4096, nz=512, n_species=8, total_time=600.0, steps=4096,
                    label="C5: 4,096 LangmuirMulti columns, n=8, nz=512, RK4 600 s / 4,096 steps",
                    flops=FLOPS_PER_CELL_STAGE["multi8"])
    if name == "c2":
        return dict(name="c2", kind="multi", n_cols=1, nz=100, n_species=2, total_time=1200.0, steps=20_000,
                    label="C2: ascorbic/erythorbic LangmuirMulti n=2, nz=100, RK4 1200 s / 20,000 steps",
                    flops=FLOPS_PER_CELL_STAGE["multi2"])
    if name == "c4":
        return dict(name="c4", kind="single", n_cols=1, nz=1 << 22, n_species=1, total_time=None, steps=1000,
                    label="C4: one fine-grid TFA column nz=2^22, RK4 1,000 steps (streaming path)",
                    flops=FLOPS_PER_CELL_STAGE["single"])
    raise SystemExit(f"unknown workload {name}")


def units_of(spec: dict, n_cols: int) -> float:
    return float(n_cols) * spec["nz"] * spec["steps"] * 4.0


def shard_range(n_cols: int, world: int, rank: int) -> tuple:
    """chrom_cuda_shard_range (api.cu): ceil(n_cols / world) contiguous columns per rank, trailing ranks may get
    fewer or none.  Restated here so the CPU tests of the N>1 host path need no GPU library; the GPU arm checks
    it against the library's own answer."""
    per = (n_cols + world - 1) // world
    c0 = min(rank * per, n_cols)
    return c0, min(per, n_cols - c0)


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
def cpu_sample(spec: dict, n_cols: int, threads: int, steps: int | None = None, cells_parallel: bool = False):
    """Time the oracle on `n_cols` columns of the workload.  Default: one column per OpenMP thread (a lone column
    runs on one thread: LangmuirSingle has no parallel path in the reference).  cells_parallel (LangmuirMulti): the
    reference's own parallel path instead — columns one after the other, the CELLS of each right-hand side spread
    over the threads, as rayon does above parallel_threshold() = 999 (langmuir_multi.rs:855-874, solver/mod.rs:385).
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
        O.solve_multi_batch(models, O.RK4, total_time, steps, n_threads=threads, want_outlet=True,
                            cells_parallel=cells_parallel)
        dt = time.perf_counter() - t0
    return len(models) * spec["nz"] * steps * 4.0 / dt, dt


def cpu_sample_size(spec: dict, cores: int, wall_s: float = 4.0) -> tuple:
    """(columns, steps) of a bounded sample: roughly `wall_s` seconds of wall clock on all cores."""
    per_core_rate = 1.0e8 if spec["kind"] == "single" else 2.0e8 / (spec["n_species"] ** 2.2)  # cell-stage/s, rough
    budget = wall_s * per_core_rate * cores
    steps = spec["steps"]
    per_col = spec["nz"] * steps * 4.0
    if per_col > budget:  # very long column (C4): cut steps instead
        return 1, max(8, int(budget / (spec["nz"] * 4.0)))
    n = max(cores, int(budget / per_col))
    n = (n // cores) * cores
    return min(n, spec["n_cols"]), steps


def cpu_baseline_of(spec: dict, wall_s: float = 4.0) -> dict:
    cores = host_cores()
    n, csteps = cpu_sample_size(spec, cores, wall_s)
    v, secs = cpu_sample(spec, n, cores, csteps)
    cpu = {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
           "sample": f"{n} of {spec['n_cols']} columns x {csteps} of {spec['steps']} time steps, one column per "
                     f"OpenMP thread, {secs:.1f} s wall"}
    if spec["kind"] == "multi" and spec["nz"] * spec["n_species"] > 999:
        # the reference's existing parallel path: cells of one right-hand side over the threads (rayon mirror)
        ncp = max(1, min(4, n))
        psteps = max(16, csteps // 16)
        vp, sp = cpu_sample(spec, ncp, cores, psteps, cells_parallel=True)
        cpu["cells_parallel"] = {"value": vp, "unit": UNIT, "cores": cores,
                                 "sample": f"{ncp} column(s) x {psteps} time steps, cells of each right-hand side "
                                           f"over {cores} OpenMP threads (nz*n = {spec['nz'] * spec['n_species']} > "
                                           f"parallel_threshold 999: langmuir_multi.rs:855-874), {sp:.1f} s wall"}
    return cpu


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
        "scaling": args.scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
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
                # a second, CPU-only group: a rank waiting in an NCCL barrier keeps a spinning kernel on ITS GPU, which
                # halves the rate of anything another process runs there (the one-context arm drives every GPU
                # from rank 0: measured 188 ms instead of 87 ms at N = 2 before this)
                self.host_group = dist.new_group(backend="gloo")
            else:
                dist.init_process_group(backend)
                self.host_group = None

    def barrier(self):
        if self.distributed:
            self.dist.barrier()
        if self.cuda:
            self.torch.cuda.synchronize()

    def host_barrier(self):
        """Wait for every rank on the host only: no kernel is queued on any GPU."""
        if self.distributed:
            if self.cuda:
                self.torch.cuda.synchronize()
            self.dist.barrier(group=self.host_group)

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
    rc = lib.chrom_cuda_host_alloc(max(n, 1) * np.dtype(dtype).itemsize, C.byref(p))
    if rc != 0:
        raise RuntimeError("chrom_cuda_host_alloc failed")
    buf = (C.c_byte * (n * np.dtype(dtype).itemsize)).from_address(p.value)
    return np.frombuffer(buf, dtype=dtype).reshape(shape), p


def rank_columns(spec: dict, rank: int, world: int, scaling: str, n_cols_override: int = 0) -> tuple:
    """(first column, count, seed) of this rank's share of the workload.  strong: ONE batch of spec.n_cols columns
    (seed 42) cut into chrom_cuda_shard_range pieces; weak: spec.n_cols columns of its own per rank (seed 42 + rank)."""
    n_total = n_cols_override or spec["n_cols"]
    if scaling == "strong" and world > 1 and n_total >= world:
        c0, cnt = shard_range(n_total, world, rank)
        return c0, cnt, 42
    return 0, n_total, 42 + (rank if scaling == "weak" else 0)


def build_inputs(spec, rank, world=1, scaling="strong", n_cols_override=0):
    """Synthetic inputs of the workload for this rank, packed for the C ABI."""
    import chromatography_b200 as cb
    from chromatography_b200 import _ffi, workloads as W
    steps = spec["steps"]
    c0, cnt, seed = rank_columns(spec, rank, world, scaling, n_cols_override)
    n_total = c0 + cnt if scaling == "weak" or world == 1 else (n_cols_override or spec["n_cols"])
    if spec["kind"] == "single":
        if spec["nz"] > 100000:
            m, total_time, _ = W.fine_grid_single(spec["nz"], steps)
            cols, tables = cb.pack_single([m], cb.RK4, total_time, steps)
            return dict(cols=cols, tables=tables, species=None, total_time=total_time, col0=0)
        a = W.tfa_sweep_arrays(n_total, seed=seed, nz=spec["nz"])
        cols = (_ffi.SingleCol * cnt)()
        arr = np.frombuffer(cols, dtype=np.float64).reshape(cnt, 7)
        for j, k in enumerate(["lambda_", "langmuir_k", "port_number", "fe", "ue", "dz"]):
            arr[:, j] = a[k][c0:c0 + cnt]
        arr[:, 6] = 0.0  # inj_table = 0 (shared Gaussian(10, 3, 0.1) table), detector at the outlet
        tables = cb.tabulate_injection(cb.TemporalInjection.gaussian(10.0, 3.0, 0.1), cb.RK4, spec["total_time"],
                                       steps)[None]
        return dict(cols=cols, tables=np.ascontiguousarray(tables), species=None, total_time=spec["total_time"], col0=c0)
    if spec["n_cols"] > 1:
        models = W.multi_batch_models(n_total, spec["n_species"], spec["nz"], seed=seed)[c0:c0 + cnt]
    else:
        models = [W.acids_multi()]
    cols, species, tables = cb.pack_multi(models, cb.RK4, spec["total_time"], steps)
    return dict(cols=cols, tables=tables, species=species, total_time=spec["total_time"], col0=c0)


def host_d2h_ceiling(ranks, local_rank: int, nbytes: int, reps: int = 3) -> dict:
    """The box's bare device->host rate for `nbytes` per rank, every rank copying at once into its own page-locked
    buffer, no kernels running (torch for the plumbing).  Whole-job GB/s = total bytes / slowest rank."""
    import torch
    n = max(8, int(nbytes))
    d = torch.empty(n, dtype=torch.uint8, device=f"cuda:{local_rank}")
    d.fill_(1)
    h = torch.empty(n, dtype=torch.uint8, pin_memory=True)
    h.fill_(0)
    best = None
    for _ in range(reps + 1):
        ranks.barrier()
        t0 = time.perf_counter()
        h.copy_(d, non_blocking=True)
        torch.cuda.synchronize()
        ms = ranks.max((time.perf_counter() - t0) * 1e3)
        best = ms if best is None else min(best, ms)
    total = ranks.sum(float(n))
    del d, h
    torch.cuda.empty_cache()
    return {"bytes_per_rank": n, "ms": best, "aggregate_gbs": total / (best * 1e-3) / 1e9,
            "how": "pinned cudaMemcpyAsync of the same volume, all ranks at once, no kernels, best of %d" % reps}


def run_workload(ctx, lib, ranks, args, spec, rank, world, local_rank, steps_k, warmup_w, arith_name="fast",
                 scaling="strong", peak_tflops=0.0, main=False, want_cpu=True, numa=(-1, 0), do_e2e=True):
    """One workload on this rank's share of the columns: value (plan-resident kernels), e2e (C-ABI call with host
    buffers), roofline, CPU baseline.  Returns (record for rank 0 or None, launches)."""
    import chromatography_b200 as cb
    barrier, max_over_ranks, sum_over_ranks = ranks.barrier, ranks.max, ranks.sum
    slen = getattr(cb, "SUMMARY_LEN", 4)
    inp = build_inputs(spec, rank, world, scaling, args.cols if main else 0)
    n_cols = len(inp["cols"])
    units_rank = units_of(spec, n_cols)
    arith = cb.ARITH_FAST
    flops = spec["flops"]
    if spec["kind"] == "multi":
        if arith_name == "dense":
            arith = cb.ARITH_FAST_DENSE  # the dense register LU: SURVEY 8d's 776 flops per cell-stage at n = 8
        else:
            # FAST = the closed-form structured elimination: per species b 2, S 2, p 1, alpha 2, 1/alpha 1, z 1, v 2,
            # q.v 2, q.z 2, x 2, stage algebra 3.25 = 20.25 flops; per cell 1+S, 1/D, 1/D^2, 1-q.v, f 2, pivot bound 4 = 10
            flops = 20.25 * spec["n_species"] + 10.0
    cfg = cb.make_cfg(cb.RK4, inp["total_time"], spec["steps"], spec["nz"], spec["n_species"], 1, 0, arith)
    want = cb.WANT_OUTLET | cb.WANT_FINAL | cb.WANT_STATUS

    # ---- value: inputs resident in HBM --------------------------------------------------------------
    plan = cb.Plan(ctx, cfg, inp["cols"], inp["tables"], species=inp["species"], want=want)
    desc = plan.describe()
    for _ in range(warmup_w):
        plan.execute()
    barrier()
    sampler = ClockSampler(local_rank) if (rank == 0 and main) else None
    kernel_ms, launches = [], 0
    wall0 = time.perf_counter()
    for _ in range(steps_k):
        ctx.flush_l2()  # cold L2 between timed iterations (outside the event-timed kernel region)
        t = plan.execute()
        kernel_ms.append(t["kernel_ms"])
        launches += int(t["kernel_launches"])
    barrier()
    wall_ms = (time.perf_counter() - wall0) * 1e3
    clocks = sampler.stop() if sampler else None
    t_dev_ms = max_over_ranks(sum(kernel_ms))
    value = aggregate_value(ranks, units_rank * steps_k, sum(kernel_ms))
    launches_value = launches
    st = plan.download().status
    n_bad = int(sum_over_ranks(float(np.count_nonzero(st))))
    plan.close()

    # ---- e2e: host buffers through the C-ABI solve call ----------------------------------------------
    e2e = None
    if do_e2e and not args.no_e2e:
        sp = (spec["n_species"],) if spec["kind"] == "multi" else ()
        tables = inp["tables"]
        dp = lambda a: None if a is None else a.ctypes.data_as(C.POINTER(C.c_double))  # noqa: E731
        final, p_final = pinned_array(lib, (n_cols,) + sp + (spec["nz"],))
        status, p_status = pinned_array(lib, (n_cols,), np.int64)
        sweep = spec["kind"] == "single" and spec["nz"] <= 8192 and spec["n_cols"] > 1
        e_steps = max(1, min(steps_k, 3))

        def timed_call(cfg_c, outlet, summ):
            def call():
                if spec["kind"] == "single":
                    rc = lib.chrom_cuda_solve_single_batch(ctx._h, C.byref(cfg_c), inp["cols"], n_cols, dp(tables),
                                                           tables.shape[0], None, dp(outlet), None, dp(final), dp(summ),
                                                           status.ctypes.data_as(C.POINTER(C.c_int64)))
                else:
                    rc = lib.chrom_cuda_solve_multi_batch(ctx._h, C.byref(cfg_c), inp["cols"], n_cols, inp["species"],
                                                          len(inp["species"]), dp(tables), tables.shape[0], None,
                                                          dp(outlet), None, dp(final), dp(summ),
                                                          status.ctypes.data_as(C.POINTER(C.c_int64)))
                ctx.check(rc)
                return ctx.timing()

            call()  # warm-up (first touch of the pinned buffers)
            barrier()
            t0 = time.perf_counter()
            n_launch = 0
            for _ in range(e_steps):
                tm = call()
                n_launch += int(tm["kernel_launches"])
            barrier()
            ms = max_over_ranks((time.perf_counter() - t0) * 1e3)
            return {"value": sum_over_ranks(units_rank) * e_steps / (ms * 1e-3), "unit": UNIT,
                    "h2d_bytes_per_step": int(sum_over_ranks(float(tm["h2d_bytes"]))),
                    "d2h_bytes_per_step": int(sum_over_ranks(float(tm["d2h_bytes"]))),
                    "ms_per_step": ms / e_steps, "steps": e_steps}, n_launch, int(tm["d2h_bytes"])

        # full-rate outlet series: what the reference's SimulationResult holds for its consumers
        outlet, p_out = pinned_array(lib, (n_cols,) + sp + (spec["steps"] + 1,))
        full, nl, full_d2h = timed_call(cfg, outlet, None)
        launches += nl
        lib.chrom_cuda_host_free(p_out)
        full["returns"] = "outlet[n_cols][steps+1] (every step: the reference's chromatogram) + final_state + status"
        if world > 1 or main:
            ceil = host_d2h_ceiling(ranks, local_rank, full_d2h)
            full["host_d2h_ceiling"] = ceil
            full["d2h_floor_ms"] = ceil["ms"]
        api = "chrom_cuda_solve_%s_batch, pinned host buffers, one process per GPU" % spec["kind"]
        if sweep:
            cfg_s = cb.make_cfg(cb.RK4, inp["total_time"], spec["steps"], spec["nz"], spec["n_species"],
                                SWEEP_OUTLET_STRIDE, 0, arith)
            n_out_s = spec["steps"] // SWEEP_OUTLET_STRIDE + 1
            outlet_s, p_out_s = pinned_array(lib, (n_cols, n_out_s))
            summ, p_summ = pinned_array(lib, (n_cols, slen))
            e2e, nl, _ = timed_call(cfg_s, outlet_s, summ)
            launches += nl
            lib.chrom_cuda_host_free(p_out_s)
            lib.chrom_cuda_host_free(p_summ)
            e2e["mode"] = "sweep"
            e2e["returns"] = (f"summary[n_cols][{slen}] (area, first moment, max, time of max, sigma, min: on device) + "
                              f"outlet every {SWEEP_OUTLET_STRIDE}th step + final_state + status")
            e2e["full_outlet"] = full
        else:
            e2e = full
            e2e["mode"] = "full_outlet"
        e2e["api"] = api
        e2e["host_numa_node"], e2e["host_cpus_bound"] = numa
        for p in (p_final, p_status):
            lib.chrom_cuda_host_free(p)

    # ---- CPU baseline (rank 0, N = 1 only) -------------------------------------------------------------
    cpu = None
    if rank == 0 and world == 1 and want_cpu and not args.no_cpu:
        cpu = cpu_baseline_of(spec, 4.0 if main else 2.0)

    if rank != 0:
        return None, launches
    per_rank_rate = units_rank * steps_k / (sum(kernel_ms) * 1e-3)
    achieved = per_rank_rate * flops / 1e12
    clk = (clocks or {}).get("sm_max_mhz") or measured_peaks().get("sm_max_mhz") or 1965.0
    peak_nominal = 148 * 64 * 2 * clk * 1e6 / 1e12
    # dominant kernel: launches of the solver kernel per execute and its average duration (CUDA events)
    solver_launches = max(1, int(launches_value / steps_k) - (4 if "single_stream" in desc else 0))
    launch_ms = (sum(kernel_ms) / steps_k) / solver_launches
    family = desc.split("<")[0]
    traffic, traffic_src = ncu_traffic(family, spec["name"] + ("_dense" if arith_name == "dense" else ""))
    # algorithmic HBM bytes per solver launch: streaming path 16 B per cell-step;
    # resident paths only write their sampled outputs (outlet series + final state)
    if "single_stream" in desc:
        fuse = int(desc.split(" fused")[0].split()[-1])
        alg_bytes = 16.0 * spec["nz"] * n_cols * fuse
    else:
        alg_bytes = 8.0 * n_cols * spec["n_species"] * (spec["steps"] + 1 + spec["nz"])
    hbm_peak = measured_peaks().get("hbm_gbs")
    hbm_gbs = alg_bytes / (launch_ms * 1e-3) / 1e9
    rec = {
        "value": value, "unit": UNIT, "steps": steps_k, "warmup": warmup_w, "ms_per_step": t_dev_ms / steps_k,
        "config": {"workload": spec["label"], "columns_total": spec["n_cols"] * (world if scaling == "weak" else 1),
                   "columns_rank0": n_cols, "kernel": desc, "columns_with_nonfinite_status": n_bad},
        "roofline": {"bound": "fp64", "achieved": achieved, "peak": peak_tflops, "unit": "TFLOP/s",
                     "frac": achieved / peak_tflops if peak_tflops else None,
                     "peak_nominal": peak_nominal, "frac_nominal": achieved / peak_nominal,
                     "traffic": traffic, "traffic_source": traffic_src,
                     "peak_source": "peak = DFMA micro-benchmark in this run (chrom_cuda_measure_fp64_peak); peak_nominal = "
                                    "148 SM x 64 DFMA/clk x 2 x %.0f MHz; MEASURED_PEAKS.json holds HBM and bf16 peaks "
                                    "only, no FP64 entry" % clk,
                     "flops_per_cell_stage": flops,
                     "kernel_launch_ms": launch_ms, "kernel_launches_per_step": solver_launches,
                     "algorithmic_bytes": alg_bytes,
                     "hbm": {"achieved": hbm_gbs, "peak": hbm_peak, "unit": "GB/s",
                             "frac": hbm_gbs / hbm_peak if hbm_peak else None,
                             "peak_source": "MEASURED_PEAKS.json hbm_gbs (burst copy)" if hbm_peak else "absent"}},
        "cpu_baseline": cpu, "e2e": e2e, "clocks": clocks, "wall_ms_per_step_incl_flush": wall_ms / steps_k,
    }
    return rec, launches


def one_context_e2e(lib, ranks, spec, world, rank) -> dict | None:
    """The library's own sharding: ONE chrom_cuda context over all `world` devices, driven by rank 0 alone through one
    call with one set of caller buffers; the other ranks wait on the host (no kernel of
    theirs sits on a GPU meanwhile).  Sweep mode (summary + strided outlet) and the full-rate outlet series."""
    import chromatography_b200 as cb
    slen = getattr(cb, "SUMMARY_LEN", 4)
    out = None
    ranks.host_barrier()
    if rank == 0:
        ctx = cb.Context(list(range(world)))
        inp = build_inputs(spec, 0, 1, "strong")
        n_cols = len(inp["cols"])
        dp = lambda a: None if a is None else a.ctypes.data_as(C.POINTER(C.c_double))  # noqa: E731
        final, p1 = pinned_array(lib, (n_cols, spec["nz"]))
        status, p2 = pinned_array(lib, (n_cols,), np.int64)
        tables = inp["tables"]

        def timed(stride, outlet, summ, reps=3):
            cfg = cb.make_cfg(cb.RK4, inp["total_time"], spec["steps"], spec["nz"], 1, stride, 0, cb.ARITH_FAST)

            def call():
                ctx.check(lib.chrom_cuda_solve_single_batch(ctx._h, C.byref(cfg), inp["cols"], n_cols, dp(tables),
                                                            tables.shape[0], None, dp(outlet), None, dp(final), dp(summ),
                                                            status.ctypes.data_as(C.POINTER(C.c_int64))))
                return ctx.timing()

            call()
            t0 = time.perf_counter()
            for _ in range(reps):
                tm = call()
            ms = (time.perf_counter() - t0) * 1e3 / reps
            return {"value": units_of(spec, n_cols) / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms, "steps": reps,
                    "h2d_bytes_per_step": int(tm["h2d_bytes"]), "d2h_bytes_per_step": int(tm["d2h_bytes"]),
                    "kernel_ms_max_over_devices": tm["kernel_ms"]}

        summ, p3 = pinned_array(lib, (n_cols, slen))
        outlet_s, p4 = pinned_array(lib, (n_cols, spec["steps"] // SWEEP_OUTLET_STRIDE + 1))
        out = timed(SWEEP_OUTLET_STRIDE, outlet_s, summ)
        out.update({"mode": "sweep", "devices": world,
                    "api": "ONE chrom_cuda context over %d devices, one caller buffer (chrom_cuda_shard_range inside), "
                           "driven by rank 0; the other ranks wait on the host" % world})
        lib.chrom_cuda_host_free(p4)
        lib.chrom_cuda_host_free(p3)
        outlet, p5 = pinned_array(lib, (n_cols, spec["steps"] + 1))
        out["full_outlet"] = timed(1, outlet, None)
        for p in (p1, p2, p5):
            lib.chrom_cuda_host_free(p)
        ctx.close()
    ranks.host_barrier()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c3", choices=["c1", "c2", "c3", "c4", "c5"])
    ap.add_argument("--scaling", default="strong", choices=["strong", "weak"],
                    help="N > 1: strong = ONE batch of the workload's columns sharded over the ranks (default, config 3 as "
                         "worded); weak = every rank its own full batch")
    ap.add_argument("--cols", type=int, default=0, help="override the batch size (debug only; invalidates the metric)")
    ap.add_argument("--arith", default="fast", choices=["fast", "dense"],
                    help="multi-species workloads: closed-form structured LU (fast, the library default) or the "
                         "dense register LU with partial pivoting (dense)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="skip the extra_workloads (configs 4 and 5) of the default run")
    ap.add_argument("--no-numa", action="store_true", help="do not bind the rank to its GPU's NUMA node")
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

    import chromatography_b200 as cb
    from chromatography_b200 import _ffi
    lib = _ffi.lib()
    ctx = cb.Context([local_rank])
    # the restated shard rule must be the library's
    c0, cnt = C.c_uint64(), C.c_uint64()
    lib.chrom_cuda_shard_range(spec["n_cols"], world, rank, C.byref(c0), C.byref(cnt))
    assert (c0.value, cnt.value) == shard_range(spec["n_cols"], world, rank)
    # multi-GPU hosts: keep this rank's thread and its page-locked buffers on the NUMA node of its GPU
    # (N = 1: nothing to bind, and the CPU baseline of this run wants every host core)
    numa = ctx.bind_host_thread(0) if (world > 1 and not args.no_numa) else (-1, 0)
    peak_tflops = ctx.fp64_peak_tflops() if rank == 0 else 0.0
    if spec["n_cols"] == 1 and world > 1:
        args.scaling = "weak"  # a single column does not shard: replicas only

    rec, launches = run_workload(ctx, lib, ranks, args, spec, rank, world, local_rank, args.steps, W_, args.arith,
                                 args.scaling, peak_tflops, main=True, numa=numa)

    extras = {}
    default_run = args.workload == "c3" and not args.cols and not args.no_extra
    if default_run:
        # configs 4 and 5 where the driver sees them: <= 3 timed steps each.  Config 5 shards like config 3 (at every
        # N); config 4 is one column = one GPU (N = 1 only), and so is the dense-LU variant of config 5.
        k = max(1, min(3, args.steps))
        plan = [("c5", "fast", True)] + ([("c5_dense", "dense", True), ("c4", "fast", True)] if world == 1 else [])
        for name, arith_name, cpu in plan:
            s = workload_spec(name.split("_")[0])
            r, nl = run_workload(ctx, lib, ranks, args, s, rank, world, local_rank, k, 3, arith_name, args.scaling,
                                 peak_tflops, main=False, want_cpu=cpu and arith_name == "fast", numa=numa)
            launches += nl
            if r is not None:
                extras[name] = r
    weak = None
    if world > 1 and default_run and args.scaling == "strong":
        # the weak-scaling figure of round 1 (every rank its own full batch, kernels only) beside the strong-scaling line
        r, nl = run_workload(ctx, lib, ranks, args, spec, rank, world, local_rank, max(1, min(3, args.steps)), 3, args.arith,
                             "weak", peak_tflops, main=False, want_cpu=False, numa=numa, do_e2e=False)
        launches += nl
        if r is not None:
            weak = {"value": r["value"], "unit": UNIT, "ms_per_step": r["ms_per_step"], "steps": r["steps"],
                    "columns_total": r["config"]["columns_total"], "columns_per_rank": r["config"]["columns_rank0"],
                    "note": "weak scaling: independent processes, every rank its own full batch, kernels only (no e2e)"}
    one_ctx = None
    if world > 1 and args.workload == "c3" and not args.no_e2e and not args.cols:
        ctx.close()
        ctx = None
        one_ctx = one_context_e2e(lib, ranks, spec, world, rank)

    if rank == 0:
        e2e = rec["e2e"]
        if e2e is not None and one_ctx is not None:
            e2e["one_context"] = one_ctx
        out_gb = 8.0 * spec["n_cols"] * spec["n_species"] * (spec["steps"] + 1) / 1e9
        cfg = rec["config"]
        cfg.update({
            "columns_total": spec["n_cols"] * (world if args.scaling == "weak" else 1) if not args.cols else args.cols,
            "sharding": ("ONE batch sharded over %d rank(s) by chrom_cuda_shard_range, no collective" % world)
            if args.scaling == "strong" else "weak: every rank its own full batch (independent processes)",
            "arith": "fast (FMA contraction, one refined MUFU reciprocal per cell-stage / pivot; <=1e-9 relative to the "
                     "reference's operation order)",
            "l2": f"256 MiB buffer written between timed iterations (L2 flush); full-rate outlet series {out_gb:.2f} GB per batch",
            "e2e_mode": (e2e or {}).get("mode")})
        if weak is not None:
            cfg["weak"] = weak
        line = {
            "metric": METRIC, "value": rec["value"], "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": W_,
            "ms_per_step": rec["ms_per_step"], "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "config": cfg, "roofline": rec["roofline"],
            "cpu_baseline": rec["cpu_baseline"], "e2e": e2e, "gpu_launches": launches, "clocks": rec["clocks"],
            "wall_ms_per_step_incl_flush": rec["wall_ms_per_step_incl_flush"],
        }
        if extras:
            line["extra_workloads"] = extras
        print(json.dumps(line), flush=True)
    if ctx is not None:
        ctx.close()
    ranks.close()


if __name__ == "__main__":
    main()
