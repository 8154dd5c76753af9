// api.cu — the C ABI of libchrom_cuda.so (include/chrom_cuda.h): contexts, plans, host-side
// helpers that restate the reference's boundary semantics (config validation, stage-time
// injection tables, time points, validate_state messages) and the batch -> device sharding.
//
// Sharding: columns are independent, so a batch is cut into contiguous column
// ranges, one per device of the context; each device has its own stream; there is no collective.
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <string>
#include <vector>

#include <sched.h>
#include <sys/syscall.h>
#include <unistd.h>

#include "kernels.h"

using namespace chrom;

namespace {

thread_local std::string g_create_error;

struct DeviceState {
  int id = 0;
  int sm_count = 0;
  cudaStream_t stream = nullptr;
  cudaStream_t stream2 = nullptr;      // second compute stream: consecutive chunks of a one-shot solve overlap
  cudaStream_t copy_stream = nullptr;  // device->host copies overlapped with compute (one-shot solves)
  void* flush_buf = nullptr;
  unsigned long long pool_threshold0 = 0;  // release threshold of the device's default pool before this context raised it
  cudaEvent_t ev[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};  // h2d0 h2d1 k0 k1 d2h0 d2h1
};

std::string format(const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  return buf;
}

}  // namespace

struct chrom_cuda_ctx {
  std::vector<DeviceState> devs;
  mutable std::string err;
  chrom_cuda_timing timing{};
};

namespace {

int32_t fail(const chrom_cuda_ctx* ctx, int32_t code, const std::string& msg) {
  if (ctx) ctx->err = msg;
  else g_create_error = msg;
  return code;
}

#define CUDA_TRY(ctx, expr)                                                                       \
  do {                                                                                            \
    cudaError_t _e = (expr);                                                                      \
    if (_e != cudaSuccess)                                                                        \
      return fail(ctx, CHROM_ERR_CUDA, format("%s failed: %s", #expr, cudaGetErrorString(_e)));   \
  } while (0)

// Device buffers come from the device's stream-ordered memory pool (cudaMallocAsync) whose release
// threshold is raised at context creation, so repeated solves reuse the same memory instead of
// paying cudaMalloc/cudaFree for gigabytes of output on every call.
template <class T>
struct DevBuf {
  T* p = nullptr;
  size_t n = 0;
  cudaStream_t s = nullptr;
  ~DevBuf() { release(); }
  void release() {
    if (p) cudaFreeAsync(p, s);
    p = nullptr;
    n = 0;
  }
  cudaError_t alloc(size_t count, cudaStream_t stream) {
    release();
    n = count;
    s = stream;
    if (count == 0) return cudaSuccess;
    return cudaMallocAsync((void**)&p, count * sizeof(T), stream);
  }
};

struct Shard {
  int dev = 0;  // index into ctx->devs
  uint64_t col0 = 0, n_cols = 0;
  DevBuf<chrom_single_col> scols;
  DevBuf<chrom_multi_col> mcols;
  DevBuf<chrom_species> species;
  DevBuf<double> tables, y0, outlet, snapshots, final_state, summary, ping, pong, scratch, series;
  DevBuf<uint32_t> out_steps, snap_steps;
  DevBuf<long long> status;
  BatchDev b{};
  LaunchGeom g;
  float kernel_ms = 0.f;
  cudaGraphExec_t graph = nullptr;  // streaming path: the per-step launches of one execute, captured once
  ~Shard() {
    if (graph) cudaGraphExecDestroy(graph);
  }
};

// src/models/injection.rs:410-446
double injection_evaluate(int32_t kind, double p0, double p1, double p2, double t) {
  switch (kind) {
    case CHROM_INJ_DIRAC:
      return (t == p0) ? p1 : 0.0;
    case CHROM_INJ_GAUSSIAN: {
      const double distance = (t - p0) / p1;
      return p2 * std::exp(((-distance) * distance) / 2.0);
    }
    case CHROM_INJ_RECTANGLE:
      return (t >= p0 && t < p1) ? p2 : 0.0;
    default:
      return 0.0;
  }
}

double now_ms() {
  using namespace std::chrono;
  return duration<double, std::milli>(steady_clock::now().time_since_epoch()).count();
}

}  // namespace

struct chrom_cuda_plan {
  chrom_cuda_ctx* ctx = nullptr;
  chrom_run_cfg cfg{};
  bool multi = false;
  uint32_t want = 0;
  uint64_t n_cols = 0, n_out = 0, n_snap = 0;
  size_t state_len = 0;  // doubles per column state: nz * n_species
  size_t series = 0;     // outlet series per column: n_species
  std::vector<uint32_t> out_steps, snap_steps;  // explicit sample lists (+ sentinel), empty = strides
  std::vector<std::unique_ptr<Shard>> shards;
  std::string desc;
  double h2d_ms = 0.0;
  uint64_t h2d_bytes = 0;
  // Whatever path leaves a plan (normal destroy, an error in the middle of a pipelined solve, a failed build): nothing
  // may still be running on, or copying out of, the buffers the shards are about to free — DevBuf frees on the main
  // stream only, and host buffers belong to the caller again the moment the call returns.
  ~chrom_cuda_plan() {
    if (!ctx) return;
    for (auto& sh : shards) {
      if (!sh) continue;
      DeviceState& d = ctx->devs[sh->dev];
      cudaSetDevice(d.id);
      cudaStreamSynchronize(d.stream);
      cudaStreamSynchronize(d.stream2);
      cudaStreamSynchronize(d.copy_stream);
      sh.reset();
    }
  }
};

namespace {

int32_t check_cfg(chrom_cuda_ctx* ctx, const chrom_run_cfg* cfg, bool multi) {
  if (!cfg) return fail(ctx, CHROM_ERR_INVALID, "cfg is NULL");
  if (const char* m = chrom_cuda_validate_cfg(cfg)) return fail(ctx, CHROM_ERR_INVALID, m);
  if (cfg->method != CHROM_EULER && cfg->method != CHROM_RK4)
    return fail(ctx, CHROM_ERR_INVALID, format("unknown method %d (0 = Euler, 1 = RK4)", cfg->method));
  if (cfg->arith != CHROM_ARITH_FAST && cfg->arith != CHROM_ARITH_EXACT &&
      !(multi && (cfg->arith == CHROM_ARITH_FAST_STRUCTURED || cfg->arith == CHROM_ARITH_FAST_DENSE)))
    return fail(ctx, CHROM_ERR_INVALID, format("unknown arith mode %d", cfg->arith));
  if (cfg->n_points < 2)
    return fail(ctx, CHROM_ERR_INVALID, format("Need at least 2 spatial points, got %u", cfg->n_points));
  if (!multi && cfg->n_species != 1)
    return fail(ctx, CHROM_ERR_INVALID, format("single-species entry point needs n_species == 1, got %u", cfg->n_species));
  if (multi && cfg->n_species < 1)
    return fail(ctx, CHROM_ERR_INVALID, "Langmuir multi species must have at least one species.");
  if (cfg->time_steps >= (1ull << 31))
    return fail(ctx, CHROM_ERR_UNSUPPORTED, "time_steps >= 2^31 is not supported");
  if (cfg->outlet_stride >= (1ull << 31) || cfg->snapshot_stride >= (1ull << 31))
    return fail(ctx, CHROM_ERR_UNSUPPORTED, "strides >= 2^31 are not supported");
  const struct {
    const uint64_t* p;
    uint64_t n;
    const char* what;
  } lists[2] = {{cfg->outlet_steps, cfg->n_outlet_steps, "outlet_steps"}, {cfg->snapshot_steps, cfg->n_snapshot_steps, "snapshot_steps"}};
  for (const auto& l : lists) {
    if (!l.p) continue;
    if (l.n == 0 || l.n > cfg->time_steps + 1)
      return fail(ctx, CHROM_ERR_INVALID, format("%s: needs 1 .. time_steps + 1 entries, got %llu", l.what, (unsigned long long)l.n));
    for (uint64_t k = 0; k < l.n; ++k)
      if (l.p[k] > cfg->time_steps || (k > 0 && l.p[k] <= l.p[k - 1]))
        return fail(ctx, CHROM_ERR_INVALID,
                    format("%s[%llu] = %llu: step indices must be strictly increasing and <= time_steps", l.what,
                           (unsigned long long)k, (unsigned long long)l.p[k]));
  }
  return CHROM_OK;
}

// Upload + allocate one shard.  `cols_*`/`species` are host pointers for the WHOLE batch.
int32_t build_shard(chrom_cuda_plan* plan, Shard& sh, const chrom_single_col* scols, const chrom_multi_col* mcols,
                    const chrom_species* species, const double* inj_tables, uint32_t n_tables, const double* y0) {
  chrom_cuda_ctx* ctx = plan->ctx;
  const chrom_run_cfg& cfg = plan->cfg;
  DeviceState& d = ctx->devs[sh.dev];
  CUDA_TRY(ctx, cudaSetDevice(d.id));
  const uint32_t stages = chrom_cuda_table_stages(cfg.method);
  const size_t table_len = (size_t)n_tables * cfg.time_steps * stages;
  const size_t state_len = plan->state_len;

  CUDA_TRY(ctx, cudaEventRecord(d.ev[0], d.stream));
  uint64_t bytes = 0;
  if (!plan->multi) {
    CUDA_TRY(ctx, sh.scols.alloc(sh.n_cols, d.stream));
    CUDA_TRY(ctx, cudaMemcpyAsync(sh.scols.p, scols + sh.col0, sh.n_cols * sizeof(chrom_single_col),
                                  cudaMemcpyHostToDevice, d.stream));
    bytes += sh.n_cols * sizeof(chrom_single_col);
  } else {
    // rebase species offsets so the shard owns a dense species array
    const uint32_t n = cfg.n_species;
    std::vector<chrom_multi_col> cols(mcols + sh.col0, mcols + sh.col0 + sh.n_cols);
    std::vector<chrom_species> sp((size_t)sh.n_cols * n);
    for (uint64_t c = 0; c < sh.n_cols; ++c) {
      std::memcpy(&sp[c * n], species + cols[c].species_off, n * sizeof(chrom_species));
      cols[c].species_off = (uint32_t)(c * n);
    }
    CUDA_TRY(ctx, sh.mcols.alloc(sh.n_cols, d.stream));
    CUDA_TRY(ctx, sh.species.alloc(sp.size(), d.stream));
    CUDA_TRY(ctx, cudaMemcpyAsync(sh.mcols.p, cols.data(), cols.size() * sizeof(chrom_multi_col),
                                  cudaMemcpyHostToDevice, d.stream));
    CUDA_TRY(ctx, cudaMemcpyAsync(sh.species.p, sp.data(), sp.size() * sizeof(chrom_species),
                                  cudaMemcpyHostToDevice, d.stream));
    CUDA_TRY(ctx, cudaStreamSynchronize(d.stream));  // the staging vectors die at scope exit
    bytes += cols.size() * sizeof(chrom_multi_col) + sp.size() * sizeof(chrom_species);
  }
  CUDA_TRY(ctx, sh.tables.alloc(table_len, d.stream));
  CUDA_TRY(ctx, cudaMemcpyAsync(sh.tables.p, inj_tables, table_len * sizeof(double), cudaMemcpyHostToDevice, d.stream));
  bytes += table_len * sizeof(double);
  if (y0) {
    CUDA_TRY(ctx, sh.y0.alloc(sh.n_cols * state_len, d.stream));
    CUDA_TRY(ctx, cudaMemcpyAsync(sh.y0.p, y0 + sh.col0 * state_len, sh.n_cols * state_len * sizeof(double),
                                  cudaMemcpyHostToDevice, d.stream));
    bytes += sh.n_cols * state_len * sizeof(double);
  }
  CUDA_TRY(ctx, cudaEventRecord(d.ev[1], d.stream));
  plan->h2d_bytes += bytes;

  const uint32_t want = plan->want;
  if ((want & CHROM_WANT_OUTLET) && plan->n_out) CUDA_TRY(ctx, sh.outlet.alloc(sh.n_cols * plan->series * plan->n_out, d.stream));
  if ((want & CHROM_WANT_SNAPSHOTS) && plan->n_snap) CUDA_TRY(ctx, sh.snapshots.alloc(sh.n_cols * plan->n_snap * state_len, d.stream));
  if (want & CHROM_WANT_FINAL) CUDA_TRY(ctx, sh.final_state.alloc(sh.n_cols * state_len, d.stream));
  if (want & CHROM_WANT_SUMMARY) CUDA_TRY(ctx, sh.summary.alloc(sh.n_cols * plan->series * CHROM_SUMMARY_LEN, d.stream));
  if (!plan->out_steps.empty() && sh.outlet.p) {
    CUDA_TRY(ctx, sh.out_steps.alloc(plan->out_steps.size(), d.stream));
    CUDA_TRY(ctx, cudaMemcpyAsync(sh.out_steps.p, plan->out_steps.data(), plan->out_steps.size() * sizeof(uint32_t),
                                  cudaMemcpyHostToDevice, d.stream));
  }
  if (!plan->snap_steps.empty() && sh.snapshots.p) {
    CUDA_TRY(ctx, sh.snap_steps.alloc(plan->snap_steps.size(), d.stream));
    CUDA_TRY(ctx, cudaMemcpyAsync(sh.snap_steps.p, plan->snap_steps.data(), plan->snap_steps.size() * sizeof(uint32_t),
                                  cudaMemcpyHostToDevice, d.stream));
  }
  CUDA_TRY(ctx, sh.status.alloc(sh.n_cols, d.stream));  // always kept: validate_state is part of the solve

  BatchDev& b = sh.b;
  b.scols = sh.scols.p;
  b.mcols = sh.mcols.p;
  b.species = sh.species.p;
  b.tables = sh.tables.p;
  b.y0 = sh.y0.p;
  b.outlet = sh.outlet.p;
  b.snapshots = sh.snapshots.p;
  b.final_state = sh.final_state.p;
  b.summary = sh.summary.p;
  b.status = sh.status.p;
  b.ping = b.pong = nullptr;
  b.scratch = nullptr;
  b.n_cols = (uint32_t)sh.n_cols;
  b.nz = cfg.n_points;
  b.n_species = cfg.n_species;
  b.steps = (uint32_t)cfg.time_steps;
  b.step_begin = 0;
  b.step_end = b.steps;
  b.stages = stages;
  b.outlet_stride = sh.outlet.p ? (uint32_t)cfg.outlet_stride : 0u;
  b.snapshot_stride = sh.snapshots.p ? (uint32_t)cfg.snapshot_stride : 0u;
  b.out_steps = sh.out_steps.p;
  b.snap_steps = sh.snap_steps.p;
  b.h_out_steps = sh.out_steps.p ? plan->out_steps.data() : nullptr;
  b.h_snap_steps = sh.snap_steps.p ? plan->snap_steps.data() : nullptr;
  b.series = nullptr;
  b.n_out = (uint32_t)plan->n_out;
  b.n_snap = (uint32_t)plan->n_snap;
  b.method = cfg.method;
  b.arith = cfg.arith;
  if (!plan->multi && cfg.arith == CHROM_ARITH_FAST && !getenv("CHROM_NO_POLY")) {
    // FAST_POLY (single_rhs.cuh) saves one FP64 instruction per cell-stage but subtracts two numbers
    // that agree to a factor 1 + B/A; use it unless some column's isotherm is extremely steep.
    double worst = 0.0;
    for (uint64_t c = 0; c < sh.n_cols; ++c) {
      const chrom_single_col& q = scols[sh.col0 + c];
      const double n_bar = q.fe / (1.0 + q.fe) * q.port_number;
      const double ratio = std::fabs(q.fe * n_bar * q.langmuir_k / (1.0 + q.fe * q.lambda));
      worst = (ratio > worst || ratio != ratio) ? (ratio != ratio ? 1e300 : ratio) : worst;
    }
    if (worst <= 1e3) b.arith = CHROM_KARITH_POLY;
  }
  if (!plan->multi && cfg.arith == CHROM_ARITH_FAST) {
    // The FAST forms fold Cg = -ue/dz into the denominator (A/Cg, B/Cg): a column at rest (ue == 0) or with a
    // non-finite coefficient would turn a state the reference keeps finite into NaN.  Such a batch runs EXACT.
    bool safe = true;
    for (uint64_t c = 0; c < sh.n_cols && safe; ++c) {
      const chrom_single_col& q = scols[sh.col0 + c];
      const double cg = -q.ue / q.dz;
      const double a = (1.0 + q.fe * q.lambda) / cg, bb = q.fe * (q.fe / (1.0 + q.fe) * q.port_number) * q.langmuir_k / cg;
      safe = (q.ue != 0.0) && std::isfinite(a) && std::isfinite(bb) && std::isfinite(q.langmuir_k);
    }
    if (!safe) b.arith = CHROM_ARITH_EXACT;
  }
  // rk4.rs:238 `dt = total_time / (time_steps as f64)`, :284 `dt / 2.0`, :311 `dt / 6.0`
  b.dt = cfg.total_time / (double)cfg.time_steps;
  b.half_dt = b.dt / 2.0;
  b.sixth_dt = b.dt / 6.0;

  std::string why;
  bool ok = false;
  if (!plan->multi) {
    ok = single_resident_choose(b, d.sm_count, &sh.g, &why);
    if (!ok) ok = single_stream_choose(b, d.sm_count, &sh.g, &why);
    if (ok && sh.g.kind == 3) {
      if (want & CHROM_WANT_SUMMARY) {  // the summary is formed from a full-rate detector series kept on the device
        CUDA_TRY(ctx, sh.series.alloc(sh.n_cols * (cfg.time_steps + 1), d.stream));
        b.series = sh.series.p;
      }
      CUDA_TRY(ctx, sh.ping.alloc(sh.n_cols * state_len, d.stream));
      CUDA_TRY(ctx, sh.pong.alloc(sh.n_cols * (state_len + single_stream_extra_doubles()), d.stream));  // + status code + coefficients per column
      b.ping = sh.ping.p;
      b.pong = sh.pong.p;
    }
  } else {
    b.arith = multi_kernel_arith((int)cfg.n_species, cfg.arith);
    const char* force_tiled = getenv("CHROM_MULTI_TILED");  // experiments / tests: tiles even when the column fits
    const bool forced = force_tiled && atoi(force_tiled) != 0;
    // FAST structured arithmetic, n <= 8, column short enough: everything in registers (multi_reg.cu)
    const char* no_reg = getenv("CHROM_MULTI_NO_REG");
    if (!forced && b.arith == CHROM_KARITH_STRUCT && !(no_reg && atoi(no_reg) != 0))
      ok = multi_reg_choose(b, d.sm_count, &sh.g, &why);
    // Outside the register-resident kernel the 1x1 / 2x2 systems are cheapest as a dense register LU (measured on
    // n = 2, nz = 8192: 1.63e11 vs 1.23e11 cell-stage/s for the closed-form elimination; the order flips at n >= 4).
    if (!ok && b.arith == CHROM_KARITH_STRUCT && cfg.n_species <= 2 && cfg.arith != CHROM_ARITH_FAST_STRUCTURED)
      b.arith = CHROM_ARITH_FAST;
    if (ok) {
      // no scratch, no streamed state
    } else if ((ok = !forced && multi_resident_choose(b, d.sm_count, &sh.g, &why))) {
      // shared-memory-resident: no scratch, no streamed state
    } else if ((ok = multi_tiled_choose(b, d.sm_count, &sh.g, &why))) {
      // long columns / many species: the state streams through HBM/L2 in tiles (multi_tiled.cu)
      CUDA_TRY(ctx, sh.ping.alloc(sh.n_cols * state_len, d.stream));
      CUDA_TRY(ctx, sh.pong.alloc(sh.n_cols * (state_len + multi_tiled_extra_doubles(cfg.n_species)), d.stream));
      b.ping = sh.ping.p;
      b.pong = sh.pong.p;
      const size_t scratch = multi_tiled_scratch_doubles(b, sh.g);
      if (scratch) {
        CUDA_TRY(ctx, sh.scratch.alloc(scratch, d.stream));
        b.scratch = sh.scratch.p;
      }
    }
  }
  if (!ok) return fail(ctx, CHROM_ERR_UNSUPPORTED, "no kernel covers this shape: " + why);
  return CHROM_OK;
}

int32_t make_plan(chrom_cuda_ctx* ctx, const chrom_run_cfg* cfg, bool multi, const chrom_single_col* scols,
                  const chrom_multi_col* mcols, uint64_t n_cols, const chrom_species* species,
                  uint64_t n_species_total, const double* inj_tables, uint32_t n_tables, const double* y0,
                  uint32_t want, chrom_cuda_plan** out) {
  if (!ctx) return fail(nullptr, CHROM_ERR_INVALID, "ctx is NULL");
  if (!out) return fail(ctx, CHROM_ERR_INVALID, "out is NULL");
  *out = nullptr;
  if (int32_t rc = check_cfg(ctx, cfg, multi)) return rc;
  if (n_cols == 0) return fail(ctx, CHROM_ERR_INVALID, "n_cols must be > 0");
  if (n_cols > 0x7fffffffull) return fail(ctx, CHROM_ERR_UNSUPPORTED, "n_cols > 2^31-1 is not supported");
  if ((!multi && !scols) || (multi && (!mcols || !species)))
    return fail(ctx, CHROM_ERR_INVALID, "column / species descriptors are NULL");
  if (!inj_tables || n_tables == 0) return fail(ctx, CHROM_ERR_INVALID, "inj_tables is NULL / n_tables == 0");
  const uint32_t n = cfg->n_species;
  if (!multi) {
    for (uint64_t c = 0; c < n_cols; ++c) {
      if (scols[c].inj_table >= n_tables)
        return fail(ctx, CHROM_ERR_INVALID, format("column %llu: inj_table %u >= n_tables %u",
                                                   (unsigned long long)c, scols[c].inj_table, n_tables));
      if (scols[c].detector_cell > cfg->n_points)
        return fail(ctx, CHROM_ERR_INVALID, format("column %llu: detector_cell %u exceeds n_points %u (0 = outlet, k = cell k - 1)",
                                                   (unsigned long long)c, scols[c].detector_cell, cfg->n_points));
    }
  } else {
    for (uint64_t c = 0; c < n_cols; ++c) {
      if (mcols[c].detector_cell > cfg->n_points)
        return fail(ctx, CHROM_ERR_INVALID, format("column %llu: detector_cell %u exceeds n_points %u (0 = outlet, k = cell k - 1)",
                                                   (unsigned long long)c, mcols[c].detector_cell, cfg->n_points));
      if ((uint64_t)mcols[c].species_off + n > n_species_total)
        return fail(ctx, CHROM_ERR_INVALID, format("column %llu: species range exceeds n_species_total", (unsigned long long)c));
      for (uint32_t k = 0; k < n; ++k)
        if (species[mcols[c].species_off + k].inj_table >= n_tables)
          return fail(ctx, CHROM_ERR_INVALID, format("column %llu species %u: inj_table >= n_tables", (unsigned long long)c, k));
    }
  }

  std::unique_ptr<chrom_cuda_plan> plan(new chrom_cuda_plan);
  plan->ctx = ctx;
  plan->cfg = *cfg;
  plan->multi = multi;
  plan->want = want;
  plan->n_cols = n_cols;
  plan->cfg.outlet_steps = plan->cfg.snapshot_steps = nullptr;  // the caller's lists are copied, never kept
  plan->n_out = chrom_cuda_cfg_n_out(cfg);
  plan->n_snap = chrom_cuda_cfg_n_snap(cfg);
  if (cfg->outlet_steps) {
    for (uint64_t k = 0; k < cfg->n_outlet_steps; ++k) plan->out_steps.push_back((uint32_t)cfg->outlet_steps[k]);
    plan->out_steps.push_back(kNoSample);
  }
  if (cfg->snapshot_steps) {
    for (uint64_t k = 0; k < cfg->n_snapshot_steps; ++k) plan->snap_steps.push_back((uint32_t)cfg->snapshot_steps[k]);
    plan->snap_steps.push_back(kNoSample);
  }
  plan->state_len = (size_t)cfg->n_points * n;
  plan->series = n;

  const uint64_t G = ctx->devs.size();
  const double t0 = now_ms();
  for (uint64_t gidx = 0; gidx < G; ++gidx) {
    uint64_t c0 = 0, cnt = 0;
    chrom_cuda_shard_range(n_cols, (uint32_t)G, (uint32_t)gidx, &c0, &cnt);
    if (cnt == 0) break;
    std::unique_ptr<Shard> sh(new Shard);
    sh->dev = (int)gidx;
    sh->col0 = c0;
    sh->n_cols = cnt;
    if (int32_t rc = build_shard(plan.get(), *sh, scols, mcols, species, inj_tables, n_tables, y0)) return rc;
    plan->shards.push_back(std::move(sh));
  }
  for (auto& sh : plan->shards) {
    DeviceState& d = ctx->devs[sh->dev];
    CUDA_TRY(ctx, cudaSetDevice(d.id));
    CUDA_TRY(ctx, cudaStreamSynchronize(d.stream));
  }
  plan->h2d_ms = now_ms() - t0;
  plan->desc = plan->shards[0]->g.name;
  if (plan->shards.size() > 1) plan->desc += format(" x%zu devices", plan->shards.size());
  *out = plan.release();
  return CHROM_OK;
}

// The descriptor of columns [c0, c0 + cnt) of a shard: every per-column pointer offset.
BatchDev sub_batch(const chrom_cuda_plan* plan, const Shard& sh, uint64_t c0, uint64_t cnt) {
  BatchDev b = sh.b;
  const size_t sl = plan->state_len, ser = plan->series;
  if (b.scols) b.scols += c0;
  if (b.mcols) b.mcols += c0;  // species offsets are absolute indices into the shard's species array
  if (b.y0) b.y0 += c0 * sl;
  if (b.outlet) b.outlet += c0 * ser * plan->n_out;
  if (b.snapshots) b.snapshots += c0 * plan->n_snap * sl;
  if (b.final_state) b.final_state += c0 * sl;
  if (b.summary) b.summary += c0 * ser * CHROM_SUMMARY_LEN;
  if (b.status) b.status += c0;
  b.n_cols = (uint32_t)cnt;
  return b;
}

int32_t launch_batch(chrom_cuda_ctx* ctx, const BatchDev& b, const LaunchGeom& g, cudaStream_t s) {
  cudaError_t e = cudaSuccess;
  switch (g.kind) {
    case 1:
    case 2:
      e = single_resident_launch(b, g, s);
      break;
    case 3:
      e = single_stream_launch(b, g, s);
      break;
    case 4:
      e = multi_resident_launch(b, g, s);
      break;
    case 5:
      e = multi_tiled_launch(b, g, s);
      break;
    case 6:
      e = multi_reg_launch(b, g, s);
      break;
    case 7:
      e = multi_reg2_launch(b, g, s);
      break;
    default:
      return fail(ctx, CHROM_ERR_UNSUPPORTED, "plan has no kernel");
  }
  if (e != cudaSuccess) return fail(ctx, CHROM_ERR_CUDA, format("kernel launch failed: %s", cudaGetErrorString(e)));
  return CHROM_OK;
}

}  // namespace

extern "C" {

const char* chrom_cuda_version(void) { return "chrom_cuda 0.2.0 (sm_100a)"; }

int32_t chrom_cuda_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return n;
}

int32_t chrom_cuda_create(const int32_t* device_ids, int32_t n_devices, chrom_cuda_ctx** out) {
  if (!out) return fail(nullptr, CHROM_ERR_INVALID, "out is NULL");
  *out = nullptr;
  const int visible = chrom_cuda_device_count();
  if (visible <= 0)
    return fail(nullptr, CHROM_ERR_NO_DEVICE, "no CUDA device visible; libchrom_cuda has no CPU fallback");
  std::vector<int> ids;
  if (device_ids) {
    for (int i = 0; i < n_devices; ++i) ids.push_back(device_ids[i]);
  } else {
    const int n = (n_devices <= 0) ? visible : n_devices;
    for (int i = 0; i < n; ++i) ids.push_back(i);
  }
  if (ids.empty()) return fail(nullptr, CHROM_ERR_INVALID, "empty device list");
  std::unique_ptr<chrom_cuda_ctx> ctx(new chrom_cuda_ctx);
  for (int id : ids) {
    if (id < 0 || id >= visible)
      return fail(nullptr, CHROM_ERR_INVALID, format("device id %d out of range (visible: %d)", id, visible));
    DeviceState d;
    d.id = id;
    CUDA_TRY(nullptr, cudaSetDevice(id));
    cudaDeviceProp prop;
    CUDA_TRY(nullptr, cudaGetDeviceProperties(&prop, id));
    if (prop.major < 10)
      return fail(nullptr, CHROM_ERR_NO_DEVICE,
                  format("device %d (%s, sm_%d%d) is not a Blackwell sm_100 part; this library ships sm_100a code only",
                         id, prop.name, prop.major, prop.minor));
    d.sm_count = prop.multiProcessorCount;
    CUDA_TRY(nullptr, cudaStreamCreateWithFlags(&d.stream, cudaStreamNonBlocking));
    CUDA_TRY(nullptr, cudaStreamCreateWithFlags(&d.stream2, cudaStreamNonBlocking));
    CUDA_TRY(nullptr, cudaStreamCreateWithFlags(&d.copy_stream, cudaStreamNonBlocking));
    {
      cudaMemPool_t pool;
      CUDA_TRY(nullptr, cudaDeviceGetDefaultMemPool(&pool, id));
      // keep freed blocks cached for the next solve; the previous threshold comes back at chrom_cuda_destroy (the
      // default pool is process-wide: other libraries on the device must not inherit this setting for good)
      CUDA_TRY(nullptr, cudaMemPoolGetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &d.pool_threshold0));
      unsigned long long keep = ~0ull;
      CUDA_TRY(nullptr, cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep));
    }
    for (auto& ev : d.ev) CUDA_TRY(nullptr, cudaEventCreate(&ev));
    ctx->devs.push_back(d);
  }
  *out = ctx.release();
  return CHROM_OK;
}

void chrom_cuda_destroy(chrom_cuda_ctx* ctx) {
  if (!ctx) return;
  for (auto& d : ctx->devs) {
    cudaSetDevice(d.id);
    for (auto& ev : d.ev)
      if (ev) cudaEventDestroy(ev);
    if (d.stream) cudaStreamDestroy(d.stream);
    if (d.stream2) cudaStreamDestroy(d.stream2);
    if (d.copy_stream) cudaStreamDestroy(d.copy_stream);
    if (d.flush_buf) cudaFree(d.flush_buf);
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, d.id) == cudaSuccess) {
      cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &d.pool_threshold0);
      cudaMemPoolTrimTo(pool, 0);  // hand the cached output buffers back
    }
  }
  delete ctx;
}

const char* chrom_cuda_last_error(const chrom_cuda_ctx* ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }

int32_t chrom_cuda_last_timing(const chrom_cuda_ctx* ctx, chrom_cuda_timing* out) {
  if (!ctx || !out) return CHROM_ERR_INVALID;
  *out = ctx->timing;
  return CHROM_OK;
}

int32_t chrom_cuda_host_alloc(size_t bytes, void** out) {
  if (!out) return CHROM_ERR_INVALID;
  *out = nullptr;
  unsigned flags = cudaHostAllocPortable;
  if (const char* e = getenv("CHROM_HOST_WC"))  // write-combined pages: device->host DMA without CPU cache snooping
    if (atoi(e) != 0) flags |= cudaHostAllocWriteCombined;
  if (cudaHostAlloc(out, bytes, flags) != cudaSuccess) {
    cudaGetLastError();
    return CHROM_ERR_CUDA;
  }
  return CHROM_OK;
}

void chrom_cuda_host_free(void* p) {
  if (p) cudaFreeHost(p);
}

int32_t chrom_cuda_bind_host_thread(chrom_cuda_ctx* ctx, int32_t device_index, int32_t* numa_node, int32_t* n_cpus) {
  if (numa_node) *numa_node = -1;
  if (n_cpus) *n_cpus = 0;
  if (!ctx || device_index < 0 || device_index >= (int)ctx->devs.size()) return CHROM_ERR_INVALID;
  char bus[32] = {0};
  CUDA_TRY(ctx, cudaDeviceGetPCIBusId(bus, sizeof bus, ctx->devs[device_index].id));
  for (char* c = bus; *c; ++c) *c = (char)tolower(*c);
  int node = -1;
  if (FILE* f = fopen(format("/sys/bus/pci/devices/%s/numa_node", bus).c_str(), "r")) {
    if (fscanf(f, "%d", &node) != 1) node = -1;
    fclose(f);
  }
  if (node < 0) return CHROM_OK;  // single-node host or no topology information: leave the thread alone
  cpu_set_t allowed, want;
  CPU_ZERO(&want);
  if (sched_getaffinity(0, sizeof allowed, &allowed) != 0) return CHROM_OK;
  if (FILE* f = fopen(format("/sys/devices/system/node/node%d/cpulist", node).c_str(), "r")) {
    int a = 0, b = 0;
    for (;;) {  // "0-15,32-47"
      if (fscanf(f, "%d", &a) != 1) break;
      b = a;
      int ch = fgetc(f);
      if (ch == '-') {
        if (fscanf(f, "%d", &b) != 1) break;
        ch = fgetc(f);
      }
      for (int c = a; c <= b && c < CPU_SETSIZE; ++c)
        if (CPU_ISSET(c, &allowed)) CPU_SET(c, &want);
      if (ch != ',') break;
    }
    fclose(f);
  }
  const int count = CPU_COUNT(&want);
  if (count == 0) return CHROM_OK;  // the process is confined to another node: nothing sensible to do
  if (sched_setaffinity(0, sizeof want, &want) != 0) return CHROM_OK;
  if (node < 1024) {  // prefer the node for this thread's allocations (MPOL_PREFERRED = 1); best effort
    unsigned long mask[1024 / (8 * sizeof(unsigned long))] = {0};
    mask[node / (8 * sizeof(unsigned long))] |= 1ul << (node % (8 * sizeof(unsigned long)));
    (void)syscall(SYS_set_mempolicy, 1, mask, 1024ul + 1);
  }
  if (numa_node) *numa_node = node;
  if (n_cpus) *n_cpus = count;
  return CHROM_OK;
}

// ---- host helpers ---------------------------------------------------------------------------

const char* chrom_cuda_validate_cfg(const chrom_run_cfg* cfg) {
  // SolverType::validate, TimeEvolution arm — src/solver/traits.rs:172-185 (same comparisons).
  if (cfg->total_time <= 0.0) return "Total time must be positive";
  if (cfg->time_steps == 0) return "TimeSteps must be greater than 0";
  return nullptr;
}

uint64_t chrom_cuda_n_samples(uint64_t time_steps, uint64_t stride) { return stride == 0 ? 0 : time_steps / stride + 1; }

uint64_t chrom_cuda_cfg_n_out(const chrom_run_cfg* cfg) {
  if (!cfg) return 0;
  return cfg->outlet_steps ? cfg->n_outlet_steps : chrom_cuda_n_samples(cfg->time_steps, cfg->outlet_stride);
}

uint64_t chrom_cuda_cfg_n_snap(const chrom_run_cfg* cfg) {
  if (!cfg) return 0;
  return cfg->snapshot_steps ? cfg->n_snapshot_steps : chrom_cuda_n_samples(cfg->time_steps, cfg->snapshot_stride);
}

// sample_indices — src/physics/traits.rs:1010-1023 (integer division, last index forced to total - 1)
uint64_t chrom_cuda_sample_indices(uint64_t total, uint64_t n, uint64_t* out, uint64_t cap) {
  auto put = [&](uint64_t k, uint64_t v) {
    if (out && k < cap) out[k] = v;
  };
  if (n == 0 || n >= total) {
    for (uint64_t i = 0; i < total; ++i) put(i, i);
    return total;
  }
  if (n == 1) {
    put(0, 0);
    return 1;
  }
  for (uint64_t i = 0; i < n; ++i) put(i, (i * (total - 1)) / (n - 1));
  put(n - 1, total - 1);
  return n;
}

// Detector::node_index — src/domain/detector.rs:285-290: `(z / dz).round() as usize` (round half away from zero,
// saturating cast), `min(n_points - 1)`
uint32_t chrom_cuda_node_index(double position, double column_length, uint32_t n_points) {
  if (n_points == 0) return 0;
  const double dz = column_length / (double)n_points;
  const double r = std::round(position / dz);
  uint64_t idx = 0;
  if (r != r || r <= 0.0) idx = 0;  // NaN and negatives saturate to 0 in a Rust `as usize`
  else if (r >= 1.8446744073709552e19) idx = ~0ull;
  else idx = (uint64_t)r;
  return (uint32_t)std::min<uint64_t>(idx, (uint64_t)n_points - 1);
}

void chrom_cuda_time_points(double total_time, uint64_t time_steps, double* out) {
  const double dt = total_time / (double)time_steps;
  out[0] = 0.0;                                                                   // rk4.rs:254
  for (uint64_t s = 0; s < time_steps; ++s) out[s + 1] = ((double)s + 1.0) * dt;  // rk4.rs:327
}

void chrom_cuda_shard_range(uint64_t n_cols, uint32_t n_devices, uint32_t device_index, uint64_t* col0,
                            uint64_t* count) {
  const uint64_t G = n_devices ? n_devices : 1;
  const uint64_t per = (n_cols + G - 1) / G;
  const uint64_t c0 = std::min<uint64_t>((uint64_t)device_index * per, n_cols);
  if (col0) *col0 = c0;
  if (count) *count = std::min<uint64_t>(per, n_cols - c0);
}

uint32_t chrom_cuda_table_stages(int32_t method) { return method == CHROM_RK4 ? 3u : 1u; }

int32_t chrom_cuda_tabulate_injection(int32_t kind, double p0, double p1, double p2, int32_t method, double total_time,
                                      uint64_t time_steps, double* out) {
  if (!out || time_steps == 0) return CHROM_ERR_INVALID;
  if (kind < CHROM_INJ_NONE || kind > CHROM_INJ_RECTANGLE) return CHROM_ERR_INVALID;
  const double dt = total_time / (double)time_steps;
  if (method == CHROM_RK4) {
    for (uint64_t step = 0; step < time_steps; ++step) {
      const double t = (double)step * dt;                                     // rk4.rs:267
      out[step * 3 + 0] = injection_evaluate(kind, p0, p1, p2, t);            // ctx1: t
      out[step * 3 + 1] = injection_evaluate(kind, p0, p1, p2, t + dt / 2.0);  // ctx2, ctx3: t + dt/2.0
      out[step * 3 + 2] = injection_evaluate(kind, p0, p1, p2, t + dt);        // ctx4: t + dt
    }
  } else if (method == CHROM_EULER) {
    for (uint64_t step = 0; step < time_steps; ++step)
      out[step] = injection_evaluate(kind, p0, p1, p2, dt * (double)step);  // euler.rs:229
  } else {
    return CHROM_ERR_INVALID;
  }
  return CHROM_OK;
}

size_t chrom_cuda_status_message(int64_t status, char* buf, size_t buf_len) {
  if (status == 0) {
    if (buf && buf_len) buf[0] = '\0';
    return 0;
  }
  std::string m;
  if (status > 0)
    m = format("NaN detected in Concentration at step %lld. This indicates numerical instability. "
               "Try reducing time step (increase time_steps parameter).", (long long)status);
  else
    m = format("Infinity detected in Concentration at step %lld. This indicates numerical overflow. "
               "Try reducing time step or check physics model for division by zero.", (long long)-status);
  if (buf && buf_len) {
    const size_t k = std::min(buf_len - 1, m.size());
    std::memcpy(buf, m.data(), k);
    buf[k] = '\0';
  }
  return m.size();
}

// ---- plans ------------------------------------------------------------------------------------

int32_t chrom_cuda_plan_single(chrom_cuda_ctx* ctx, const chrom_run_cfg* cfg, const chrom_single_col* cols,
                               uint64_t n_cols, const double* inj_tables, uint32_t n_tables, const double* y0,
                               uint32_t want, chrom_cuda_plan** out) {
  return make_plan(ctx, cfg, false, cols, nullptr, n_cols, nullptr, 0, inj_tables, n_tables, y0, want, out);
}

int32_t chrom_cuda_plan_multi(chrom_cuda_ctx* ctx, const chrom_run_cfg* cfg, const chrom_multi_col* cols,
                              uint64_t n_cols, const chrom_species* species, uint64_t n_species_total,
                              const double* inj_tables, uint32_t n_tables, const double* y0, uint32_t want,
                              chrom_cuda_plan** out) {
  return make_plan(ctx, cfg, true, nullptr, cols, n_cols, species, n_species_total, inj_tables, n_tables, y0, want, out);
}

int32_t chrom_cuda_plan_execute(chrom_cuda_plan* plan) {
  if (!plan) return CHROM_ERR_INVALID;
  chrom_cuda_ctx* ctx = plan->ctx;
  const double t0 = now_ms();
  uint64_t launches = 0;
  for (auto& sh : plan->shards) {
    DeviceState& d = ctx->devs[sh->dev];
    CUDA_TRY(ctx, cudaSetDevice(d.id));
    CUDA_TRY(ctx, cudaMemsetAsync(sh->status.p, 0, sh->n_cols * sizeof(long long), d.stream));
    CUDA_TRY(ctx, cudaEventRecord(d.ev[2], d.stream));
    if (sh->g.kind == 3 || (sh->g.kind == 5 && sh->g.launches_per_execute > 3)) {
      // one launch per (few) time step(s): replay them as a CUDA graph (captured on first use)
      if (!sh->graph) {
        cudaGraph_t graph = nullptr;
        CUDA_TRY(ctx, cudaStreamBeginCapture(d.stream, cudaStreamCaptureModeThreadLocal));
        const int32_t rc = launch_batch(ctx, sh->b, sh->g, d.stream);
        const cudaError_t ce = cudaStreamEndCapture(d.stream, &graph);
        if (rc != CHROM_OK) return rc;
        CUDA_TRY(ctx, ce);
        const cudaError_t ie = cudaGraphInstantiate(&sh->graph, graph, 0);
        cudaGraphDestroy(graph);
        CUDA_TRY(ctx, ie);
        CUDA_TRY(ctx, cudaEventRecord(d.ev[2], d.stream));  // instantiation is not part of the solve
      }
      CUDA_TRY(ctx, cudaGraphLaunch(sh->graph, d.stream));
    } else if (int32_t rc = launch_batch(ctx, sh->b, sh->g, d.stream)) {
      return rc;
    }
    CUDA_TRY(ctx, cudaEventRecord(d.ev[3], d.stream));
    launches += sh->g.launches_per_execute;
  }
  float kmax = 0.f;
  for (auto& sh : plan->shards) {
    DeviceState& d = ctx->devs[sh->dev];
    CUDA_TRY(ctx, cudaSetDevice(d.id));
    CUDA_TRY(ctx, cudaEventSynchronize(d.ev[3]));
    CUDA_TRY(ctx, cudaEventElapsedTime(&sh->kernel_ms, d.ev[2], d.ev[3]));
    kmax = std::max(kmax, sh->kernel_ms);
  }
  ctx->timing = chrom_cuda_timing{};
  ctx->timing.kernel_ms = kmax;
  ctx->timing.kernel_launches = launches;
  ctx->timing.h2d_ms = plan->h2d_ms;
  ctx->timing.h2d_bytes = plan->h2d_bytes;
  ctx->timing.total_ms = now_ms() - t0;
  return CHROM_OK;
}

int32_t chrom_cuda_plan_download(chrom_cuda_plan* plan, double* outlet, double* snapshots, double* final_state,
                                 double* summary, int64_t* status) {
  if (!plan) return CHROM_ERR_INVALID;
  chrom_cuda_ctx* ctx = plan->ctx;
  const double t0 = now_ms();
  uint64_t bytes = 0;
  auto copy = [&](Shard& sh, cudaStream_t s, void* host, const void* dev, size_t per_col_bytes) -> cudaError_t {
    if (!host || !dev) return cudaSuccess;
    bytes += sh.n_cols * per_col_bytes;
    return cudaMemcpyAsync((char*)host + sh.col0 * per_col_bytes, dev, sh.n_cols * per_col_bytes,
                           cudaMemcpyDeviceToHost, s);
  };
  for (auto& sh : plan->shards) {
    DeviceState& d = ctx->devs[sh->dev];
    CUDA_TRY(ctx, cudaSetDevice(d.id));
    CUDA_TRY(ctx, copy(*sh, d.stream, outlet, sh->outlet.p, plan->series * plan->n_out * sizeof(double)));
    CUDA_TRY(ctx, copy(*sh, d.stream, snapshots, sh->snapshots.p, plan->n_snap * plan->state_len * sizeof(double)));
    CUDA_TRY(ctx, copy(*sh, d.stream, final_state, sh->final_state.p, plan->state_len * sizeof(double)));
    CUDA_TRY(ctx, copy(*sh, d.stream, summary, sh->summary.p, plan->series * CHROM_SUMMARY_LEN * sizeof(double)));
    CUDA_TRY(ctx, copy(*sh, d.stream, status, sh->status.p, sizeof(long long)));
  }
  for (auto& sh : plan->shards) {
    DeviceState& d = ctx->devs[sh->dev];
    CUDA_TRY(ctx, cudaSetDevice(d.id));
    CUDA_TRY(ctx, cudaStreamSynchronize(d.stream));
  }
  ctx->timing.d2h_ms = now_ms() - t0;
  ctx->timing.d2h_bytes = bytes;
  return CHROM_OK;
}

int32_t chrom_cuda_plan_rsf(chrom_cuda_plan* a, chrom_cuda_plan* b, double* rsf) {
  if (!a || !b || !rsf) return CHROM_ERR_INVALID;
  chrom_cuda_ctx* ctx = a->ctx;
  if (b->ctx != ctx) return fail(ctx, CHROM_ERR_INVALID, "plan_rsf: the two plans belong to different contexts");
  if (a->n_cols != b->n_cols || a->series != b->series || a->n_out != b->n_out || a->shards.size() != b->shards.size() ||
      a->cfg.total_time != b->cfg.total_time || a->cfg.time_steps != b->cfg.time_steps ||
      a->cfg.outlet_stride != b->cfg.outlet_stride || a->out_steps != b->out_steps)
    return fail(ctx, CHROM_ERR_INVALID,
                "plan_rsf: the two plans must hold the same columns, species and outlet sampling (one time grid)");
  if (a->n_out < 2) return fail(ctx, CHROM_ERR_INVALID, "plan_rsf: needs at least two outlet samples");
  const double dt = a->cfg.total_time / (double)a->cfg.time_steps;
  std::vector<DevBuf<double>> out(a->shards.size());
  for (size_t si = 0; si < a->shards.size(); ++si) {
    Shard& sa = *a->shards[si];
    Shard& sb = *b->shards[si];
    if (!sa.outlet.p || !sb.outlet.p)
      return fail(ctx, CHROM_ERR_INVALID, "plan_rsf: both plans must have been built with CHROM_WANT_OUTLET");
    DeviceState& d = ctx->devs[sa.dev];
    CUDA_TRY(ctx, cudaSetDevice(d.id));
    const size_t n_series = sa.n_cols * a->series;
    CUDA_TRY(ctx, out[si].alloc(n_series, d.stream));
    CUDA_TRY(ctx, rsf_launch(sa.outlet.p, sb.outlet.p, n_series, (uint32_t)a->n_out, sa.out_steps.p,
                             (uint32_t)a->cfg.outlet_stride, dt, out[si].p, d.stream));
    CUDA_TRY(ctx, cudaMemcpyAsync(rsf + sa.col0 * a->series, out[si].p, n_series * sizeof(double), cudaMemcpyDeviceToHost,
                                  d.stream));
  }
  for (auto& sh : a->shards) {
    DeviceState& d = ctx->devs[sh->dev];
    CUDA_TRY(ctx, cudaSetDevice(d.id));
    CUDA_TRY(ctx, cudaStreamSynchronize(d.stream));
  }
  return CHROM_OK;
}

int32_t chrom_cuda_plan_describe(const chrom_cuda_plan* plan, char* buf, size_t buf_len) {
  if (!plan || !buf || buf_len == 0) return CHROM_ERR_INVALID;
  snprintf(buf, buf_len, "%s", plan->desc.c_str());
  return CHROM_OK;
}

void chrom_cuda_plan_destroy(chrom_cuda_plan* plan) { delete plan; }  // ~chrom_cuda_plan drains the streams first

// ---- one-shot solves --------------------------------------------------------------------------

// One-shot solve with large outputs: the shard is cut into wave-aligned column chunks; chunk i's
// outputs go back to the host on the copy stream while chunk i+1 computes (north_star: "copied
// back on a side stream").  Returns CHROM_OK, an error, or 1 when the batch is not worth chunking.
static int32_t solve_pipelined(chrom_cuda_ctx* ctx, chrom_cuda_plan* plan, double* outlet, double* snapshots,
                               double* final_state, double* summary, int64_t* status) {
  const size_t sl = plan->state_len, ser = plan->series;
  const size_t per_col = (outlet ? ser * plan->n_out : 0) + (snapshots ? plan->n_snap * sl : 0) +
                         (final_state ? sl : 0) + (summary ? ser * CHROM_SUMMARY_LEN : 0);
  bool worth = false;
  for (auto& sh : plan->shards) {
    const bool big = sh->n_cols * per_col * sizeof(double) >= ((size_t)64 << 20);
    // several waves of columns: chunks;  a register-resident single-species batch of any size: its (only) chunk is
    // advanced in time slices, the outlet samples of a slice going back while the next one computes
    worth |= big && sh->g.cols_per_wave > 0 && sh->g.kind != 3 && sh->n_cols >= 2ull * sh->g.cols_per_wave;
    worth |= big && (sh->g.kind == 1 || sh->g.kind == 2) && outlet && final_state && !snapshots && plan->out_steps.empty();
  }
  if (!worth) return 1;
  const double t0 = now_ms();
  uint64_t launches = 0, bytes = 0;
  // events of the chunks and slices, per shard: destroyed on EVERY exit path (CUDA_TRY returns from inside the loops;
  // the plan's destructor, which solve_common runs next, drains the streams the events were recorded on)
  struct EventBag {
    std::vector<std::vector<cudaEvent_t>> per_shard;
    explicit EventBag(size_t n) : per_shard(n) {}
    ~EventBag() {
      for (auto& v : per_shard)
        for (cudaEvent_t e : v) cudaEventDestroy(e);
    }
  } bag(plan->shards.size());
  std::vector<std::vector<cudaEvent_t>>& events = bag.per_shard;
  int32_t rc = CHROM_OK;
  // round-robin over shards so every device has work queued before any host-side wait
  std::vector<uint64_t> next(plan->shards.size(), 0);
  for (size_t si = 0; si < plan->shards.size(); ++si) {
    Shard& sh = *plan->shards[si];
    DeviceState& d = ctx->devs[sh.dev];
    CUDA_TRY(ctx, cudaSetDevice(d.id));
    CUDA_TRY(ctx, cudaMemsetAsync(sh.status.p, 0, sh.n_cols * sizeof(long long), d.stream));
    CUDA_TRY(ctx, cudaEventRecord(d.ev[2], d.stream));
    CUDA_TRY(ctx, cudaStreamWaitEvent(d.stream2, d.ev[2], 0));  // the status memset precedes every chunk
  }
  std::vector<uint64_t> chunk_idx(plan->shards.size(), 0);
  bool more = true;
  while (more && rc == CHROM_OK) {
    more = false;
    for (size_t si = 0; si < plan->shards.size() && rc == CHROM_OK; ++si) {
      Shard& sh = *plan->shards[si];
      if (next[si] >= sh.n_cols) continue;
      DeviceState& d = ctx->devs[sh.dev];
      CUDA_TRY(ctx, cudaSetDevice(d.id));
      const uint64_t waves = sh.g.cols_per_wave ? (sh.n_cols + sh.g.cols_per_wave - 1) / sh.g.cols_per_wave : 1;
      const uint64_t chunk = std::max<uint64_t>(1, (uint64_t)sh.g.cols_per_wave * std::max<uint64_t>(1, (waves + 7) / 8));
      const uint64_t c0 = next[si], cnt = std::min<uint64_t>(chunk, sh.n_cols - c0);
      next[si] = c0 + cnt;
      more |= next[si] < sh.n_cols;
      // Consecutive chunks alternate between two compute streams: a kernel boundary in ONE stream is a full
      // barrier (the next wave cannot start until the last block of this one has ended); across two streams the
      // blocks of chunk i+1 fill the SMs as those of chunk i drain.
      cudaStream_t cs = (chunk_idx[si]++ & 1u) ? d.stream2 : d.stream;
      const BatchDev b = sub_batch(plan, sh, c0, cnt);
      // The copy of the LAST chunk is the one nothing overlaps.  Where the kernel can continue from a carried
      // state (register-resident single-species kernels, no on-device summary, no snapshots), that chunk is
      // advanced in kSlices time slices and the outlet samples of a slice go back while the next one computes.
      const bool last_chunk = next[si] >= sh.n_cols;
      // (a batch that is one chunk has nothing else to hide its copy behind: more, shorter slices)
      const uint32_t kSlices = (c0 == 0 && last_chunk) ? 8u : 4u;
      if (last_chunk && (sh.g.kind == 1 || sh.g.kind == 2) && !snapshots && outlet && b.outlet && !b.out_steps &&
          b.final_state && b.outlet_stride > 0 && b.steps >= 64u * kSlices) {
        const size_t n_out = plan->n_out;
        for (uint32_t k = 0; k < kSlices && rc == CHROM_OK; ++k) {
          BatchDev bs = b;
          bs.step_begin = (uint32_t)((uint64_t)b.steps * k / kSlices);
          bs.step_end = (uint32_t)((uint64_t)b.steps * (k + 1) / kSlices);
          if (k > 0) bs.y0 = b.final_state;  // every slice leaves its state there
          rc = launch_batch(ctx, bs, sh.g, cs);
          if (rc != CHROM_OK) break;
          launches += sh.g.launches_per_execute;
          cudaEvent_t evs;
          CUDA_TRY(ctx, cudaEventCreateWithFlags(&evs, cudaEventDisableTiming));
          events[si].push_back(evs);
          CUDA_TRY(ctx, cudaEventRecord(evs, cs));
          CUDA_TRY(ctx, cudaStreamWaitEvent(d.copy_stream, evs, 0));
          const size_t slot_a = k == 0 ? 0 : bs.step_begin / b.outlet_stride + 1;  // slot 0 = the initial state
          const size_t slot_b = bs.step_end / b.outlet_stride + 1;
          if (slot_b > slot_a) {
            CUDA_TRY(ctx, cudaMemcpy2DAsync(outlet + (sh.col0 + c0) * n_out + slot_a, n_out * sizeof(double),
                                            b.outlet + slot_a, n_out * sizeof(double), (slot_b - slot_a) * sizeof(double),
                                            cnt, cudaMemcpyDeviceToHost, d.copy_stream));
            bytes += cnt * (slot_b - slot_a) * sizeof(double);
          }
        }
        if (rc != CHROM_OK) break;
        if (final_state) {
          CUDA_TRY(ctx, cudaMemcpyAsync(final_state + (sh.col0 + c0) * sl, b.final_state, cnt * sl * sizeof(double),
                                        cudaMemcpyDeviceToHost, d.copy_stream));
          bytes += cnt * sl * sizeof(double);
        }
        if (status) {
          CUDA_TRY(ctx, cudaMemcpyAsync(status + sh.col0 + c0, b.status, cnt * sizeof(long long), cudaMemcpyDeviceToHost,
                                        d.copy_stream));
          bytes += cnt * sizeof(long long);
        }
        if (summary && b.summary) {  // finished by the last slice (the earlier ones parked their raw sums there)
          CUDA_TRY(ctx, cudaMemcpyAsync(summary + (sh.col0 + c0) * ser * CHROM_SUMMARY_LEN, b.summary,
                                        cnt * ser * CHROM_SUMMARY_LEN * sizeof(double), cudaMemcpyDeviceToHost, d.copy_stream));
          bytes += cnt * ser * CHROM_SUMMARY_LEN * sizeof(double);
        }
        continue;
      }
      rc = launch_batch(ctx, b, sh.g, cs);
      if (rc != CHROM_OK) break;
      launches += sh.g.launches_per_execute;
      cudaEvent_t ev;
      CUDA_TRY(ctx, cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
      events[si].push_back(ev);
      CUDA_TRY(ctx, cudaEventRecord(ev, cs));
      CUDA_TRY(ctx, cudaStreamWaitEvent(d.copy_stream, ev, 0));
      auto copy = [&](void* host, const void* dev, size_t per_col_bytes) -> cudaError_t {
        if (!host || !dev) return cudaSuccess;
        bytes += cnt * per_col_bytes;
        return cudaMemcpyAsync((char*)host + (sh.col0 + c0) * per_col_bytes, dev, cnt * per_col_bytes,
                               cudaMemcpyDeviceToHost, d.copy_stream);
      };
      CUDA_TRY(ctx, copy(outlet, b.outlet, ser * plan->n_out * sizeof(double)));
      CUDA_TRY(ctx, copy(snapshots, b.snapshots, plan->n_snap * sl * sizeof(double)));
      CUDA_TRY(ctx, copy(final_state, b.final_state, sl * sizeof(double)));
      CUDA_TRY(ctx, copy(summary, b.summary, ser * CHROM_SUMMARY_LEN * sizeof(double)));
      CUDA_TRY(ctx, copy(status, b.status, sizeof(long long)));
    }
  }
  float kmax = 0.f;
  for (size_t si = 0; si < plan->shards.size(); ++si) {
    Shard& sh = *plan->shards[si];
    DeviceState& d = ctx->devs[sh.dev];
    cudaSetDevice(d.id);
    // the kernel interval ends when the last chunk of EITHER compute stream has finished
    for (size_t k = events[si].size(); k > 0 && k + 6 > events[si].size(); --k) cudaStreamWaitEvent(d.stream, events[si][k - 1], 0);
    cudaEventRecord(d.ev[3], d.stream);
    cudaError_t e = cudaStreamSynchronize(d.copy_stream);
    if (e == cudaSuccess) e = cudaEventSynchronize(d.ev[3]);
    if (e == cudaSuccess) e = cudaStreamSynchronize(d.stream2);
    if (e == cudaSuccess) cudaEventElapsedTime(&sh.kernel_ms, d.ev[2], d.ev[3]);
    kmax = std::max(kmax, sh.kernel_ms);
    for (cudaEvent_t ev : events[si]) cudaEventDestroy(ev);
    events[si].clear();
    if (e != cudaSuccess && rc == CHROM_OK)
      rc = fail(ctx, CHROM_ERR_CUDA, format("pipelined solve failed: %s", cudaGetErrorString(e)));
  }
  ctx->timing = chrom_cuda_timing{};
  ctx->timing.kernel_ms = kmax;
  ctx->timing.kernel_launches = launches;
  ctx->timing.h2d_ms = plan->h2d_ms;
  ctx->timing.h2d_bytes = plan->h2d_bytes;
  ctx->timing.d2h_ms = now_ms() - t0;  // overlapped with the kernels: wall time of compute + copy-back
  ctx->timing.d2h_bytes = bytes;
  return rc;
}

static int32_t solve_common(chrom_cuda_ctx* ctx, chrom_cuda_plan* plan, double* outlet, double* snapshots,
                            double* final_state, double* summary, int64_t* status, double t0) {
  int32_t rc = solve_pipelined(ctx, plan, outlet, snapshots, final_state, summary, status);
  if (rc == 1) {
    rc = chrom_cuda_plan_execute(plan);
    if (rc == CHROM_OK) rc = chrom_cuda_plan_download(plan, outlet, snapshots, final_state, summary, status);
  }
  chrom_cuda_plan_destroy(plan);
  ctx->timing.total_ms = now_ms() - t0;
  return rc;
}

static uint32_t want_mask(const chrom_run_cfg* cfg, const double* outlet, const double* snapshots,
                          const double* final_state, const double* summary) {
  uint32_t w = CHROM_WANT_STATUS;
  if (outlet && cfg && (cfg->outlet_stride || cfg->outlet_steps)) w |= CHROM_WANT_OUTLET;
  if (snapshots && cfg && (cfg->snapshot_stride || cfg->snapshot_steps)) w |= CHROM_WANT_SNAPSHOTS;
  if (final_state) w |= CHROM_WANT_FINAL;
  if (summary) w |= CHROM_WANT_SUMMARY;
  return w;
}

int32_t chrom_cuda_solve_single_batch(chrom_cuda_ctx* ctx, const chrom_run_cfg* cfg, const chrom_single_col* cols,
                                      uint64_t n_cols, const double* inj_tables, uint32_t n_tables, const double* y0,
                                      double* outlet, double* snapshots, double* final_state, double* summary,
                                      int64_t* status) {
  const double t0 = now_ms();
  chrom_cuda_plan* plan = nullptr;
  int32_t rc = chrom_cuda_plan_single(ctx, cfg, cols, n_cols, inj_tables, n_tables, y0,
                                      want_mask(cfg, outlet, snapshots, final_state, summary), &plan);
  if (rc != CHROM_OK) return rc;
  return solve_common(ctx, plan, outlet, snapshots, final_state, summary, status, t0);
}

int32_t chrom_cuda_solve_multi_batch(chrom_cuda_ctx* ctx, const chrom_run_cfg* cfg, const chrom_multi_col* cols,
                                     uint64_t n_cols, const chrom_species* species, uint64_t n_species_total,
                                     const double* inj_tables, uint32_t n_tables, const double* y0, double* outlet,
                                     double* snapshots, double* final_state, double* summary, int64_t* status) {
  const double t0 = now_ms();
  chrom_cuda_plan* plan = nullptr;
  int32_t rc = chrom_cuda_plan_multi(ctx, cfg, cols, n_cols, species, n_species_total, inj_tables, n_tables, y0,
                                     want_mask(cfg, outlet, snapshots, final_state, summary), &plan);
  if (rc != CHROM_OK) return rc;
  return solve_common(ctx, plan, outlet, snapshots, final_state, summary, status, t0);
}

// ---- measurement ------------------------------------------------------------------------------

int32_t chrom_cuda_measure_fp64_peak(chrom_cuda_ctx* ctx, int32_t device_index, double* tflops, double* ms) {
  if (!ctx || !tflops || !ms || device_index < 0 || device_index >= (int)ctx->devs.size()) return CHROM_ERR_INVALID;
  DeviceState& d = ctx->devs[device_index];
  CUDA_TRY(ctx, cudaSetDevice(d.id));
  CUDA_TRY(ctx, measure_fp64_peak(d.sm_count, d.stream, tflops, ms));
  return CHROM_OK;
}

int32_t chrom_cuda_flush_l2(chrom_cuda_ctx* ctx) {
  if (!ctx) return CHROM_ERR_INVALID;
  const size_t bytes = (size_t)256 << 20;
  for (auto& d : ctx->devs) {
    CUDA_TRY(ctx, cudaSetDevice(d.id));
    if (!d.flush_buf) CUDA_TRY(ctx, cudaMalloc(&d.flush_buf, bytes));
    CUDA_TRY(ctx, cudaMemsetAsync(d.flush_buf, 0x5a, bytes, d.stream));
  }
  for (auto& d : ctx->devs) {
    CUDA_TRY(ctx, cudaSetDevice(d.id));
    CUDA_TRY(ctx, cudaStreamSynchronize(d.stream));
  }
  return CHROM_OK;
}

int32_t chrom_cuda_measure_copy_bw(chrom_cuda_ctx* ctx, int32_t device_index, size_t bytes, double* gbs) {
  if (!ctx || !gbs || device_index < 0 || device_index >= (int)ctx->devs.size()) return CHROM_ERR_INVALID;
  DeviceState& d = ctx->devs[device_index];
  CUDA_TRY(ctx, cudaSetDevice(d.id));
  CUDA_TRY(ctx, measure_copy_bw(bytes, d.stream, gbs));
  return CHROM_OK;
}

int32_t chrom_cuda_probe_rcp(chrom_cuda_ctx* ctx, double lo, double hi, uint32_t n, double* max_rel_err) {
  if (!ctx || !max_rel_err || n == 0) return CHROM_ERR_INVALID;
  DeviceState& d = ctx->devs[0];
  CUDA_TRY(ctx, cudaSetDevice(d.id));
  CUDA_TRY(ctx, probe_rcp(lo, hi, n, d.stream, max_rel_err));
  return CHROM_OK;
}

}  // extern "C"
