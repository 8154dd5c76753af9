// single_resident.cu — register-resident LangmuirSingle time integrator (RK4 / Euler).
//
// Replaces, for a whole batch of columns and ALL time steps in one launch:
//   RK4Solver::solve / EulerSolver::solve   src/solver/methods/rk4.rs:266-332, euler.rs:227-269
//   LangmuirSingle::compute_physics          src/models/langmuir_single.rs:445-516
//   PhysicalState Add / Mul<f64>             src/physics/traits.rs:454-491
//   validate_state                           src/solver/mod.rs:514-553
//
// Layout: a column of nz cells is cut into L contiguous chunks of M cells; lane p of the column
// owns cells [p*M, p*M+M) in registers (y = state at t_n, u = current stage state, acc = running
// weighted slope).  The upwind neighbour of a lane's first cell is the previous lane's last cell:
// one 64-bit shuffle per lane per stage; the column's first lane takes the tabulated inlet value.
// A warp carries cpw = 32/L columns (L need not divide 32; spare lanes idle).  When a column
// needs more than one warp (MW) the warp-boundary value crosses through shared memory with one
// __syncthreads per stage.  HBM is touched for: parameters + y0 (once), the inlet table (3 doubles
// per step, L1/L2 resident), sampled outputs, final state.
//
// In-place stage update: cells are visited in DESCENDING order, so u[j-1] is still the current
// stage's value when cell j reads it, and u[j] can be overwritten with the next stage's state at once.
#include <algorithm>
#include <cstdio>
#include <cstdlib>

#include "arith.cuh"
#include "kernels.h"
#include "single_rhs.cuh"

namespace chrom {
namespace {

const char* arith_name(int a) { return a == CHROM_ARITH_EXACT ? "exact" : (a == CHROM_KARITH_POLY ? "fast-poly" : "fast"); }

constexpr int kBlockThreads = 64;     // single-warp-per-column variant: 2 independent warps per CTA ...
constexpr int kMaxBlockThreads = 128;  // ... or 4 (split launches: one warp per SM sub-partition and CTA)

// resident 64-thread CTAs per SM the register file allows (255 / 168 / 128 registers per thread)
constexpr int min_blocks_for(int M) { return M >= 17 ? 4 : (M >= 11 ? 6 : 8); }
// multi-warp variant: threads per CTA the register file allows for M cells per lane
constexpr int mw_max_threads(int M) { return M <= 8 ? 512 : 256; }

template <int M, int METHOD, int ARITH, bool MW>
__global__ void __launch_bounds__(MW ? mw_max_threads(M) : kMaxBlockThreads, MW ? 1 : min_blocks_for(M) / 2)
    k_single_resident(const BatchDev b, const int L, const int cpw) {
  constexpr int ST = (METHOD == CHROM_RK4) ? 3 : 1;
  using Q = Coef<ARITH>;
  __shared__ double halo[2][32];

  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  int p;
  uint32_t col;
  bool active;
  unsigned gmask;  // lanes of this lane's column (single-warp variant)
  if (MW) {
    col = blockIdx.x;
    p = threadIdx.x;
    active = true;
    gmask = 0xffffffffu;
  } else {
    const uint32_t gw = blockIdx.x * (blockDim.x >> 5) + warp;
    const int sub = lane / L;
    p = lane - sub * L;
    col = gw * cpw + sub;
    active = (sub < cpw) && (col < b.n_cols);
    gmask = (L >= 32) ? 0xffffffffu : (((1u << L) - 1u) << (sub * L));
  }
  const uint32_t colc = active ? col : 0u;
  const uint32_t nz = b.nz;
  const chrom_single_col cp = b.scols[colc];
  Q q;
  q.init(cp);

  const int cell0 = p * M;
  const int nv = max(0, min(M, (int)nz - cell0));  // real cells on this lane
  const int det = (int)detector_of(cp.detector_cell, nz);  // where outlet and summary are sampled (default: nz - 1)
  const int p_out = det / M;
  const int j_out = det - p_out * M;
  const bool out_lane = active && (p == p_out);

  double y[M], u[M], acc[M];
#pragma unroll
  for (int j = 0; j < M; ++j) {
    y[j] = (b.y0 != nullptr && j < nv) ? b.y0[(size_t)colc * nz + cell0 + j] : 0.0;
    u[j] = 0.0;
    acc[j] = 0.0;
  }

  const double dt = b.dt, h2 = b.half_dt, h6 = b.sixth_dt;
  const uint32_t steps = b.steps;
  const double* __restrict__ tab = b.tables + (size_t)cp.inj_table * steps * ST;

  auto outlet_value = [&]() -> double {
    double v = y[M - 1];
    if (j_out != M - 1) {  // (the usual detector - the outlet of a column whose length is a multiple of M - skips this)
#pragma unroll
      for (int j = 0; j < M - 1; ++j) v = (j == j_out) ? y[j] : v;
    }
    return v;
  };
  auto store_profile = [&](double* dst) {  // dst -> this column's [nz] slot
#pragma unroll
    for (int j = 0; j < M; ++j)
      if (j < nv) dst[cell0 + j] = y[j];
  };

  // This launch advances steps [s0, s1); s0 > 0 continues a solve from the carried state in y0 (the last chunk
  // of a one-shot solve is time-sliced so that its outlet series can go back to the host slice by slice; the raw
  // sums of the on-device summary are parked in the summary buffer between slices).
  const uint32_t s0 = b.step_begin, s1 = b.step_end;
  // The running summary lives in shared memory, one slot per column of the CTA, touched by the detector lane only:
  // seven more doubles per lane would not fit the register file of the widest variants.
  __shared__ Summ summ_sh[MW ? 1 : (kMaxBlockThreads / 32) * 32];
  Summ& sm = summ_sh[MW ? 0 : (warp * 32 + lane / L)];
  // step 0 samples (initial condition: rk4.rs:254-255)
  if (s0 == 0) {
    const double v0 = outlet_value();
    if (out_lane) summ_init(sm, v0);
    if (out_lane && b.outlet && sample_has_initial(b.out_steps, b.outlet_stride)) b.outlet[(size_t)col * b.n_out] = v0;
    if (active && b.snapshots && sample_has_initial(b.snap_steps, b.snapshot_stride))
      store_profile(b.snapshots + (size_t)col * b.n_snap * nz);
  } else if (out_lane && b.summary) {
    summ_resume(sm, b.summary + (size_t)col * CHROM_SUMMARY_LEN, outlet_value());
  }
  SampleCursor oc = sample_begin(b.out_steps, b.n_out, b.outlet ? b.outlet_stride : 0u, s0);
  SampleCursor sc_ = sample_begin(b.snap_steps, b.n_snap, b.snapshots ? b.snapshot_stride : 0u, s0);
  if (!b.outlet) oc.next = kNoSample;
  if (!b.snapshots) sc_.next = kNoSample;
  bool done = (s0 > 0) && (b.status != nullptr) && (b.status[colc] != 0);  // validate_state fired in an earlier slice

  double n0 = 0.0, n1 = 0.0, n2 = 0.0;
  n0 = tab[(size_t)s0 * ST];
  if (ST == 3) {
    n1 = tab[(size_t)s0 * ST + 1];
    n2 = tab[(size_t)s0 * ST + 2];
  }

  // upwind value for the first cell of this lane at the current stage
  auto exchange = [&](double last, double inlet, int parity) -> double {
    double up = __shfl_up_sync(0xffffffffu, last, 1);
    if (MW) {
      if (lane == 31) halo[parity][warp] = last;
      __syncthreads();
      if (lane == 0 && warp > 0) up = halo[parity][warp - 1];
    }
    return (p == 0) ? inlet : up;
  };

  for (uint32_t step = s0; step < s1; ++step) {
    const double cin0 = n0, cin1 = n1, cin2 = n2;
    if (step + 1 < steps) {  // prefetch the next step's inlet values
      const double* nx = tab + (size_t)(step + 1) * ST;
      n0 = nx[0];
      if (ST == 3) {
        n1 = nx[1];
        n2 = nx[2];
      }
    }

    unsigned mx = 0u;  // max |high word| of the new state: the non-finite test, on the integer pipe
    if (METHOD == CHROM_RK4) {
      // stage 1: k1 = f(y, t)            acc = k1            u = y + k1*(dt/2)
      double up = exchange(y[M - 1], cin0, 0);
#pragma unroll
      for (int j = M - 1; j >= 0; --j) {
        const double k = q.rhs(y[j], j > 0 ? y[j - 1] : up);
        acc[j] = k;
        u[j] = Q::axpy(y[j], k, h2);
      }
      // stage 2: k2 = f(u, t+dt/2)       acc = acc + k2*2    u = y + k2*(dt/2)
      up = exchange(u[M - 1], cin1, 1);
#pragma unroll
      for (int j = M - 1; j >= 0; --j) {
        const double k = q.rhs(u[j], j > 0 ? u[j - 1] : up);
        acc[j] = Q::axpy(acc[j], k, 2.0);
        u[j] = Q::axpy(y[j], k, h2);
      }
      // stage 3: k3 = f(u, t+dt/2)       acc = acc + k3*2    u = y + k3*dt
      up = exchange(u[M - 1], cin1, 0);
#pragma unroll
      for (int j = M - 1; j >= 0; --j) {
        const double k = q.rhs(u[j], j > 0 ? u[j - 1] : up);
        acc[j] = Q::axpy(acc[j], k, 2.0);
        u[j] = Q::axpy(y[j], k, dt);
      }
      // stage 4: k4 = f(u, t+dt)         y = y + (acc + k4)*(dt/6)
      up = exchange(u[M - 1], cin2, 1);
#pragma unroll
      for (int j = M - 1; j >= 0; --j) {
        const double k = q.rhs(u[j], j > 0 ? u[j - 1] : up);
        y[j] = Q::axpy(y[j], Q::add(acc[j], k), h6);
        mx = max(mx, abs_hi(y[j]));
      }
    } else {
      // forward Euler: y = y + f(y, t)*dt
      const double up = exchange(y[M - 1], cin0, (int)(step & 1u));
#pragma unroll
      for (int j = M - 1; j >= 0; --j) {
        const double k = q.rhs(y[j], j > 0 ? y[j - 1] : up);
        y[j] = Q::axpy(y[j], k, dt);
        mx = max(mx, abs_hi(y[j]));
      }
    }

    // ---- samples ------------------------------------------------------------------------
    const uint32_t done_steps = step + 1u;
    const bool want_out = (done_steps == oc.next);
    if (b.summary != nullptr || want_out) {
      const double v = outlet_value();
      if (want_out) {
        if (out_lane) b.outlet[(size_t)col * b.n_out + oc.k] = v;
        sample_advance(oc, b.out_steps, b.outlet_stride);
      }
      if (b.summary != nullptr && out_lane) summ_step(sm, v, step, dt);
    }
    if (done_steps == sc_.next) {
      if (active) store_profile(b.snapshots + ((size_t)col * b.n_snap + sc_.k) * nz);
      sample_advance(sc_, b.snap_steps, b.snapshot_stride);
    }

    // ---- validate_state (solver/mod.rs:514-553) ---------------------------------------------------
    // Fast test over every cell of the lane, padding included (mx, accumulated inside the last stage);
    // only when it fires is the state re-examined cell by cell, real cells only, NaN before Inf.
    const bool maybe = hi_nonfinite(mx);
    if (MW) {
      if (__syncthreads_or(maybe)) {
        bool bad = false, has_nan = false;
#pragma unroll
        for (int j = 0; j < M; ++j) {
          bad |= (j < nv) && hi_nonfinite(abs_hi(y[j]));
          has_nan |= (j < nv) && (y[j] != y[j]);
        }
        const int any_bad = __syncthreads_or(bad);
        const int any_nan = __syncthreads_or(has_nan);
        if (any_bad && !done) {
          if (threadIdx.x == 0 && b.status)
            b.status[col] = any_nan ? (long long)(step + 1) : -(long long)(step + 1);
          done = true;
        }
      }
    } else {
      if (__ballot_sync(0xffffffffu, maybe) != 0u) {
        bool bad = false, has_nan = false;
#pragma unroll
        for (int j = 0; j < M; ++j) {
          bad |= (j < nv) && hi_nonfinite(abs_hi(y[j]));
          has_nan |= (j < nv) && (y[j] != y[j]);
        }
        const unsigned bad_ballot = __ballot_sync(0xffffffffu, bad && active);
        const unsigned nan_ballot = __ballot_sync(0xffffffffu, has_nan && active);
        if ((bad_ballot & gmask) != 0u && !done) {
          if (p == 0 && active && b.status)
            b.status[col] = (nan_ballot & gmask) ? (long long)(step + 1) : -(long long)(step + 1);
          done = true;
        }
      }
    }
  }

  if (active && b.final_state) store_profile(b.final_state + (size_t)col * nz);
  if (out_lane && b.summary) {
    double* o = b.summary + (size_t)col * CHROM_SUMMARY_LEN;
    if (s1 >= steps) summ_store(o, sm);
    else summ_park(o, sm);  // a later slice of the same solve resumes from the raw sums
  }
}

// ---- instantiation tables -----------------------------------------------------------------
typedef void (*kern_t)(const BatchDev, const int, const int);

template <int M, bool MW>
kern_t pick(int method, int arith) {
  if (method == CHROM_RK4) {
    if (arith == CHROM_ARITH_EXACT) return k_single_resident<M, CHROM_RK4, 1, MW>;
    return arith == CHROM_KARITH_POLY ? k_single_resident<M, CHROM_RK4, 2, MW> : k_single_resident<M, CHROM_RK4, 0, MW>;
  }
  if (arith == CHROM_ARITH_EXACT) return k_single_resident<M, CHROM_EULER, 1, MW>;
  return arith == CHROM_KARITH_POLY ? k_single_resident<M, CHROM_EULER, 2, MW> : k_single_resident<M, CHROM_EULER, 0, MW>;
}

const int kM[] = {2, 4, 5, 6, 7, 8, 10, 12, 13, 14, 16, 17, 20, 25, 28, 32};
const int kMmw[] = {8, 12, 16, 20, 24, 28, 32};

kern_t lookup(int M, bool mw, int method, int arith) {
  if (mw) {
    switch (M) {
      case 8: return pick<8, true>(method, arith);
      case 12: return pick<12, true>(method, arith);
      case 16: return pick<16, true>(method, arith);
      case 20: return pick<20, true>(method, arith);
      case 24: return pick<24, true>(method, arith);
      case 28: return pick<28, true>(method, arith);
      case 32: return pick<32, true>(method, arith);
    }
    return nullptr;
  }
  switch (M) {
    case 2: return pick<2, false>(method, arith);
    case 4: return pick<4, false>(method, arith);
    case 5: return pick<5, false>(method, arith);
    case 6: return pick<6, false>(method, arith);
    case 7: return pick<7, false>(method, arith);
    case 8: return pick<8, false>(method, arith);
    case 10: return pick<10, false>(method, arith);
    case 12: return pick<12, false>(method, arith);
    case 13: return pick<13, false>(method, arith);
    case 14: return pick<14, false>(method, arith);
    case 16: return pick<16, false>(method, arith);
    case 17: return pick<17, false>(method, arith);
    case 20: return pick<20, false>(method, arith);
    case 25: return pick<25, false>(method, arith);
    case 28: return pick<28, false>(method, arith);
    case 32: return pick<32, false>(method, arith);
  }
  return nullptr;
}

}  // namespace

// Largest nz the resident path holds (256 threads x 32 cells, 512 x 16, 1024 x 8).
static const uint32_t kMaxResidentNz = 8192;

bool single_resident_choose(const BatchDev& b, int sm_count, LaunchGeom* g, std::string* why) {
  const uint32_t nz = b.nz;
  if (nz > kMaxResidentNz) {
    if (why) *why = "nz exceeds the register-resident capacity (8192 cells per CTA)";
    return false;
  }
  char buf[360];
  const char* force = getenv("CHROM_FORCE_M");  // experiments only
  const uint64_t fill_warps = (uint64_t)sm_count * 8u;

  // ---- candidate 1: a single warp per column (nz <= 1024) ------------------------------------------------
  // Cost model, in "cells a lane steps through per time step" on the busiest SM sub-partition (the kernel is bound
  // by the FP64 pipe of its sub-partition; measured on B200, profiles/r2_single_shapes_8192.jsonl):
  //   warps = ceil(n_cols / cpw);  a sub-partition holds wps(M) resident warps (register file), the machine
  //   slots = 4 SMs sub-partitions;  full waves cost wps * M / eff_m each, the last partial wave ceil(rest / slots) *
  //   M / eff_m.  eff_m = measured fraction of the FP64 peak at a perfect fit: flat 0.69-0.72 for 6 <= M <= 25, lower
  //   for very short chunks (per-lane overhead) and for M >= 28 (255 registers).
  // For a batch of many waves this is the old "lane utilisation x eff_m" rule; for a batch that cannot fill the machine
  // it sees the quantisation (8192 columns of nz = 100: M = 25 puts 2 warps on 432 of 592 sub-partitions and 1 on
  // the rest, 24.7 ms instead of the 20.6 ms of a perfect spread), and for a handful of columns the shortest chunk.
  auto eff_of = [](int M) { return M < 6 ? 0.645 : (M == 12 || M == 25) ? 0.72 : M <= 25 ? 0.70 : M == 28 ? 0.675 : 0.65; };
  auto wps_of = [](int M) { return min_blocks_for(M) * (kBlockThreads / 32) / 4; };
  const uint64_t slots = (uint64_t)sm_count * 4u;
  double sw_cost = 1e300;
  int swM = 0, swL = 0, swCpw = 0;
  bool latency_bound = false;
  if (nz <= 32u * 32u) {
    for (int M : kM) {
      if (force && atoi(force) != M) continue;
      const int L = (int)((nz + M - 1) / M);
      if (L > 32) continue;
      const int cpw = 32 / L;
      const uint64_t warps = ((uint64_t)b.n_cols + cpw - 1) / cpw;
      const uint64_t cap = slots * (uint64_t)wps_of(M);
      const uint64_t full = warps / cap, rest = warps % cap;
      const double cost = ((double)full * wps_of(M) + (double)((rest + slots - 1) / slots)) * M / eff_of(M);
      if (cost < sw_cost) {
        sw_cost = cost;
        swM = M;
        swL = L;
        swCpw = cpw;
        latency_bound = warps < fill_warps / 2;
      }
    }
  }
  // A sub-wave batch: two launches with different chunk sizes, side by side (one warp of each per sub-partition).
  // 8192 columns of nz = 100: 4736 columns at M = 25 (8 per warp) + 3456 at M = 20 (6 per warp) = 45 cells per lane
  // instead of 50; measured 22.8 ms against 24.7 ms (scripts/split_experiment.py).
  int spA = 0, spB = 0;
  double sp_cost = 1e300;
  if (swM && !force && !getenv("CHROM_NO_SPLIT") && nz <= 32u * 32u) {
    const uint64_t warps1 = ((uint64_t)b.n_cols + swCpw - 1) / swCpw;
    if (warps1 > slots && warps1 <= 2 * slots) {  // the single launch puts 2 warps on some sub-partitions, 1 on others
      for (int MA : kM)
        for (int MB : kM) {
          if (MB > MA || MA < 11 || MB < 11) continue;  // 255 / 168 registers: one CTA of each per SM
          const int LA = (int)((nz + MA - 1) / MA), LB = (int)((nz + MB - 1) / MB);
          if (LA > 32 || LB > 32) continue;
          const uint64_t capA = slots * (uint64_t)(32 / LA), capB = slots * (uint64_t)(32 / LB);
          if (capA + capB < b.n_cols || capA >= b.n_cols) continue;
          const double cost = MA / eff_of(MA) + MB / eff_of(MB);
          if (cost < sp_cost) {
            sp_cost = cost;
            spA = MA;
            spB = MB;
          }
        }
      if (!(sp_cost < 0.96 * sw_cost)) spA = spB = 0;
    }
  }
  const double sw_score = swM ? 1.0 / sw_cost : -1.0;

  // ---- candidate 2: several warps per column, one CTA per column (nz > 512) ------------------------------
  // score = (fraction of lane-slots doing real cells) x (measured efficiency of the chunk size at a perfect fit)
  // x (balance of the resident warps over the 4 SM sub-partitions: 6 warps per SM run at 6/8).
  double mw_score = -1.0;
  int mwM = 0, mwW = 0;
  if (nz > 512u && !(latency_bound && swM)) {
    for (int M : kMmw) {
      if (force && atoi(force) != M) continue;
      const int W = (int)((nz + 32u * M - 1) / (32u * M));
      if (W * 32 > mw_max_threads(M)) continue;
      const double eff_m = M <= 8 ? 0.55 : M <= 12 ? 0.647 : M <= 16 ? 0.691 : M <= 20 ? 0.695 : M <= 24 ? 0.711 : 0.65;
      int ctas = 1;
      if (kern_t k = lookup(M, true, b.method, b.arith)) {
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctas, (const void*)k, W * 32, 0) != cudaSuccess || ctas < 1) {
          cudaGetLastError();
          ctas = 1;
        }
      }
      const int wt = W * ctas;
      const double balance = (double)wt / (4.0 * ((wt + 3) / 4));
      const double score = (double)nz / ((double)W * 32.0 * M) * eff_m * balance;  // fraction of the FP64 peak
      if (score >= mw_score) {
        mw_score = score;
        mwM = M;
        mwW = W;
      }
    }
  }

  // the single-warp candidate as a fraction of the FP64 peak too: useful cell-steps per lane-slot over its cost
  const double sw_frac = swM ? ((double)b.n_cols * nz / ((double)slots * 32.0)) / sw_cost : -1.0;
  (void)sw_score;
  if (swM && (sw_frac >= mw_score || !mwM || b.n_cols < (uint32_t)sm_count)) {
    const uint64_t warps = ((uint64_t)b.n_cols + swCpw - 1) / swCpw;
    g->kind = 1;
    g->M = swM;
    g->L = swL;
    g->cpw = swCpw;
    g->warps_per_col = 1;
    g->block = dim3(kBlockThreads);
    g->grid = dim3((unsigned)((warps + (kBlockThreads / 32) - 1) / (kBlockThreads / 32)));
    g->smem = 0;
    g->cols_per_wave = (uint32_t)(sm_count * min_blocks_for(swM) * (kBlockThreads / 32) * swCpw);
    snprintf(buf, sizeof buf, "single_resident<M=%d,%s,%s> lanes/col=%d cols/warp=%d grid=%u block=%d",
             swM, b.method == CHROM_RK4 ? "RK4" : "Euler", arith_name(b.arith),
             swL, swCpw, g->grid.x, kBlockThreads);
    g->name = buf;
    if (spA) {  // two launches side by side: 128-thread CTAs, one warp per sub-partition each
      g->M = spA;
      g->L = (int)((nz + spA - 1) / spA);
      g->cpw = 32 / g->L;
      g->M2 = spB;
      g->L2 = (int)((nz + spB - 1) / spB);
      g->cpw2 = 32 / g->L2;
      g->cols_first = (uint32_t)std::min<uint64_t>(b.n_cols, slots * (uint64_t)g->cpw);
      g->block = dim3(kMaxBlockThreads);
      g->cols_per_wave = b.n_cols;  // one wave by construction: never chunked
      snprintf(buf, sizeof buf,
               "single_resident<M=%d+%d,%s,%s> split launch: %u columns at %d cells/lane (%d cols/warp) + %u at %d (%d cols/warp), "
               "one warp of each per SM sub-partition, block=%d",
               spA, spB, b.method == CHROM_RK4 ? "RK4" : "Euler", arith_name(b.arith), g->cols_first, spA, g->cpw,
               b.n_cols - g->cols_first, spB, g->cpw2, kMaxBlockThreads);
      g->name = buf;
    }
    return true;
  }
  if (mwM == 0) return false;
  g->kind = 2;
  g->M = mwM;
  g->L = mwW * 32;
  g->cpw = 1;
  g->warps_per_col = mwW;
  g->block = dim3(mwW * 32);
  g->grid = dim3(b.n_cols);
  g->smem = 0;
  g->cols_per_wave = (uint32_t)(sm_count * std::max(1, 512 / (mwW * 32)));
  snprintf(buf, sizeof buf, "single_resident_mw<M=%d,%s,%s> warps/col=%d grid=%u block=%d", mwM,
           b.method == CHROM_RK4 ? "RK4" : "Euler", arith_name(b.arith), mwW,
           g->grid.x, mwW * 32);
  g->name = buf;
  return true;
}

namespace {
// columns [c0, c0 + cnt) of a single-species batch: every per-column pointer offset
BatchDev offset_cols(BatchDev b, uint64_t c0, uint64_t cnt) {
  const size_t nz = b.nz;
  if (b.scols) b.scols += c0;
  if (b.y0) b.y0 += c0 * nz;
  if (b.outlet) b.outlet += c0 * b.n_out;
  if (b.snapshots) b.snapshots += c0 * b.n_snap * nz;
  if (b.final_state) b.final_state += c0 * nz;
  if (b.summary) b.summary += c0 * CHROM_SUMMARY_LEN;
  if (b.status) b.status += c0;
  b.n_cols = (uint32_t)cnt;
  return b;
}
cudaStream_t side_stream() {  // one helper stream per device, created on first use
  static cudaStream_t streams[64] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64) return nullptr;
  if (!streams[dev] && cudaStreamCreateWithFlags(&streams[dev], cudaStreamNonBlocking) != cudaSuccess) {
    cudaGetLastError();
    streams[dev] = nullptr;
  }
  return streams[dev];
}
}  // namespace

cudaError_t single_resident_launch(const BatchDev& b, const LaunchGeom& g, cudaStream_t s) {
  kern_t k = lookup(g.M, g.kind == 2, b.method, b.arith);
  if (!k) return cudaErrorInvalidValue;
  if (g.kind == 1 && g.M2 > 0 && b.n_cols > g.cols_first) {
    // split launch (see single_resident_choose): the first cols_first columns at M cells per lane on `s`, the rest
    // at M2 on a side stream that forks from and joins back into `s`
    kern_t k2 = lookup(g.M2, false, b.method, b.arith);
    cudaStream_t s2 = side_stream();
    if (!k2 || !s2) return cudaErrorInvalidValue;
    cudaEvent_t fork = nullptr, join = nullptr;
    cudaError_t e = cudaEventCreateWithFlags(&fork, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&join, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventRecord(fork, s);
    if (e == cudaSuccess) e = cudaStreamWaitEvent(s2, fork, 0);
    if (e == cudaSuccess) {
      const int wpb = kMaxBlockThreads / 32;
      const BatchDev ba = offset_cols(b, 0, g.cols_first), bb = offset_cols(b, g.cols_first, b.n_cols - g.cols_first);
      const uint64_t wa = ((uint64_t)ba.n_cols + g.cpw - 1) / g.cpw, wb = ((uint64_t)bb.n_cols + g.cpw2 - 1) / g.cpw2;
      k<<<dim3((unsigned)((wa + wpb - 1) / wpb)), dim3(kMaxBlockThreads), 0, s>>>(ba, g.L, g.cpw);
      k2<<<dim3((unsigned)((wb + wpb - 1) / wpb)), dim3(kMaxBlockThreads), 0, s2>>>(bb, g.L2, g.cpw2);
      e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaEventRecord(join, s2);
    if (e == cudaSuccess) e = cudaStreamWaitEvent(s, join, 0);
    if (fork) cudaEventDestroy(fork);
    if (join) cudaEventDestroy(join);
    return e;
  }
  // grid from b.n_cols (a chunk of the planned batch may be launched on its own)
  dim3 grid = g.grid;
  if (g.kind == 1) {
    const uint64_t warps = ((uint64_t)b.n_cols + g.cpw - 1) / g.cpw;
    grid = dim3((unsigned)((warps + (kBlockThreads / 32) - 1) / (kBlockThreads / 32)));
  } else {
    grid = dim3(b.n_cols);
  }
  k<<<grid, g.kind == 1 ? dim3(kBlockThreads) : g.block, g.smem, s>>>(b, g.L, g.cpw);
  return cudaGetLastError();
}

}  // namespace chrom
