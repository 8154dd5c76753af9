// multi_reg2.cuh — register-resident LangmuirMulti time integrator, second generation (n <= 8, FAST arithmetic).
//
// Same reference code as multi_reg.cuh (rk4.rs:266-332, euler.rs:227-269, langmuir_multi.rs:627-686, 717-899,
// solver/mod.rs:514-553) and the same per-cell mathematics (closed-form partial-pivoting elimination of
// diag(alpha) - p q^T, dense row-exchanging LU for a cell that needs an exchange).  What changed is everything
// AROUND the arithmetic, because ncu showed k_multi_reg bound by the shared-memory pipe and the CTA barrier, not by
// FP64 (profiles/r1_ncu_full_multi_reg_c5.txt: FP64 pipe 40 %, l1tex 67 %, ~80 shared-memory wavefronts per warp and
// cell-stage against 270 FP64-pipe cycles spread over 4 sub-partitions):
//   * a thread owns M = 2 contiguous cells with y, the stage state u AND the running slope acc in registers
//     (no private shared-memory slots): 256 threads for a 512-cell column, 8 warps per SM at <= 255 registers;
//   * the species constants {K, feNK, a0} are read once per column with uniform addresses: ptxas keeps them in
//     UNIFORM registers (DFMA R, R, UR, R), so they cost neither shared-memory loads nor vector registers;
//   * the upstream neighbour of a thread's first cell comes by warp shuffle; only the value that crosses a WARP
//     boundary goes through shared memory, into a ring of 8 stage slots guarded by a per-warp release/acquire
//     counter - a warp waits for its upstream warp only, never for the whole CTA.  Cells run in descending order,
//     so the value the downstream warp needs (the LAST cell) is ready first and is posted at once, and the value a
//     warp needs from upstream (for its FIRST cell) is awaited last: a full stage of slack between the two;
//   * ONE __syncthreads per time step (it bounds how far a warp may run ahead, which makes the 8-slot ring safe, and
//     publishes the step's inlet row) instead of one per stage plus a vote;
//   * the hot path is straight-line code: the sufficient no-row-exchange bound only raises a flag, the per-step
//     pivoting test and the dense fallback sit behind one warp vote per cell, and the fallback takes its operands
//     through a shared-memory scratch block (a by-value vector across a call makes ptxas keep the hot result array in
//     local memory as well: 16 local-memory accesses per stage in the first version of this kernel);
//   * validate_state: every thread remembers the first step at which one of its own cells turned non-finite
//     (a non-finite cell never becomes finite again); the column's status is the minimum of those keys.
// Measured on B200 (C5, 4096 columns x nz 512 x n 8 x 4096 RK4 steps): 645 ms -> 471 ms.  Alternatives measured and
// dropped (same workload): a rolled loop over the four stages 586 ms, one cell per thread with 16 warps 550 ms, the
// two cells solved side by side 564 ms, no per-step barrier (warps skewed behind a consumed-counter) 599 ms, the n
// reciprocals started from D instead of 1/D 501 ms, the constants through shared memory 603 ms.
#pragma once
#include "multi_reg.cuh"

namespace chrom {
namespace {

__device__ __forceinline__ unsigned ld_acquire_cta(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.cta.shared.u32 %0, [%1];" : "=r"(v) : "r"((unsigned)__cvta_generic_to_shared(p)) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_cta(unsigned* p, unsigned v) {
  asm volatile("st.release.cta.shared.u32 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(p)), "r"(v) : "memory");
}

// doubles a thread keeps live: y, u, acc (3 M N), z and v of the cell being solved (2 N), the upstream cell (N)
// ->  registers ~ 2x + overhead; thread limit in steps of 128 (see reg_max_threads).
constexpr int reg2_budget(int N, int M) { return 2 * (3 * M * N + 3 * N) + 30; }
constexpr int reg2_max_threads(int N, int M) {
  const int w = 512 / reg2_budget(N, M);
  return w < 1 ? 128 : (w > 8 ? 1024 : 128 * w);
}

template <int N>
struct Consts {
  double K[N], feNK[N], a0[N];
};

// Dense fallback for a cell whose elimination needs a row exchange: LU with partial pivoting of I + Fe*M on a
// shared-memory scratch block, by ONE thread, with run-time loops.  It is rare (strongly overloaded cells) and it
// must not cost the hot path anything: no by-value vectors (those travel through the stack and ptxas then keeps
// the hot result array in local memory too), a handful of registers, so the call needs no spills around it.
// w: [c (n) | up (n) | A (n x n, row-major) | b (n)]; the solution of A x = c - up overwrites w[0 .. n).
// A zero pivot leaves x = c - up (the reference's identity fallback, langmuir_multi.rs:679-686).
__device__ __noinline__ void dense_fallback_smem(const SpeciesConst* sc, double* w, int n) {
  double* c = w;
  double* up = w + n;
  double* A = w + 2 * n;
  double* rhs = A + n * n;
  double S = 0.0;
  for (int s = 0; s < n; ++s) S = fma(sc[s].K, c[s], S);
  const double D = 1.0 + S;
  const double rD2 = 1.0 / (D * D);
  for (int i = 0; i < n; ++i) {
    const double p = sc[i].feNK * c[i] * rD2;
    for (int j = 0; j < n; ++j) A[i * n + j] = -p * sc[j].K;
    A[i * n + i] += fma(sc[i].feNK, D * rD2, sc[i].a0);
    rhs[i] = c[i] - up[i];
  }
  bool singular = false;
  for (int k = 0; k < n && !singular; ++k) {
    int piv = k;
    double best = fabs(A[k * n + k]);
    for (int r = k + 1; r < n; ++r) {
      const double v = fabs(A[r * n + k]);
      if (v > best) {
        best = v;
        piv = r;
      }
    }
    if (best == 0.0) {
      singular = true;
      break;
    }
    if (piv != k) {
      for (int j = 0; j < n; ++j) {
        const double t = A[k * n + j];
        A[k * n + j] = A[piv * n + j];
        A[piv * n + j] = t;
      }
      const double t = rhs[k];
      rhs[k] = rhs[piv];
      rhs[piv] = t;
    }
    const double rp = 1.0 / A[k * n + k];
    for (int r = k + 1; r < n; ++r) {
      const double m = A[r * n + k] * rp;
      for (int j = k + 1; j < n; ++j) A[r * n + j] = fma(-m, A[k * n + j], A[r * n + j]);
      rhs[r] = fma(-m, rhs[k], rhs[r]);
    }
  }
  if (singular) {
    for (int i = 0; i < n; ++i) w[i] = c[i] - up[i];
    return;
  }
  for (int i = n - 1; i >= 0; --i) {
    double t = rhs[i];
    for (int j = i + 1; j < n; ++j) t = fma(-A[i * n + j], rhs[j], t);
    rhs[i] = t / A[i * n + i];
  }
  for (int i = 0; i < n; ++i) w[i] = rhs[i];
}

// One cell: x solves (diag(alpha) - p q^T) x = c - up (the factor -ue/dz lives in the stage coefficients).
// Returns true when the sufficient no-row-exchange bound fails (the caller then runs pivot_test on the cell).
// The three dot products (S = q.c, q.v, q.z) are accumulated in kSums interleaved partial sums.  Two measured best on
// C5 (465 ms; four partial sums 478 ms: the shorter chains do not pay for the additions that join them).
#ifndef CHROM_REG2_SUMS
#define CHROM_REG2_SUMS 2
#endif
template <int N>
__device__ __forceinline__ bool solve_cell(const Consts<N>& kc, double kmax2, double a0min, const double (&c)[N],
                                           const double (&up)[N], double (&x)[N]) {
  constexpr int kSums = (CHROM_REG2_SUMS < N) ? CHROM_REG2_SUMS : (N > 1 ? N : 1);
  auto join = [](const double (&p)[kSums]) {
    double t = p[0];
#pragma unroll
    for (int k = 1; k < kSums; ++k) t += p[k];
    return t;
  };
  double v[N];
  double S[kSums] = {};
#pragma unroll
  for (int s = 0; s < N; ++s) S[s % kSums] = fma(kc.K[s], c[s], S[s % kSums]);
  const double rD = fast_rcp(1.0 + join(S));
  const double rD2 = rD * rD;
  unsigned pmax_hi = 0u;  // max over the high words of feNK_s c_s (monotone for non-negative doubles; a negative or
                          // non-finite entry reads as huge): one integer max per species
  double qv[kSums] = {}, qz[kSums] = {};
#pragma unroll
  for (int i = 0; i < N; ++i) {
    const double tc = kc.feNK[i] * c[i];                        // p_i D^2
    const double ra = fast_rcp(fma(kc.feNK[i], rD, kc.a0[i]));  // 1 / alpha_i
    pmax_hi = max(pmax_hi, (unsigned)__double2hiint(tc));
    x[i] = (c[i] - up[i]) * ra;  // z_i
    v[i] = tc * ra;              // v_i D^2
    qv[i % kSums] = fma(kc.K[i], v[i], qv[i % kSums]);
    qz[i % kSums] = fma(kc.K[i], x[i], qz[i % kSums]);
  }
  const double hn = fma(-rD2, join(qv), 1.0);  // 1 - q.v
  const double g = (join(qz) * rD2) * fast_rcp(hn);
#pragma unroll
  for (int i = 0; i < N; ++i) x[i] = fma(v[i], g, x[i]);
  // Sufficient for "no row exchange anywhere": 2 max q max|p| < min alpha (1 - q.v), with alpha_i >= a0_i when every
  // c >= 0.
  const bool nonneg = pmax_hi < 0x7ff00000u;
  const double pmax2 = __hiloint2double((int)(pmax_hi + 1u), 0) * rD2;
  return !(nonneg && kmax2 * pmax2 < a0min * hn);
}

// The per-step partial-pivoting test of multi_kernels.cuh::scan_cell for a cell whose sufficient bound failed,
// recomputed from the cell (rare; inline, so it is evaluated by all flagged lanes of a warp at once), walked from
// the LAST species down so that it needs two running scalars only: h_i = 1 - sum_{j<i} q_j v_j = h_n + sum_{j>=i}
// q_j v_j, below_i = max_{r>i} |p_r|; step i exchanges rows when q_i below_i > |alpha_i h_i - p_i q_i|.
template <int N>
__device__ __forceinline__ bool pivot_test(const SpeciesConst* sc, const double (&c)[N]) {
  double S = 0.0;
#pragma unroll
  for (int s = 0; s < N; ++s) S = fma(sc[s].K, c[s], S);
  const double rD = fast_rcp(1.0 + S);
  const double rD2 = rD * rD;
  double qv = 0.0;
#pragma unroll
  for (int i = 0; i < N; ++i) {
    const double alpha = fma(sc[i].feNK, rD, sc[i].a0);
    qv = fma(sc[i].K, ((sc[i].feNK * c[i]) * rD2) * fast_rcp(alpha), qv);
  }
  const double hn = 1.0 - qv;
  double suffix = 0.0, below = 0.0;
  bool redo = false;
#pragma unroll
  for (int i = N - 1; i >= 0; --i) {
    const double Ki = sc[i].K;
    const double alpha = fma(sc[i].feNK, rD, sc[i].a0);
    const double p = (sc[i].feNK * c[i]) * rD2;
    suffix = fma(Ki, p * fast_rcp(alpha), suffix);
    const double h = hn + suffix;
    const double dpiv = fabs(fma(alpha, h, -(p * Ki)));
    redo |= !(h > 0.0) || (Ki * below > dpiv) || (dpiv == 0.0);
    below = fmax(below, fabs(p));
  }
  return redo;
}

template <int N, int M, int METHOD>
__global__ void __launch_bounds__(reg2_max_threads(N, M), 1) k_multi_reg2(const BatchDev b) {
  constexpr int ST = (METHOD == CHROM_RK4) ? 3 : 1;
  constexpr int XS = N | 1;  // seam rows: odd stride
  __shared__ SpeciesConst sc[N];
  __shared__ double inlet[2][3][N];  // [step parity][stage row][species]
  __shared__ Summ summ[N];           // running summary of the detector signal (detector owner only)
  __shared__ const double* tabs[N];
  __shared__ double colc[2];              // 2 max_i K_i, min_i (1 + fe lambda_i) of the current column
  __shared__ unsigned flags[32];          // per warp: stage inputs posted so far in this column
  __shared__ unsigned long long badkey;   // min over threads of (first bad step << 1) | (NaN ? 0 : 1); ~0 = clean
  extern __shared__ double seam[];        // [8 slots][W warps][XS]: the last cell of each warp, per stage input;
                                          // then cold[W][3 N + N N]: the dense-fallback scratch block of each warp

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int W = blockDim.x >> 5;
  double* const cold = seam + (size_t)8 * W * XS;
  const uint32_t nz = b.nz;
  const int cell0 = tid * M;
  const int nv = max(0, min(M, (int)nz - cell0));  // real cells of this thread
  const double dt = b.dt, h2 = b.half_dt, h6 = b.sixth_dt;
  const uint32_t steps = b.steps;

  for (uint32_t col = blockIdx.x; col < b.n_cols; col += gridDim.x) {  // persistent CTAs over columns
    const chrom_multi_col cp = b.mcols[col];
    const double fe = cp.fe;
    const double cg = -cp.ue / cp.dz;
    const double cdt = cg * dt, ch2 = cg * h2, ch6 = cg * h6;  // x is solved without the factor -ue/dz
    const double cfin = (METHOD == CHROM_RK4) ? ch6 : cdt;      // coefficient of the closing update
    const int det = (int)detector_of(cp.detector_cell, nz);     // where outlet and summary are sampled
    const int t_out = det / M, j_out = det - t_out * M;
    const bool out_thread = (tid == t_out);
    __syncthreads();  // the previous column's readers of shared memory are done
    if (tid < N) {
      const chrom_species sp = b.species[cp.species_off + tid];
      SpeciesConst q;
      q.lambda = sp.lambda;
      q.K = sp.langmuir_k;
      q.nbar = cp.stationary_fraction * (double)sp.port_number;
      q.a0 = 1.0 + fe * sp.lambda;
      q.feNK = fe * (q.nbar * sp.langmuir_k);
      q.feNK2 = q.feNK;
      q.a0rK = q.a0 / sp.langmuir_k;
      q.feN = fe * q.nbar;
      q.cgrK = cg / sp.langmuir_k;
      q.pad_ = 0.0;
      sc[tid] = q;
      tabs[tid] = b.tables + (size_t)sp.inj_table * steps * ST;
    }
    if (tid < 32) flags[tid] = 0u;
    if (tid == 0) badkey = ~0ull;
    double y[M][N], u[M][N], acc[M][N];
#pragma unroll
    for (int j = 0; j < M; ++j)
#pragma unroll
      for (int s = 0; s < N; ++s) {
        y[j][s] = (b.y0 != nullptr && j < nv) ? b.y0[((size_t)col * N + s) * nz + cell0 + j] : 0.0;
        u[j][s] = y[j][s];
        acc[j][s] = 0.0;
      }
    auto out_value = [&](int s) -> double {  // species s of the detector cell (meaningful on out_thread)
      double v = y[M - 1][s];
#pragma unroll
      for (int j = 0; j < M - 1; ++j) v = (j == j_out) ? y[j][s] : v;
      return v;
    };
    auto store_profile = [&](double* dst) {  // dst -> this column's [N][nz] slot
#pragma unroll
      for (int j = 0; j < M; ++j)
        if (j < nv)
#pragma unroll
          for (int s = 0; s < N; ++s) dst[(size_t)s * nz + cell0 + j] = y[j][s];
    };
    __syncthreads();
    Consts<N> kc;  // uniform addresses: these land in uniform registers
#pragma unroll
    for (int s = 0; s < N; ++s) {
      kc.K[s] = sc[s].K;
      kc.feNK[s] = sc[s].feNK;
      kc.a0[s] = sc[s].a0;
    }
    if (tid == 0) {
      double km = 0.0, am = 1e300;
#pragma unroll
      for (int s = 0; s < N; ++s) {
        km = fmax(km, sc[s].K);
        am = fmin(am, sc[s].a0);
      }
      colc[0] = 2.0 * km;
      colc[1] = am;
    }
    if (out_thread) {
#pragma unroll
      for (int s = 0; s < N; ++s) {
        const double v0 = out_value(s);
        if (b.outlet && sample_has_initial(b.out_steps, b.outlet_stride)) b.outlet[((size_t)col * N + s) * b.n_out] = v0;
        summ_init(summ[s], v0);
      }
    }
    if (b.snapshots && sample_has_initial(b.snap_steps, b.snapshot_stride))
      store_profile(b.snapshots + (size_t)col * b.n_snap * N * nz);
    SampleCursor oc = sample_begin(b.out_steps, b.n_out, b.outlet ? b.outlet_stride : 0u, 0u);
    SampleCursor sc_ = sample_begin(b.snap_steps, b.n_snap, b.snapshots ? b.snapshot_stride : 0u, 0u);
    if (!b.outlet) oc.next = kNoSample;
    if (!b.snapshots) sc_.next = kNoSample;
    unsigned posted = 0u;  // stage inputs this warp has posted (uniform over the CTA: every warp posts every one)
    bool recorded = false;

    // The last cell of a stage's OUTPUT (= the next stage's input) goes to ring slot `slot` (step parity, stage) for
    // the downstream warp: lane 31 stores, then publishes with a release store of the new count.  The per-step
    // barrier keeps the ring safe: a slot is written again two steps later.
    auto post = [&](const double (&last)[N], int slot) {
      ++posted;
      if (lane == 31 && warp + 1 < W) {
        double* dst = seam + ((size_t)slot * W + warp) * XS;
#pragma unroll
        for (int s = 0; s < N; ++s) dst[s] = last[s];
        st_release_cta(&flags[warp], posted);
      }
    };

    // One stage over this thread's cells.  Invariant at the top of a step: u == y, acc == 0.  Every stage solves
    // k = f(u), accumulates acc += w k (w = 1, 2, 2, 1: rk4.rs:311) and forms the next stage state u = y + cu k
    // (cu = dt/2, dt/2, dt); the last one forms y += acc dt/6 (Euler: y += k dt) and restores the invariant.
    // (acc = k1 is 0 + 1 k1 exactly, so this is the arithmetic of rk4.rs:284-311 term by term.)
    auto stage = [&](int irow, int spar, int slot, int slot_next, double wk, double cuk, bool last) {
      double up[N];
#pragma unroll
      for (int s = 0; s < N; ++s) up[s] = __shfl_up_sync(0xffffffffu, u[M - 1][s], 1);
      const double kmax2 = colc[0], a0min = colc[1];
      auto cell = [&](int j, const double (&upv)[N]) {
        double x[N];
        const bool suspect = solve_cell<N>(kc, kmax2, a0min, u[j], upv, x);
        if (__any_sync(0xffffffffu, suspect)) {
          // rare: the per-step pivoting test, then the dense row-exchanging LU, one flagged cell at a time on the
          // warp's scratch block
          bool redo = false;
          if (suspect) redo = pivot_test<N>(sc, u[j]);
          double* w = cold + (size_t)warp * (2 * N + N * N + N);
          for (unsigned m = __ballot_sync(0xffffffffu, redo); m != 0u; m &= m - 1u) {
            const int src_lane = __ffs((int)m) - 1;
            if (lane == src_lane) {
#pragma unroll
              for (int s = 0; s < N; ++s) {
                w[s] = u[j][s];
                w[N + s] = upv[s];
              }
              dense_fallback_smem(sc, w, N);
#pragma unroll
              for (int s = 0; s < N; ++s) x[s] = w[s];
            }
            __syncwarp();
          }
        }
#pragma unroll
        for (int s = 0; s < N; ++s) {
          acc[j][s] = fma(x[s], wk, acc[j][s]);
          if (!last) {
            u[j][s] = fma(x[s], cuk, y[j][s]);
          } else {
            y[j][s] = fma(acc[j][s], cfin, y[j][s]);
            u[j][s] = y[j][s];
            acc[j][s] = 0.0;
          }
        }
      };
      // cells in DESCENDING order: cell j reads u[j - 1] before that cell is overwritten in place; the last cell is
      // posted as soon as it is done
#pragma unroll
      for (int j = M - 1; j >= 1; --j) {
        cell(j, u[j - 1]);
        if (j == M - 1) post(u[M - 1], slot_next);
      }
      if (lane == 0) {  // the first lane's upstream cell: the inlet (warp 0) or the upstream warp's last cell
        if (warp == 0) {
#pragma unroll
          for (int s = 0; s < N; ++s) up[s] = inlet[spar][irow][s];
        } else {
          const unsigned need = posted - (M > 1 ? 1u : 0u);  // inputs this warp had posted when the stage began
          while ((int)(ld_acquire_cta(&flags[warp - 1]) - need) < 0) {
          }
          const double* srcp = seam + ((size_t)slot * W + (warp - 1)) * XS;
#pragma unroll
          for (int s = 0; s < N; ++s) up[s] = srcp[s];
        }
      }
      cell(0, up);
      if (M == 1) post(u[0], slot_next);
    };

    post(y[M - 1], 0);  // input of stage 1 of step 0
    for (uint32_t step = 0; step < steps; ++step) {
      const int spar = (int)(step & 1u);
      for (int i = tid; i < N * ST; i += (int)blockDim.x) {  // inlet values of this step: [stage row][species]
        const int s = i % N, r = i / N;
        inlet[spar][r][s] = tabs[s][(size_t)step * ST + r];
      }
      // The ONE barrier of the step: publishes the inlet row and keeps every warp within one step of the others
      // (ring slots are reused two steps later).
      __syncthreads();
      const int s0 = spar * 4, s1 = (spar ^ 1) * 4;
      if (METHOD == CHROM_RK4) {
        stage(0, spar, s0 + 0, s0 + 1, 1.0, ch2, false);
        stage(1, spar, s0 + 1, s0 + 2, 2.0, ch2, false);
        stage(1, spar, s0 + 2, s0 + 3, 2.0, cdt, false);
        stage(2, spar, s0 + 3, s1 + 0, 1.0, ch2, true);
      } else {
        stage(0, spar, s0 + 0, s1 + 0, 1.0, cdt, true);
      }
      // validate_state after this step (solver/mod.rs:514-553): a thread records the first step at which one of its
      // real cells is non-finite, NaN winning over Inf within that step; the column's status is the minimum key.
      unsigned mx = 0u;
#pragma unroll
      for (int j = 0; j < M; ++j)
#pragma unroll
        for (int s = 0; s < N; ++s) mx = max(mx, abs_hi(y[j][s]));
      if (hi_nonfinite(mx) && !recorded) {
        bool bad = false, has_nan = false;
#pragma unroll
        for (int j = 0; j < M; ++j)
#pragma unroll
          for (int s = 0; s < N; ++s) {
            bad |= (j < nv) && hi_nonfinite(abs_hi(y[j][s]));
            has_nan |= (j < nv) && (y[j][s] != y[j][s]);
          }
        if (bad) {
          atomicMin(&badkey, (((unsigned long long)step + 1ull) << 1) | (has_nan ? 0ull : 1ull));
          recorded = true;
        }
      }

      // ---- samples --------------------------------------------------------------------------------
      const bool out_now = (step + 1u == oc.next);
      if (out_thread && (out_now || b.summary != nullptr)) {
#pragma unroll
        for (int s = 0; s < N; ++s) {
          const double v = out_value(s);
          if (out_now) b.outlet[((size_t)col * N + s) * b.n_out + oc.k] = v;
          if (b.summary) summ_step(summ[s], v, step, dt);
        }
      }
      if (out_now) sample_advance(oc, b.out_steps, b.outlet_stride);
      if (step + 1u == sc_.next) {
        store_profile(b.snapshots + ((size_t)col * b.n_snap + sc_.k) * N * nz);
        sample_advance(sc_, b.snap_steps, b.snapshot_stride);
      }
    }
    if (b.final_state) store_profile(b.final_state + (size_t)col * N * nz);
    if (out_thread && b.summary) {
#pragma unroll
      for (int s = 0; s < N; ++s) summ_store(b.summary + ((size_t)col * N + s) * CHROM_SUMMARY_LEN, summ[s]);
    }
    __syncthreads();  // every thread's atomicMin on badkey has landed
    if (tid == 0 && b.status && badkey != ~0ull) {
      const long long s = (long long)(badkey >> 1);
      b.status[col] = (badkey & 1ull) ? -s : s;
    }
  }
}

template <int N>
rkern_t pick_reg2_n(int method) {
  if constexpr (N >= 1 && N <= 8) return method == CHROM_RK4 ? k_multi_reg2<N, 2, CHROM_RK4> : k_multi_reg2<N, 2, CHROM_EULER>;
  return nullptr;
}

template <int N>
int reg2_tmax_rt() {
  if constexpr (N >= 1 && N <= 8) return reg2_max_threads(N, 2);
  return 0;
}

}  // namespace
}  // namespace chrom
