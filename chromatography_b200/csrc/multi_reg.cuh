// multi_reg.cuh — register-resident LangmuirMulti time integrator (RK4 / Euler), n <= 16, FAST arithmetic.
//
// Replaces, for a batch of columns and ALL time steps in one launch, the same reference code as
// multi_kernels.cuh (rk4.rs:266-332, euler.rs:227-269, langmuir_multi.rs:627-686, 717-899,
// solver/mod.rs:514-553).  What differs is where the state lives.  The shared-memory kernel of
// multi_kernels.cuh moves ~110 shared-memory wavefronts per warp and cell-stage (c, upstream, y, acc in; stage
// state, acc out; species constants) and is bound by that pipe, not by FP64.  Here a thread owns M CONTIGUOUS
// cells with y, acc and the stage state u of all n species in registers (3*M*n doubles), exactly the scheme of
// single_resident.cu.  The
// upstream neighbour of a thread's first cell is the previous thread's last cell: every thread posts that cell
// (n doubles) to a double-buffered shared array xch[parity][species][thread + 1], slot 0 holding the inlet, and
// reads slot `thread` after the ONE __syncthreads of the stage: n stores + n loads, conflict-free, no lane
// special-casing (a shuffle would cost 2 instructions per double plus the warp-boundary handling).  Shared memory
// otherwise only carries the species constants (broadcast loads), the inlet row, the outlet summary and, for the
// wider shapes, the per-thread slots of the running slope acc.
//
// Per cell the solve is scan_row: the elimination of A = diag(alpha) - p q^T in closed form (pivot
// alpha_i - p_i q_i / h_i, h_i = 1 - sum_{j<i} p_j q_j / alpha_j; x = z + v (q.z)/(1 - q.v)), n independent
// reciprocals instead of n chained ones, with the partial-pivoting test of multi_kernels.cuh::scan_cell; a warp
// with a cell that needs a row exchange redoes it with the dense row-exchanging register LU (fast_row).
#pragma once
#include "multi_kernels.cuh"

namespace chrom {
namespace {

// doubles a thread keeps live: state 3*M*N, solver temporaries ~3*N  ->  registers ~ 2x + overhead.
// The register file is 4 x 16K (one per SM sub-partition) and a CTA's warps are dealt round-robin to the
// sub-partitions, so the thread limit comes in steps of 128: w warps per sub-partition leave 512 / w registers.
// acc (the running weighted slope) moves to shared memory, one private slot per thread, once y/u/acc together
// would not fit: 12 shared accesses per cell-stage and species instead of register spills.
__host__ __device__ constexpr bool reg_acc_shared(int N, int M) { return 3 * M * N > 18; }
constexpr int reg_budget(int N, int M) {
  return 2 * ((reg_acc_shared(N, M) ? 2 : 3) * M * N + 3 * N + 2 * N * (M - 1)) + 32;
}
constexpr int reg_max_threads(int N, int M) {
  const int w = 512 / reg_budget(N, M);
  return w < 1 ? 128 : (w > 8 ? 1024 : 128 * w);
}

// scan_cell of multi_kernels.cuh on register operands.  Returns true when the cell needs a row exchange.
template <int N>
// x solves A x = c - up: the factor -ue/dz of the right-hand side is folded into the stage coefficients by the caller.
__device__ __forceinline__ bool scan_row(const SpeciesConst* sc, double kmax, double a0min,
                                         const double (&c)[N], const double (&up)[N], double (&x)[N]) {
  double pv[N];  // feNK_s c_s, then v_s = p_s / alpha_s
  double S0 = 0.0, S1 = 0.0;
  unsigned pmax_hi = 0u;  // max over the HIGH WORDS of feNK_s c_s, as unsigned: one integer max per species (a
                          // double max costs 6 instructions here); a negative or non-finite entry reads as huge
#pragma unroll
  for (int s = 0; s < N; ++s) {
    const double2 kf = *reinterpret_cast<const double2*>(&sc[s].K);  // {K, feNK}
    x[s] = c[s] - up[s];
    if (s & 1) S1 = fma(kf.x, c[s], S1);
    else S0 = fma(kf.x, c[s], S0);
    pv[s] = kf.y * c[s];
    pmax_hi = max(pmax_hi, (unsigned)__double2hiint(pv[s]));
  }
  const double rD = fast_rcp(1.0 + (S0 + S1));
  const double rD2 = rD * rD;
  // upper bound of max_r |p_r| (next high word up); all c >= 0 iff the sign bit is clear
  const bool nonneg = pmax_hi < 0x7ff00000u;
  const double pmax2 = __hiloint2double((int)(pmax_hi + 1u), 0) * rD2;
  double qv0 = 0.0, qv1 = 0.0, qz0 = 0.0, qz1 = 0.0;
#pragma unroll
  for (int i = 0; i < N; ++i) {
    const double2 af = *reinterpret_cast<const double2*>(&sc[i].a0);  // {a0, feNK}
    const double ra = fast_rcp(fma(af.y, rD, af.x));                    // 1 / alpha_i
    const double Ki = sc[i].K;
    x[i] *= ra;                   // z_i
    pv[i] = (pv[i] * rD2) * ra;   // v_i
    if (i & 1) {
      qv1 = fma(Ki, pv[i], qv1);
      qz1 = fma(Ki, x[i], qz1);
    } else {
      qv0 = fma(Ki, pv[i], qv0);
      qz0 = fma(Ki, x[i], qz0);
    }
  }
  const double hn = 1.0 - (qv0 + qv1);
  const double f = (qz0 + qz1) * fast_rcp(hn);
  bool redo = false;
  // Sufficient for "no row exchange anywhere": 2 max q max|p| < min alpha (1 - q.v), with alpha_i >= a0_i when
  // every c >= 0 (h_i >= h_n then).  Anything else takes the per-step test.
  if (!(nonneg && (2.0 * kmax) * pmax2 < a0min * hn)) {
    double run = 0.0;
#pragma unroll
    for (int i = 0; i < N; ++i) {
      const double2 af = *reinterpret_cast<const double2*>(&sc[i].a0);
      const double alpha = fma(af.y, rD, af.x);
      const double Ki = sc[i].K;
      const double h = 1.0 - run;
      const double p = pv[i] * alpha;  // v_i alpha_i = p_i
      double below = 0.0;              // max_{r>i} |p_r|
#pragma unroll
      for (int r = i + 1; r < N; ++r) {
        const double2 ar = *reinterpret_cast<const double2*>(&sc[r].a0);
        below = fmax(below, fabs(pv[r] * fma(ar.y, rD, ar.x)));
      }
      const double dpiv = fabs(fma(alpha, h, -(p * Ki)));
      redo |= !(h > 0.0) || (Ki * below > dpiv) || (dpiv == 0.0);
      run = fma(Ki, pv[i], run);
    }
  }
#pragma unroll
  for (int i = 0; i < N; ++i) x[i] = fma(pv[i], f, x[i]);
  return redo;
}

// dense row-exchanging register LU for a flagged cell, out of line (its matrix must not inflate the hot path).
// Operands and result travel BY VALUE: no array of the hot path ever has its address taken.
template <int N>
struct VecN {
  double v[N];
};

template <int N>
__device__ __noinline__ VecN<N> cold_row(const SpeciesConst* sc, double fe, double cg, const VecN<N> c, const VecN<N> up) {
  VecN<N> k;
  fast_row<N>(sc, fe, cg, c.v, up.v, k.v);  // warp-synchronous inside (pivot votes): called by whole warps
  return k;
}

template <int N, int M, int METHOD>
__global__ void __launch_bounds__(reg_max_threads(N, M), 1) k_multi_reg(const BatchDev b) {
  constexpr int ST = (METHOD == CHROM_RK4) ? 3 : 1;
  __shared__ SpeciesConst sc[N];
  __shared__ double inlet[2][3][N];   // [step parity][stage row][species]
  __shared__ Summ summ[N];            // running summary of the detector signal (detector owner only)
  __shared__ const double* tabs[N];
  __shared__ double colc[2];          // max_i K_i, min_i (1 + fe lambda_i) of the current column
  constexpr bool ACC_SH = reg_acc_shared(N, M);
  extern __shared__ double dyn_sh[];  // xch[2][T + 1][N | 1], then (ACC_SH) acc[T][(M * N) | 1]

  const int tid = threadIdx.x;
  const uint32_t nz = b.nz;
  const int cell0 = tid * M;
  const int nv = max(0, min(M, (int)nz - cell0));  // real cells of this thread
  const double dt = b.dt, h2 = b.half_dt, h6 = b.sixth_dt;
  const uint32_t steps = b.steps;

  for (uint32_t col = blockIdx.x; col < b.n_cols; col += gridDim.x) {  // persistent CTAs over columns
    const chrom_multi_col cp = b.mcols[col];
    const double fe = cp.fe;
    const double cg = -cp.ue / cp.dz;
    const double cdt = cg * dt, ch2 = cg * h2, ch6 = cg * h6;  // k is solved without the factor -ue/dz (scan_row)
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
    double y[M][N], u[M][N], acc[ACC_SH ? 1 : M][ACC_SH ? 1 : N];
    const int T = blockDim.x;
    // Shared-memory layouts keep the species index fastest with an ODD row stride (in doubles): a thread's row is
    // addressed with compile-time offsets from one base register and consecutive lanes hit distinct banks.
    constexpr int XS = N | 1;        // xch[2][T + 1][XS]: slot t + 1 = thread t's last cell, slot 0 = the inlet
    constexpr int AS = (M * N) | 1;  // acc[T][AS] (ACC_SH): thread-private slots
    double* const xch = dyn_sh;
    double* const my_acc = dyn_sh + 2 * (T + 1) * XS + tid * AS;
#pragma unroll
    for (int j = 0; j < M; ++j)
#pragma unroll
      for (int s = 0; s < N; ++s) {
        y[j][s] = (b.y0 != nullptr && j < nv) ? b.y0[((size_t)col * N + s) * nz + cell0 + j] : 0.0;
        u[j][s] = 0.0;
        if (!ACC_SH) acc[j][s] = 0.0;
      }
    auto out_value = [&](int s) -> double {  // species s of the outlet cell (meaningful on out_thread)
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
    if (tid == 0) {
      double km = 0.0, am = 1e300;
#pragma unroll
      for (int s = 0; s < N; ++s) {
        km = fmax(km, sc[s].K);
        am = fmin(am, sc[s].a0);
      }
      colc[0] = km;
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
    bool done = false;

    // One stage over this thread's cells: k = f(src), then the stage algebra `which`
    // (0 Euler; 1..4 RK4).  src/dst are register arrays; stages 2-4 run in place on u.
    auto stage = [&](double (&src)[M][N], int irow, int spar, int par, int which) {
      double upx[N];
      double* const post = xch + ((size_t)par * (T + 1) + tid) * XS;  // row of slot `tid`; the thread's own row is next
#pragma unroll
      for (int s = 0; s < N; ++s) post[XS + s] = src[M - 1][s];
      if (tid < N) xch[(size_t)par * (T + 1) * XS + tid] = inlet[spar][irow][tid];  // slot 0: the inlet of this stage
      __syncthreads();
#pragma unroll
      for (int s = 0; s < N; ++s) upx[s] = post[s];
      // All M cells are solved before any is updated: their chains are independent and interleave (one vote for
      // the rare dense-LU redo instead of one per cell), and the update needs no ordering.
      double k[M][N];
      bool redo[M];
      bool any_redo = false;
#pragma unroll
      for (int j = 0; j < M; ++j) {
        double upv[N];
#pragma unroll
        for (int s = 0; s < N; ++s) upv[s] = (j > 0) ? src[j - 1][s] : upx[s];
        redo[j] = scan_row<N>(sc, colc[0], colc[1], src[j], upv, k[j]);
        any_redo |= redo[j];
      }
      if (__any_sync(0xffffffffu, any_redo)) {
#pragma unroll
        for (int j = 0; j < M; ++j) {
          if (!__any_sync(0xffffffffu, redo[j])) continue;
          VecN<N> cc, uu;
#pragma unroll
          for (int s = 0; s < N; ++s) {
            cc.v[s] = src[j][s];
            uu.v[s] = (j > 0) ? src[j - 1][s] : upx[s];
          }
          const VecN<N> kk = cold_row<N>(sc, fe, 1.0, cc, uu);  // unscaled like scan_row
#pragma unroll
          for (int s = 0; s < N; ++s) k[j][s] = kk.v[s];
        }
      }
#pragma unroll
      for (int j = 0; j < M; ++j) {
#pragma unroll
        for (int s = 0; s < N; ++s) {
          const double kv = k[j][s];
          if (which == 0) {
            y[j][s] = fma(kv, cdt, y[j][s]);
          } else if (which == 1) {
            if (ACC_SH) my_acc[j * N + s] = kv;
            else acc[j][s] = kv;
            u[j][s] = fma(kv, ch2, y[j][s]);
          } else if (which == 2 || which == 3) {
            if (ACC_SH) my_acc[j * N + s] = fma(kv, 2.0, my_acc[j * N + s]);
            else acc[j][s] = fma(kv, 2.0, acc[j][s]);
            u[j][s] = fma(kv, which == 2 ? ch2 : cdt, y[j][s]);
          } else {
            const double a = ACC_SH ? my_acc[j * N + s] : acc[j][s];
            y[j][s] = fma(a + kv, ch6, y[j][s]);
          }
        }
      }
    };

    bool maybe = false;  // some cell of this thread may be non-finite after the previous step
    for (uint32_t step = 0; step <= steps; ++step) {
      const int spar = (int)(step & 1u);
      if (step < steps) {  // inlet values of this step: [stage row][species]; T may be < N * ST (nz <= 32, n >= 11)
        for (int i = tid; i < N * ST; i += (int)blockDim.x) {
          const int s = i % N, r = i / N;
          inlet[spar][r][s] = tabs[s][(size_t)step * ST + r];
        }
      }
      // one barrier: publishes the inlet row and carries the previous step's non-finite vote
      const int any_maybe = __syncthreads_or(maybe);
      if (any_maybe && !done) {  // validate_state after step `step` (solver/mod.rs:514-553): NaN wins over Inf
        bool bad = false, has_nan = false;
#pragma unroll
        for (int j = 0; j < M; ++j)
#pragma unroll
          for (int s = 0; s < N; ++s) {
            bad |= (j < nv) && hi_nonfinite(abs_hi(y[j][s]));
            has_nan |= (j < nv) && (y[j][s] != y[j][s]);
          }
        const int any_bad = __syncthreads_or(bad);
        const int any_nan = __syncthreads_or(has_nan);
        if (any_bad) {
          if (tid == 0 && b.status) b.status[col] = any_nan ? (long long)step : -(long long)step;
          done = true;
        }
      }
      if (step == steps) break;

      if (METHOD == CHROM_RK4) {
        stage(y, 0, spar, 0, 1);
        stage(u, 1, spar, 1, 2);
        stage(u, 1, spar, 0, 3);
        stage(u, 2, spar, 1, 4);
      } else {
        stage(y, 0, spar, spar, 0);
      }
      unsigned mx = 0u;
#pragma unroll
      for (int j = 0; j < M; ++j)
#pragma unroll
        for (int s = 0; s < N; ++s) mx = max(mx, abs_hi(y[j][s]));
      maybe = hi_nonfinite(mx);

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
  }
}

typedef void (*rkern_t)(const BatchDev);

// cells per thread: 1 for every n <= 16, 2 for n <= 12, 4 for n <= 4 (register file)
template <int N>
rkern_t pick_reg_n(int method, int m) {
  if constexpr (N >= 1 && N <= 16) {
    const bool rk4 = method == CHROM_RK4;
    if (m == 1) return rk4 ? k_multi_reg<N, 1, CHROM_RK4> : k_multi_reg<N, 1, CHROM_EULER>;
    if constexpr (N <= 12) {
      if (m == 2) return rk4 ? k_multi_reg<N, 2, CHROM_RK4> : k_multi_reg<N, 2, CHROM_EULER>;
    }
    if constexpr (N <= 4) {
      if (m == 4) return rk4 ? k_multi_reg<N, 4, CHROM_RK4> : k_multi_reg<N, 4, CHROM_EULER>;
    }
  }
  return nullptr;
}

template <int N>
bool reg_acc_shared_rt(int m) {
  if constexpr (N >= 1 && N <= 16) {
    if (m == 1) return reg_acc_shared(N, 1);
    if (m == 2 && N <= 12) return reg_acc_shared(N, 2);
    if (m == 4 && N <= 4) return reg_acc_shared(N, 4);
  }
  return false;
}

template <int N>
int reg_max_threads_rt(int m) {
  if constexpr (N >= 1 && N <= 16) {
    if (m == 1) return reg_max_threads(N, 1);
    if (m == 2 && N <= 12) return reg_max_threads(N, 2);
    if (m == 4 && N <= 4) return reg_max_threads(N, 4);
  }
  return 0;
}

}  // namespace
}  // namespace chrom
