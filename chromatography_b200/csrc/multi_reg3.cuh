// multi_reg3.cuh — register-resident LangmuirMulti time integrator, third generation (n <= 8, FAST arithmetic).
//
// Same reference code as multi_reg2.cuh (rk4.rs:266-332, euler.rs:227-269, langmuir_multi.rs:627-686, 717-899,
// solver/mod.rs:514-553), the same per-cell mathematics (solve_cell / pivot_test / dense_fallback_smem of
// multi_reg2.cuh) and the same register layout (a thread owns 2 contiguous cells with y, u and acc in registers,
// species constants in uniform registers, upstream neighbour by shuffle).  Two structural changes, both aimed at
// what ncu showed k_multi_reg2 waiting on (profiles/r2_ncu_full_multi_reg2_c5.txt: branch_resolving 0.35,
// no_instruction 0.21 and barrier 0.20 stall cycles per issued instruction; timing-only ablations: guard 7 %,
// per-step barrier + seam protocol 10 %):
//
//  1. The partial-pivoting guard is decided per COLUMN RUN, not per cell.  The hot run solves every cell with the
//     closed-form elimination and only ORs the sufficient no-row-exchange bound into a sticky per-thread flag: no
//     warp vote, no branch, no cold code between two cells (the step loop shrinks from ~4650 to ~2000 SASS
//     instructions and ptxas may overlap the tail of one cell with the head of the next).  The flag is folded into
//     the step's CTA barrier (`__syncthreads_or`); if any cell of the column ever fails the bound, the CTA abandons
//     the run and integrates the column again from its initial state with the guarded code of multi_reg2.cuh
//     (per-step pivoting test, dense row-exchanging LU for a cell that needs an exchange).  A column whose cells all
//     pass the bound takes exactly the arithmetic it took before, so results are bit-identical to k_multi_reg2;
//     a column that fails it is computed by the guarded code alone, also bit-identical to k_multi_reg2.
//
//  2. SYNC selects the warp-seam protocol.  Six were measured on C5 (profiles/r2_c5_experiments.txt); three are built:
//       0  st.release / ld.acquire counters + one CTA barrier per step: multi_reg2.cuh's protocol, kept as the A/B
//          baseline (the release store compiles to MEMBAR.ALL.CTA + STS, the wait is a polling loop).  438 ms.
//       3  one mbarrier per (ring slot, seam): the producer's lane 31 stores the row (STS.128) and arrives
//          (SYNCS.ARRIVE: release semantics, no MEMBAR), the consumer waits with mbarrier.try_wait (the warp sleeps
//          in hardware instead of polling through the LSU).  Store and arrival are PREDICATED on lane 31 rather than
//          inside a divergent branch, and on the consumer side every lane waits and the row load is predicated on
//          lane 0, so a stage has no divergent branch at all.  The CTA barrier per step stays and bounds the skew
//          (the 8-slot ring needs no back-pressure).  392 ms (the same with the seam code in branches: 408 ms).
//       4  protocol 3 with a SPLIT step barrier (the default): instead of `__syncthreads` at the top of a step every
//          warp ARRIVES on a CTA-wide mbarrier there (count = warps) and WAITS for that phase only just before the
//          step's last post (3.5 stages later, when the others have long arrived).  The ring stays safe - slot 0 of
//          the next step is rewritten after everyone has finished the previous step, slots 1..3 were covered by the
//          previous step's wait - but no warp stands still at the top of a step; warp 0 publishes the inlet row for
//          itself and the redo flag is polled with a CTA barrier every 64 steps.  372 ms.
//     Measured and removed: full AND empty mbarriers with no step barrier at all (a warp runs ahead until the ring
//     is full: 517 ms - slack has to be bounded), and protocol 4 with the next inlet row requested by cp.async (382 ms).
//     The mbarriers are initialised ONCE per kernel.  (Re-initialising them per run with mbarrier.inval +
//     mbarrier.init stalled on B200: C5 and an abandoned Euler run never finished, first GPU session of this
//     kernel.)  Instead the phase bookkeeping runs on a CTA-wide step counter that continues across runs and
//     columns: input i of global step g lives in ring slot (g & 1) * 4 + i and is use g >> 1 of that slot's
//     barrier, and a run ends by topping every barrier up to the same number of completed phases (`settle`).
#pragma once
#include <type_traits>

#include "multi_reg2.cuh"

namespace chrom {
namespace {

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned long long* bar) {  // .release.cta by default
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {  // .acquire.cta by default
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "CHROM_MBAR_WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@!p bra CHROM_MBAR_WAIT_%=;\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}

// predicated forms (SYNC 3, 4): the seam store and the arrival are issued by every lane under a predicate instead of
// inside a divergent branch (no BSSY / BRA / BSYNC around five instructions)
__device__ __forceinline__ void st_shared_v2_if(bool p, double* dst, double a, double b) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %0, 0;\n"
      "@p st.shared.v2.f64 [%1], {%2, %3};\n"
      "}\n" ::"r"((int)p),
      "r"(smem_u32(dst)), "d"(a), "d"(b)
      : "memory");
}
__device__ __forceinline__ void st_shared_if(bool p, double* dst, double a) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %0, 0;\n"
      "@p st.shared.f64 [%1], %2;\n"
      "}\n" ::"r"((int)p),
      "r"(smem_u32(dst)), "d"(a)
      : "memory");
}
__device__ __forceinline__ void mbar_arrive_if(bool p, unsigned long long* bar) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %0, 0;\n"
      "@p mbarrier.arrive.shared::cta.b64 _, [%1];\n"
      "}\n" ::"r"((int)p),
      "r"(smem_u32(bar))
      : "memory");
}

constexpr int kReg3CheckEvery = 64;  // SYNC 4: steps between two CTA-wide polls of the redo flag

template <int N, int M, int METHOD, int SYNC>
__global__ void __launch_bounds__(reg2_max_threads(N, M), 1) k_multi_reg3(const BatchDev b) {
  constexpr int ST = (METHOD == CHROM_RK4) ? 3 : 1;
  constexpr int XS = (SYNC == 0) ? (N | 1) : ((N + 1) & ~1);  // seam rows: odd stride / 16-byte rows (STS.128)
  constexpr int kMaxW = 32;
  __shared__ SpeciesConst sc[N];
  __shared__ __align__(16) double inlet[2][3][N];  // [step parity][stage row][species]
  __shared__ Summ summ[N];           // running summary of the detector signal (detector owner only)
  __shared__ const double* tabs[N];
  __shared__ double colc[2];              // 2 max_i K_i, min_i (1 + fe lambda_i) of the current column
  __shared__ unsigned flags[32];          // SYNC 0: per warp, stage inputs posted so far in this run
  __shared__ unsigned long long badkey;   // min over threads of (first bad step << 1) | (NaN ? 0 : 1); ~0 = clean
  __shared__ __align__(8) unsigned long long bars[(SYNC == 0) ? 1 : 8 * kMaxW + 1];  // [slot][seam], then the step barrier
  static_assert(SYNC == 0 || SYNC == 3 || SYNC == 4, "seam protocols built: 0, 3, 4");
  constexpr bool FLAT = (SYNC >= 3);   // mbarrier seam, predicated seam code
  constexpr bool SPLIT = (SYNC == 4);  // split step barrier: arrive at the top of a step, wait before its last post
  extern __shared__ __align__(16) double seam3[];  // [8 slots][W warps][XS]: the last cell of each warp, per stage
                                                  // input; then cold[W][3 N + N N]: dense-fallback scratch per warp

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int W = blockDim.x >> 5;
  double* const seam = seam3;
  double* const cold = seam + (size_t)8 * W * XS;
  const uint32_t nz = b.nz;
  const int cell0 = tid * M;
  const int nv = max(0, min(M, (int)nz - cell0));  // real cells of this thread
  const double dt = b.dt, h2 = b.half_dt, h6 = b.sixth_dt;
  const uint32_t steps = b.steps;
  uint32_t gbase = 0u;  // even; global step number at which the next run starts (seam phase bookkeeping)
  auto full_bar = [&](int slot, int w) { return &bars[slot * kMaxW + w]; };
  unsigned long long* const step_bar = &bars[(SYNC == 0) ? 0 : 8 * kMaxW];  // SPLIT only
  if (FLAT) {
    for (int i = tid; i < 8 * kMaxW; i += (int)blockDim.x) mbar_init(&bars[i], 1u);
    if (SPLIT && tid == 0) mbar_init(step_bar, (unsigned)W);  // one arrival per warp and step
    __syncthreads();
  }

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

    // One integration of the column from its initial state.  GUARD = false: closed-form elimination on every cell,
    // the no-row-exchange bound only raises `sus`; returns true (CTA-uniform) when the column has to be integrated
    // again with GUARD = true, which resolves every suspect cell on the spot as k_multi_reg2 does.
    auto run = [&](auto guard_tag) -> bool {
      constexpr bool GUARD = decltype(guard_tag)::value;
      __syncthreads();  // nobody is still inside the previous run's seam protocol / reading badkey
      if (!FLAT && tid < 32) flags[tid] = 0u;
      const uint32_t G0 = gbase;  // this run's step t is global step G0 + t
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
      __syncthreads();  // barriers / flags / badkey initialised, colc published
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
      unsigned posted = 0u;  // SYNC 0: stage inputs this warp has posted (uniform over the CTA)
      bool recorded = false;
      bool sus = false;  // GUARD = false: some cell of this thread failed the sufficient bound (sticky)

      // The last cell of a stage's OUTPUT (= the input of stage `slot & 3` of global step g) goes to ring slot `slot`
      // for the downstream warp; its arrival completes phase g >> 1 of the slot's barrier (SYNC 0: the warp's counter
      // of posted inputs moves on).  wait_step: this is the step's last post (SPLIT waits for the step barrier first).
      auto post = [&](const double (&last)[N], int slot, uint32_t g, bool wait_step = false, uint32_t g_step = 0u) {
        if (FLAT) {
          if (SPLIT && wait_step) mbar_wait(step_bar, g_step & 1u);  // everyone has finished the previous step
          const bool poster = (lane == 31) && (warp + 1 < W);
          double* dst = seam + ((size_t)slot * W + warp) * XS;
          if (N % 2 == 0) {
#pragma unroll
            for (int s = 0; s + 1 < N; s += 2) st_shared_v2_if(poster, dst + s, last[s], last[s + 1]);
          } else {
#pragma unroll
            for (int s = 0; s < N; ++s) st_shared_if(poster, dst + s, last[s]);
          }
          mbar_arrive_if(poster, full_bar(slot, warp));
        } else {
          ++posted;
          if (lane == 31 && warp + 1 < W) {
            double* dst = seam + ((size_t)slot * W + warp) * XS;
#pragma unroll
            for (int s = 0; s < N; ++s) dst[s] = last[s];
            st_release_cta(&flags[warp], posted);
          }
        }
      };

      // One stage over this thread's cells.  Invariant at the top of a step: u == y, acc == 0.  Every stage solves
      // k = f(u), accumulates acc += w k (w = 1, 2, 2, 1: rk4.rs:311) and forms the next stage state u = y + cu k
      // (cu = dt/2, dt/2, dt); the last one forms y += acc dt/6 (Euler: y += k dt) and restores the invariant.
      auto stage = [&](int irow, int spar, int slot, int slot_next, uint32_t g, uint32_t g_next, double wk, double cuk,
                       bool last) {
        double up[N];
#pragma unroll
        for (int s = 0; s < N; ++s) up[s] = __shfl_up_sync(0xffffffffu, u[M - 1][s], 1);
        const double kmax2 = colc[0], a0min = colc[1];
        auto cell = [&](int j, const double (&upv)[N]) {
          double x[N];
          const bool suspect = solve_cell<N>(kc, kmax2, a0min, u[j], upv, x);
          if (!GUARD) {
            sus |= suspect;
          } else if (__any_sync(0xffffffffu, suspect)) {
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
        // cells in DESCENDING order: cell j reads u[j - 1] before that cell is overwritten in place; the last cell
        // is posted as soon as it is done
#pragma unroll
        for (int j = M - 1; j >= 1; --j) {
          cell(j, u[j - 1]);
          if (j == M - 1) post(u[M - 1], slot_next, g_next, last, g);
        }
        if (FLAT) {  // warp-uniform: every lane waits and loads the row (a broadcast), lane 0 keeps it
          const double* srcp = &inlet[spar][irow][0];
          if (warp != 0) {
            mbar_wait(full_bar(slot, warp - 1), (g >> 1) & 1u);
            srcp = seam + ((size_t)slot * W + (warp - 1)) * XS;
          }
          if (N % 2 == 0) {
#pragma unroll
            for (int s = 0; s + 1 < N; s += 2) {
              const double2 t = reinterpret_cast<const double2*>(srcp)[s / 2];
              up[s] = (lane == 0) ? t.x : up[s];
              up[s + 1] = (lane == 0) ? t.y : up[s + 1];
            }
          } else {
#pragma unroll
            for (int s = 0; s < N; ++s) {
              const double t = srcp[s];
              up[s] = (lane == 0) ? t : up[s];
            }
          }
        } else if (lane == 0) {  // the first lane's upstream cell: the inlet (warp 0) or the upstream warp's last cell
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
        if (M == 1) post(u[0], slot_next, g_next, last, g);
      };

      post(y[M - 1], 0, G0);  // input of stage 1 of step 0 (G0 is even: slot 0)
      bool abandon = false;
      uint32_t done = steps;  // steps whose stage inputs were all posted and consumed when the run ended
      for (uint32_t step = 0; step < steps; ++step) {
        const uint32_t g = G0 + step;
        const int spar = (int)(g & 1u);
        if (!SPLIT) {
          for (int i = tid; i < N * ST; i += (int)blockDim.x) {  // inlet values of this step: [stage row][species]
            const int s = i % N, r = i / N;
            inlet[spar][r][s] = tabs[s][(size_t)step * ST + r];
          }
          // The ONE barrier of the step: publishes the inlet row and keeps every warp within one step of the others
          // (ring slots are reused two steps later); the hot run folds the redo flag into it.
          if (!GUARD) {
            if (__syncthreads_or(sus ? 1 : 0)) {
              abandon = true;
              done = step;
              break;
            }
          } else {
            __syncthreads();
          }
        } else {
          if (warp == 0) {  // only warp 0 reads the inlet: it publishes the row for itself
            for (int i = lane; i < N * ST; i += 32) {
              const int s = i % N, r = i / N;
              inlet[spar][r][s] = tabs[s][(size_t)step * ST + r];
            }
            __syncwarp();
          }
          if (!GUARD && (step % kReg3CheckEvery) == kReg3CheckEvery - 1) {
            if (__syncthreads_or(sus ? 1 : 0)) {
              abandon = true;
              done = step;
              break;
            }
          }
          mbar_arrive_if(lane == 0, step_bar);  // "this warp has finished step g - 1": one of the W arrivals of phase g
        }
        const int s0 = spar * 4, s1 = (spar ^ 1) * 4;
        if (METHOD == CHROM_RK4) {
          stage(0, spar, s0 + 0, s0 + 1, g, g, 1.0, ch2, false);
          stage(1, spar, s0 + 1, s0 + 2, g, g, 2.0, ch2, false);
          stage(1, spar, s0 + 2, s0 + 3, g, g, 2.0, cdt, false);
          stage(2, spar, s0 + 3, s1 + 0, g, g + 1u, 1.0, ch2, true);
        } else {
          stage(0, spar, s0 + 0, s1 + 0, g, g + 1u, 1.0, cdt, true);
        }
        // validate_state after this step (solver/mod.rs:514-553): a thread records the first step at which one of
        // its real cells is non-finite, NaN winning over Inf within that step; the column's status is the minimum.
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
      // every thread's last cells and atomicMin on badkey have landed; the hot run learns here about a cell that
      // failed the bound after the last poll
      bool again = abandon;
      if (!abandon) {
        if (!GUARD)
          again = __syncthreads_or(sus ? 1 : 0) != 0;
        else
          __syncthreads();
      }
      if (FLAT) {
        // settle: nobody waits on a seam barrier any more (CTA barrier above).  Posted so far: input 0 of every global
        // step <= gend, inputs 1..3 of every step < gend.  Top each barrier up to the phases of all steps < G1 (the
        // next even step number > gend), where the next run starts.
        const uint32_t gend = G0 + done, G1 = (gend + 2u) & ~1u;
        constexpr int kIn = (METHOD == CHROM_RK4) ? 4 : 1;  // stage inputs per step that cross a seam
        if (lane == 31 && warp + 1 < W)
          for (uint32_t g2 = gend; g2 < G1; ++g2)
            for (int i = (g2 == gend ? 1 : 0); i < kIn; ++i) mbar_arrive(full_bar((int)(g2 & 1u) * 4 + i, warp));
        if (SPLIT)  // the step barrier has one phase per global step: arrivals exist for every g < gend
          for (uint32_t g2 = gend; g2 < G1; ++g2) mbar_arrive_if(lane == 0, step_bar);
        gbase = G1;
      }
      if (again) return true;
      if (b.final_state) store_profile(b.final_state + (size_t)col * N * nz);
      if (out_thread && b.summary) {
#pragma unroll
        for (int s = 0; s < N; ++s) summ_store(b.summary + ((size_t)col * N + s) * CHROM_SUMMARY_LEN, summ[s]);
      }
      if (tid == 0 && b.status && badkey != ~0ull) {
        const long long s = (long long)(badkey >> 1);
        b.status[col] = (badkey & 1ull) ? -s : s;
      }
      return false;
    };

    if (run(std::false_type{})) run(std::true_type{});
  }
}

template <int N>
rkern_t pick_reg3_n(int method, int sync) {
  if constexpr (N >= 1 && N <= 8) {
    const bool rk4 = method == CHROM_RK4;
    switch (sync) {
      case 0: return rk4 ? k_multi_reg3<N, 2, CHROM_RK4, 0> : k_multi_reg3<N, 2, CHROM_EULER, 0>;
      case 3: return rk4 ? k_multi_reg3<N, 2, CHROM_RK4, 3> : k_multi_reg3<N, 2, CHROM_EULER, 3>;
      case 4: return rk4 ? k_multi_reg3<N, 2, CHROM_RK4, 4> : k_multi_reg3<N, 2, CHROM_EULER, 4>;
    }
  }
  return nullptr;
}

}  // namespace
}  // namespace chrom
