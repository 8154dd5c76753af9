// single_stream.cu — LangmuirSingle columns too long for the register-resident path (nz > 8192,
// e.g. the nz = 2^22 spatial-convergence column of BASELINE config 4): the state streams through
// HBM/L2, FUSE time steps per pass (temporal blocking, RK4 and Euler), all four RK4 stages of each step fused.
//
// Replaces per launch FUSE iterations of the time loop (rk4.rs:266-332 / euler.rs:227-269) with
// LangmuirSingle::compute_physics (langmuir_single.rs:445-516) evaluated 4 (or 1) times in registers.
//
// Tiling: a warp owns 32*M consecutive cells, M per lane (same register scheme as
// single_resident.cu: one shuffle per lane per stage).  RK4 needs the stage values of the four cells
// upstream of a tile per time step; instead of exchanging them the tile starts 4*FUSE cells early and
// discards its first 4*FUSE results (they are recomputed, correctly, by the previous tile): no
// inter-CTA communication.  Default M = 16, FUSE = 8: halo 32 of 512 cells (6.3 % redundant work),
// the state crosses L2/HBM once per 8 steps.  Tile 0 starts at the inlet and keeps everything.
// Algorithmic HBM traffic: 16 B per cell-step (read y_n, write y_n+1); both ping-pong buffers of
// the 2^22 column (67 MB) fit the 126 MB L2.  Per-column coefficients are derived once per solve
// (k_stream_coef); which sub-steps sample the outlet is decided on the host (no 64-bit division
// in the kernel).
#include <algorithm>
#include <cstdio>
#include <cstdlib>

#include "arith.cuh"
#include "kernels.h"
#include "single_rhs.cuh"

namespace chrom {
namespace {

constexpr int kCoefSlots = 8;  // doubles reserved per column for Coef<ARITH> (sizeof <= 64)
constexpr int kMaxWarps = 4;  // warps per CTA (upper bound; the launch picks 1..4)

// status word while streaming: ~0 = clean, else (step << 1) | (NaN ? 0 : 1), merged with atomicMin: the first
// bad step wins, NaN over Inf within it (solver/mod.rs:519-548).
//
// FUSE consecutive time steps are advanced per launch (temporal blocking): the state crosses HBM/L2
// once per FUSE steps, the halo grows to 4*FUSE cells (FUSE = 4: 6.5 % redundant work at M = 8).
// Outlet samples and the non-finite scan are taken after every sub-step, on cells no halo touches.
// kM cells per lane (a warp tile is 32*kM cells): larger kM = more independent cells per lane for the
// FP64 pipe and a smaller halo fraction, at 6 registers per cell.
template <int METHOD, int ARITH, int FUSE, int kM>
__global__ void __launch_bounds__(kMaxWarps * 32)
    k_single_stream_step(const BatchDev b, const double* __restrict__ src, double* __restrict__ dst, uint32_t step,
                         unsigned long long* __restrict__ code, int nsub, int tiles_per_col, uint32_t out_mask,
                         uint32_t out_slot0) {
  constexpr int ST = (METHOD == CHROM_RK4) ? 3 : 1;
  constexpr int kTile = 32 * kM;
  // RK4 stage values reach 4 cells upstream per step, a forward-Euler step 1; an unfused Euler step needs no halo
  // (the upstream neighbour of the tile's first cell is an input of the step)
  constexpr int H = (METHOD == CHROM_RK4) ? 4 * FUSE : (FUSE > 1 ? FUSE : 0);
  constexpr int V = kTile - H;                             // cells a non-first tile contributes
  using Q = Coef<ARITH>;
  const int lane = threadIdx.x & 31;
  const int tile = (int)(blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5));
  const uint32_t col = blockIdx.y;
  if (tile >= tiles_per_col) return;
  const uint32_t nz = b.nz;
  // Tile 0 covers [0, kTile); tile t >= 1 starts H cells early: kTile + (t-1)*V - H = t*V since
  // V = kTile - H.  Written without a select on purpose: `(tile == 0) ? 0 : 64-bit expr` made
  // ptxas 12.9 emit CS2R R, SRZ + predicated IMAD.WIDE + an under-stalled consumer (tile 0 read a
  // stale register in the Euler/fast variant; found by the parity tests).
  const long long base = (long long)tile * V;
  const long long cell0 = base + (long long)lane * kM;
  const int first_valid = (tile == 0) ? 0 : H;  // tile-local index of the first cell this tile owns
  const int local0 = lane * kM;                 // tile-local index of this lane's first cell

  // per-column coefficients, derived once per solve by k_stream_coef (their IEEE divisions would cost ~9 %
  // of a launch if every thread repeated them): kCoefSlots doubles per column behind the status codes
  const Q q = *reinterpret_cast<const Q*>(reinterpret_cast<const double*>(code + b.n_cols) + (size_t)col * kCoefSlots);
  const double* __restrict__ tab = b.tables + ((size_t)b.scols[col].inj_table * b.steps + step) * ST;
  const double dt = b.dt, h2 = b.half_dt, h6 = b.sixth_dt;

  const double* __restrict__ s = src + (size_t)col * nz;
  double* __restrict__ d = dst + (size_t)col * nz;
  double y[kM], u[kM], acc[kM];
  const bool vec = ((nz & 1u) == 0u) && (cell0 + kM <= (long long)nz);
  if (vec) {
    const double2* sv = reinterpret_cast<const double2*>(s + cell0);
#pragma unroll
    for (int j = 0; j < kM / 2; ++j) {
      const double2 v = sv[j];
      y[2 * j] = v.x;
      y[2 * j + 1] = v.y;
    }
  } else {
#pragma unroll
    for (int j = 0; j < kM; ++j) y[j] = (cell0 + j < (long long)nz) ? s[cell0 + j] : 0.0;
  }
  const bool inlet_lane = (tile == 0) && (lane == 0);
  double euler_up = 0.0;  // unfused Euler: the upstream neighbour of the tile's first cell is an input
  if (METHOD == CHROM_EULER && FUSE == 1 && lane == 0 && tile > 0) euler_up = s[cell0 - 1];

  auto exchange = [&](double last, double inlet) -> double {
    const double up = __shfl_up_sync(0xffffffffu, last, 1);
    return inlet_lane ? inlet : up;  // lane 0 of a later tile gets its own value back: halo, discarded
  };

  // cells [j_lo, j_hi) of this lane are owned by the tile and inside the column
  const int j_lo = min(kM, max(0, first_valid - local0));
  const int j_hi = (int)min((long long)kM, max(0ll, (long long)nz - cell0));
  const bool full_lane = (j_lo == 0) && (j_hi == kM);
  // the detector cell (default: the outlet) lives on exactly one lane of one tile
  const long long oj = (long long)detector_of(b.scols[col].detector_cell, nz) - cell0;
  const int out_j = ((b.outlet != nullptr || b.series != nullptr) && oj >= (long long)j_lo && oj < (long long)j_hi) ? (int)oj : -1;

  bool bad = false, has_nan = false;
  uint32_t bad_step = 0;
  for (int f = 0; f < FUSE; ++f) {
    if (f >= nsub) break;
    const double cin0 = tab[f * ST];
    const double cin1 = (ST == 3) ? tab[f * ST + 1] : 0.0;
    const double cin2 = (ST == 3) ? tab[f * ST + 2] : 0.0;
    if (METHOD == CHROM_RK4) {
      double up = exchange(y[kM - 1], cin0);
#pragma unroll
      for (int j = kM - 1; j >= 0; --j) {
        const double k = q.rhs(y[j], j > 0 ? y[j - 1] : up);
        acc[j] = k;
        u[j] = Q::axpy(y[j], k, h2);
      }
      up = exchange(u[kM - 1], cin1);
#pragma unroll
      for (int j = kM - 1; j >= 0; --j) {
        const double k = q.rhs(u[j], j > 0 ? u[j - 1] : up);
        acc[j] = Q::axpy(acc[j], k, 2.0);
        u[j] = Q::axpy(y[j], k, h2);
      }
      up = exchange(u[kM - 1], cin1);
#pragma unroll
      for (int j = kM - 1; j >= 0; --j) {
        const double k = q.rhs(u[j], j > 0 ? u[j - 1] : up);
        acc[j] = Q::axpy(acc[j], k, 2.0);
        u[j] = Q::axpy(y[j], k, dt);
      }
      up = exchange(u[kM - 1], cin2);
#pragma unroll
      for (int j = kM - 1; j >= 0; --j) {
        const double k = q.rhs(u[j], j > 0 ? u[j - 1] : up);
        y[j] = Q::axpy(y[j], Q::add(acc[j], k), h6);
      }
    } else {
      double up = exchange(y[kM - 1], cin0);
      if (FUSE == 1 && lane == 0 && tile > 0) up = euler_up;
#pragma unroll
      for (int j = kM - 1; j >= 0; --j) {
        const double k = q.rhs(y[j], j > 0 ? y[j - 1] : up);
        y[j] = Q::axpy(y[j], k, dt);
      }
    }
    // ---- per sub-step: outlet sample and non-finite scan on the cells this tile owns ------------------
    // (which sub-steps sample the outlet, and into which slot, is decided on the host: out_mask/out_slot0)
    const uint32_t done_steps = step + (uint32_t)f + 1u;
    bool sub_bad;
    if (full_lane) {  // every cell of this lane is owned and inside the column: one max over the high words
      unsigned m = 0u;
#pragma unroll
      for (int j = 0; j < kM; ++j) m = max(m, abs_hi(y[j]));
      sub_bad = hi_nonfinite(m);
    } else {  // halo lanes of a tile and the ragged end of the column
      sub_bad = false;
#pragma unroll
      for (int j = 0; j < kM; ++j)
        if (j >= j_lo && j < j_hi) sub_bad |= hi_nonfinite(abs_hi(y[j]));
    }
    if (out_j >= 0 && (b.series != nullptr || ((out_mask >> f) & 1u))) {  // the one lane that holds the detector cell
      double v = y[0];
#pragma unroll
      for (int j = 1; j < kM; ++j) v = (out_j == j) ? y[j] : v;
      if ((out_mask >> f) & 1u)
        b.outlet[(size_t)col * b.n_out + out_slot0 + (uint32_t)__popc(out_mask & ((1u << f) - 1u))] = v;
      if (b.series != nullptr) b.series[(size_t)col * (b.steps + 1) + done_steps] = v;  // full rate, for the summary
    }
    if (sub_bad && !bad) {  // remember the FIRST bad sub-step of this lane (rare: the NaN scan lives here)
      bool sub_nan = false;
#pragma unroll
      for (int j = 0; j < kM; ++j)
        if (j >= j_lo && j < j_hi) sub_nan |= (y[j] != y[j]);
      bad = true;
      has_nan = sub_nan;
      bad_step = done_steps;
    }
  }

  // ---- store the cells this tile owns ---------------------------------------------------------------
  if (vec && local0 >= first_valid) {
    double2* dv = reinterpret_cast<double2*>(d + cell0);
#pragma unroll
    for (int j = 0; j < kM / 2; ++j) dv[j] = make_double2(y[2 * j], y[2 * j + 1]);
  } else {
#pragma unroll
    for (int j = 0; j < kM; ++j) {
      const long long cell = cell0 + j;
      if (local0 + j >= first_valid && cell < (long long)nz) d[cell] = y[j];
    }
  }
  if (__any_sync(0xffffffffu, bad)) {
    // key = (step << 1) | (NaN ? 0 : 1): the minimum over lanes, tiles and launches is the first bad
    // step, NaN winning over Inf within it (validate_state order).  ~0 = clean.
    unsigned long long key = bad ? ((((unsigned long long)bad_step) << 1) | (has_nan ? 0ull : 1ull)) : ~0ull;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const unsigned long long other = __shfl_xor_sync(0xffffffffu, key, o);
      key = other < key ? other : key;
    }
    if (lane == 0) atomicMin(&code[col], key);
  }
}

// After the last step: status codes -> the ABI's signed convention, summaries from the full-rate detector series.
__global__ void k_stream_finalize(const BatchDev b, const unsigned long long* __restrict__ code) {
  const uint32_t col = blockIdx.x * blockDim.x + threadIdx.x;
  if (col >= b.n_cols) return;
  if (b.status) {
    const unsigned long long c = code[col];  // ~0 clean, else (step << 1) | (NaN ? 0 : 1)
    const long long s = (long long)(c >> 1);
    b.status[col] = (c == ~0ull) ? 0ll : ((c & 1ull) ? -s : s);
  }
  if (b.summary && b.series) {
    const double* o = b.series + (size_t)col * (b.steps + 1);
    Summ sm;
    summ_init(sm, o[0]);
    for (uint32_t st = 0; st < b.steps; ++st) summ_step(sm, o[st + 1], st, b.dt);
    summ_store(b.summary + (size_t)col * CHROM_SUMMARY_LEN, sm);
  }
}

// Sample slot 0 (the initial state at the detector cell) of the outlet buffer and of the full-rate series.
__global__ void k_stream_sample0(const BatchDev b, const double* __restrict__ state, int to_outlet) {
  const uint32_t col = blockIdx.x * blockDim.x + threadIdx.x;
  if (col >= b.n_cols) return;
  const double v = state[(size_t)col * b.nz + detector_of(b.scols[col].detector_cell, b.nz)];
  if (to_outlet) b.outlet[(size_t)col * b.n_out] = v;
  if (b.series) b.series[(size_t)col * (b.steps + 1)] = v;
}

__global__ void k_stream_init(const BatchDev b, double* __restrict__ ping, unsigned long long* __restrict__ code) {
  const size_t n = (size_t)b.n_cols * b.nz;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) ping[i] = b.y0 ? b.y0[i] : 0.0;
  if (blockIdx.x == 0)
    for (uint32_t c = threadIdx.x; c < b.n_cols; c += blockDim.x) code[c] = ~0ull;
}

template <int ARITH>
__global__ void k_stream_coef(const BatchDev b, unsigned long long* __restrict__ code) {
  static_assert(sizeof(Coef<ARITH>) <= kCoefSlots * sizeof(double), "Coef does not fit its slot");
  const uint32_t col = blockIdx.x * blockDim.x + threadIdx.x;
  if (col >= b.n_cols) return;
  Coef<ARITH> q;
  q.init(b.scols[col]);
  *reinterpret_cast<Coef<ARITH>*>(reinterpret_cast<double*>(code + b.n_cols) + (size_t)col * kCoefSlots) = q;
}

typedef void (*skern_t)(const BatchDev, const double*, double*, uint32_t, unsigned long long*, int, int, uint32_t,
                        uint32_t);

template <int FUSE, int M>
skern_t pick_stream_f(int method, int arith) {
  if (method == CHROM_RK4) {
    if (arith == CHROM_ARITH_EXACT) return k_single_stream_step<CHROM_RK4, 1, FUSE, M>;
    return arith == CHROM_KARITH_POLY ? k_single_stream_step<CHROM_RK4, 2, FUSE, M> : k_single_stream_step<CHROM_RK4, 0, FUSE, M>;
  }
  if (arith == CHROM_ARITH_EXACT) return k_single_stream_step<CHROM_EULER, 1, FUSE, M>;
  return arith == CHROM_KARITH_POLY ? k_single_stream_step<CHROM_EULER, 2, FUSE, M> : k_single_stream_step<CHROM_EULER, 0, FUSE, M>;
}

template <int M>
skern_t pick_stream_m(int method, int arith, int fuse) {
  if (fuse == 1) return pick_stream_f<1, M>(method, arith);
  if (fuse == 2) return pick_stream_f<2, M>(method, arith);
  return fuse == 4 ? pick_stream_f<4, M>(method, arith) : pick_stream_f<8, M>(method, arith);
}

skern_t pick_stream(int method, int arith, int fuse, int m) {
  return m == 16 ? pick_stream_m<16>(method, arith, fuse) : pick_stream_m<8>(method, arith, fuse);
}

// Steps fused per launch: snapshots are taken from the state between launches, so the snapshot
// stride must be a multiple of the fusion depth.  CHROM_STREAM_FUSE overrides (experiments).
int stream_fuse(const BatchDev& b) {
  int f = 8;
  if (const char* e = getenv("CHROM_STREAM_FUSE")) f = atoi(e) >= 8 ? 8 : (atoi(e) >= 4 ? 4 : (atoi(e) >= 2 ? 2 : 1));
  while (f > 1 && b.snapshots && !b.h_snap_steps && b.snapshot_stride && (b.snapshot_stride % f) != 0) f /= 2;
  if (b.snapshots && b.h_snap_steps)  // an explicit list: every listed step must fall on a launch boundary
    for (uint32_t k = 0; k < b.n_snap; ++k)
      while (f > 1 && (b.h_snap_steps[k] % (uint32_t)f) != 0u && b.h_snap_steps[k] != b.steps) f /= 2;
  while (f > 1 && (uint32_t)f > b.steps) f /= 2;
  return f;
}

// Cells per lane and warps per CTA.  CHROM_STREAM_M / CHROM_STREAM_WARPS override (experiments).
int stream_m() {
  int m = 16;
  if (const char* e = getenv("CHROM_STREAM_M")) m = atoi(e) >= 16 ? 16 : 8;
  return m;
}
int stream_warps() {
  int w = kMaxWarps;
  if (const char* e = getenv("CHROM_STREAM_WARPS")) w = std::max(1, std::min(kMaxWarps, atoi(e)));
  return w;
}

}  // namespace

size_t single_stream_extra_doubles() { return 1 + kCoefSlots; }

bool single_stream_choose(const BatchDev& b, int sm_count, LaunchGeom* g, std::string* why) {
  (void)sm_count;
  (void)why;
  const int fuse = stream_fuse(b);
  const int kM = stream_m(), kWarps = stream_warps(), kTile = 32 * kM;
  const int H = b.method == CHROM_RK4 ? 4 * fuse : (fuse > 1 ? fuse : 0);
  const int V = kTile - H;
  const long long rest = (long long)b.nz - kTile;
  const int tiles = 1 + (rest > 0 ? (int)((rest + V - 1) / V) : 0);
  g->kind = 3;
  g->M = kM;
  g->L = 32;
  g->cpw = fuse;
  g->warps_per_col = tiles;
  g->block = dim3(kWarps * 32);
  g->grid = dim3((tiles + kWarps - 1) / kWarps, b.n_cols);
  g->smem = 0;
  g->launches_per_execute = ((uint64_t)b.steps + fuse - 1) / fuse + 4;  // + init, coefficients, sample 0, finalize
  char buf[240];
  snprintf(buf, sizeof buf,
           "single_stream<M=%d,%s,%s> tiles/col=%d halo=%d grid=(%u,%u) block=%d, %d fused step(s) per launch, CUDA graph",
           kM, b.method == CHROM_RK4 ? "RK4" : "Euler",
           (b.arith == CHROM_ARITH_EXACT ? "exact" : (b.arith == CHROM_KARITH_POLY ? "fast-poly" : "fast")), tiles, H,
           g->grid.x, g->grid.y, kWarps * 32, fuse);
  g->name = buf;
  return true;
}

// b.ping: [n_cols][nz]; b.pong: [n_cols][nz] followed by n_cols 64-bit status codes and n_cols
// coefficient slots (the caller allocates n_cols*(nz + 1 + single_stream_extra_doubles()) doubles).  A summary needs
// the full-rate detector series b.series [n_cols][steps + 1] (device scratch, allocated by the caller).

cudaError_t single_stream_launch(const BatchDev& b, const LaunchGeom& g, cudaStream_t s) {
  const int fuse = g.cpw > 0 ? g.cpw : 1;
  skern_t k = pick_stream(b.method, b.arith, fuse, g.M);
  unsigned long long* code = reinterpret_cast<unsigned long long*>(b.pong + (size_t)b.n_cols * b.nz);
  const size_t plane = (size_t)b.n_cols * b.nz;
  k_stream_init<<<296, 256, 0, s>>>(b, b.ping, code);
  {
    const dim3 cg((b.n_cols + 127) / 128), cb(128);
    if (b.arith == CHROM_ARITH_EXACT) k_stream_coef<1><<<cg, cb, 0, s>>>(b, code);
    else if (b.arith == CHROM_KARITH_POLY) k_stream_coef<2><<<cg, cb, 0, s>>>(b, code);
    else k_stream_coef<0><<<cg, cb, 0, s>>>(b, code);
  }
  double* src = b.ping;
  double* dst = b.pong;
  const dim3 cgrid((b.n_cols + 127) / 128), cblock(128);
  // sample schedules, walked on the host: which sub-steps of a launch sample the outlet (and into which slot), which
  // launch boundaries take a snapshot
  SampleCursor oc = sample_begin(b.h_out_steps, b.n_out, b.outlet ? b.outlet_stride : 0u, 0u);
  SampleCursor sc = sample_begin(b.h_snap_steps, b.n_snap, b.snapshots ? b.snapshot_stride : 0u, 0u);
  if (!b.outlet) oc.next = kNoSample;
  if (!b.snapshots) sc.next = kNoSample;
  const bool out0 = b.outlet && sample_has_initial(b.h_out_steps, b.outlet_stride);
  k_stream_sample0<<<cgrid, cblock, 0, s>>>(b, src, out0 ? 1 : 0);
  auto snapshot = [&](const double* state, uint32_t slot) -> cudaError_t {  // one strided copy for the whole batch
    return cudaMemcpy2DAsync(b.snapshots + (size_t)slot * b.nz, (size_t)b.n_snap * b.nz * sizeof(double), state,
                             (size_t)b.nz * sizeof(double), (size_t)b.nz * sizeof(double), b.n_cols,
                             cudaMemcpyDeviceToDevice, s);
  };
  if (b.snapshots && sample_has_initial(b.h_snap_steps, b.snapshot_stride)) {
    cudaError_t e = snapshot(src, 0);
    if (e != cudaSuccess) return e;
  }
  for (uint32_t step = 0; step < b.steps; step += (uint32_t)fuse) {
    const int nsub = (int)std::min<uint32_t>((uint32_t)fuse, b.steps - step);
    uint32_t out_mask = 0u, out_slot0 = 0u;  // sub-steps of this launch that sample the outlet, first slot
    for (int f = 0; f < nsub; ++f) {
      const uint32_t ds = step + (uint32_t)f + 1u;
      if (ds == oc.next) {
        if (!out_mask) out_slot0 = oc.k;
        out_mask |= 1u << f;
        sample_advance(oc, b.h_out_steps, b.outlet_stride);
      }
    }
    k<<<g.grid, g.block, 0, s>>>(b, src, dst, step, code, nsub, g.warps_per_col, out_mask, out_slot0);
    const uint32_t done = step + (uint32_t)nsub;
    while (sc.next != kNoSample && sc.next < done) sample_advance(sc, b.h_snap_steps, b.snapshot_stride);  // not on a boundary
    if (done == sc.next) {
      cudaError_t e = snapshot(dst, sc.k);
      if (e != cudaSuccess) return e;
      sample_advance(sc, b.h_snap_steps, b.snapshot_stride);
    }
    double* t = src;
    src = dst;
    dst = t;
  }
  if (b.final_state) {
    cudaError_t e = cudaMemcpyAsync(b.final_state, src, plane * sizeof(double), cudaMemcpyDeviceToDevice, s);
    if (e != cudaSuccess) return e;
  }
  k_stream_finalize<<<cgrid, cblock, 0, s>>>(b, code);
  return cudaGetLastError();
}

}  // namespace chrom
