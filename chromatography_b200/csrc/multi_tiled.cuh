// multi_tiled.cuh — LangmuirMulti beyond the shared-memory-resident capacity: columns whose
// 4 * n * nz doubles do not fit one CTA (long columns, many species) and species counts above 64.
//
// Replaces the same reference code as multi_kernels.cuh (rk4.rs:266-332, euler.rs:227-269,
// langmuir_multi.rs:627-686, 717-899, solver/mod.rs:514-553); only the mapping onto the machine differs:
//
//  * TILES.  A column is cut along z into tiles of `tw` cells; the state lives in HBM/L2 as two
//    ping-pong buffers [col][species][nz] and one launch advances every tile by `nsub` time steps with
//    all RK4 stages fused.  The upwind stencil reaches one cell upstream per stage, so a tile starts
//    H = 4 * nsub cells early (Euler: nsub) and discards its first H results — they are recomputed,
//    correctly, by the previous tile: no inter-CTA communication.  Tile 0 starts at the inlet and keeps
//    everything.  A column that fits ONE tile runs all its time steps in a single launch (this is how
//    n > 64 runs resident).  Algorithmic HBM traffic: 16 * n bytes per cell-step.
//  * THREAD PER CELL (WJ == 0, n <= 8): the per-cell solvers of multi_kernels.cuh, unchanged.
//  * WARP PER CELL (WJ > 0, n <= 32 * WJ, every n >= 9): a warp owns one cell, lane l holds the
//    species l, l + 32, ...  FAST: the elimination of A = diag(alpha) - p q^T collapses to scans —
//    eliminating species 0..i-1 leaves the pivot d_i = alpha_i - p_i q_i / h_i with
//    h_i = 1 - sum_{j<i} p_j q_j / alpha_j (an exclusive prefix sum over the warp), the partial-pivoting
//    test q_i max_{r>i}|p_r| > |alpha_i h_i - p_i q_i| needs a suffix maximum, and the solve is two warp
//    reductions; a cell that needs a row exchange (or has a zero pivot) is redone by the warp with a dense
//    row-exchanging LU on a scratch matrix, lanes over rows, pivot search by shuffle arg-max.
//    EXACT: nalgebra's lu::try_invert_to + gemv restated with lanes over rows / columns; every matrix
//    element sees the reference's operations in the reference's order, so the result is bit-identical.
#pragma once
#include "multi_kernels.cuh"

namespace chrom {
namespace {

constexpr unsigned kFull = 0xffffffffu;

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
  return v;  // xor butterfly: every lane ends with the same bits
}

// stage algebra of ONE species of one cell (apply_stage of multi_kernels.cuh, one entry)
template <bool EXACT>
__device__ __forceinline__ bool stage_one(const StageCtx& st, size_t idx, double k) {
  auto axpy = [](double y, double kk, double h) { return EXACT ? x_add(y, x_mul(kk, h)) : fma(kk, h, y); };
  if (st.stage == 0) {
    const double yn = axpy(st.Y[idx], k, st.dt);
    st.Y[idx] = yn;
    st.DST[idx] = yn;
    return hi_nonfinite(abs_hi(yn));
  } else if (st.stage == 1) {
    st.ACC[idx] = k;
    st.DST[idx] = axpy(st.Y[idx], k, st.h2);
  } else if (st.stage == 2) {
    st.ACC[idx] = axpy(st.ACC[idx], k, 2.0);
    st.DST[idx] = axpy(st.Y[idx], k, st.h2);
  } else if (st.stage == 3) {
    st.ACC[idx] = axpy(st.ACC[idx], k, 2.0);
    st.DST[idx] = axpy(st.Y[idx], k, st.dt);
  } else {
    const double w = EXACT ? x_add(st.ACC[idx], k) : st.ACC[idx] + k;
    const double yn = axpy(st.Y[idx], w, st.h6);
    st.Y[idx] = yn;
    return hi_nonfinite(abs_hi(yn));
  }
  return false;
}

// ---- warp per cell, FAST: dense LU with partial pivoting on a scratch matrix ------------------------
// A: n*n doubles column-major (lanes over rows: coalesced) followed by n doubles (right-hand side ->
// solution).  Reached only when the scan form reports that a row exchange is needed.
__device__ __noinline__ void warp_dense_fast(int n, const SpeciesConst* sc, double cg, const double* SRC,
                                             const double* inlet_row, uint32_t pitch, uint32_t cc, double* A,
                                             int lane) {
  double* vec = A + (size_t)n * n;
  double S = 0.0;
  for (int s = lane; s < n; s += 32) S = fma(sc[s].K, SRC[(size_t)s * pitch + cc], S);
  S = warp_sum(S);
  const double D = 1.0 + S, rD2 = 1.0 / (D * D);
  for (int r = lane; r < n; r += 32) {
    const double c = SRC[(size_t)r * pitch + cc];
    const double up = (cc > 0) ? SRC[(size_t)r * pitch + cc - 1] : inlet_row[r];
    const double p = sc[r].feNK * c * rD2;
    for (int q = 0; q < n; ++q) A[r + (size_t)q * n] = -p * sc[q].K;
    A[r + (size_t)r * n] = fma(sc[r].feNK, D * rD2, sc[r].a0) + A[r + (size_t)r * n];
    vec[r] = (c - up) * cg;
  }
  __syncwarp();
  bool singular = false;
  for (int i = 0; i < n; ++i) {
    double best = -1.0;
    int bi = n;
    for (int r = i + lane; r < n; r += 32) {
      const double v = fabs(A[r + (size_t)i * n]);
      if (v > best) {
        best = v;
        bi = r;
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const double ob = __shfl_xor_sync(kFull, best, o);
      const int oi = __shfl_xor_sync(kFull, bi, o);
      if (ob > best || (ob == best && oi < bi)) {
        best = ob;
        bi = oi;
      }
    }
    const int piv = (bi < n) ? bi : i;
    if (piv != i) {
      for (int q = i + lane; q < n; q += 32) {
        const double t = A[i + (size_t)q * n];
        A[i + (size_t)q * n] = A[piv + (size_t)q * n];
        A[piv + (size_t)q * n] = t;
      }
      if (lane == 0) {
        const double t = vec[i];
        vec[i] = vec[piv];
        vec[piv] = t;
      }
    }
    __syncwarp();
    const double d = A[i + (size_t)i * n];
    singular |= (d == 0.0);
    const double rp = 1.0 / d, bi_v = vec[i];
    for (int r = i + 1 + lane; r < n; r += 32) {
      const double l = A[r + (size_t)i * n] * rp;
      for (int q = i + 1; q < n; ++q) A[r + (size_t)q * n] = fma(-l, A[i + (size_t)q * n], A[r + (size_t)q * n]);
      vec[r] = fma(-l, bi_v, vec[r]);
    }
    __syncwarp();
  }
  for (int i = n - 1; i >= 0; --i) {  // column-oriented back substitution
    const double x = vec[i] / A[i + (size_t)i * n];
    __syncwarp();
    if (lane == 0) vec[i] = x;
    for (int r = lane; r < i; r += 32) vec[r] = fma(-A[r + (size_t)i * n], x, vec[r]);
    __syncwarp();
  }
  if (singular) {  // exact zero pivot: the reference falls back to the identity (langmuir_multi.rs:790-796)
    for (int r = lane; r < n; r += 32) {
      const double c = SRC[(size_t)r * pitch + cc];
      const double up = (cc > 0) ? SRC[(size_t)r * pitch + cc - 1] : inlet_row[r];
      vec[r] = (c - up) * cg;
    }
    __syncwarp();
  }
}

// ---- warp per cell, FAST: the structured elimination in closed form ---------------------------------------
// Lane `lane` holds species lane + 32*j, j < WJ.  K/feNK/a0 are the lane's species constants (cached in
// registers for the whole tile); kmax = max_i K_i and a0min = min_i (1 + fe lambda_i) of the column.  C cells
// are solved side by side: their shuffle chains are independent, so interleaving them hides the shuffle /
// reciprocal latency a single cell would expose.
// Solution: x = z + v (q.z)/(1 - q.v), z = b/alpha, v = p/alpha — three butterfly reductions (S, q.v, q.z).
// Partial pivoting: a sufficient bound, 2 max q max|p| < min alpha (1 - q.v) with every c >= 0 (one integer
// REDUX over the high words), rules a row exchange out for the whole cell; only a cell that fails it runs the
// per-step test (exclusive prefix sums h_i and suffix maxima of |p| by warp scans).
// Returns the (warp-uniform) mask of cells where partial pivoting would exchange rows (their k[] is invalid).
template <int WJ, int C>
__device__ __forceinline__ unsigned warp_cells_fast(int n, int lane, const double (&K)[WJ], const double (&feNK)[WJ],
                                                    const double (&a0)[WJ], double kmax, double a0min, double cg,
                                                    const double (&c)[C][WJ], const double (&up)[C][WJ],
                                                    double (&k)[C][WJ]) {
  double S[C];
#pragma unroll
  for (int q = 0; q < C; ++q) {
    S[q] = 0.0;
#pragma unroll
    for (int j = 0; j < WJ; ++j) S[q] = fma(K[j], c[q][j], S[q]);  // absent species carry K = 0
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1)
#pragma unroll
    for (int q = 0; q < C; ++q) S[q] += __shfl_xor_sync(kFull, S[q], o);
  double rD[C], rD2[C], qv[C], qz[C];
  double p[C][WJ], z[C][WJ], v[C][WJ];
  unsigned phi[C];  // max over the high words of feNK c (unsigned: a negative or non-finite entry reads as huge)
#pragma unroll
  for (int q = 0; q < C; ++q) {
    rD[q] = fast_rcp(1.0 + S[q]);
    rD2[q] = rD[q] * rD[q];
    qv[q] = 0.0;
    qz[q] = 0.0;
    phi[q] = 0u;
#pragma unroll
    for (int j = 0; j < WJ; ++j) {
      const double pw = feNK[j] * c[q][j];
      phi[q] = max(phi[q], (unsigned)__double2hiint(pw));
      p[q][j] = pw * rD2[q];                                  // fe nbar K c / D^2
      const double ra = fast_rcp(fma(feNK[j], rD[q], a0[j]));  // 1 / alpha, alpha = 1 + fe (lambda + nbar K / D)
      z[q][j] = (c[q][j] - up[q][j]) * cg * ra;
      v[q][j] = p[q][j] * ra;
      qv[q] = fma(K[j], v[q][j], qv[q]);
      qz[q] = fma(K[j], z[q][j], qz[q]);
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1)
#pragma unroll
    for (int q = 0; q < C; ++q) {
      qv[q] += __shfl_xor_sync(kFull, qv[q], o);
      qz[q] += __shfl_xor_sync(kFull, qz[q], o);
    }
  unsigned suspect = 0u;  // warp-uniform
#pragma unroll
  for (int q = 0; q < C; ++q) {
    const double hn = 1.0 - qv[q];
    const double f = qz[q] * fast_rcp(hn);  // (q.z) / (1 - q.v)
#pragma unroll
    for (int j = 0; j < WJ; ++j) k[q][j] = fma(v[q][j], f, z[q][j]);
    const unsigned pmax_hi = __reduce_max_sync(kFull, phi[q]);
    const double pmax = __hiloint2double((int)(pmax_hi + 1u), 0) * rD2[q];  // upper bound of max_r |p_r|
    if (!(pmax_hi < 0x7ff00000u && (2.0 * kmax) * pmax < a0min * hn)) suspect |= 1u << q;
  }
  if (suspect == 0u) return 0u;

  // ---- rare: the per-step partial-pivoting test, q_i * max_{r>i}|p_r| > |alpha_i h_i - p_i q_i| (h_i > 0) ----
  unsigned redo = 0u;
#pragma unroll
  for (int q = 0; q < C; ++q) {
    if (!((suspect >> q) & 1u)) continue;  // warp-uniform
    double h[WJ], run = 0.0;
#pragma unroll
    for (int j = 0; j < WJ; ++j) {  // h_i = 1 - sum_{r<i} p_r q_r / alpha_r: inclusive scan per block of 32 species
      const double t = K[j] * v[q][j];
      double incl = t;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const double u = __shfl_up_sync(kFull, incl, o);
        if (lane >= o) incl += u;
      }
      double excl = __shfl_up_sync(kFull, incl, 1);
      if (lane == 0) excl = 0.0;
      h[j] = 1.0 - (run + excl);
      run += __shfl_sync(kFull, incl, 31);
    }
    double runmax = 0.0;
#pragma unroll
    for (int j = WJ - 1; j >= 0; --j) {
      double sfx = fabs(p[q][j]);  // inclusive suffix maximum inside block j
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const double u = __shfl_down_sync(kFull, sfx, o);
        if (lane + o < 32) sfx = fmax(sfx, u);
      }
      double excl = __shfl_down_sync(kFull, sfx, 1);
      if (lane == 31) excl = 0.0;
      const double below = fmax(runmax, excl);
      runmax = fmax(runmax, __shfl_sync(kFull, sfx, 0));
      if (lane + 32 * j < n) {
        const double alpha = fma(feNK[j], rD[q], a0[j]);
        const double dpiv = fabs(fma(alpha, h[j], -(p[q][j] * K[j])));
        if (!(h[j] > 0.0) || (K[j] * below > dpiv) || (dpiv == 0.0)) redo |= 1u << q;
      }
    }
  }
  return __reduce_or_sync(kFull, redo);
}

// ---- warp per cell, EXACT: compute_row with nalgebra's LU inverse and gemv -----------------------------
// scratch: A (the matrix, then its LU) column-major n*n, O (inverse) ROW-major n*n.  Lanes run over rows in
// the elimination and over columns of O in the triangular solves; each element sees the operations of
// exact_row / the oracle in the same order.
template <int WJ>
__device__ __noinline__ void warp_cell_exact(int n, int lane, const SpeciesConst* sc, double fe, double ue, double dz,
                                             const double* SRC, const double* inlet_row, uint32_t pitch, uint32_t cc,
                                             double* scratch, double (&k)[WJ]) {
  const size_t nn = (size_t)n * n;
  double* A = scratch;          // LU, column-major
  double* O = scratch + nn;     // inverse, row-major
  auto cval = [&](int s) { return SRC[(size_t)s * pitch + cc]; };
  auto upval = [&](int s) { return (cc > 0) ? SRC[(size_t)s * pitch + cc - 1] : inlet_row[s]; };
  double sum_kc = 0.0;
  for (int s = 0; s < n; ++s) sum_kc = x_add(sum_kc, x_mul(sc[s].K, cval(s)));  // sequential, as the reference
  const double denom = x_add(1.0, sum_kc);
  const double denom2 = x_mul(denom, denom);
  for (int i = lane; i < n; i += 32) {
    const double nb = sc[i].nbar, Ki = sc[i].K, ci = cval(i);
    for (int j = 0; j < n; ++j) {
      double J;
      if (i == j) {
        const double num = x_mul(x_mul(nb, Ki), x_sub(denom, x_mul(Ki, ci)));
        J = x_add(sc[i].lambda, x_div(num, denom2));
      } else {
        const double num = x_mul(x_mul(x_mul(-nb, Ki), sc[j].K), ci);
        J = x_div(num, denom2);
      }
      A[i + (size_t)j * n] = x_add((i == j) ? 1.0 : 0.0, x_mul(fe, J));
      O[(size_t)i * n + j] = (i == j) ? 1.0 : 0.0;
    }
  }
  __syncwarp();
  bool ok = true;
  for (int i = 0; i < n; ++i) {
    // icamax: first row holding the largest |entry| (strict '>' in ascending order; NaN never wins,
    // a NaN on the diagonal keeps the diagonal)
    const double d0 = fabs(A[i + (size_t)i * n]);
    double best = -1.0;
    int bi = n;
    for (int r = i + lane; r < n; r += 32) {
      const double v = fabs(A[r + (size_t)i * n]);
      if (v > best) {
        best = v;
        bi = r;
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const double ob = __shfl_xor_sync(kFull, best, o);
      const int oi = __shfl_xor_sync(kFull, bi, o);
      if (ob > best || (ob == best && oi < bi)) {
        best = ob;
        bi = oi;
      }
    }
    const int piv = (d0 != d0 || bi >= n) ? i : bi;
    const double diag = A[piv + (size_t)i * n];
    if (diag == 0.0) {
      ok = false;
      break;
    }
    if (piv != i) {
      for (int j = lane; j < n; j += 32) {
        const double t = O[(size_t)i * n + j];
        O[(size_t)i * n + j] = O[(size_t)piv * n + j];
        O[(size_t)piv * n + j] = t;
        const double u = A[i + (size_t)j * n];
        A[i + (size_t)j * n] = A[piv + (size_t)j * n];
        A[piv + (size_t)j * n] = u;
      }
    }
    __syncwarp();
    const double inv_diag = x_div(1.0, diag);
    for (int r = i + 1 + lane; r < n; r += 32) {
      const double l = x_mul(A[r + (size_t)i * n], inv_diag);
      A[r + (size_t)i * n] = l;
      for (int q = i + 1; q < n; ++q) {
        const double neg_pivot = -A[i + (size_t)q * n];
        A[r + (size_t)q * n] = x_add(x_mul(neg_pivot, l), A[r + (size_t)q * n]);
      }
    }
    __syncwarp();
  }
  if (ok)
    for (int i = 0; i < n; ++i) ok &= (A[i + (size_t)i * n] != 0.0);
  if (ok) {
    for (int col = lane; col < n; col += 32) {  // each lane solves whole columns of the inverse
      for (int i = 0; i + 1 < n; ++i) {
        const double coeff = O[(size_t)i * n + col];
        for (int r = i + 1; r < n; ++r)
          O[(size_t)r * n + col] = x_add(x_mul(-coeff, A[r + (size_t)i * n]), O[(size_t)r * n + col]);
      }
      for (int i = n - 1; i >= 0; --i) {
        const double coeff = x_div(O[(size_t)i * n + col], A[i + (size_t)i * n]);
        O[(size_t)i * n + col] = coeff;
        for (int r = 0; r < i; ++r)
          O[(size_t)r * n + col] = x_add(x_mul(-coeff, A[r + (size_t)i * n]), O[(size_t)r * n + col]);
      }
    }
  } else {  // identity fallback, langmuir_multi.rs:790-796
    for (int i = lane; i < n; i += 32)
      for (int j = 0; j < n; ++j) O[(size_t)i * n + j] = (i == j) ? 1.0 : 0.0;
  }
  __syncwarp();
  // grad, gemv (column by column, j ascending), scale
#pragma unroll
  for (int jj = 0; jj < WJ; ++jj) {
    const int i = lane + 32 * jj;
    double acc = 0.0;
    if (i < n) {
      acc = x_mul(O[(size_t)i * n], x_div(x_sub(cval(0), upval(0)), dz));
      for (int j = 1; j < n; ++j) acc = x_add(x_mul(O[(size_t)i * n + j], x_div(x_sub(cval(j), upval(j)), dz)), acc);
      acc = x_mul(-ue, acc);
    }
    k[jj] = acc;
  }
  __syncwarp();  // the scratch matrices are reused by the warp's next cell
}

// Visits every (species, cell) pair of a tile's cells [j0, j1) with the whole CTA.  A wide tile (thread per cell)
// walks species in the outer loop: coalesced, no index arithmetic.  A narrow tile (warp per cell: a few cells, up
// to 256 species) would leave most of the CTA idle that way, so it runs over the flattened index instead.
template <class F>
__device__ __forceinline__ void for_tile(int n, uint32_t j0, uint32_t j1, int tid, int T, F f) {
  const uint32_t w = j1 - j0;
  if (w >= (uint32_t)T) {
    for (int s = 0; s < n; ++s)
      for (uint32_t j = j0 + tid; j < j1; j += T) f((uint32_t)s, j);
  } else {
    for (uint32_t i = tid; i < (uint32_t)n * w; i += T) {
      const uint32_t s = i / w;
      f(s, j0 + (i - s * w));
    }
  }
}

// ---- kernel -----------------------------------------------------------------------------------------------
// threads per CTA of the warp-per-cell variants: 16 warps where the lane state is small (FAST, <= 2 species per lane)
constexpr int tiled_warp_threads(int WJ, int ARITH) { return (ARITH != CHROM_ARITH_EXACT && WJ <= 2) ? 512 : 256; }

template <int N, int METHOD, int ARITH, int WJ>
__global__ void __launch_bounds__(WJ > 0 ? tiled_warp_threads(WJ, ARITH) : multi_max_threads(N, ARITH), 1)
    k_multi_tiled(const BatchDev b, double* scratch, const TiledArgs a) {
  constexpr int ST = (METHOD == CHROM_RK4) ? 3 : 1;
  constexpr bool EXACT = (ARITH == CHROM_ARITH_EXACT);
  constexpr int WJC = WJ > 0 ? WJ : 1;
  extern __shared__ double smem[];
  const int n = N > 0 ? N : (int)b.n_species;
  const uint32_t nz = b.nz, pitch = a.pitch;
  const int T = blockDim.x;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, W = T >> 5;
  const size_t plane = (size_t)n * pitch;
  double* Y = smem;
  double* ACC = Y + plane;
  double* U0 = ACC + plane;
  double* U1 = U0 + plane;
  double* inlet = U1 + plane;                        // [3][n]
  double* summ = inlet + 3 * n;                               // [n] Summ (kSummSlots doubles each)
  SpeciesConst* sc = (SpeciesConst*)(summ + kSummSlots * n);  // [n]
  const double** tabs = (const double**)(sc + n);    // [n]
  double* warp_scratch = nullptr;                    // warp-per-cell scratch: a contiguous block per warp
  if (WJ > 0 && scratch)
    warp_scratch = scratch + ((size_t)blockIdx.x * W + warp) * ((size_t)(EXACT ? 2 : 1) * n * n + n);

  const double dt = b.dt, h2 = b.half_dt, h6 = b.sixth_dt;
  const uint32_t V = a.tw - a.halo;
  const uint32_t items = b.n_cols * a.tiles_per_col;

  for (uint32_t item = blockIdx.x; item < items; item += gridDim.x) {  // persistent CTAs over (column, tile)
    const uint32_t col = item / a.tiles_per_col, tile = item - col * a.tiles_per_col;
    const uint32_t base = tile * V;                       // tile 0: [0, tw); tile t: [t*V, t*V + tw), first `halo` discarded
    const uint32_t ncell = min(a.tw, nz - base);
    const uint32_t first_valid = tile ? a.halo : 0u;
    const uint32_t det = detector_of(b.mcols[col].detector_cell, nz);  // the detector cell, owned by exactly one tile
    const bool owns_outlet = (det >= base + first_valid) && (det < base + ncell);
    const uint32_t jd = det - base;  // its tile-local index (meaningful when owns_outlet)
    const chrom_multi_col cp = b.mcols[col];
    const double fe = cp.fe, ue = cp.ue, dz = cp.dz;
    const double cg = -ue / dz;
    __syncthreads();  // the previous item's readers of shared memory are done
    for (int s = tid; s < n; s += T) {
      const chrom_species sp = b.species[cp.species_off + s];
      SpeciesConst q;
      q.lambda = sp.lambda;
      q.K = sp.langmuir_k;
      q.nbar = EXACT ? x_mul(cp.stationary_fraction, (double)sp.port_number)
                     : cp.stationary_fraction * (double)sp.port_number;
      q.a0 = 1.0 + fe * sp.lambda;
      q.feNK = fe * (q.nbar * sp.langmuir_k);
      q.feNK2 = q.feNK;
      q.a0rK = q.a0 / sp.langmuir_k;
      q.feN = fe * q.nbar;
      q.cgrK = cg / sp.langmuir_k;
      q.pad_ = 0.0;
      sc[s] = q;
      tabs[s] = b.tables + (size_t)sp.inj_table * b.steps * ST;
    }
    const double* gsrc = a.src + (size_t)col * n * nz + base;
    for_tile(n, 0u, ncell, tid, T, [&](uint32_t s, uint32_t j) {
      const double v = gsrc[(size_t)s * nz + j];
      Y[(size_t)s * pitch + j] = v;
      if (METHOD == CHROM_EULER) U0[(size_t)s * pitch + j] = v;
    });
    __syncthreads();
    if (owns_outlet)
      for (int s = tid; s < n; s += T) {
        Summ& sm = reinterpret_cast<Summ*>(summ)[s];
        if (a.step0 == 0) summ_init(sm, Y[(size_t)s * pitch + jd]);
        else sm = reinterpret_cast<const Summ*>(a.gsumm)[(size_t)col * n + s];
      }
    SampleCursor oc = sample_begin(b.out_steps, b.n_out, b.outlet ? b.outlet_stride : 0u, a.step0);
    SampleCursor sc_ = sample_begin(b.snap_steps, b.n_snap, b.snapshots ? b.snapshot_stride : 0u, a.step0);
    if (!b.outlet) oc.next = kNoSample;
    if (!b.snapshots) sc_.next = kNoSample;
    // warp per cell: the lane's species constants stay in registers for the whole tile
    double wK[WJC], wF[WJC], wA[WJC];
    double w_kmax = 0.0, w_a0min = 0.0;  // max_i K_i and min_i (1 + fe lambda_i) of the column (pivot bound)
    if (WJ > 0) {
      unsigned kh = 0u, ah = 0xffffffffu;
#pragma unroll
      for (int j = 0; j < WJC; ++j) {
        const int s = lane + 32 * j;
        wK[j] = (s < n) ? sc[s].K : 0.0;
        wF[j] = (s < n) ? sc[s].feNK : 0.0;
        wA[j] = (s < n) ? sc[s].a0 : 1.0;
        if (s < n) {  // positive constants: the high words order like the doubles
          kh = max(kh, (unsigned)__double2hiint(wK[j]));
          ah = min(ah, (unsigned)__double2hiint(wA[j]));
        }
      }
      kh = __reduce_max_sync(kFull, kh);
      ah = __reduce_min_sync(kFull, ah);
      w_kmax = __hiloint2double((int)(kh + 1u), 0);  // rounded up / down: the bound stays sufficient
      w_a0min = __hiloint2double((int)ah, 0);
    }

    // One stage over the tile.  Returns the non-finite flag of the cells this tile owns.
    auto run_stage = [&](const double* SRC, double* DST, int irow, int stage) -> bool {
      bool bad = false;
      StageCtx st{Y, ACC, DST, pitch, METHOD == CHROM_EULER ? 0 : stage, dt, h2, h6};
      const double* inlet_row = inlet + irow * n;
      if constexpr (WJ > 0) {
        if constexpr (EXACT) {
          for (uint32_t cc = (uint32_t)warp; cc < ncell; cc += (uint32_t)W) {
            double k[WJC];
            warp_cell_exact<WJC>(n, lane, sc, fe, ue, dz, SRC, inlet_row, pitch, cc, warp_scratch, k);
            bool cell_bad = false;
#pragma unroll
            for (int j = 0; j < WJC; ++j) {
              const int s = lane + 32 * j;
              if (s < n) cell_bad |= stage_one<true>(st, (size_t)s * pitch + cc, k[j]);
            }
            bad |= cell_bad && (cc >= first_valid);
          }
        } else {
          constexpr int C = WJC == 1 ? 4 : (WJC == 2 ? 2 : 1);  // cells solved side by side per warp
          for (uint32_t cc0 = (uint32_t)(warp * C); cc0 < ncell; cc0 += (uint32_t)(W * C)) {
            double c[C][WJC], up[C][WJC], k[C][WJC];
#pragma unroll
            for (int q = 0; q < C; ++q) {
              const uint32_t cc = min(cc0 + (uint32_t)q, ncell - 1u);  // past the tile: recomputed, never stored
#pragma unroll
              for (int j = 0; j < WJC; ++j) {
                const int s = lane + 32 * j;
                c[q][j] = (s < n) ? SRC[(size_t)s * pitch + cc] : 0.0;
                up[q][j] = (s < n) ? ((cc > 0) ? SRC[(size_t)s * pitch + cc - 1] : inlet_row[s]) : 0.0;
              }
            }
            const unsigned redo = warp_cells_fast<WJC, C>(n, lane, wK, wF, wA, w_kmax, w_a0min, cg, c, up, k);
#pragma unroll
            for (int q = 0; q < C; ++q) {
              const uint32_t cc = cc0 + (uint32_t)q;
              if (cc < ncell) {
                if ((redo >> q) & 1u) {  // warp-uniform: redo this cell with the dense row-exchanging LU
                  warp_dense_fast(n, sc, cg, SRC, inlet_row, pitch, cc, warp_scratch, lane);
                  const double* vec = warp_scratch + (size_t)n * n;
#pragma unroll
                  for (int j = 0; j < WJC; ++j) k[q][j] = (lane + 32 * j < n) ? vec[lane + 32 * j] : 0.0;
                  __syncwarp();
                }
                bool cell_bad = false;
#pragma unroll
                for (int j = 0; j < WJC; ++j) {
                  const int s = lane + 32 * j;
                  if (s < n) cell_bad |= stage_one<false>(st, (size_t)s * pitch + cc, k[q][j]);
                }
                bad |= cell_bad && (cc >= first_valid);
              }
            }
          }
        }
      } else if constexpr (N > 0) {  // thread per cell: the n x n system of a cell in one thread's registers
        const int cpt = (int)((a.tw + T - 1) / T);
        unsigned redo_mask = 0u;
        for (int ci = 0; ci < cpt; ++ci) {
          const uint32_t cell = (uint32_t)tid + (uint32_t)ci * T;
          const bool live = cell < ncell;
          const uint32_t cc = live ? cell : ncell - 1;  // idle lanes recompute the last cell, never store
          const bool owned = live && cc >= first_valid;
          double k[N];
          if constexpr (ARITH == CHROM_KARITH_STRUCT) {
            const bool redo = scan_cell<N>(sc, cg, SRC, inlet_row, pitch, cc, k);
            if (redo) redo_mask |= 1u << ci;
            else if (live) bad |= apply_stage<N, false>(st, cc, k, n) && owned;
          } else {
            double c[N], up[N];
#pragma unroll
            for (int s = 0; s < N; ++s) {
              c[s] = SRC[(size_t)s * pitch + cc];
              up[s] = (cc > 0) ? SRC[(size_t)s * pitch + cc - 1] : inlet_row[s];
            }
            if constexpr (EXACT) {
              exact_row<N>(sc, fe, ue, dz, c, up, k);
            } else {
              const bool redo = fast_row_spec<N>(sc, c, up, k);
              if (__any_sync(kFull, redo)) {
                bad |= cold_cell<N>(sc, fe, cg, SRC, inlet_row, cc, live, st) && owned;
                continue;
              }
            }
            if (live) bad |= apply_stage<N, EXACT>(st, cc, k, n) && owned;
          }
        }
        if constexpr (ARITH == CHROM_KARITH_STRUCT) {
          if (__any_sync(kFull, redo_mask != 0u)) {
            for (int ci = 0; ci < cpt; ++ci) {
              const bool mine = (redo_mask >> ci) & 1u;
              if (!__any_sync(kFull, mine)) continue;
              const uint32_t cell = (uint32_t)tid + (uint32_t)ci * T;
              const bool live = cell < ncell;
              const uint32_t cc = live ? cell : ncell - 1;
              bad |= cold_cell<N>(sc, fe, cg, SRC, inlet_row, cc, mine && live, st) && (live && cc >= first_valid);
            }
          }
        }
      }
      return bad;
    };

    bool flagged = false;
    for (uint32_t f = 0; f < a.nsub; ++f) {
      const uint32_t step = a.step0 + f;
      for (int i = tid; i < n * ST; i += T) {
        const int s = i % n, r = i / n;
        inlet[r * n + s] = tabs[s][(size_t)step * ST + r];
      }
      __syncthreads();
      bool bad;
      if (METHOD == CHROM_RK4) {
        run_stage(Y, U0, 0, 1);
        __syncthreads();
        run_stage(U0, U1, 1, 2);
        __syncthreads();
        run_stage(U1, U0, 1, 3);
        __syncthreads();
        bad = run_stage(U0, nullptr, 2, 4);
      } else {
        bad = (f & 1u) ? run_stage(U1, U0, 0, 0) : run_stage(U0, U1, 0, 0);
      }
      const int any_bad = __syncthreads_or(bad);

      // ---- samples (cells this tile owns) ----------------------------------------------------------
      const uint32_t done = step + 1u;
      const bool out_now = (done == oc.next);
      if (owns_outlet && (out_now || b.summary)) {
        for (int s = tid; s < n; s += T) {
          const double v = Y[(size_t)s * pitch + jd];
          if (out_now) b.outlet[((size_t)col * n + s) * b.n_out + oc.k] = v;
          if (b.summary) summ_step(reinterpret_cast<Summ*>(summ)[s], v, step, dt);
        }
      }
      if (out_now) sample_advance(oc, b.out_steps, b.outlet_stride);
      if (done == sc_.next) {
        double* gdst = b.snapshots + ((size_t)col * b.n_snap + sc_.k) * n * nz + base;
        for_tile(n, first_valid, ncell, tid, T,
                 [&](uint32_t s, uint32_t j) { gdst[(size_t)s * nz + j] = Y[(size_t)s * pitch + j]; });
        sample_advance(sc_, b.snap_steps, b.snapshot_stride);
      }
      if (any_bad && !flagged) {  // validate_state: the first bad step wins, NaN over Inf within it
        bool has_nan = false;
        for_tile(n, first_valid, ncell, tid, T, [&](uint32_t s, uint32_t j) {
          const double v = Y[(size_t)s * pitch + j];
          has_nan |= (v != v);
        });
        const int any_nan = __syncthreads_or(has_nan);
        if (tid == 0) atomicMin(&a.code[col], (((unsigned long long)done) << 1) | (any_nan ? 0ull : 1ull));
        flagged = true;
      }
    }
    __syncthreads();
    double* gdst = a.dst + (size_t)col * n * nz + base;
    for_tile(n, first_valid, ncell, tid, T,
             [&](uint32_t s, uint32_t j) { gdst[(size_t)s * nz + j] = Y[(size_t)s * pitch + j]; });
    if (owns_outlet && b.summary)
      for (int s = tid; s < n; s += T)
        reinterpret_cast<Summ*>(a.gsumm)[(size_t)col * n + s] = reinterpret_cast<const Summ*>(summ)[s];
  }
}

// Initial state into the first buffer, clean status codes, sample slot 0 of outlet / snapshots.
__global__ void k_tiled_init(const BatchDev b, double* __restrict__ ping, unsigned long long* __restrict__ code) {
  const size_t plane = (size_t)b.n_species * b.nz;
  const size_t total = (size_t)b.n_cols * plane;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const double v = b.y0 ? b.y0[i] : 0.0;
    ping[i] = v;
    const size_t col = i / plane, r = i - col * plane;
    if (b.snapshots && sample_has_initial(b.snap_steps, b.snapshot_stride)) b.snapshots[col * b.n_snap * plane + r] = v;
    if (b.outlet && sample_has_initial(b.out_steps, b.outlet_stride) &&
        (r % b.nz) == detector_of(b.mcols[col].detector_cell, b.nz))
      b.outlet[(col * b.n_species + r / b.nz) * b.n_out] = v;
  }
  for (size_t c = (size_t)blockIdx.x * blockDim.x + threadIdx.x; c < b.n_cols; c += stride) code[c] = ~0ull;
}

// After the last launch: status codes -> the ABI's signed convention; running summaries -> summary[].
__global__ void k_tiled_finalize(const BatchDev b, const unsigned long long* __restrict__ code,
                                 const double* __restrict__ gsumm) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < b.n_cols && b.status) {
    const unsigned long long c = code[i];
    const long long s = (long long)(c >> 1);
    b.status[i] = (c == ~0ull) ? 0ll : ((c & 1ull) ? -s : s);
  }
  if (i < b.n_cols * b.n_species && b.summary) {
    summ_store(b.summary + (size_t)i * CHROM_SUMMARY_LEN, reinterpret_cast<const Summ*>(gsumm)[i]);
  }
}

typedef void (*tkern_t)(const BatchDev, double*, const TiledArgs);

// thread per cell (wj == 0) for N = 1..8 (every mode) and 9..16 (closed form only); warp per cell (wj = 1, 2, 4,
// 8) in the runtime-n object (N == 0)
template <int N>
tkern_t pick_tiled_n(int method, int arith, int wj) {
  if (wj == 0) {
    if constexpr (N >= 1 && N <= 8) {
      if (method == CHROM_RK4) {
        if (arith == CHROM_ARITH_EXACT) return k_multi_tiled<N, CHROM_RK4, CHROM_ARITH_EXACT, 0>;
        if (arith == CHROM_KARITH_STRUCT) return k_multi_tiled<N, CHROM_RK4, CHROM_KARITH_STRUCT, 0>;
        return k_multi_tiled<N, CHROM_RK4, CHROM_ARITH_FAST, 0>;
      }
      if (arith == CHROM_ARITH_EXACT) return k_multi_tiled<N, CHROM_EULER, CHROM_ARITH_EXACT, 0>;
      if (arith == CHROM_KARITH_STRUCT) return k_multi_tiled<N, CHROM_EULER, CHROM_KARITH_STRUCT, 0>;
      return k_multi_tiled<N, CHROM_EULER, CHROM_ARITH_FAST, 0>;
    } else if constexpr (N >= 9 && N <= 16) {
      if (arith != CHROM_KARITH_STRUCT) return nullptr;
      return method == CHROM_RK4 ? k_multi_tiled<N, CHROM_RK4, CHROM_KARITH_STRUCT, 0>
                                 : k_multi_tiled<N, CHROM_EULER, CHROM_KARITH_STRUCT, 0>;
    }
    return nullptr;
  }
  if constexpr (N == 0) {
    const bool ex = arith == CHROM_ARITH_EXACT;
#define CHROM_WPC(J)                                                                                          \
  if (wj == J) {                                                                                              \
    if (method == CHROM_RK4)                                                                                  \
      return ex ? k_multi_tiled<0, CHROM_RK4, CHROM_ARITH_EXACT, J> : k_multi_tiled<0, CHROM_RK4, CHROM_KARITH_STRUCT, J>; \
    return ex ? k_multi_tiled<0, CHROM_EULER, CHROM_ARITH_EXACT, J> : k_multi_tiled<0, CHROM_EULER, CHROM_KARITH_STRUCT, J>; \
  }
    CHROM_WPC(1)
    CHROM_WPC(2)
    CHROM_WPC(4)
    CHROM_WPC(8)
#undef CHROM_WPC
  }
  return nullptr;
}

}  // namespace
}  // namespace chrom
