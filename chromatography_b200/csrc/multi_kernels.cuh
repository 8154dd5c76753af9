// multi_kernels.cuh — device code of the shared-memory-resident LangmuirMulti time integrator (RK4 / Euler).
//
// Replaces, for a batch of columns and ALL time steps in one launch:
//   RK4Solver::solve / EulerSolver::solve        src/solver/methods/rk4.rs:266-332, euler.rs:227-269
//   LangmuirMulti::compute_physics / compute_row src/models/langmuir_multi.rs:717-899 (770-840)
//   LangmuirMulti::jacobian                      src/models/langmuir_multi.rs:627-668
//   inverse_propagation + `inv * grad`           src/models/langmuir_multi.rs:674-686, 838 (nalgebra 0.34)
//   validate_state                               src/solver/mod.rs:514-553
//
// Layout: one CTA owns one column.  Its state lives in shared memory as [array][species][nz]
// (species-major, the reference's nalgebra column-major nz x n matrix): y (state at t_n), acc
// (running weighted slope) and two stage buffers u0/u1 (ping-pong, so a cell may read its upstream
// neighbour while that neighbour's thread writes the next stage).  One thread processes one cell
// at a time — all n species — with the n x n matrix I + Fe*M in REGISTERS:
//   FAST : LU with partial pivoting applied directly to the gradient (no explicit inverse);
//          pivot search on the integer pipe, row swaps only when some lane of the warp needs one.
//   EXACT: nalgebra's try_inverse restated operation by operation (closed forms n<=4, partial-pivot
//          LU + two triangular solves on the permuted identity n>=5), then gemv in nalgebra's order.
// Cells are strided over threads (cell = tid + c*T): conflict-free shared-memory access.
// One __syncthreads per stage; the last one of a step carries the non-finite flag.
#pragma once
#include <algorithm>
#include <cstdio>
#include <cstdlib>

#include "arith.cuh"
#include "kernels.h"

namespace chrom {
namespace {


struct alignas(16) SpeciesConst {  // per species, shared memory; {K, feNK} and {a0, feNK2} load as 128-bit pairs
  double K;      // langmuir_k
  double feNK;   // FAST: fe * nbar * K
  double a0;     // FAST: 1 + fe*lambda
  double feNK2;  // = feNK (second copy so that the forward sweep needs one 128-bit load)
  double lambda;
  double nbar;   // stationary_fraction * (double)port_number
  double a0rK;   // FAST dense: (1 + fe*lambda) / K      } one 128-bit load
  double feN;    // FAST dense: fe * nbar (= feNK / K)   }
  double cgrK;   // FAST dense: (-ue/dz) / K
  double pad_;
};
static_assert(sizeof(SpeciesConst) == 10 * sizeof(double), "multi_resident.cu sizes shared memory with 10 doubles per species");

// ------------------------------------------------------------------------------------------------
// EXACT: nalgebra try_inverse + gemv, column-major a[i + j*N], one IEEE op per reference op.
// ------------------------------------------------------------------------------------------------
template <int N>
__device__ bool exact_try_inverse(const double* m, double* out) {
  if constexpr (N == 1) {
    if (m[0] == 0.0) return false;
    out[0] = x_div(1.0, m[0]);
    return true;
  } else if constexpr (N == 2) {
    const double m11 = m[0], m21 = m[1], m12 = m[2], m22 = m[3];
    const double det = x_sub(x_mul(m11, m22), x_mul(m21, m12));
    if (det == 0.0) return false;
    out[0] = x_div(m22, det);
    out[2] = x_div(-m12, det);
    out[1] = x_div(-m21, det);
    out[3] = x_div(m11, det);
    return true;
  } else if constexpr (N == 3) {
    const double m11 = m[0], m21 = m[1], m31 = m[2], m12 = m[3], m22 = m[4], m32 = m[5], m13 = m[6], m23 = m[7],
                 m33 = m[8];
    const double minor_m12_m23 = x_sub(x_mul(m22, m33), x_mul(m32, m23));
    const double minor_m11_m23 = x_sub(x_mul(m21, m33), x_mul(m31, m23));
    const double minor_m11_m22 = x_sub(x_mul(m21, m32), x_mul(m31, m22));
    const double det =
        x_add(x_sub(x_mul(m11, minor_m12_m23), x_mul(m12, minor_m11_m23)), x_mul(m13, minor_m11_m22));
    if (det == 0.0) return false;
    out[0] = x_div(minor_m12_m23, det);
    out[3] = x_div(x_sub(x_mul(m13, m32), x_mul(m33, m12)), det);
    out[6] = x_div(x_sub(x_mul(m12, m23), x_mul(m22, m13)), det);
    out[1] = x_div(-minor_m11_m23, det);
    out[4] = x_div(x_sub(x_mul(m11, m33), x_mul(m31, m13)), det);
    out[7] = x_div(x_sub(x_mul(m13, m21), x_mul(m23, m11)), det);
    out[2] = x_div(minor_m11_m22, det);
    out[5] = x_div(x_sub(x_mul(m12, m31), x_mul(m32, m11)), det);
    out[8] = x_div(x_sub(x_mul(m11, m22), x_mul(m21, m12)), det);
    return true;
  } else if constexpr (N == 4) {
    // do_inverse4: six-term cofactors a*b*c -/+ ..., evaluated left to right
#define T3(s, a, b, c) (s x_mul(x_mul(a, b), c))
#define COF(t0, t1, t2, t3, t4, t5) x_add(x_add(x_add(x_add(x_add(t0, t1), t2), t3), t4), t5)
    const double c00 = COF(T3(+, m[5], m[10], m[15]), T3(-, m[5], m[11], m[14]), T3(-, m[9], m[6], m[15]),
                           T3(+, m[9], m[7], m[14]), T3(+, m[13], m[6], m[11]), T3(-, m[13], m[7], m[10]));
    const double c01 = COF(T3(+, -m[4], m[10], m[15]), T3(+, m[4], m[11], m[14]), T3(+, m[8], m[6], m[15]),
                           T3(-, m[8], m[7], m[14]), T3(-, m[12], m[6], m[11]), T3(+, m[12], m[7], m[10]));
    const double c02 = COF(T3(+, m[4], m[9], m[15]), T3(-, m[4], m[11], m[13]), T3(-, m[8], m[5], m[15]),
                           T3(+, m[8], m[7], m[13]), T3(+, m[12], m[5], m[11]), T3(-, m[12], m[7], m[9]));
    const double c03 = COF(T3(+, -m[4], m[9], m[14]), T3(+, m[4], m[10], m[13]), T3(+, m[8], m[5], m[14]),
                           T3(-, m[8], m[6], m[13]), T3(-, m[12], m[5], m[10]), T3(+, m[12], m[6], m[9]));
    const double det =
        x_add(x_add(x_add(x_mul(m[0], c00), x_mul(m[1], c01)), x_mul(m[2], c02)), x_mul(m[3], c03));
    if (det == 0.0) return false;
    double inv[16];
    inv[0] = c00;
    inv[1] = COF(T3(+, -m[1], m[10], m[15]), T3(+, m[1], m[11], m[14]), T3(+, m[9], m[2], m[15]),
                 T3(-, m[9], m[3], m[14]), T3(-, m[13], m[2], m[11]), T3(+, m[13], m[3], m[10]));
    inv[2] = COF(T3(+, m[1], m[6], m[15]), T3(-, m[1], m[7], m[14]), T3(-, m[5], m[2], m[15]),
                 T3(+, m[5], m[3], m[14]), T3(+, m[13], m[2], m[7]), T3(-, m[13], m[3], m[6]));
    inv[3] = COF(T3(+, -m[1], m[6], m[11]), T3(+, m[1], m[7], m[10]), T3(+, m[5], m[2], m[11]),
                 T3(-, m[5], m[3], m[10]), T3(-, m[9], m[2], m[7]), T3(+, m[9], m[3], m[6]));
    inv[4] = c01;
    inv[5] = COF(T3(+, m[0], m[10], m[15]), T3(-, m[0], m[11], m[14]), T3(-, m[8], m[2], m[15]),
                 T3(+, m[8], m[3], m[14]), T3(+, m[12], m[2], m[11]), T3(-, m[12], m[3], m[10]));
    inv[6] = COF(T3(+, -m[0], m[6], m[15]), T3(+, m[0], m[7], m[14]), T3(+, m[4], m[2], m[15]),
                 T3(-, m[4], m[3], m[14]), T3(-, m[12], m[2], m[7]), T3(+, m[12], m[3], m[6]));
    inv[7] = COF(T3(+, m[0], m[6], m[11]), T3(-, m[0], m[7], m[10]), T3(-, m[4], m[2], m[11]),
                 T3(+, m[4], m[3], m[10]), T3(+, m[8], m[2], m[7]), T3(-, m[8], m[3], m[6]));
    inv[8] = c02;
    inv[9] = COF(T3(+, -m[0], m[9], m[15]), T3(+, m[0], m[11], m[13]), T3(+, m[8], m[1], m[15]),
                 T3(-, m[8], m[3], m[13]), T3(-, m[12], m[1], m[11]), T3(+, m[12], m[3], m[9]));
    inv[10] = COF(T3(+, m[0], m[5], m[15]), T3(-, m[0], m[7], m[13]), T3(-, m[4], m[1], m[15]),
                  T3(+, m[4], m[3], m[13]), T3(+, m[12], m[1], m[7]), T3(-, m[12], m[3], m[5]));
    inv[11] = COF(T3(+, -m[0], m[5], m[11]), T3(+, m[0], m[7], m[9]), T3(+, m[4], m[1], m[11]),
                  T3(-, m[4], m[3], m[9]), T3(-, m[8], m[1], m[7]), T3(+, m[8], m[3], m[5]));
    inv[12] = c03;
    inv[13] = COF(T3(+, m[0], m[9], m[14]), T3(-, m[0], m[10], m[13]), T3(-, m[8], m[1], m[14]),
                  T3(+, m[8], m[2], m[13]), T3(+, m[12], m[1], m[10]), T3(-, m[12], m[2], m[9]));
    inv[14] = COF(T3(+, -m[0], m[5], m[14]), T3(+, m[0], m[6], m[13]), T3(+, m[4], m[1], m[14]),
                  T3(-, m[4], m[2], m[13]), T3(-, m[12], m[1], m[6]), T3(+, m[12], m[2], m[5]));
    inv[15] = COF(T3(+, m[0], m[5], m[10]), T3(-, m[0], m[6], m[9]), T3(-, m[4], m[1], m[10]),
                  T3(+, m[4], m[2], m[9]), T3(+, m[8], m[1], m[6]), T3(-, m[8], m[2], m[5]));
#undef T3
#undef COF
    const double inv_det = x_div(1.0, det);
#pragma unroll
    for (int k = 0; k < 16; ++k) out[k] = x_mul(inv[k], inv_det);
    return true;
  } else {
  // N >= 5: lu::try_invert_to
  double a[N * N];
#pragma unroll
  for (int k = 0; k < N * N; ++k) {
    a[k] = m[k];
    out[k] = ((k % N) == (k / N)) ? 1.0 : 0.0;
  }
  for (int i = 0; i < N; ++i) {
    int piv = i;
    double the_max = fabs(a[i + i * N]);
    for (int r = i + 1; r < N; ++r) {
      const double v = fabs(a[r + i * N]);
      if (v > the_max) {
        the_max = v;
        piv = r;
      }
    }
    const double diag = a[piv + i * N];
    if (diag == 0.0) return false;
    if (piv != i) {
      for (int j = 0; j < N; ++j) {
        const double t = out[i + j * N];
        out[i + j * N] = out[piv + j * N];
        out[piv + j * N] = t;
      }
      for (int j = 0; j <= i; ++j) {  // columns ..i of the L part, and the coefficient column i
        const double t = a[i + j * N];
        a[i + j * N] = a[piv + j * N];
        a[piv + j * N] = t;
      }
    }
    const double inv_diag = x_div(1.0, diag);
    for (int r = i + 1; r < N; ++r) a[r + i * N] = x_mul(a[r + i * N], inv_diag);
    for (int k = i + 1; k < N; ++k) {
      if (piv != i) {
        const double t = a[i + k * N];
        a[i + k * N] = a[piv + k * N];
        a[piv + k * N] = t;
      }
      const double neg_pivot = -a[i + k * N];
      for (int r = i + 1; r < N; ++r) a[r + k * N] = x_add(x_mul(neg_pivot, a[r + i * N]), a[r + k * N]);
    }
  }
  for (int c = 0; c < N; ++c)
    for (int i = 0; i + 1 < N; ++i) {
      const double coeff = out[i + c * N];  // b[i] / 1.0
      for (int r = i + 1; r < N; ++r) out[r + c * N] = x_add(x_mul(-coeff, a[r + i * N]), out[r + c * N]);
    }
  for (int c = 0; c < N; ++c)
    for (int i = N - 1; i >= 0; --i) {
      const double diag = a[i + i * N];
      if (diag == 0.0) return false;
      const double coeff = x_div(out[i + c * N], diag);
      out[i + c * N] = coeff;
      for (int r = 0; r < i; ++r) out[r + c * N] = x_add(x_mul(-coeff, a[r + i * N]), out[r + c * N]);
    }
  return true;
  }
}

// compute_row in the reference's order.  c[], up[] -> k[] = dC/dt of this cell.
template <int N>
__device__ void exact_row(const SpeciesConst* sc, double fe, double ue, double dz, const double* c, const double* up,
                          double* k) {
  double sum_kc = 0.0;
#pragma unroll
  for (int s = 0; s < N; ++s) sum_kc = x_add(sum_kc, x_mul(sc[s].K, c[s]));
  const double denom = x_add(1.0, sum_kc);
  const double denom2 = x_mul(denom, denom);
  double mat[N * N], inv[N * N];
#pragma unroll
  for (int i = 0; i < N; ++i) {
    const double nb = sc[i].nbar, Ki = sc[i].K;
#pragma unroll
    for (int j = 0; j < N; ++j) {
      double J;
      if (i == j) {
        const double num = x_mul(x_mul(nb, Ki), x_sub(denom, x_mul(Ki, c[i])));
        J = x_add(sc[i].lambda, x_div(num, denom2));
      } else {
        const double num = x_mul(x_mul(x_mul(-nb, Ki), sc[j].K), c[i]);
        J = x_div(num, denom2);
      }
      mat[i + j * N] = x_add((i == j) ? 1.0 : 0.0, x_mul(fe, J));
    }
  }
  if (!exact_try_inverse<N>(mat, inv)) {  // identity fallback, langmuir_multi.rs:790-796
#pragma unroll
    for (int q = 0; q < N * N; ++q) inv[q] = ((q % N) == (q / N)) ? 1.0 : 0.0;
  }
  double grad[N], dv[N];
#pragma unroll
  for (int s = 0; s < N; ++s) grad[s] = x_div(x_sub(c[s], up[s]), dz);
#pragma unroll
  for (int i = 0; i < N; ++i) dv[i] = x_mul(inv[i], grad[0]);
#pragma unroll
  for (int j = 1; j < N; ++j)
#pragma unroll
    for (int i = 0; i < N; ++i) dv[i] = x_add(x_mul(inv[i + j * N], grad[j]), dv[i]);
#pragma unroll
  for (int s = 0; s < N; ++s) k[s] = x_mul(-ue, dv[s]);
}

// ------------------------------------------------------------------------------------------------
// FAST: A = I + fe*M = diag(alpha) - p q^T built in registers, LU with partial pivoting applied to
// rhs = (-ue/dz)*(c - up), back substitution.  k = A^-1 rhs.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned long long abs_bits(double v) {
  return (unsigned long long)__double_as_longlong(v) & 0x7fffffffffffffffull;
}

// One elimination step with a compile-time pivot column I (template recursion instead of a loop:
// every register index is a constant, so the matrix never falls back to local memory).
template <int N, int I>
__device__ __forceinline__ void lu_step(double (&a)[N][N], double (&b)[N], bool& singular) {
  if constexpr (I + 1 < N) {
    // Partial pivoting.  Cheap test first (integer pipe): can any row below hold a larger |entry|?
    // High words only, '>=' so that ties in the high word are never missed.  The matrices are
    // close to diagonally dominant, so the swap path below is almost never entered.
    unsigned below = 0u;
#pragma unroll
    for (int r = I + 1; r < N; ++r) below = max(below, abs_hi(a[r][I]));
    if (__any_sync(0xffffffffu, below >= abs_hi(a[I][I]))) {
      // Exact path: compare-and-swap row I against every row below.  Afterwards
      // |a[I][I]| = max_r |a[r][I]|: a valid partial pivot.
#pragma unroll
      for (int r = I + 1; r < N; ++r) {
        const bool sw = abs_bits(a[r][I]) > abs_bits(a[I][I]);
#pragma unroll
        for (int q = I; q < N; ++q) {
          const double t = a[I][q];
          a[I][q] = sw ? a[r][q] : t;
          a[r][q] = sw ? t : a[r][q];
        }
        const double t = b[I];
        b[I] = sw ? b[r] : t;
        b[r] = sw ? t : b[r];
      }
    }
  }
  singular |= (a[I][I] == 0.0);
  const double rp = fast_rcp(a[I][I]);
  a[I][I] = rp;  // keep the reciprocal for the back substitution
#pragma unroll
  for (int r = I + 1; r < N; ++r) {
    const double l = a[r][I] * rp;
#pragma unroll
    for (int q = I + 1; q < N; ++q) a[r][q] = fma(-l, a[I][q], a[r][q]);
    b[r] = fma(-l, b[I], b[r]);
  }
  if constexpr (I + 1 < N) lu_step<N, I + 1>(a, b, singular);
}

template <int N, int I>
__device__ __forceinline__ void back_step(const double (&a)[N][N], const double (&b)[N], double* k) {
  double s = b[I];
#pragma unroll
  for (int q = I + 1; q < N; ++q) s = fma(-a[I][q], k[q], s);
  k[I] = s * a[I][I];
  if constexpr (I > 0) back_step<N, I - 1>(a, b, k);
}

template <int N>
__device__ __forceinline__ void fast_row(const SpeciesConst* sc, double fe, double cg, const double* c,
                                         const double* up, double* k) {
  (void)fe;
  double S = 0.0;
#pragma unroll
  for (int s = 0; s < N; ++s) S = fma(sc[s].K, c[s], S);
  const double D = 1.0 + S;
  const double rD2 = fast_rcp(D * D);
  const double DrD2 = D * rD2;
  double a[N][N], b[N];
#pragma unroll
  for (int i = 0; i < N; ++i) {
    const double p = sc[i].feNK * c[i] * rD2;  // fe * nbar_i K_i c_i / D^2
#pragma unroll
    for (int j = 0; j < N; ++j) a[i][j] = -p * sc[j].K;
    a[i][i] = fma(sc[i].feNK, DrD2, sc[i].a0) + a[i][i];  // 1 + fe*(lambda_i + nbar_i K_i (D - K_i c_i)/D^2)
    b[i] = (c[i] - up[i]) * cg;
  }
  bool singular = false;
  lu_step<N, 0>(a, b, singular);
  back_step<N, N - 1>(a, b, k);
  if (singular) {  // exact zero pivot: the reference falls back to the identity (A^-1 := I)
#pragma unroll
    for (int s = 0; s < N; ++s) k[s] = (c[s] - up[s]) * cg;
  }
}

// ------------------------------------------------------------------------------------------------
// FAST dense, speculative form (the hot path for n <= 8).  Two changes against fast_row, same LU:
//  * column scaling x_j = y_j / K_j:  (diag(alpha) - p q^T) x = r  <=>  (diag(alpha_i/K_i) - p 1^T) y = r,
//    so the off-diagonal entries of row i are all -p_i (no n^2 products to build the matrix) and the
//    scaling by cg = -ue/dz moves into the final product x_j = y_j * (cg / K_j).  Column scaling
//    commutes with Gaussian elimination: the multipliers and the pivot ROWS are those of the unscaled
//    matrix.
//  * the elimination runs without row exchanges and without votes or branches (straight-line code the
//    scheduler can overlap: reciprocal chain of step i+1 under the rank-1 update of step i) while the
//    partial-pivoting test of every step is evaluated on the integer pipe; a cell where a row below the
//    diagonal could hold a larger entry (or a pivot is exactly zero) returns true and is redone by the
//    caller with the row-exchanging fast_row.  The matrices are close to diagonally dominant
//    (|a_ri|/|a_ii| <= (nbar_r/nbar_i) K_r c_r / (1 + sum K c)), so this is rare.
// ------------------------------------------------------------------------------------------------
template <int N, int I>
__device__ __forceinline__ void lu_step_spec(double (&a)[N][N], double (&b)[N], bool& redo) {
  if constexpr (I + 1 < N) {
    unsigned below = 0u;
#pragma unroll
    for (int r = I + 1; r < N; ++r) below = max(below, abs_hi(a[r][I]));
    redo |= below >= abs_hi(a[I][I]);  // '>=' on the high words: never misses a needed exchange
  }
  redo |= (a[I][I] == 0.0);
  const double rp = fast_rcp(a[I][I]);
  a[I][I] = rp;
#pragma unroll
  for (int r = I + 1; r < N; ++r) {
    const double l = a[r][I] * rp;
#pragma unroll
    for (int q = I + 1; q < N; ++q) a[r][q] = fma(-l, a[I][q], a[r][q]);
    b[r] = fma(-l, b[I], b[r]);
  }
  if constexpr (I + 1 < N) lu_step_spec<N, I + 1>(a, b, redo);
}

// Column-oriented back substitution: as soon as k[I] is known every row above takes its contribution,
// so the dependent chain is one multiply + one FMA per unknown (the row-oriented form of back_step
// chains a whole dot product per unknown).
template <int N, int I>
__device__ __forceinline__ void back_axpy(const double (&a)[N][N], double (&b)[N], double* k) {
  k[I] = b[I] * a[I][I];
#pragma unroll
  for (int r = 0; r < I; ++r) b[r] = fma(-a[r][I], k[I], b[r]);
  if constexpr (I > 0) back_axpy<N, I - 1>(a, b, k);
}

template <int N>
__device__ __forceinline__ bool fast_row_spec(const SpeciesConst* sc, const double* c, const double* up, double* k) {
  double S0 = 0.0, S1 = 0.0;  // two chains: the sum is not on the critical path for long
#pragma unroll
  for (int s = 0; s < N; s += 2) {
    S0 = fma(sc[s].K, c[s], S0);
    if (s + 1 < N) S1 = fma(sc[s + 1].K, c[s + 1], S1);
  }
  const double D = 1.0 + (S0 + S1);
  const double rD2 = fast_rcp(D * D);
  const double DrD2 = D * rD2;
  double a[N][N], b[N];
#pragma unroll
  for (int i = 0; i < N; ++i) {
    const double np = -(sc[i].feNK * (c[i] * rD2));  // -p_i = -fe nbar_i K_i c_i / D^2
#pragma unroll
    for (int j = 0; j < N; ++j) a[i][j] = np;
    a[i][i] = fma(sc[i].feN, DrD2, sc[i].a0rK) + np;  // alpha_i / K_i - p_i
    b[i] = c[i] - up[i];
  }
  bool redo = false;
  lu_step_spec<N, 0>(a, b, redo);
  back_axpy<N, N - 1>(a, b, k);
#pragma unroll
  for (int s = 0; s < N; ++s) k[s] *= sc[s].cgrK;
  return redo;
}

// ------------------------------------------------------------------------------------------------
// STRUCTURED: Gaussian elimination with partial pivoting on A = diag(alpha) - p q^T WITHOUT forming
// the matrix.  Eliminating column i of diag(alpha) - g p q^T leaves diag(alpha') - g' p' q'^T on the
// trailing block with g' = g * alpha_i / d_i, d_i = alpha_i - g p_i q_i (the pivot), so L, U and the
// solve need O(n) scalars:  U[i][c] = -g_i p_i q_c,  L[r][i] = -g_i p_r q_i / d_i.
// The pivot test of partial pivoting, |L[r][i] d_i| > |d_i| for some r > i, becomes
// g_i q_i max_{r>i} |p_r| > |d_i|; when it fires the caller redoes the cell with the dense
// row-exchanging LU (register LU for n <= 8, scratch-memory LU beyond).  Same pivots, same factors,
// O(n) registers: this is what makes n > 8 fit a thread and lets two warps more per SM fly.
// ------------------------------------------------------------------------------------------------

// The same elimination in closed form (compile-time n, registers).  Eliminating species 0..i-1 of
// diag(alpha) - p q^T leaves the pivot alpha_i - p_i q_i / h_i with h_i = 1 - sum_{j<i} p_j q_j / alpha_j, and
// the solution is x = z + v (q.z) / (1 - q.v), z = b / alpha, v = p / alpha: the n reciprocals are independent
// (the sweep above chains n dependent ones) and the work drops to ~18 FP64 instructions per species.
// Partial pivoting: rows would be exchanged at step i iff q_i max_{r>i}|p_r| > |alpha_i h_i - p_i q_i|.  A cheap
// sufficient condition rules that out for the whole cell, 2 max_i q_i max_r|p_r| < min_i alpha_i (1 - q.v)
// (h_i >= h_n = 1 - q.v for p >= 0); only a cell that fails it evaluates the per-step test, and only a cell
// that fails THAT is redone with the dense row-exchanging LU by the caller (returns true).
template <int N>
__device__ __forceinline__ bool scan_cell(const SpeciesConst* sc, double cg, const double* SRC, const double* inlet_row,
                                          uint32_t nz, uint32_t cc, double (&x)[N]) {
  double K[N], pw[N], v[N];
  double S = 0.0, kmax = 0.0;
  unsigned pmax_hi = 0u;  // integer max over the high words: negative / non-finite entries read as huge
#pragma unroll
  for (int s = 0; s < N; ++s) {
    const double c = SRC[(size_t)s * nz + cc];
    const double up = (cc > 0) ? SRC[(size_t)s * nz + cc - 1] : inlet_row[s];
    const double2 kf = *reinterpret_cast<const double2*>(&sc[s].K);  // {K, feNK}
    K[s] = kf.x;
    x[s] = (c - up) * cg;
    S = fma(kf.x, c, S);
    pw[s] = kf.y * c;  // feNK_s * c_s; p_s = pw_s / D^2
    pmax_hi = max(pmax_hi, (unsigned)__double2hiint(pw[s]));
    kmax = fmax(kmax, kf.x);
  }
  const bool nonneg = pmax_hi < 0x7ff00000u;
  const double pmax = __hiloint2double((int)(pmax_hi + 1u), 0);  // upper bound of max |feNK c|
  const double rD = fast_rcp(1.0 + S);
  const double rD2 = rD * rD;
  double qv = 0.0, qz = 0.0, amin = 1e300;
#pragma unroll
  for (int i = 0; i < N; ++i) {
    const double2 af = *reinterpret_cast<const double2*>(&sc[i].a0);  // {a0, feNK}
    const double alpha = fma(af.y, rD, af.x);                        // 1 + fe*(lambda_i + nbar_i K_i / D)
    const double ra = fast_rcp(alpha);
    amin = fmin(amin, alpha);
    x[i] *= ra;                  // z_i
    v[i] = (pw[i] * rD2) * ra;   // v_i
    qv = fma(K[i], v[i], qv);
    qz = fma(K[i], x[i], qz);
  }
  const double hn = 1.0 - qv;
  const double f = qz * fast_rcp(hn);
  bool redo = false;
  if (!(nonneg && (2.0 * kmax) * (pmax * rD2) < amin * hn)) {
    // per-step test (rare): h_i by a running sum, suffix maximum of |p| from the back
    double sfx[N], run = 0.0;
    double m = 0.0;
#pragma unroll
    for (int s = N - 1; s >= 0; --s) {
      sfx[s] = m;
      m = fmax(m, fabs(pw[s]));
    }
#pragma unroll
    for (int i = 0; i < N; ++i) {
      const double2 af = *reinterpret_cast<const double2*>(&sc[i].a0);
      const double alpha = fma(af.y, rD, af.x);
      const double h = 1.0 - run;
      const double p = pw[i] * rD2;
      const double dpiv = fabs(fma(alpha, h, -(p * K[i])));
      redo |= !(h > 0.0) || (K[i] * (sfx[i] * rD2) > dpiv) || (dpiv == 0.0);
      run = fma(K[i], v[i], run);
    }
  }
#pragma unroll
  for (int i = 0; i < N; ++i) x[i] = fma(v[i], f, x[i]);
  return redo;
}

// Stage algebra of one cell (all species) on the shared-memory state.  Returns the non-finite flag of
// the new state where the stage produces one.
struct StageCtx {
  double* Y;
  double* ACC;
  double* DST;
  uint32_t nz;
  int stage;  // 0 Euler, 1..4 RK4
  double dt, h2, h6;
};

template <int N, bool EXACT>
__device__ __forceinline__ bool apply_stage(const StageCtx& st, uint32_t cc, const double* k, int /*n*/) {
  constexpr int CAP = N;
  constexpr int n = N;
  auto axpy = [](double y, double kk, double h) { return EXACT ? x_add(y, x_mul(kk, h)) : fma(kk, h, y); };
  bool bad = false;
#pragma unroll
  for (int s = 0; s < CAP; ++s)
    if (s < n) {
      const size_t idx = (size_t)s * st.nz + cc;
      if (st.stage == 0) {
        const double yn = axpy(st.Y[idx], k[s], st.dt);
        st.Y[idx] = yn;  // Euler reads SRC = U0/U1 (a copy of Y), so Y can be overwritten
        st.DST[idx] = yn;
        bad |= hi_nonfinite(abs_hi(yn));
      } else if (st.stage == 1) {
        st.ACC[idx] = k[s];
        st.DST[idx] = axpy(st.Y[idx], k[s], st.h2);
      } else if (st.stage == 2) {
        st.ACC[idx] = axpy(st.ACC[idx], k[s], 2.0);
        st.DST[idx] = axpy(st.Y[idx], k[s], st.h2);
      } else if (st.stage == 3) {
        st.ACC[idx] = axpy(st.ACC[idx], k[s], 2.0);
        st.DST[idx] = axpy(st.Y[idx], k[s], st.dt);
      } else {
        const double w = EXACT ? x_add(st.ACC[idx], k[s]) : st.ACC[idx] + k[s];
        const double yn = axpy(st.Y[idx], w, st.h6);
        st.Y[idx] = yn;
        bad |= hi_nonfinite(abs_hi(yn));
      }
    }
  return bad;
}

// The dense fallback of the structured kernels, kept out of line so that its 64-entry matrix does not
// raise the register count of the hot path (it reloads the cell from shared memory itself).
template <int N>
__device__ __noinline__ bool cold_cell(const SpeciesConst* sc, double fe, double cg, const double* SRC,
                                       const double* inlet_row, uint32_t cc, bool write, StageCtx st) {
  double c[N], up[N], k[N];
#pragma unroll
  for (int s = 0; s < N; ++s) {
    c[s] = SRC[(size_t)s * st.nz + cc];
    up[s] = (cc > 0) ? SRC[(size_t)s * st.nz + cc - 1] : inlet_row[s];
  }
  fast_row<N>(sc, fe, cg, c, up, k);  // warp-synchronous inside (pivot votes): called by whole warps
  return write ? apply_stage<N, false>(st, cc, k, N) : false;
}

// ---- kernel ---------------------------------------------------------------------------------------
// ARITH: CHROM_ARITH_FAST (dense register LU) and CHROM_ARITH_EXACT for 1 <= N <= 8 (the n x n system in one
// thread's registers); CHROM_KARITH_STRUCT (closed-form elimination, 4 N doubles per thread) for 1 <= N <= 16.
// Other species counts / modes run one warp per cell (multi_tiled.cuh).
constexpr int multi_max_threads(int N, int ARITH) {
  return ARITH == CHROM_KARITH_STRUCT ? (N <= 8 ? 512 : 256) : (N <= 2 ? 512 : 256);
}

template <int N, int METHOD, int ARITH>
__global__ void __launch_bounds__(multi_max_threads(N, ARITH), 1) k_multi_resident(const BatchDev b) {
  static_assert(N >= 1 && (N <= 8 || (N <= 16 && ARITH == CHROM_KARITH_STRUCT)),
                "thread-per-cell kernels: dense / exact n <= 8 (matrix in registers), closed form n <= 16");
  constexpr int ST = (METHOD == CHROM_RK4) ? 3 : 1;
  constexpr bool EXACT = (ARITH == CHROM_ARITH_EXACT);
  constexpr int CAP = N;
  extern __shared__ double smem[];
  constexpr int n = N;
  const uint32_t nz = b.nz;
  const int T = blockDim.x;
  const int tid = threadIdx.x;
  const size_t plane = (size_t)n * nz;  // doubles per state array
  double* Y = smem;
  double* ACC = Y + plane;
  double* U0 = ACC + plane;
  double* U1 = U0 + plane;
  double* inlet = U1 + plane;                        // [3][n] current step's inlet values
  double* summ = inlet + 3 * n;                               // [n] Summ (kSummSlots doubles each)
  SpeciesConst* sc = (SpeciesConst*)(summ + kSummSlots * n);  // [n]
  __shared__ const double* tabs[CAP];

  const double dt = b.dt, h2 = b.half_dt, h6 = b.sixth_dt;
  const uint32_t steps = b.steps;
  const int cpt = (int)((nz + T - 1) / T);  // cells per thread
  auto axpy = [](double y, double k, double h) { return EXACT ? x_add(y, x_mul(k, h)) : fma(k, h, y); };

  for (uint32_t col = blockIdx.x; col < b.n_cols; col += gridDim.x) {  // persistent CTAs over columns
    const chrom_multi_col cp = b.mcols[col];
    const double fe = cp.fe, ue = cp.ue, dz = cp.dz;
    const double cg = -ue / dz;  // FAST only
    __syncthreads();             // the previous column's readers of sc / tabs / Y are done
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
    for (size_t i = tid; i < plane; i += T) {
      Y[i] = b.y0 ? b.y0[(size_t)col * plane + i] : 0.0;
      ACC[i] = 0.0;
    }
    __syncthreads();

    // initial samples
    const uint32_t det = detector_of(cp.detector_cell, nz);  // where outlet and summary are sampled
    for (int s = tid; s < n; s += T) {
      const double v0 = Y[(size_t)s * nz + det];
      if (b.outlet && sample_has_initial(b.out_steps, b.outlet_stride)) b.outlet[((size_t)col * n + s) * b.n_out] = v0;
      summ_init(reinterpret_cast<Summ*>(summ)[s], v0);
    }
    if (b.snapshots && sample_has_initial(b.snap_steps, b.snapshot_stride))
      for (size_t i = tid; i < plane; i += T) b.snapshots[(size_t)col * b.n_snap * plane + i] = Y[i];
    SampleCursor oc = sample_begin(b.out_steps, b.n_out, b.outlet ? b.outlet_stride : 0u, 0u);
    SampleCursor sc_ = sample_begin(b.snap_steps, b.n_snap, b.snapshots ? b.snapshot_stride : 0u, 0u);
    if (!b.outlet) oc.next = kNoSample;
    if (!b.snapshots) sc_.next = kNoSample;
    bool done = false;

    // One stage: K = f(SRC, inlet row `irow`), then the stage algebra selected by `stage`.
    auto run_stage = [&](const double* SRC, double* DST, int irow, int stage) -> bool {
      bool bad = false;
      StageCtx st{Y, ACC, DST, nz, METHOD == CHROM_EULER ? 0 : stage, dt, h2, h6};
      unsigned redo_mask = 0u;  // structured kernels: cells whose LU needs a row exchange
      for (int ci = 0; ci < cpt; ++ci) {
        const uint32_t cell = (uint32_t)tid + (uint32_t)ci * T;
        const bool live = cell < nz;
        const uint32_t cc = live ? cell : nz - 1;  // idle lanes recompute the last cell, never store
        double k[CAP];
        if constexpr (ARITH == CHROM_KARITH_STRUCT) {
          const bool redo = scan_cell<N>(sc, cg, SRC, inlet + irow * n, nz, cc, k);
          if (redo) redo_mask |= 1u << ci;
          else if (live) bad |= apply_stage<N, false>(st, cc, k, n);
        } else {
          double c[CAP], up[CAP];
#pragma unroll
          for (int s = 0; s < CAP; ++s) {
            c[s] = SRC[(size_t)s * nz + cc];
            up[s] = (cc > 0) ? SRC[(size_t)s * nz + cc - 1] : inlet[irow * n + s];
          }
          if constexpr (EXACT) {
            exact_row<N>(sc, fe, ue, dz, c, up, k);
          } else {
            // speculative exchange-free LU; a warp with a flagged cell redoes it with the row-exchanging LU
            const bool redo = fast_row_spec<N>(sc, c, up, k);
            if (__any_sync(0xffffffffu, redo)) {
              bad |= cold_cell<N>(sc, fe, cg, SRC, inlet + irow * n, cc, live, st);
              continue;
            }
          }
          if (live) bad |= apply_stage<N, EXACT>(st, cc, k, n);
        }
      }
      if constexpr (ARITH == CHROM_KARITH_STRUCT) {
        // cold pass: redo the flagged cells with the dense row-exchanging LU (whole warps, out of line)
        if (__any_sync(0xffffffffu, redo_mask != 0u)) {
          for (int ci = 0; ci < cpt; ++ci) {
            const bool mine = (redo_mask >> ci) & 1u;
            if (!__any_sync(0xffffffffu, mine)) continue;
            const uint32_t cell = (uint32_t)tid + (uint32_t)ci * T;
            const bool live = cell < nz;
            bad |= cold_cell<N>(sc, fe, cg, SRC, inlet + irow * n, live ? cell : nz - 1, mine && live, st);
          }
        }
      }
      return bad;
    };

    if (METHOD == CHROM_EULER) {
      for (size_t i = tid; i < plane; i += T) U0[i] = Y[i];
    }

    for (uint32_t step = 0; step < steps; ++step) {
      for (int i = tid; i < n * ST; i += T) {  // inlet values of this step: [stage row][species]
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
        // ping-pong: even steps read U0 / write U1, odd steps the reverse; Y always holds the new state
        bad = (step & 1u) ? run_stage(U1, U0, 0, 0) : run_stage(U0, U1, 0, 0);
      }
      const int any_bad = __syncthreads_or(bad);

      // ---- samples --------------------------------------------------------------------------
      const bool out_now = (step + 1u == oc.next);
      if (out_now || b.summary)
        for (int s = tid; s < n; s += T) {
          const double v = Y[(size_t)s * nz + det];
          if (out_now) b.outlet[((size_t)col * n + s) * b.n_out + oc.k] = v;
          if (b.summary) summ_step(reinterpret_cast<Summ*>(summ)[s], v, step, dt);
        }
      if (out_now) sample_advance(oc, b.out_steps, b.outlet_stride);
      if (step + 1u == sc_.next) {
        double* dst = b.snapshots + ((size_t)col * b.n_snap + sc_.k) * plane;
        for (size_t i = tid; i < plane; i += T) dst[i] = Y[i];
        sample_advance(sc_, b.snap_steps, b.snapshot_stride);
      }
      if (any_bad && !done) {  // validate_state: NaN anywhere wins over Inf (solver/mod.rs:519-548)
        bool has_nan = false;
        for (size_t i = tid; i < plane; i += T) has_nan |= (Y[i] != Y[i]);
        const int any_nan = __syncthreads_or(has_nan);
        if (tid == 0 && b.status) b.status[col] = any_nan ? (long long)(step + 1) : -(long long)(step + 1);
        done = true;
      }
    }
    __syncthreads();
    if (b.final_state)
      for (size_t i = tid; i < plane; i += T) b.final_state[(size_t)col * plane + i] = Y[i];
    if (b.summary)
      for (int s = tid; s < n; s += T)
        summ_store(b.summary + ((size_t)col * n + s) * CHROM_SUMMARY_LEN, reinterpret_cast<const Summ*>(summ)[s]);
  }
}

typedef void (*mkern_t)(const BatchDev);

template <int N>
mkern_t pick_n(int method, int arith) {
  if constexpr (N >= 1 && N <= 8) {
    if (method == CHROM_RK4) {
      if (arith == CHROM_ARITH_EXACT) return k_multi_resident<N, CHROM_RK4, CHROM_ARITH_EXACT>;
      if (arith == CHROM_KARITH_STRUCT) return k_multi_resident<N, CHROM_RK4, CHROM_KARITH_STRUCT>;
      return k_multi_resident<N, CHROM_RK4, CHROM_ARITH_FAST>;
    }
    if (arith == CHROM_ARITH_EXACT) return k_multi_resident<N, CHROM_EULER, CHROM_ARITH_EXACT>;
    if (arith == CHROM_KARITH_STRUCT) return k_multi_resident<N, CHROM_EULER, CHROM_KARITH_STRUCT>;
    return k_multi_resident<N, CHROM_EULER, CHROM_ARITH_FAST>;
  } else if constexpr (N >= 9 && N <= 16) {
    if (arith != CHROM_KARITH_STRUCT) return nullptr;
    return method == CHROM_RK4 ? k_multi_resident<N, CHROM_RK4, CHROM_KARITH_STRUCT>
                               : k_multi_resident<N, CHROM_EULER, CHROM_KARITH_STRUCT>;
  }
  return nullptr;
}

}  // namespace
}  // namespace chrom
