// single_rhs.cuh — the LangmuirSingle right-hand side of one cell, in both arithmetic policies.
// Shared by the register-resident kernels (single_resident.cu) and the streaming kernel (single_stream.cu).
#pragma once
#include "../../include/chrom_cuda.h"
#include "arith.cuh"

namespace chrom {

// ARITH: 0 = FAST (no cancellation anywhere), 1 = EXACT (reference order), 2 = FAST_POLY (one instruction
// fewer; chosen by the host when the isotherm is not extremely steep, see CHROM_KARITH_POLY in kernels.h)
template <int ARITH>
struct Coef;

// EXACT: langmuir_single.rs:391-396, 409-412, 503-510, one IEEE op per reference op.
template <>
struct Coef<1> {
  double K, nK, lambda, fe, ue, dz;
  __device__ __forceinline__ void init(const chrom_single_col& c) {
    // n_bar = fe / (1.0 + fe) * port_number is loop-invariant: hoisting it gives the same bits.
    const double n_bar = x_mul(x_div(c.fe, x_add(1.0, c.fe)), c.port_number);
    K = c.langmuir_k;
    nK = x_mul(n_bar, c.langmuir_k);
    lambda = c.lambda;
    fe = c.fe;
    ue = c.ue;
    dz = c.dz;
  }
  __device__ __forceinline__ double rhs(double c, double up) const {
    const double denom = x_add(1.0, x_mul(K, c));
    const double deriv = x_add(lambda, x_div(nK, x_mul(denom, denom)));
    const double sigma = x_div(1.0, x_add(1.0, x_mul(fe, deriv)));
    const double dc_dz = x_div(x_sub(c, up), dz);
    return x_mul(x_mul(-sigma, ue), dc_dz);
  }
  static __device__ __forceinline__ double axpy(double y, double k, double h) { return x_add(y, x_mul(k, h)); }
  static __device__ __forceinline__ double add(double a, double b) { return x_add(a, b); }
};

// FAST: sigma*(-ue/dz)*(c-up) = (c-up)*d2 / (A'*d2 + B'),  d = 1+K*c, d2 = d*d,
//       A' = (1+fe*lambda)/Cg, B' = fe*n_bar*K/Cg, Cg = -ue/dz.   9 DFMA-pipe ops + 1 MUFU.
template <>
struct Coef<0> {
  double K, Ap, Bp;
  __device__ __forceinline__ void init(const chrom_single_col& c) {
    const double n_bar = c.fe / (1.0 + c.fe) * c.port_number;
    const double A = 1.0 + c.fe * c.lambda;
    const double B = c.fe * (n_bar * c.langmuir_k);
    const double Cg = -c.ue / c.dz;
    K = c.langmuir_k;
    Ap = A / Cg;
    Bp = B / Cg;
  }
  __device__ __forceinline__ double rhs(double c, double up) const {
    const double d = fma(K, c, 1.0);
    const double d2 = d * d;
    const double den = fma(Ap, d2, Bp);
    const double r = fast_rcp(den);
    const double g = c - up;
    return (g * d2) * r;
  }
  static __device__ __forceinline__ double axpy(double y, double k, double h) { return fma(k, h, y); }
  static __device__ __forceinline__ double add(double a, double b) { return a + b; }
};

// FAST_POLY: the same function with the denominator as a Horner polynomial in c and the quotient
// split off its constant part:
//   sigma' = d2/(A'd2 + B') = c1 - 1/P(c),  c1 = 1/A',  P(c) = (A'd2 + B')/c2,  c2 = B'/A',
//   P(c) = p0 + c*(p1 + p2*c),  p0 = (A'+B')/c2, p1 = 2A'K/c2, p2 = A'K^2/c2.
// 8 FP64-pipe ops + 1 MUFU per cell-stage (2 FMA for P, 3 for the reciprocal, c1 - x, c - up, product).
// c1 - x cancels by a factor 1 + B/(A d2) <= 1 + B/A, so the host selects this form only when
// B/A = fe*n_bar*K/(1+fe*lambda) <= 1e3 (relative error <= ~1e-13, far inside 1e-9).
template <>
struct Coef<2> {
  double p0, p1, p2, c1;
  __device__ __forceinline__ void init(const chrom_single_col& c) {
    const double n_bar = c.fe / (1.0 + c.fe) * c.port_number;
    const double A = 1.0 + c.fe * c.lambda;
    const double B = c.fe * (n_bar * c.langmuir_k);
    const double Cg = -c.ue / c.dz;
    const double Ap = A / Cg, Bp = B / Cg, K = c.langmuir_k;
    c1 = 1.0 / Ap;
    if (B != 0.0) {
      const double c2 = Bp / Ap;
      p0 = (Ap + Bp) / c2;
      p1 = 2.0 * Ap * K / c2;
      p2 = Ap * K * K / c2;
    } else {  // linear isotherm: sigma' = c1 exactly; 1/P underflows to nothing
      p0 = -1e300;
      p1 = 0.0;
      p2 = 0.0;
    }
  }
  __device__ __forceinline__ double rhs(double c, double up) const {
    const double P = fma(c, fma(p2, c, p1), p0);
    const double x = fast_rcp(P);
    const double g = c - up;
    return g * (c1 - x);
  }
  static __device__ __forceinline__ double axpy(double y, double k, double h) { return fma(k, h, y); }
  static __device__ __forceinline__ double add(double a, double b) { return a + b; }
};

}  // namespace chrom
