// single_rhs.cuh — the LangmuirSingle right-hand side of one cell, in both arithmetic policies.
// Shared by the register-resident kernels (single_resident.cu) and the streaming kernel (single_stream.cu).
#pragma once
#include "../../include/chrom_cuda.h"
#include "arith.cuh"

namespace chrom {

template <bool EXACT>
struct Coef;

// EXACT: langmuir_single.rs:391-396, 409-412, 503-510, one IEEE op per reference op.
template <>
struct Coef<true> {
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
struct Coef<false> {
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

}  // namespace chrom
