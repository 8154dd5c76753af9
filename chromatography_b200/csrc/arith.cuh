// arith.cuh — the two arithmetic policies of the kernels.
//
// EXACT: each IEEE-754 double operation of the reference, in the reference's order, through
//        the __d*_rn intrinsics (round-to-nearest, never contracted into an FMA by nvcc).
// FAST : FMA contraction + one reciprocal per cell-stage (MUFU.RCP64H seed refined by DFMAs).
#pragma once
#include <cuda_runtime.h>

namespace chrom {

// ---- exact (reference-order) primitives --------------------------------------------------
__device__ __forceinline__ double x_add(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double x_sub(double a, double b) { return __dsub_rn(a, b); }
__device__ __forceinline__ double x_mul(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double x_div(double a, double b) { return __ddiv_rn(a, b); }

// ---- fast reciprocal -----------------------------------------------------------------------
// Seed: rcp.approx.ftz.f64 (SASS MUFU.RCP64H, runs on the XU pipe, not the FP64 pipe).
// RCP_REFINE = 0: one cubic step   e = 1 - d*x ; x += x*(e + e*e)              (3 DFMA)
// RCP_REFINE = 1: cubic + one Newton step                                      (5 DFMA)
// Measured on B200 (chrom_cuda_probe_rcp, 4M samples of +-[1e-3, 1e6]): seed 9.9e-7 (2^-20),
// cubic 2.2e-16 (<= 1 ulp), cubic+Newton 0 (= IEEE 1/x).  The kernels use the cubic step alone.
// The operands here are O(1) and far from the subnormal range, so no special-case path.
#ifndef CHROM_RCP_REFINE
#define CHROM_RCP_REFINE 0
#endif

__device__ __forceinline__ double rcp_seed(double d) {
  double x;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(x) : "d"(d));
  return x;
}

template <int REFINE = CHROM_RCP_REFINE>
__device__ __forceinline__ double fast_rcp(double d) {
  double x = rcp_seed(d);
  double e = fma(-d, x, 1.0);
  double t = fma(e, e, e);
  x = fma(x, t, x);
  if (REFINE >= 1) {
    e = fma(-d, x, 1.0);
    x = fma(x, e, x);
  }
  return x;
}

// Non-finite test on the integer pipe: exponent field all ones.
__device__ __forceinline__ unsigned abs_hi(double v) { return (unsigned)__double2hiint(v) & 0x7fffffffu; }
__device__ __forceinline__ bool hi_nonfinite(unsigned ahi) { return ahi >= 0x7ff00000u; }

}  // namespace chrom
