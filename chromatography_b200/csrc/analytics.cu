// analytics.cu — on-device chromatogram analytics that need a whole series at once.
//
// rsf: the surface resolution of the reference's validation suite (validation/dissimilarity.rs:173-203, used for the
// Euler-vs-RK4 criterion validation/main.rs:430-440) between two outlet series that sit in HBM on ONE time grid:
//   Y = c / trapezoid(t, c)   (normalize, :55-67),   Rsf = trapezoid(t, |Y_a - Y_b|)   (:81-88).
// With one grid the reference's merge / resample steps are the identity (interpolate_at returns the node value with
// alpha = 0), so this is the same function; the sums are tree-reduced instead of sequential (differences ~1e-16).
// One warp per series, two passes over it (areas, then the integral): ~2 x 16 B per sample, HBM-bound.
#include <cfloat>

#include "kernels.h"

namespace chrom {
namespace {

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__global__ void k_rsf(const double* __restrict__ a, const double* __restrict__ b, size_t n_series, uint32_t n_out,
                      const uint32_t* __restrict__ steps_list, uint32_t stride, double dt, double* __restrict__ out) {
  const size_t series = (size_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (series >= n_series) return;
  const double* pa = a + series * n_out;
  const double* pb = b + series * n_out;
  auto t_of = [&](uint32_t k) -> double {  // time_points of the sampled steps: rk4.rs:254,327
    const uint32_t step = steps_list ? steps_list[k] : k * stride;
    return (double)step * dt;
  };
  double area_a = 0.0, area_b = 0.0;
  for (uint32_t k = lane; k + 1 < n_out; k += 32) {
    const double w = 0.5 * (t_of(k + 1) - t_of(k));
    area_a = fma(w, pa[k] + pa[k + 1], area_a);
    area_b = fma(w, pb[k] + pb[k + 1], area_b);
  }
  area_a = warp_sum(area_a);
  area_b = warp_sum(area_b);
  double r = 0.0;
  for (uint32_t k = lane; k + 1 < n_out; k += 32) {
    const double w = 0.5 * (t_of(k + 1) - t_of(k));
    const double d0 = fabs(pa[k] / area_a - pb[k] / area_b), d1 = fabs(pa[k + 1] / area_a - pb[k + 1] / area_b);
    r = fma(w, d0 + d1, r);
  }
  r = warp_sum(r);
  // the reference panics on a signal of (near) zero area ("cannot normalise a zero signal"): NaN here
  if (!(fabs(area_a) > DBL_EPSILON) || !(fabs(area_b) > DBL_EPSILON)) r = __longlong_as_double(0x7ff8000000000000ll);
  if (lane == 0) out[series] = r;
}

}  // namespace

cudaError_t rsf_launch(const double* a, const double* b, size_t n_series, uint32_t n_out, const uint32_t* steps_list,
                       uint32_t stride, double dt, double* out, cudaStream_t s) {
  const int warps = 8;
  k_rsf<<<(unsigned)((n_series + warps - 1) / warps), warps * 32, 0, s>>>(a, b, n_series, n_out, steps_list, stride, dt, out);
  return cudaGetLastError();
}

}  // namespace chrom
