// probe.cu — measurement kernels: FP64 (DFMA) peak, device copy bandwidth, reciprocal accuracy.
// MEASURED_PEAKS.json (driver-written) carries HBM and bf16 peaks only; the FP64 roofline
// denominator for the register-resident kernels is measured here, in the same run as the bench.
#include <vector>

#include "arith.cuh"
#include "kernels.h"

namespace chrom {
namespace {

constexpr int kChains = 16;

__global__ void __launch_bounds__(256) k_dfma_peak(double* sink, int iters, double a, double b) {
  double x[kChains];
#pragma unroll
  for (int c = 0; c < kChains; ++c) x[c] = (double)(threadIdx.x + c) * 1e-3;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int c = 0; c < kChains; ++c) x[c] = fma(x[c], a, b);
  }
  double s = 0.0;
#pragma unroll
  for (int c = 0; c < kChains; ++c) s += x[c];
  if (s == 123.456) sink[0] = s;  // never true; keeps the chains alive
}

__global__ void __launch_bounds__(256) k_copy(const double2* __restrict__ src, double2* __restrict__ dst, size_t n) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) dst[i] = src[i];
}

__global__ void k_probe_rcp(double lo, double hi, uint32_t n, double* out /* [3][blocks] */) {
  __shared__ double red[3][256];
  double m0 = 0.0, m1 = 0.0, m2 = 0.0;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const double f = (double)i / (double)(n > 1 ? n - 1 : 1);
    double d = lo + (hi - lo) * f;
    if ((i & 1u) != 0u) d = -d;  // the kernels feed negative denominators too
    const double ref = 1.0 / d;
    const double e0 = fabs(rcp_seed(d) - ref) / fabs(ref);
    const double e1 = fabs(fast_rcp<0>(d) - ref) / fabs(ref);
    const double e2 = fabs(fast_rcp<1>(d) - ref) / fabs(ref);
    m0 = fmax(m0, e0);
    m1 = fmax(m1, e1);
    m2 = fmax(m2, e2);
  }
  red[0][threadIdx.x] = m0;
  red[1][threadIdx.x] = m1;
  red[2][threadIdx.x] = m2;
  __syncthreads();
  for (int s = blockDim.x / 2; s > 0; s >>= 1) {
    if ((int)threadIdx.x < s)
      for (int k = 0; k < 3; ++k) red[k][threadIdx.x] = fmax(red[k][threadIdx.x], red[k][threadIdx.x + s]);
    __syncthreads();
  }
  if (threadIdx.x == 0)
    for (int k = 0; k < 3; ++k) out[k * gridDim.x + blockIdx.x] = red[k][0];
}

}  // namespace

cudaError_t measure_fp64_peak(int sm_count, cudaStream_t s, double* tflops, double* ms) {
  double* sink = nullptr;
  cudaError_t e = cudaMalloc(&sink, sizeof(double));
  if (e != cudaSuccess) return e;
  cudaEvent_t a, b;
  cudaEventCreate(&a);
  cudaEventCreate(&b);
  const int blocks = sm_count * 8, threads = 256, iters = 20000;
  k_dfma_peak<<<blocks, threads, 0, s>>>(sink, 2000, 1.0000001, 1e-9);  // warm-up
  float best = 1e30f;
  for (int rep = 0; rep < 3; ++rep) {
    cudaEventRecord(a, s);
    k_dfma_peak<<<blocks, threads, 0, s>>>(sink, iters, 1.0000001, 1e-9);
    cudaEventRecord(b, s);
    e = cudaEventSynchronize(b);
    if (e != cudaSuccess) break;
    float t = 0.f;
    cudaEventElapsedTime(&t, a, b);
    if (t < best) best = t;
  }
  cudaEventDestroy(a);
  cudaEventDestroy(b);
  cudaFree(sink);
  if (e != cudaSuccess) return e;
  const double flops = 2.0 * (double)blocks * threads * (double)iters * kChains;
  *tflops = flops / ((double)best * 1e-3) / 1e12;
  *ms = best;
  return cudaGetLastError();
}

cudaError_t measure_copy_bw(size_t bytes, cudaStream_t s, double* gbs) {
  const size_t n = bytes / sizeof(double2);
  double2 *src = nullptr, *dst = nullptr;
  cudaError_t e = cudaMalloc(&src, n * sizeof(double2));
  if (e != cudaSuccess) return e;
  e = cudaMalloc(&dst, n * sizeof(double2));
  if (e != cudaSuccess) {
    cudaFree(src);
    return e;
  }
  cudaMemsetAsync(src, 0, n * sizeof(double2), s);
  cudaEvent_t a, b;
  cudaEventCreate(&a);
  cudaEventCreate(&b);
  k_copy<<<148 * 16, 256, 0, s>>>(src, dst, n);
  float best = 1e30f;
  for (int rep = 0; rep < 5; ++rep) {
    cudaEventRecord(a, s);
    k_copy<<<148 * 16, 256, 0, s>>>(src, dst, n);
    cudaEventRecord(b, s);
    e = cudaEventSynchronize(b);
    if (e != cudaSuccess) break;
    float t = 0.f;
    cudaEventElapsedTime(&t, a, b);
    if (t < best) best = t;
  }
  cudaEventDestroy(a);
  cudaEventDestroy(b);
  cudaFree(src);
  cudaFree(dst);
  if (e != cudaSuccess) return e;
  *gbs = 2.0 * (double)(n * sizeof(double2)) / ((double)best * 1e-3) / 1e9;
  return cudaGetLastError();
}

cudaError_t probe_rcp(double lo, double hi, uint32_t n, cudaStream_t s, double* max_rel_err) {
  const int blocks = 64;
  double* d = nullptr;
  cudaError_t e = cudaMalloc(&d, 3 * blocks * sizeof(double));
  if (e != cudaSuccess) return e;
  k_probe_rcp<<<blocks, 256, 0, s>>>(lo, hi, n, d);
  std::vector<double> h(3 * blocks);
  e = cudaMemcpyAsync(h.data(), d, h.size() * sizeof(double), cudaMemcpyDeviceToHost, s);
  if (e == cudaSuccess) e = cudaStreamSynchronize(s);
  cudaFree(d);
  if (e != cudaSuccess) return e;
  for (int k = 0; k < 3; ++k) {
    double m = 0.0;
    for (int i = 0; i < blocks; ++i) m = h[k * blocks + i] > m ? h[k * blocks + i] : m;
    max_rel_err[k] = m;
  }
  return cudaSuccess;
}

}  // namespace chrom
