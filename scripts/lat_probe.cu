// lat_probe.cu — dependent-issue latencies on the FP64 path of a B200 SM (one warp, clock64 around an unrolled chain):
// DFMA -> DFMA, DMUL -> DMUL, DADD -> DADD, MUFU.RCP64H -> DFMA, and SHFL.UP (64-bit) round trip.  The numbers size the
// instruction-level parallelism the register-resident LangmuirMulti kernels need per warp.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/_bin/lat_probe scripts/lat_probe.cu
#include <cstdio>
#include <cuda_runtime.h>

constexpr int CH = 256;

__global__ void k_lat(double a, double b, long long* out, double* sink) {
  double x = a;
  long long t0 = clock64();
#pragma unroll
  for (int i = 0; i < CH; ++i) x = fma(x, b, a);
  long long t1 = clock64();
  double y = x;
#pragma unroll
  for (int i = 0; i < CH; ++i) y = y * b;
  long long t2 = clock64();
  double z = y;
#pragma unroll
  for (int i = 0; i < CH; ++i) z = z + b;
  long long t3 = clock64();
  double r = z;
#pragma unroll
  for (int i = 0; i < CH; ++i) {
    double s;
    asm volatile("rcp.approx.ftz.f64 %0, %1;" : "=d"(s) : "d"(r));
    r = fma(s, b, a);
  }
  long long t4 = clock64();
  double w = r;
#pragma unroll
  for (int i = 0; i < CH; ++i) w = __shfl_up_sync(0xffffffffu, w, 1) + b;
  long long t5 = clock64();
  // two independent DFMA chains: issue-limited rate of one warp
  double p = w, q = a;
#pragma unroll
  for (int i = 0; i < CH; ++i) {
    p = fma(p, b, a);
    q = fma(q, b, a);
  }
  long long t6 = clock64();
  double e[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) e[k] = p + k;
#pragma unroll
  for (int i = 0; i < CH; ++i)
#pragma unroll
    for (int k = 0; k < 8; ++k) e[k] = fma(e[k], b, a);
  long long t7 = clock64();
  if (threadIdx.x == 0) {
    out[0] = t1 - t0;
    out[1] = t2 - t1;
    out[2] = t3 - t2;
    out[3] = t4 - t3;
    out[4] = t5 - t4;
    out[5] = t6 - t5;
    out[6] = t7 - t6;
  }
  double acc = q;
#pragma unroll
  for (int k = 0; k < 8; ++k) acc += e[k];
  sink[threadIdx.x] = acc;
}

int main() {
  long long* out;
  double* sink;
  cudaMallocManaged(&out, 8 * sizeof(long long));
  cudaMalloc(&sink, 1024 * sizeof(double));
  for (int warps = 1; warps <= 4; warps *= 2) {
    k_lat<<<1, 32 * warps>>>(1.0000001, 0.9999999, out, sink);
    cudaDeviceSynchronize();
    k_lat<<<1, 32 * warps>>>(1.0000001, 0.9999999, out, sink);
    if (cudaDeviceSynchronize() != cudaSuccess) return 1;
    printf("{\"warps_in_cta\": %d, \"dfma_dep_cycles\": %.2f, \"dmul_dep_cycles\": %.2f, \"dadd_dep_cycles\": %.2f, "
           "\"mufu_rcp64h_plus_dfma_cycles\": %.2f, \"shfl64_plus_dadd_cycles\": %.2f, \"dfma_2chains_cycles_per_pair\": %.2f, "
           "\"dfma_8chains_cycles_per_8\": %.2f}\n",
           warps, out[0] / (double)CH, out[1] / (double)CH, out[2] / (double)CH, out[3] / (double)CH, out[4] / (double)CH,
           out[5] / (double)CH, out[6] / (double)CH);
  }
  return 0;
}
