// multi_resident.cu — host side of the shared-memory-resident LangmuirMulti integrator: kernel lookup,
// launch geometry and launch.  The device code lives in multi_kernels.cuh and is instantiated per species
// count by multi_inst.cu (one object per n, built in parallel).
#include <algorithm>
#include <cstdio>
#include <cstdlib>

#include "kernels.h"

namespace chrom {

typedef void (*mkern_t)(const BatchDev);
mkern_t multi_pick_1(int method, int arith);
mkern_t multi_pick_2(int method, int arith);
mkern_t multi_pick_3(int method, int arith);
mkern_t multi_pick_4(int method, int arith);
mkern_t multi_pick_5(int method, int arith);
mkern_t multi_pick_6(int method, int arith);
mkern_t multi_pick_7(int method, int arith);
mkern_t multi_pick_8(int method, int arith);
mkern_t multi_pick_9(int method, int arith);
mkern_t multi_pick_10(int method, int arith);
mkern_t multi_pick_11(int method, int arith);
mkern_t multi_pick_12(int method, int arith);
mkern_t multi_pick_13(int method, int arith);
mkern_t multi_pick_14(int method, int arith);
mkern_t multi_pick_15(int method, int arith);
mkern_t multi_pick_16(int method, int arith);

namespace {

// bytes of one SpeciesConst in shared memory (multi_kernels.cuh: 10 doubles, 16-byte aligned)
constexpr size_t kSpeciesConstBytes = 10 * sizeof(double);

constexpr int multi_max_threads(int N, int ARITH) {  // = multi_kernels.cuh
  return ARITH == CHROM_KARITH_STRUCT ? (N <= 8 ? 512 : 256) : (N <= 2 ? 512 : 256);
}

mkern_t lookup_multi(int n, int method, int arith) {
  switch (n) {
    case 1: return multi_pick_1(method, arith);
    case 2: return multi_pick_2(method, arith);
    case 3: return multi_pick_3(method, arith);
    case 4: return multi_pick_4(method, arith);
    case 5: return multi_pick_5(method, arith);
    case 6: return multi_pick_6(method, arith);
    case 7: return multi_pick_7(method, arith);
    case 8: return multi_pick_8(method, arith);
    case 9: return multi_pick_9(method, arith);
    case 10: return multi_pick_10(method, arith);
    case 11: return multi_pick_11(method, arith);
    case 12: return multi_pick_12(method, arith);
    case 13: return multi_pick_13(method, arith);
    case 14: return multi_pick_14(method, arith);
    case 15: return multi_pick_15(method, arith);
    case 16: return multi_pick_16(method, arith);
  }
  return nullptr;
}

size_t multi_smem_bytes(int n, uint32_t nz) {
  return ((size_t)4 * n * nz + 3 * n + kSummSlots * n) * sizeof(double) + (size_t)n * kSpeciesConstBytes;
}

}  // namespace

// Resolves the kernel arithmetic of a multi-species batch.  FAST (and FAST_STRUCTURED) = the closed-form
// structured elimination for every n; FAST_DENSE = the dense register LU, which exists for n <= 8 only.
int multi_kernel_arith(int n, int arith) {
  if (arith == CHROM_ARITH_EXACT) return arith;
  if (arith == CHROM_ARITH_FAST_DENSE && n <= 8) return CHROM_ARITH_FAST;  // kernel code 0 = dense register LU
  return CHROM_KARITH_STRUCT;
}

bool multi_resident_choose(const BatchDev& b, int sm_count, LaunchGeom* g, std::string* why) {
  const int n = (int)b.n_species;
  // thread per cell: dense / exact n <= 8 (the n x n system in registers), closed-form elimination n <= 16
  if (n < 1 || n > 16 || (n > 8 && b.arith != CHROM_KARITH_STRUCT)) {
    if (why) *why = "this species count / arithmetic runs one warp per cell (multi_tiled)";
    return false;
  }
  const size_t smem = multi_smem_bytes(n, b.nz);
  if (smem > 227u * 1024u) {
    if (why) *why = "nz * n_species exceeds the shared-memory-resident capacity (4*n*nz doubles <= 227 KB)";
    return false;
  }
  const int karith = b.arith;  // already resolved by multi_kernel_arith
  const int tmax = multi_max_threads(n, karith);
  int T = (int)std::min<uint32_t>((b.nz + 31u) / 32u * 32u, (uint32_t)tmax);
  // balance: fewest threads that keep the same number of cells per thread
  const int cpt = (int)((b.nz + T - 1) / T);
  T = (int)(((b.nz + cpt - 1) / cpt + 31u) / 32u * 32u);
  int ctas_per_sm;
  {
    const int regs = karith == CHROM_KARITH_STRUCT ? (n <= 8 ? 128 : 255) : (n >= 8 ? 232 : (n >= 5 ? 200 : 128));
    const int by_smem = (int)((227u * 1024u) / (smem + 1024));
    const int by_regs = 65536 / (T * regs);
    ctas_per_sm = std::max(1, std::min(by_smem, by_regs));
  }
  g->kind = 4;
  g->M = cpt;
  g->L = T;
  g->cpw = 1;
  g->warps_per_col = T / 32;
  g->block = dim3(T);
  g->cols_per_wave = (uint32_t)(sm_count * ctas_per_sm);
  g->grid = dim3(std::min<uint32_t>(b.n_cols, g->cols_per_wave));
  g->smem = smem;
  char buf[240];
  snprintf(buf, sizeof buf, "multi_resident<n=%d,%s,%s> threads/col=%d cells/thread=%d smem=%zuB persistent grid=%u", n,
           b.method == CHROM_RK4 ? "RK4" : "Euler",
           karith == CHROM_ARITH_EXACT ? "exact" : (karith == CHROM_KARITH_STRUCT ? "fast-structured-LU" : "fast-dense-LU"),
           T, cpt, smem, g->grid.x);
  g->name = buf;
  return true;
}

cudaError_t multi_resident_launch(const BatchDev& b, const LaunchGeom& g, cudaStream_t s) {
  mkern_t k = lookup_multi((int)b.n_species, b.method, b.arith);
  if (!k) return cudaErrorInvalidValue;
  cudaError_t e = cudaFuncSetAttribute((const void*)k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)g.smem);
  if (e != cudaSuccess) return e;
  const uint32_t grid = std::min<uint32_t>(b.n_cols, g.cols_per_wave ? g.cols_per_wave : b.n_cols);
  k<<<dim3(grid), g.block, g.smem, s>>>(b);
  return cudaGetLastError();
}

}  // namespace chrom
