// multi_reg.cu — host side of the register-resident LangmuirMulti integrator (device code: multi_reg.cuh).
#include <algorithm>
#include <cstdio>
#include <cstdlib>

#include "kernels.h"

namespace chrom {

typedef void (*rkern_t)(const BatchDev);
#define CHROM_DECL(N)                        \
  rkern_t multi_reg_pick_##N(int method, int m); \
  int multi_reg_tmax_##N(int m);
CHROM_DECL(1) CHROM_DECL(2) CHROM_DECL(3) CHROM_DECL(4) CHROM_DECL(5) CHROM_DECL(6) CHROM_DECL(7) CHROM_DECL(8)
CHROM_DECL(9) CHROM_DECL(10) CHROM_DECL(11) CHROM_DECL(12) CHROM_DECL(13) CHROM_DECL(14) CHROM_DECL(15) CHROM_DECL(16)
#undef CHROM_DECL
#define CHROM_DECL2(N)                                \
  rkern_t multi_reg2_pick_##N(int method);            \
  rkern_t multi_reg3_pick_##N(int method, int sync);  \
  int multi_reg2_tmax_##N();
CHROM_DECL2(1) CHROM_DECL2(2) CHROM_DECL2(3) CHROM_DECL2(4) CHROM_DECL2(5) CHROM_DECL2(6) CHROM_DECL2(7) CHROM_DECL2(8)
#undef CHROM_DECL2

namespace {

// second-generation kernels (multi_reg2.cuh), n <= 8: 2 cells per thread
rkern_t lookup_reg2(int n, int method) {
  switch (n) {
    case 1: return multi_reg2_pick_1(method);
    case 2: return multi_reg2_pick_2(method);
    case 3: return multi_reg2_pick_3(method);
    case 4: return multi_reg2_pick_4(method);
    case 5: return multi_reg2_pick_5(method);
    case 6: return multi_reg2_pick_6(method);
    case 7: return multi_reg2_pick_7(method);
    case 8: return multi_reg2_pick_8(method);
  }
  return nullptr;
}

// third-generation kernels (multi_reg3.cuh): the geometry of the second, guard per column run, seam protocol `sync`
rkern_t lookup_reg3(int n, int method, int sync) {
  switch (n) {
    case 1: return multi_reg3_pick_1(method, sync);
    case 2: return multi_reg3_pick_2(method, sync);
    case 3: return multi_reg3_pick_3(method, sync);
    case 4: return multi_reg3_pick_4(method, sync);
    case 5: return multi_reg3_pick_5(method, sync);
    case 6: return multi_reg3_pick_6(method, sync);
    case 7: return multi_reg3_pick_7(method, sync);
    case 8: return multi_reg3_pick_8(method, sync);
  }
  return nullptr;
}

int tmax_reg2(int n) {
  switch (n) {
    case 1: return multi_reg2_tmax_1();
    case 2: return multi_reg2_tmax_2();
    case 3: return multi_reg2_tmax_3();
    case 4: return multi_reg2_tmax_4();
    case 5: return multi_reg2_tmax_5();
    case 6: return multi_reg2_tmax_6();
    case 7: return multi_reg2_tmax_7();
    case 8: return multi_reg2_tmax_8();
  }
  return 0;
}

rkern_t lookup_reg(int n, int method, int m) {
  switch (n) {
    case 1: return multi_reg_pick_1(method, m);
    case 2: return multi_reg_pick_2(method, m);
    case 3: return multi_reg_pick_3(method, m);
    case 4: return multi_reg_pick_4(method, m);
    case 5: return multi_reg_pick_5(method, m);
    case 6: return multi_reg_pick_6(method, m);
    case 7: return multi_reg_pick_7(method, m);
    case 8: return multi_reg_pick_8(method, m);
    case 9: return multi_reg_pick_9(method, m);
    case 10: return multi_reg_pick_10(method, m);
    case 11: return multi_reg_pick_11(method, m);
    case 12: return multi_reg_pick_12(method, m);
    case 13: return multi_reg_pick_13(method, m);
    case 14: return multi_reg_pick_14(method, m);
    case 15: return multi_reg_pick_15(method, m);
    case 16: return multi_reg_pick_16(method, m);
  }
  return nullptr;
}

int tmax_reg(int n, int m) {
  switch (n) {
    case 1: return multi_reg_tmax_1(m);
    case 2: return multi_reg_tmax_2(m);
    case 3: return multi_reg_tmax_3(m);
    case 4: return multi_reg_tmax_4(m);
    case 5: return multi_reg_tmax_5(m);
    case 6: return multi_reg_tmax_6(m);
    case 7: return multi_reg_tmax_7(m);
    case 8: return multi_reg_tmax_8(m);
    case 9: return multi_reg_tmax_9(m);
    case 10: return multi_reg_tmax_10(m);
    case 11: return multi_reg_tmax_11(m);
    case 12: return multi_reg_tmax_12(m);
    case 13: return multi_reg_tmax_13(m);
    case 14: return multi_reg_tmax_14(m);
    case 15: return multi_reg_tmax_15(m);
    case 16: return multi_reg_tmax_16(m);
  }
  return 0;
}

}  // namespace

constexpr int kReg2MinSpecies = 5;

// Generation of the 2-cells-per-thread kernels (n <= 8): 3 = multi_reg3.cuh (guard per column run), 2 = multi_reg2.cuh.
// CHROM_REG2 = 2 / 3 forces one (experiments, A/B tests), CHROM_REG3_SYNC picks the seam protocol of the third.
constexpr int kReg3DefaultSync = 4;
static int reg_generation() {
  const char* v = getenv("CHROM_REG2");
  return v ? atoi(v) : 3;
}
static int reg3_sync() {
  const char* v = getenv("CHROM_REG3_SYNC");
  const int s = v ? atoi(v) : kReg3DefaultSync;
  return (s == 0 || s == 3 || s == 4) ? s : kReg3DefaultSync;  // the protocols multi_reg3.cuh builds
}

static bool multi_reg2_choose(const BatchDev& b, int sm_count, LaunchGeom* g) {
  const int n = (int)b.n_species;
  if (n < 1 || n > 8) return false;
  const int M = 2;
  const int gen = reg_generation() == 2 ? 2 : 3;
  const int sync = reg3_sync();
  const int tmax = tmax_reg2(n);
  const int T = (int)(((b.nz + M - 1) / M + 31u) / 32u * 32u);
  if (tmax == 0 || T > tmax) return false;
  rkern_t k = gen == 3 ? lookup_reg3(n, b.method, sync) : lookup_reg2(n, b.method);
  if (!k) return false;
  // seam ring: 8 slots x warps x row (n | 1 doubles, or n rounded up to even: n + 2 covers both); dense-fallback
  // scratch: warps x (3 n + n n)
  const size_t smem = ((size_t)8 * (T / 32) * (n + 2) + (size_t)(T / 32) * (3 * n + n * n)) * sizeof(double);
  int per_sm = 1;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, (const void*)k, T, smem) != cudaSuccess || per_sm < 1) {
    cudaGetLastError();
    per_sm = 1;
  }
  g->kind = 7;
  g->wj = gen == 3 ? 1 + sync : 0;  // which kernel multi_reg2_launch starts
  g->M = M;
  g->L = T;
  g->cpw = 1;
  g->warps_per_col = T / 32;
  g->block = dim3(T);
  g->cols_per_wave = (uint32_t)(sm_count * per_sm);
  g->grid = dim3(std::min<uint32_t>(b.n_cols, g->cols_per_wave));
  g->smem = smem;
  char buf[240];
  snprintf(buf, sizeof buf,
           "multi_reg%d<n=%d,%s,fast-scan-LU%s> threads/col=%d cells/thread=%d CTAs/SM=%d persistent grid=%u", gen, n,
           b.method == CHROM_RK4 ? "RK4" : "Euler",
           gen == 3 ? (sync == 0 ? ",seam=flags" : sync == 3 ? ",seam=mbarrier" : ",seam=mbarrier+split-step-barrier") : "", T, M,
           per_sm, g->grid.x);
  g->name = buf;
  return true;
}

cudaError_t multi_reg2_launch(const BatchDev& b, const LaunchGeom& g, cudaStream_t s) {
  rkern_t k = g.wj >= 1 ? lookup_reg3((int)b.n_species, b.method, g.wj - 1) : lookup_reg2((int)b.n_species, b.method);
  if (!k) return cudaErrorInvalidValue;
  const uint32_t grid = std::min<uint32_t>(b.n_cols, g.cols_per_wave ? g.cols_per_wave : b.n_cols);
  k<<<dim3(grid), g.block, g.smem, s>>>(b);
  return cudaGetLastError();
}

// Cells per thread M in {1, 2, 4}; CHROM_REG_M overrides (experiments).
bool multi_reg_choose(const BatchDev& b, int sm_count, LaunchGeom* g, std::string* why) {
  const int n = (int)b.n_species;
  if (n < 1 || n > 16) {
    if (why) *why = "register-resident multi kernel covers n_species 1..16";
    return false;
  }
  // Two cells per thread with acc in registers and the shuffle/seam exchange (multi_reg3.cuh; multi_reg2.cuh) where it
  // measured faster on B200 than the first generation below (profiles/r2_reg3_shapes.jsonl, cell-stage/s, third
  // generation vs first): every n >= 5 (the first generation has only one cell per thread there), and for batches that
  // fill the machine n = 4 (nz 200: 1.23e11 vs 1.09e11, nz 512: 1.57e11 vs 1.34e11), n = 3 (nz 100: 1.38e11 vs
  // 1.37e11) and n = 1 (2.74e11 vs 2.53e11); n = 2 stays with the first generation (nz 100: 1.83e11 vs 1.87e11, nz
  // 512: 2.47e11 vs 2.49e11), and so do batches too small to fill the machine (latency-bound: they want the most
  // threads per column, 1 cell per thread).  CHROM_REG2 = 0 forces the first generation, 2 / 3 the second / third for
  // every n <= 8 (experiments, A/B tests).
  {
    const int gen = reg_generation();
    const bool fills_machine = (uint64_t)b.n_cols >= (uint64_t)sm_count * 2u;
    const bool second = getenv("CHROM_REG2") ? gen != 0 : (n >= kReg2MinSpecies || (fills_machine && n != 2));
    if (second && multi_reg2_choose(b, sm_count, g)) return true;
  }
  const char* force = getenv("CHROM_REG_M");
  int bestM = 0, bestT = 0;
  // Measured on B200 (gpurun_out/multi_shapes_c.jsonl): a batch that fills the machine runs best with 2 cells per
  // thread for n <= 4 and 1 for n >= 5 (register pressure); a batch that cannot fill it is latency-bound and wants
  // the most threads per column (1 cell per thread).
  const bool fills = (uint64_t)b.n_cols >= (uint64_t)sm_count * 2u;
  const int order_small[3] = {2, 1, 4}, order_large[3] = {1, 2, 4}, order_latency[3] = {1, 2, 4};
  const int* order = !fills ? order_latency : (n <= 4 ? order_small : order_large);  // n > 12: only M = 1 exists
  for (int oi = 0; oi < 3; ++oi) {
    const int M = order[oi];
    if (force && atoi(force) != M) continue;
    const int tmax = std::abs(tmax_reg(n, M));
    if (tmax == 0) continue;
    const int T = (int)(((b.nz + M - 1) / M + 31u) / 32u * 32u);
    if (T > tmax) continue;
    bestM = M;
    bestT = T;
    break;
  }
  if (!bestM) {
    if (why) *why = "column too long for the register-resident multi kernel (3*M*n doubles per thread)";
    return false;
  }
  rkern_t k = lookup_reg(n, b.method, bestM);
  if (!k) return false;
  // neighbour exchange xch[2][T + 1][n | 1], then the acc slots [T][(M n) | 1] where acc lives in shared memory
  const size_t smem = ((size_t)2 * (bestT + 1) * (n | 1) +
                       (tmax_reg(n, bestM) < 0 ? (size_t)bestT * ((bestM * n) | 1) : 0)) * sizeof(double);
  if (smem > 48u * 1024u &&
      cudaFuncSetAttribute((const void*)k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  int per_sm = 1;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, (const void*)k, bestT, smem) != cudaSuccess || per_sm < 1) {
    cudaGetLastError();
    per_sm = 1;
  }
  g->kind = 6;
  g->M = bestM;
  g->L = bestT;
  g->cpw = 1;
  g->warps_per_col = bestT / 32;
  g->block = dim3(bestT);
  g->cols_per_wave = (uint32_t)(sm_count * per_sm);
  g->grid = dim3(std::min<uint32_t>(b.n_cols, g->cols_per_wave));
  g->smem = smem;
  char buf[240];
  snprintf(buf, sizeof buf, "multi_reg<n=%d,%s,fast-scan-LU> threads/col=%d cells/thread=%d CTAs/SM=%d persistent grid=%u", n,
           b.method == CHROM_RK4 ? "RK4" : "Euler", bestT, bestM, per_sm, g->grid.x);
  g->name = buf;
  return true;
}

cudaError_t multi_reg_launch(const BatchDev& b, const LaunchGeom& g, cudaStream_t s) {
  rkern_t k = lookup_reg((int)b.n_species, b.method, g.M);
  if (!k) return cudaErrorInvalidValue;
  const uint32_t grid = std::min<uint32_t>(b.n_cols, g.cols_per_wave ? g.cols_per_wave : b.n_cols);
  k<<<dim3(grid), g.block, g.smem, s>>>(b);
  return cudaGetLastError();
}

}  // namespace chrom
