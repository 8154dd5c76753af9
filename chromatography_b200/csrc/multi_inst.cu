// multi_inst.cu — one translation unit per species count (compiled with -DCHROM_MULTI_N=0..16, 0 = runtime n)
// so that the nine instantiation sets of multi_kernels.cuh / multi_tiled.cuh build in parallel.
#include "multi_reg.cuh"
#include "multi_reg2.cuh"
#include "multi_reg3.cuh"
#include "multi_tiled.cuh"

#define CHROM_CAT2(a, b) a##b
#define CHROM_CAT(a, b) CHROM_CAT2(a, b)

namespace chrom {
typedef void (*mkern_t)(const BatchDev);
typedef void (*tkern_t)(const BatchDev, double*, const TiledArgs);
typedef void (*rkern_t)(const BatchDev);
// shared-memory-resident and tiled kernels: n = 1..8 per thread (9..16: closed form only), runtime n (0) per warp
mkern_t CHROM_CAT(multi_pick_, CHROM_MULTI_N)(int method, int arith) { return pick_n<CHROM_MULTI_N>(method, arith); }
tkern_t CHROM_CAT(multi_tiled_pick_, CHROM_MULTI_N)(int method, int arith, int wj) {
  return pick_tiled_n<CHROM_MULTI_N>(method, arith, wj);
}
#if CHROM_MULTI_N > 0  // register-resident kernels: n = 1..16
rkern_t CHROM_CAT(multi_reg_pick_, CHROM_MULTI_N)(int method, int m) { return pick_reg_n<CHROM_MULTI_N>(method, m); }
int CHROM_CAT(multi_reg_tmax_, CHROM_MULTI_N)(int m) {  // threads per CTA; negative: acc lives in shared memory
  const int t = reg_max_threads_rt<CHROM_MULTI_N>(m);
  return reg_acc_shared_rt<CHROM_MULTI_N>(m) ? -t : t;
}
#endif
#if CHROM_MULTI_N > 0 && CHROM_MULTI_N <= 8  // second-generation register-resident kernels: n = 1..8
rkern_t CHROM_CAT(multi_reg2_pick_, CHROM_MULTI_N)(int method) { return pick_reg2_n<CHROM_MULTI_N>(method); }
int CHROM_CAT(multi_reg2_tmax_, CHROM_MULTI_N)() { return reg2_tmax_rt<CHROM_MULTI_N>(); }
// third generation (multi_reg3.cuh): same geometry, the pivoting guard per column run; sync = the seam protocol
rkern_t CHROM_CAT(multi_reg3_pick_, CHROM_MULTI_N)(int method, int sync) { return pick_reg3_n<CHROM_MULTI_N>(method, sync); }
#endif
#if CHROM_MULTI_N == 0
void multi_tiled_init_launch(const BatchDev& b, double* ping, unsigned long long* code, cudaStream_t s) {
  k_tiled_init<<<296, 256, 0, s>>>(b, ping, code);
}
void multi_tiled_finalize_launch(const BatchDev& b, const unsigned long long* code, const double* gsumm, cudaStream_t s) {
  const uint32_t items = b.n_cols * b.n_species;
  k_tiled_finalize<<<(items + 127) / 128, 128, 0, s>>>(b, code, gsumm);
}
#endif
}  // namespace chrom
