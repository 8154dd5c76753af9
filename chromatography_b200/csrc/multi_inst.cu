// multi_inst.cu — one translation unit per species count (compiled with -DCHROM_MULTI_N=0..8, 0 = runtime n)
// so that the nine instantiation sets of multi_kernels.cuh build in parallel.
#include "multi_kernels.cuh"

#define CHROM_CAT2(a, b) a##b
#define CHROM_CAT(a, b) CHROM_CAT2(a, b)

namespace chrom {
typedef void (*mkern_t)(const BatchDev, double*);
mkern_t CHROM_CAT(multi_pick_, CHROM_MULTI_N)(int method, int arith) { return pick_n<CHROM_MULTI_N>(method, arith); }
}  // namespace chrom
