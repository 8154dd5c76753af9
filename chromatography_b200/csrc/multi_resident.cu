// multi_resident.cu — placeholder until the multi-species kernel lands.
#include "kernels.h"
namespace chrom {
bool multi_resident_choose(const BatchDev&, int, LaunchGeom*, std::string* why) {
  if (why) *why = "multi-species path not built yet";
  return false;
}
cudaError_t multi_resident_launch(const BatchDev&, const LaunchGeom&, cudaStream_t) { return cudaErrorNotSupported; }
}  // namespace chrom
