// single_stream.cu — placeholder until the streaming kernel lands.
#include "kernels.h"
namespace chrom {
bool single_stream_choose(const BatchDev&, int, LaunchGeom*, std::string* why) {
  if (why) *why += "; streaming path not built yet";
  return false;
}
cudaError_t single_stream_launch(const BatchDev&, const LaunchGeom&, cudaStream_t) { return cudaErrorNotSupported; }
}  // namespace chrom
