// TEST INFRASTRUCTURE (see cuda_emul.h): the pieces of the library that are not emulated.
#include <string>

#include "pmp_cloud_build.h"

namespace pmp {
// the shared-memory arena of the one CTA that runs at a time (1 MB); the kernels' block-scope
// `extern float4 sm[]` declarations name this variable of the enclosing namespace
float4 sm[1 << 16];

// the cloud index is built with cub::DeviceRadixSort on the device: not available in the emulation
int cloud_build(const float*, int64_t, float, float, int, CloudDev*, CloudBuffers*, CloudScratch*, int* n_launches, std::string* err) {
  *n_launches = 0;
  *err = "point clouds are not supported by the SIMT emulation";
  return -1;
}
void cloud_free(CloudBuffers*) {}
void cloud_free(CloudScratch*) {}
}  // namespace pmp
