// pmp_cloud_build.h — host entry point of the CloudSDF device build (pmp_cloud_build.cu).
#pragma once
#include <string>

#include "pmp_cloud.cuh"

namespace pmp {

struct CloudBuffers {  // device allocations owned by the scene handle
  float4* pts = nullptr;
  float4* nodes = nullptr;
  float* dist = nullptr;
};

// Builds the Morton-sorted points, the implicit 8-ary box tree and the distance-bound field on the
// current device.  Returns 0 on success; on failure everything is freed and *err is set.
int cloud_build(const float* host_pts, int64_t n, float rho, float cell, int num_sms, CloudDev* out, CloudBuffers* bufs,
                int* n_launches, std::string* err);
void cloud_free(CloudBuffers* bufs);

}  // namespace pmp
