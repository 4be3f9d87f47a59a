// pmp_cloud_build.h — host entry point of the CloudSDF device build (pmp_cloud_build.cu).
#pragma once
#include <string>

#include "pmp_cloud.cuh"

namespace pmp {

struct CloudBuffers {  // device allocations owned by the scene handle; grow-only across rebuilds
  float4* pts = nullptr;
  float4* nodes = nullptr;
  float* dist = nullptr;
  size_t cap_pts = 0, cap_nodes = 0, cap_dist = 0;  // bytes
};

struct CloudScratch {  // build-time temporaries, kept by the scene handle so that a re-index of a new
  float* raw = nullptr;  // depth frame does not pay (erratic, synchronising) cudaMalloc / cudaFree
  unsigned *keys = nullptr, *keys2 = nullptr, *vals = nullptr, *vals2 = nullptr;
  void* tmp = nullptr;
  unsigned char* cls = nullptr;  // radius class per point (only when the scene merges clouds of different radii)
  size_t cap_raw = 0, cap_keys = 0, cap_keys2 = 0, cap_vals = 0, cap_vals2 = 0, cap_tmp = 0, cap_cls = 0;
};

// Builds the Morton-sorted points, the implicit 8-ary box tree and the distance-bound field on the
// current device into `bufs` (grown when too small), using `scratch` for the temporaries.  Returns 0
// on success; on failure *err is set (the buffers stay owned by the caller).
// `point_radii` (optional, one per point): the CloudSDF radius each point came from; at most
// kCloudMaxClasses distinct values.  nullptr = every point has radius `rho`.
int cloud_build(const float* host_pts, int64_t n, float rho, const float* point_radii, float cell, int num_sms, CloudDev* out,
                CloudBuffers* bufs, CloudScratch* scratch, int* n_launches, std::string* err);
void cloud_free(CloudBuffers* bufs);
void cloud_free(CloudScratch* scratch);

}  // namespace pmp
