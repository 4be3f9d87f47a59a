// pmp_cloud_build.cu — device-side construction of the CloudSDF structure (see pmp_cloud.cuh):
// Morton sort (cub::DeviceRadixSort, setup time only), implicit 8-ary box tree, distance-bound field.
// Reference counterpart: KDTree::KDTree / build (collision/kdtree.cpp:20-55) and the CloudSDF
// constructor (collision/primitive_sdf.hpp:571-583), both CPU and setup-time as well.
#include <algorithm>
#include <cmath>
#include <cub/device/device_radix_sort.cuh>
#include <string>
#include <vector>

#include "pmp_cloud_build.h"

namespace pmp {
namespace {

__device__ __forceinline__ unsigned spread10(unsigned v) {  // 10 bits -> every third bit
  v &= 0x3ffu;
  v = (v | (v << 16)) & 0x030000ffu;
  v = (v | (v << 8)) & 0x0300f00fu;
  v = (v | (v << 4)) & 0x030c30c3u;
  v = (v | (v << 2)) & 0x09249249u;
  return v;
}

__global__ void k_cloud_morton(const float* __restrict__ pts3, int64_t n, float3 lo, float3 scale, unsigned* keys,
                               unsigned* vals) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float x = pts3[3 * i], y = pts3[3 * i + 1], z = pts3[3 * i + 2];
    const unsigned qx = (unsigned)fminf(fmaxf((x - lo.x) * scale.x, 0.f), 1023.f);
    const unsigned qy = (unsigned)fminf(fmaxf((y - lo.y) * scale.y, 0.f), 1023.f);
    const unsigned qz = (unsigned)fminf(fmaxf((z - lo.z) * scale.z, 0.f), 1023.f);
    keys[i] = spread10(qx) | (spread10(qy) << 1) | (spread10(qz) << 2);
    vals[i] = (unsigned)i;
  }
}

// sorted[i] = point vals[i] (w = its original index); the tail up to a multiple of 8 is +inf
__global__ void k_cloud_gather(const float* __restrict__ pts3, int64_t n, int64_t n_pad, const unsigned* __restrict__ vals,
                               const unsigned char* __restrict__ cls, float4* sorted) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n_pad; i += (int64_t)gridDim.x * blockDim.x) {
    if (i < n) {
      const unsigned s = vals[i];
      const unsigned w = s | (cls ? ((unsigned)cls[s] << kCloudClassShift) : 0u);  // radius class in bits 28..30
      sorted[i] = make_float4(pts3[3 * (int64_t)s], pts3[3 * (int64_t)s + 1], pts3[3 * (int64_t)s + 2], __int_as_float((int)w));
    } else {
      sorted[i] = make_float4(INFINITY, INFINITY, INFINITY, __int_as_float(-1));
    }
  }
}

// level 0: boxes of 8 consecutive points; padding nodes are (+inf, -inf)
__global__ void k_cloud_leaf_boxes(const float4* __restrict__ pts, int64_t n, int n_nodes, int n_nodes_pad, float4* nodes) {
  for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < n_nodes_pad; j += gridDim.x * blockDim.x) {
    float3 lo = make_float3(INFINITY, INFINITY, INFINITY), hi = make_float3(-INFINITY, -INFINITY, -INFINITY);
    if (j < n_nodes)
      for (int k = 0; k < 8; k++) {
        const int64_t i = 8 * (int64_t)j + k;
        if (i >= n) break;
        const float4 p = pts[i];
        lo.x = fminf(lo.x, p.x); lo.y = fminf(lo.y, p.y); lo.z = fminf(lo.z, p.z);
        hi.x = fmaxf(hi.x, p.x); hi.y = fmaxf(hi.y, p.y); hi.z = fmaxf(hi.z, p.z);
      }
    nodes[2 * j] = make_float4(lo.x, lo.y, lo.z, 0.f);
    nodes[2 * j + 1] = make_float4(hi.x, hi.y, hi.z, 0.f);
  }
}

// level k: union of 8 consecutive level k-1 boxes
__global__ void k_cloud_level_boxes(const float4* __restrict__ child, int n_child, int n_nodes, int n_nodes_pad, float4* nodes) {
  for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < n_nodes_pad; j += gridDim.x * blockDim.x) {
    float3 lo = make_float3(INFINITY, INFINITY, INFINITY), hi = make_float3(-INFINITY, -INFINITY, -INFINITY);
    if (j < n_nodes)
      for (int k = 0; k < 8; k++) {
        const int c = 8 * j + k;
        if (c >= n_child) break;
        const float4 a = child[2 * c], b = child[2 * c + 1];
        lo.x = fminf(lo.x, a.x); lo.y = fminf(lo.y, a.y); lo.z = fminf(lo.z, a.z);
        hi.x = fmaxf(hi.x, b.x); hi.y = fmaxf(hi.y, b.y); hi.z = fmaxf(hi.z, b.z);
      }
    nodes[2 * j] = make_float4(lo.x, lo.y, lo.z, 0.f);
    nodes[2 * j + 1] = make_float4(hi.x, hi.y, hi.z, 0.f);
  }
}

// dist[cell] = exact NN distance of the cell centre, capped at dist_cap
__global__ void k_cloud_field(CloudDev C, float* dist) {
  const int64_t n_cells = (int64_t)C.nx * C.ny * C.nz;
  const float cap2 = C.dist_cap * C.dist_cap;
  for (int64_t cell = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; cell < n_cells;
       cell += (int64_t)gridDim.x * blockDim.x) {
    const int cx = (int)(cell % C.nx), cy = (int)((cell / C.nx) % C.ny), cz = (int)(cell / ((int64_t)C.nx * C.ny));
    const float3 p = make_float3(C.ox + (cx + 0.5f) * C.h, C.oy + (cy + 0.5f) * C.h, C.oz + (cz + 0.5f) * C.h);
    const float d2 = cloud_nn_within(C, p, cap2, nullptr);
    dist[cell] = (d2 <= cap2) ? fminf(sqrtf(d2), C.dist_cap) : C.dist_cap;
  }
}

}  // namespace

void cloud_free(CloudBuffers* b) {
  if (b->pts) cudaFree(b->pts);
  if (b->nodes) cudaFree(b->nodes);
  if (b->dist) cudaFree(b->dist);
  *b = CloudBuffers{};
}

void cloud_free(CloudScratch* s) {
  if (s->raw) cudaFree(s->raw);
  if (s->keys) cudaFree(s->keys);
  if (s->keys2) cudaFree(s->keys2);
  if (s->vals) cudaFree(s->vals);
  if (s->vals2) cudaFree(s->vals2);
  if (s->tmp) cudaFree(s->tmp);
  if (s->cls) cudaFree(s->cls);
  *s = CloudScratch{};
}

namespace {
template <class T>
cudaError_t reserve(T** p, size_t* cap, size_t bytes) {  // grow-only; 25 % headroom so that frames of
  if (*cap >= bytes) return cudaSuccess;                 // slightly different sizes do not reallocate
  if (*p) cudaFree(*p);
  *p = nullptr;
  *cap = 0;
  const size_t want = bytes + bytes / 4 + 256;
  cudaError_t e = cudaMalloc(reinterpret_cast<void**>(p), want);
  if (e == cudaSuccess) *cap = want;
  return e;
}
}  // namespace

int cloud_build(const float* host_pts, int64_t n, float rho, const float* point_radii, float cell, int num_sms, CloudDev* Cout,
                CloudBuffers* bufs, CloudScratch* scr, int* n_launches, std::string* err) {
  auto bad = [&](const std::string& m, cudaError_t e) {
    *err = m + (e == cudaSuccess ? "" : std::string(": ") + cudaGetErrorString(e));
    return -1;
  };
  *n_launches = 0;
  CloudDev C{};
  float lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
  for (int64_t i = 0; i < n; i++)
    for (int a = 0; a < 3; a++) {
      const float v = host_pts[3 * i + a];
      if (!std::isfinite(v)) return bad("non-finite cloud point", cudaSuccess);
      lo[a] = std::min(lo[a], v);
      hi[a] = std::max(hi[a], v);
    }
  // radius classes (clouds of different radii merged into this index): ascending distinct radii
  std::vector<float> classes;
  std::vector<unsigned char> cls_host;
  if (point_radii) {
    for (int64_t i = 0; i < n; i++) {
      const float r = point_radii[i];
      if (!std::isfinite(r)) return bad("non-finite cloud point radius", cudaSuccess);
      if (std::find(classes.begin(), classes.end(), r) == classes.end()) {
        if ((int)classes.size() == kCloudMaxClasses) return bad("more than 4 distinct cloud radii in one scene", cudaSuccess);
        classes.push_back(r);
      }
    }
    std::sort(classes.begin(), classes.end());
  }
  if (classes.size() > 1) {
    if (n >= ((int64_t)1 << kCloudClassShift)) return bad("too many cloud points for radius classes", cudaSuccess);
    cls_host.resize((size_t)n);
    for (int64_t i = 0; i < n; i++)
      cls_host[(size_t)i] = (unsigned char)(std::find(classes.begin(), classes.end(), point_radii[i]) - classes.begin());
    rho = classes.back();
  } else if (classes.size() == 1) {
    rho = classes[0];
  }
  C.n_classes = classes.size() > 1 ? (int)classes.size() : 1;
  C.rho_min = classes.size() > 1 ? classes.front() : rho;
  for (int k = 0; k < kCloudMaxClasses; k++) C.rho_cls[k] = k < (int)classes.size() ? classes[(size_t)k] : rho;
  // (with several classes the AABB is the point box grown by the LARGEST radius: a superset of the union of
  // the leaves' own AABBs — the AABB test is a conservative cull — and lb + rho is still the exact point box)
  C.rho = rho;
  C.lbx = lo[0] - rho; C.lby = lo[1] - rho; C.lbz = lo[2] - rho;
  C.ubx = hi[0] + rho; C.uby = hi[1] + rho; C.ubz = hi[2] + rho;
  C.n_pts = n;

  // ---- tree shape ----
  std::vector<int> cnt;
  cnt.push_back((int)((n + 7) / 8));
  while (cnt.back() > 1) cnt.push_back((cnt.back() + 7) / 8);
  if ((int)cnt.size() > kCloudMaxLevels) return bad("cloud too large for the box tree", cudaSuccess);
  C.n_levels = (int)cnt.size();
  int total_nodes = 0;
  std::vector<int> pad(cnt.size());
  for (size_t l = 0; l < cnt.size(); l++) {
    C.level_off[l] = total_nodes;
    pad[l] = (cnt[l] + 7) / 8 * 8;
    total_nodes += pad[l];
  }
  const int64_t n_pad = (n + 7) / 8 * 8;

  // ---- sort + boxes ----
  size_t tmp_bytes = 0;
  cudaError_t e = reserve(&scr->raw, &scr->cap_raw, (size_t)n * 12);
  if (e == cudaSuccess) e = cudaMemcpy(scr->raw, host_pts, (size_t)n * 12, cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = reserve(&scr->keys, &scr->cap_keys, (size_t)n * 4);
  if (e == cudaSuccess) e = reserve(&scr->keys2, &scr->cap_keys2, (size_t)n * 4);
  if (e == cudaSuccess) e = reserve(&scr->vals, &scr->cap_vals, (size_t)n * 4);
  if (e == cudaSuccess) e = reserve(&scr->vals2, &scr->cap_vals2, (size_t)n * 4);
  if (e == cudaSuccess && !cls_host.empty()) e = reserve(&scr->cls, &scr->cap_cls, (size_t)n);
  if (e == cudaSuccess && !cls_host.empty()) e = cudaMemcpy(scr->cls, cls_host.data(), (size_t)n, cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = reserve(&bufs->pts, &bufs->cap_pts, (size_t)n_pad * sizeof(float4));
  if (e == cudaSuccess) e = reserve(&bufs->nodes, &bufs->cap_nodes, (size_t)total_nodes * 2 * sizeof(float4));
  if (e == cudaSuccess)
    e = cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, scr->keys, scr->keys2, scr->vals, scr->vals2, (int)n, 0, 30);
  if (e == cudaSuccess) e = reserve(&scr->tmp, &scr->cap_tmp, std::max<size_t>(tmp_bytes, 16));
  if (e != cudaSuccess) return bad("cloud build allocation", e);
  float* raw = scr->raw;
  unsigned *keys = scr->keys, *keys2 = scr->keys2, *vals = scr->vals, *vals2 = scr->vals2;
  void* tmp = scr->tmp;
  const int grid = num_sms * 8;
  const float3 qlo = make_float3(lo[0], lo[1], lo[2]);
  const float3 qsc = make_float3(1023.999f / std::max(hi[0] - lo[0], 1e-20f), 1023.999f / std::max(hi[1] - lo[1], 1e-20f),
                                 1023.999f / std::max(hi[2] - lo[2], 1e-20f));
  k_cloud_morton<<<grid, 256>>>(raw, n, qlo, qsc, keys, vals);
  e = cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, keys, keys2, vals, vals2, (int)n, 0, 30);
  k_cloud_gather<<<grid, 256>>>(raw, n, n_pad, vals2, cls_host.empty() ? nullptr : scr->cls, bufs->pts);
  k_cloud_leaf_boxes<<<grid, 256>>>(bufs->pts, n, cnt[0], pad[0], bufs->nodes + 2 * (size_t)C.level_off[0]);
  *n_launches += 4;
  for (size_t l = 1; l < cnt.size(); l++) {
    k_cloud_level_boxes<<<std::max(1, std::min(grid, (pad[l] + 255) / 256)), 256>>>(
        bufs->nodes + 2 * (size_t)C.level_off[l - 1], cnt[l - 1], cnt[l], pad[l], bufs->nodes + 2 * (size_t)C.level_off[l]);
    *n_launches += 1;
  }
  if (e == cudaSuccess) e = cudaDeviceSynchronize();
  if (e != cudaSuccess) return bad("cloud tree build", e);
  C.pts = bufs->pts;
  C.nodes = bufs->nodes;

  // ---- distance-bound field ----
  const float margin = 0.35f;  // larger than any group-sphere query radius of the stock robots
  float h = cell > 0.f ? cell : 0.01f;
  const double ext[3] = {hi[0] - lo[0] + 2.0 * margin, hi[1] - lo[1] + 2.0 * margin, hi[2] - lo[2] + 2.0 * margin};
  while ((ext[0] / h + 1) * (ext[1] / h + 1) * (ext[2] / h + 1) > 64.0e6) h *= 1.25f;
  C.h = h;
  C.inv_h = 1.f / h;
  C.e = 0.5f * std::sqrt(3.f) * h * 1.0001f;
  C.dist_cap = margin;
  C.ox = lo[0] - margin; C.oy = lo[1] - margin; C.oz = lo[2] - margin;
  C.nx = (int)std::ceil(ext[0] / h) + 1;
  C.ny = (int)std::ceil(ext[1] / h) + 1;
  C.nz = (int)std::ceil(ext[2] / h) + 1;
  const int64_t n_cells = (int64_t)C.nx * C.ny * C.nz;
  e = reserve(&bufs->dist, &bufs->cap_dist, (size_t)n_cells * sizeof(float));
  if (e != cudaSuccess) return bad("cloud field allocation", e);
  k_cloud_field<<<num_sms * 16, 128>>>(C, bufs->dist);
  *n_launches += 1;
  e = cudaDeviceSynchronize();
  if (e != cudaSuccess) return bad("cloud field build", e);
  C.dist = bufs->dist;
  C.enabled = 1;
  *Cout = C;
  return 0;
}

}  // namespace pmp
