// pmp_cloud.cuh — CloudSDF on the GPU (reference: collision/primitive_sdf.hpp:568-603 and the
// kd-tree it wraps, collision/kdtree.cpp:20-96).
//
// The reference answers sdf(p) = sqrt(NN^2(p)) - rho with a recursive kd-tree descent (pointer
// chasing, one query at a time).  Any exact nearest-neighbour structure gives identical values
// (SURVEY App. A.7), so the device structure is chosen for the GPU instead:
//
//   * a uniform grid over the cloud's bounding box inflated by `margin`; points are counting-
//     sorted by cell on the device (build kernels below) -> cell_start[] + sorted float4 points
//     (w = original index), read with coalesced 16-byte loads;
//   * a per-cell distance bound field dist[cell] = exact NN distance of the CELL CENTRE,
//     computed on the device at build time.  For a query p in that cell
//         dist - e <= NN(p) <= dist + e,   e = h * sqrt(3) / 2,
//     which answers "any point within d of p" in O(1) for every query outside the +-e band and
//     bounds the exact search radius inside it.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace pmp {

struct CloudDev {
  const float4* pts;        // sorted by cell; w = original index (as int bits)
  const uint32_t* cell_start;  // n_cells + 1
  const float* dist;        // n_cells: NN distance of the cell centre (capped at dist_cap)
  int64_t n_pts;
  int nx, ny, nz;
  float ox, oy, oz;  // grid origin (bbox low corner - margin)
  float h, inv_h;
  float e;           // half cell diagonal
  float dist_cap;
  float rho;         // CloudSDF::radius_
  float lbx, lby, lbz, ubx, uby, ubz;  // CloudSDF AABB = point bbox +- rho (primitive_sdf.hpp:573-582)
  int enabled;
};

__device__ __forceinline__ int cloud_cell_coord(float v, float o, float inv_h, int n) {
  int c = (int)floorf((v - o) * inv_h);
  return c < 0 ? 0 : (c >= n ? n - 1 : c);
}

// exact: squared distance to the nearest cloud point if it is <= bound2, else returns a value
// > bound2 (not necessarily the true NN).  Scans the cells overlapping the cube of half-width
// sqrt(bound2) around p.  `idx` receives the original index of the best point (or -1).
__device__ inline float cloud_nn_within(const CloudDev& C, const float3 p, float bound2, int* idx) {
  const float b = sqrtf(bound2);
  const int x0 = cloud_cell_coord(p.x - b, C.ox, C.inv_h, C.nx), x1 = cloud_cell_coord(p.x + b, C.ox, C.inv_h, C.nx);
  const int y0 = cloud_cell_coord(p.y - b, C.oy, C.inv_h, C.ny), y1 = cloud_cell_coord(p.y + b, C.oy, C.inv_h, C.ny);
  const int z0 = cloud_cell_coord(p.z - b, C.oz, C.inv_h, C.nz), z1 = cloud_cell_coord(p.z + b, C.oz, C.inv_h, C.nz);
  float best = INFINITY;
  int best_i = -1;
  for (int z = z0; z <= z1; z++)
    for (int y = y0; y <= y1; y++) {
      // cells of one x-row are contiguous: one range per row
      const int64_t row = ((int64_t)z * C.ny + y) * C.nx;
      const uint32_t s = __ldg(C.cell_start + row + x0), t = __ldg(C.cell_start + row + x1 + 1);
      for (uint32_t k = s; k < t; k++) {
        const float4 q = __ldg(C.pts + k);
        const float dx = q.x - p.x, dy = q.y - p.y, dz = q.z - p.z;
        const float d2 = dx * dx + dy * dy + dz * dz;
        if (d2 < best) {
          best = d2;
          best_i = __float_as_int(q.w);
        }
      }
    }
  if (idx) *idx = best_i;
  return best;
}

// any point with |x - p|^2 <= bound2 ?  (early exit)
__device__ inline bool cloud_any_within(const CloudDev& C, const float3 p, float bound2) {
  const float b = sqrtf(bound2);
  const int x0 = cloud_cell_coord(p.x - b, C.ox, C.inv_h, C.nx), x1 = cloud_cell_coord(p.x + b, C.ox, C.inv_h, C.nx);
  const int y0 = cloud_cell_coord(p.y - b, C.oy, C.inv_h, C.ny), y1 = cloud_cell_coord(p.y + b, C.oy, C.inv_h, C.ny);
  const int z0 = cloud_cell_coord(p.z - b, C.oz, C.inv_h, C.nz), z1 = cloud_cell_coord(p.z + b, C.oz, C.inv_h, C.nz);
  for (int z = z0; z <= z1; z++)
    for (int y = y0; y <= y1; y++) {
      const int64_t row = ((int64_t)z * C.ny + y) * C.nx;
      const uint32_t s = __ldg(C.cell_start + row + x0), t = __ldg(C.cell_start + row + x1 + 1);
      for (uint32_t k = s; k < t; k++) {
        const float4 q = __ldg(C.pts + k);
        const float dx = q.x - p.x, dy = q.y - p.y, dz = q.z - p.z;
        if (dx * dx + dy * dy + dz * dz <= bound2) return true;
      }
    }
  return false;
}

// Lower / upper bounds of NN(p) from the distance field.  Outside the grid the lower bound is the
// distance to the grid box plus nothing (points all lie inside it) and the upper bound is +inf.
__device__ __forceinline__ void cloud_bounds(const CloudDev& C, const float3 p, float& lo, float& hi) {
  const float fx = (p.x - C.ox) * C.inv_h, fy = (p.y - C.oy) * C.inv_h, fz = (p.z - C.oz) * C.inv_h;
  if (fx < 0.f || fy < 0.f || fz < 0.f || fx >= (float)C.nx || fy >= (float)C.ny || fz >= (float)C.nz) {
    // distance from p to the point bounding box (AABB minus rho) is a valid lower bound
    const float bx = fmaxf(fmaxf(C.lbx + C.rho - p.x, p.x - (C.ubx - C.rho)), 0.f);
    const float by = fmaxf(fmaxf(C.lby + C.rho - p.y, p.y - (C.uby - C.rho)), 0.f);
    const float bz = fmaxf(fmaxf(C.lbz + C.rho - p.z, p.z - (C.ubz - C.rho)), 0.f);
    lo = sqrtf(bx * bx + by * by + bz * bz);
    hi = INFINITY;
    return;
  }
  const int64_t cell = ((int64_t)(int)fz * C.ny + (int)fy) * C.nx + (int)fx;
  const float d = __ldg(C.dist + cell);
  lo = d - C.e;
  hi = (d >= C.dist_cap) ? INFINITY : d + C.e;
}

// CloudSDF::is_outside(p, r) preceded by is_outside_aabb(p, r), exact:
// outside  <=>  aabb_out  ||  sqrt(NN^2) - rho > r.
__device__ inline bool cloud_is_outside(const CloudDev& C, const float3 p, const float r) {
  if (p.x < C.lbx - r || p.x > C.ubx + r || p.y < C.lby - r || p.y > C.uby + r || p.z < C.lbz - r ||
      p.z > C.ubz + r)
    return true;
  const float d = r + C.rho;
  float lo, hi;
  cloud_bounds(C, p, lo, hi);
  // keep a small guard so that FP32 rounding of the field never decides a borderline query
  if (lo > d + 1e-5f) return true;
  if (hi < d - 1e-5f) return false;
  return !cloud_any_within(C, p, d * d);
}

// Conservative group-level cull: true only if certainly no point within r + rho.
__device__ __forceinline__ bool cloud_surely_outside(const CloudDev& C, const float3 p, const float r) {
  if (p.x < C.lbx - r || p.x > C.ubx + r || p.y < C.lby - r || p.y > C.uby + r || p.z < C.lbz - r ||
      p.z > C.ubz + r)
    return true;
  float lo, hi;
  cloud_bounds(C, p, lo, hi);
  return lo > r + C.rho + 1e-5f;
}

// Exact NN squared distance, unbounded (KDTree::sqdist).  Expands the search radius from the
// field's upper bound; outside the grid it starts from the distance to the far corner.
__device__ inline float cloud_nn_exact(const CloudDev& C, const float3 p, int* idx) {
  float lo, hi;
  cloud_bounds(C, p, lo, hi);
  if (!(hi < INFINITY)) {
    // far from every point (or outside the grid): bound by the distance to the farthest corner
    const float bx = fmaxf(fabsf(p.x - (C.lbx + C.rho)), fabsf(p.x - (C.ubx - C.rho)));
    const float by = fmaxf(fabsf(p.y - (C.lby + C.rho)), fabsf(p.y - (C.uby - C.rho)));
    const float bz = fmaxf(fabsf(p.z - (C.lbz + C.rho)), fabsf(p.z - (C.ubz - C.rho)));
    hi = sqrtf(bx * bx + by * by + bz * bz);
  }
  const float b = hi * (1.f + 1e-6f) + 1e-6f;
  return cloud_nn_within(C, p, b * b, idx);
}

// ---- build kernels ------------------------------------------------------------------------------
__global__ void k_cloud_count(const float* __restrict__ pts3, int64_t n, CloudDev C, uint32_t* counts,
                              uint32_t* cell_of) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float x = pts3[3 * i], y = pts3[3 * i + 1], z = pts3[3 * i + 2];
    const int cx = cloud_cell_coord(x, C.ox, C.inv_h, C.nx), cy = cloud_cell_coord(y, C.oy, C.inv_h, C.ny),
              cz = cloud_cell_coord(z, C.oz, C.inv_h, C.nz);
    const uint32_t cell = (uint32_t)(((int64_t)cz * C.ny + cy) * C.nx + cx);
    cell_of[i] = cell;
    atomicAdd(counts + cell, 1u);
  }
}

// single-block exclusive scan over n_cells counts (setup-time only); out has n_cells + 1 entries
__global__ void k_cloud_scan(const uint32_t* __restrict__ counts, int64_t n_cells, uint32_t* out) {
  __shared__ uint32_t part[1024];
  __shared__ uint32_t carry;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (int64_t base = 0; base < n_cells; base += 1024) {
    const int64_t i = base + threadIdx.x;
    const uint32_t v = i < n_cells ? counts[i] : 0u;
    part[threadIdx.x] = v;
    __syncthreads();
    for (int off = 1; off < 1024; off <<= 1) {
      uint32_t t = threadIdx.x >= off ? part[threadIdx.x - off] : 0u;
      __syncthreads();
      part[threadIdx.x] += t;
      __syncthreads();
    }
    if (i < n_cells) out[i] = carry + part[threadIdx.x] - v;
    __syncthreads();
    if (threadIdx.x == 1023) carry += part[1023];
    __syncthreads();
  }
  if (threadIdx.x == 0) out[n_cells] = carry;
}

__global__ void k_cloud_scatter(const float* __restrict__ pts3, int64_t n, const uint32_t* __restrict__ cell_of,
                                const uint32_t* __restrict__ cell_start, uint32_t* cursor, float4* sorted) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const uint32_t cell = cell_of[i];
    const uint32_t k = cell_start[cell] + atomicAdd(cursor + cell, 1u);
    sorted[k] = make_float4(pts3[3 * i], pts3[3 * i + 1], pts3[3 * i + 2], __int_as_float((int)i));
  }
}

// dist[cell] = exact NN distance of the cell centre, by expanding cube rings of cells until the
// ring's inner distance exceeds the best found (capped at C.dist_cap).
__global__ void k_cloud_field(CloudDev C, float* dist) {
  const int64_t n_cells = (int64_t)C.nx * C.ny * C.nz;
  for (int64_t cell = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; cell < n_cells;
       cell += (int64_t)gridDim.x * blockDim.x) {
    const int cx = (int)(cell % C.nx), cy = (int)((cell / C.nx) % C.ny), cz = (int)(cell / ((int64_t)C.nx * C.ny));
    const float px = C.ox + (cx + 0.5f) * C.h, py = C.oy + (cy + 0.5f) * C.h, pz = C.oz + (cz + 0.5f) * C.h;
    float best = C.dist_cap * C.dist_cap;
    const int kmax = (int)ceilf(C.dist_cap * C.inv_h) + 1;
    for (int k = 0; k <= kmax; k++) {
      // every point in ring k is at least (k - 0.5) h away from the centre
      const float inner = (k - 0.5f) * C.h;
      if (k > 0 && inner * inner >= best) break;
      const int z0 = max(cz - k, 0), z1 = min(cz + k, C.nz - 1);
      const int y0 = max(cy - k, 0), y1 = min(cy + k, C.ny - 1);
      const int x0 = max(cx - k, 0), x1 = min(cx + k, C.nx - 1);
      for (int z = z0; z <= z1; z++)
        for (int y = y0; y <= y1; y++) {
          const bool shell_row = (abs(z - cz) == k) || (abs(y - cy) == k);
          const int64_t row = ((int64_t)z * C.ny + y) * C.nx;
          if (shell_row) {
            const uint32_t s = C.cell_start[row + x0], t = C.cell_start[row + x1 + 1];
            for (uint32_t i = s; i < t; i++) {
              const float4 q = C.pts[i];
              const float dx = q.x - px, dy = q.y - py, dz = q.z - pz;
              best = fminf(best, dx * dx + dy * dy + dz * dz);
            }
          } else {
            // only the two end cells of this row belong to the shell
            if (cx - k >= 0) {
              const uint32_t s = C.cell_start[row + cx - k], t = C.cell_start[row + cx - k + 1];
              for (uint32_t i = s; i < t; i++) {
                const float4 q = C.pts[i];
                const float dx = q.x - px, dy = q.y - py, dz = q.z - pz;
                best = fminf(best, dx * dx + dy * dy + dz * dz);
              }
            }
            if (cx + k < C.nx && k > 0) {
              const uint32_t s = C.cell_start[row + cx + k], t = C.cell_start[row + cx + k + 1];
              for (uint32_t i = s; i < t; i++) {
                const float4 q = C.pts[i];
                const float dx = q.x - px, dy = q.y - py, dz = q.z - pz;
                best = fminf(best, dx * dx + dy * dy + dz * dz);
              }
            }
          }
        }
    }
    dist[cell] = fminf(sqrtf(best), C.dist_cap);
  }
}

}  // namespace pmp
