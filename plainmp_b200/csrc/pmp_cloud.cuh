// pmp_cloud.cuh — CloudSDF on the GPU (reference: collision/primitive_sdf.hpp:568-603 and the
// kd-tree it wraps, collision/kdtree.cpp:20-96).
//
// The reference answers sdf(p) = sqrt(NN^2(p)) - rho with a recursive kd-tree descent (pointer
// chasing, one query at a time, node order fixed by nth_element).  Any exact nearest-neighbour
// structure gives identical values, so the device structure is chosen for the GPU:
//
//   * the points are sorted by 30-bit Morton code on the device (pmp_cloud_build.cu) and covered
//     by an IMPLICIT complete 8-ary tree of tight bounding boxes: level 0 = boxes of 8 consecutive
//     points, level k = boxes of 8 consecutive level k-1 nodes.  No topology is stored; children
//     of node n are nodes 8n .. 8n+7 of the level below, so a per-lane depth-first traversal needs
//     only (level, node, one 8-bit pending mask per level) — all of it lives in registers (the
//     masks are packed into one 64-bit word), no stack in local memory.  Tight boxes (rather than a
//     space partition) prune a query that does not touch the surface within 1-3 node visits
//     whatever the point density;
//   * a distance-bound field dist[cell] = exact NN distance of the cell CENTRE of a uniform grid,
//     computed with the same tree at build time.  For a query p in that cell
//         dist - e <= NN(p) <= dist + e,   e = h * sqrt(3) / 2,
//     which answers "any point within d of p" in O(1) for every query outside the +-e band; only
//     the band goes to the tree.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace pmp {

constexpr int kCloudMaxLevels = 8;  // 8^8 * 8 points
// Several CloudSDF leaves of one scene share ONE index.  Leaves of equal radius are simply one point set;
// leaves of different radii become radius CLASSES: the class of a point rides in bits 28..30 of its w
// (the original index keeps bits 0..27), the pruning tests use the largest radius (conservative for every
// class), "certainly hit" decisions the smallest, and the exact per-point tests the point's own.
constexpr int kCloudMaxClasses = 4;
constexpr int kCloudClassShift = 28;

struct CloudDev {
  const float4* pts;    // Morton-sorted, padded to a multiple of 8 with +inf; w = original index (int bits)
  const float4* nodes;  // per node: lo, hi; levels stored one after another, each padded to a multiple of 8 nodes
  const float* dist;    // n_cells: NN distance of the cell centre (capped at dist_cap)
  int64_t n_pts;
  int n_levels;                    // level n_levels-1 is the root (one node)
  int level_off[kCloudMaxLevels];  // first node of each level
  int nx, ny, nz;
  float ox, oy, oz;  // grid origin (bbox low corner - margin)
  float h, inv_h;
  float e;           // half cell diagonal
  float dist_cap;
  float rho;         // CloudSDF::radius_ (the LARGEST one when the scene's clouds have different radii)
  float lbx, lby, lbz, ubx, uby, ubz;  // CloudSDF AABB = point bbox +- rho (primitive_sdf.hpp:573-582)
  int enabled;
  int n_classes;     // 1 unless CloudSDFs of different radii were merged
  float rho_min;     // smallest class radius (= rho for one class)
  float rho_cls[kCloudMaxClasses];
};

__device__ __forceinline__ int cloud_point_class(const float w) { return (__float_as_int(w) >> kCloudClassShift) & 7; }
// (r + rho of the point's class)^2: the exact per-point bound of an "any point within" query of radius r
__device__ __forceinline__ float cloud_point_bound2(const CloudDev& C, const float w, const float r, const float bound2) {
  if (C.n_classes == 1) return bound2;
  const float d = r + C.rho_cls[cloud_point_class(w) & (kCloudMaxClasses - 1)];
  return d * d;
}

__device__ __forceinline__ float box_dist2(const float4 lo, const float4 hi, const float3 p) {
  const float dx = fmaxf(fmaxf(lo.x - p.x, p.x - hi.x), 0.f);
  const float dy = fmaxf(fmaxf(lo.y - p.y, p.y - hi.y), 0.f);
  const float dz = fmaxf(fmaxf(lo.z - p.z, p.z - hi.z), 0.f);
  return dx * dx + dy * dy + dz * dz;
}

// 8-bit mask of the nodes first .. first+7 of `level` whose box is within sqrt(bound2) of p.
// Padding nodes are (+inf, -inf) boxes: their distance is +inf and never passes.
template <bool STRICT>
__device__ __forceinline__ unsigned cloud_children_mask(const CloudDev& C, const int level, const int first, const float3 p,
                                                        const float bound2) {
  const float4* b = C.nodes + 2 * (size_t)(C.level_off[level] + first);
  unsigned m = 0;
#pragma unroll
  for (int j = 0; j < 8; j++) {
    const float d2 = box_dist2(__ldg(b + 2 * j), __ldg(b + 2 * j + 1), p);
    if (STRICT ? (d2 < bound2) : (d2 <= bound2)) m |= 1u << j;
  }
  return m;
}

// any point x with |x - p| <= r + rho(x) ?  (depth-first, early exit; boxes are pruned with the largest rho)
__device__ inline bool cloud_any_within(const CloudDev& C, const float3 p, const float r) {
  const float bound2 = (r + C.rho) * (r + C.rho);
  const int top = C.n_levels - 1;
  int l = top, n = 0;
  unsigned long long masks = 0;
  if (top == 0) {
    masks = 1ull << 8;  // the root is a leaf: treat it as the only child of a virtual level-1 node
    l = 1;
  } else {
    masks = (unsigned long long)cloud_children_mask<false>(C, l - 1, 0, p, bound2) << (8 * l);
  }
  const int vtop = l;
  for (;;) {
    const unsigned m = (unsigned)(masks >> (8 * l)) & 0xffu;
    if (!m) {
      if (l == vtop) return false;
      n >>= 3;
      l++;
      continue;
    }
    const int j = __ffs(m) - 1;
    masks &= ~(1ull << (8 * l + j));
    const int child = n * 8 + j;
    if (l == 1) {  // child is a leaf box: its 8 points
      const float4* q = C.pts + 8 * (size_t)child;
#pragma unroll
      for (int k = 0; k < 8; k++) {
        const float4 x = __ldg(q + k);
        const float dx = x.x - p.x, dy = x.y - p.y, dz = x.z - p.z;
        if (dx * dx + dy * dy + dz * dz <= cloud_point_bound2(C, x.w, r, bound2)) return true;
      }
    } else {
      l--;
      n = child;
      masks = (masks & ~(0xffull << (8 * l))) |
              ((unsigned long long)cloud_children_mask<false>(C, l - 1, n * 8, p, bound2) << (8 * l));
    }
  }
}

// exact: squared distance to the nearest cloud point if it is <= bound2, else a value > bound2
// (+inf).  `idx` receives the original index of the best point (or -1).
// `cls` >= 0 restricts the search to the points of one radius class (the boxes still cover all classes).
__device__ inline float cloud_nn_within(const CloudDev& C, const float3 p, const float bound2, int* idx,
                                        float3* nearest = nullptr, const int cls = -1) {
  const int top = C.n_levels - 1;
  // accept d2 <= bound2: search with a strict test against a bound nudged up by one ulp
  float best = __uint_as_float(__float_as_uint(fminf(fmaxf(bound2, 0.f), 3.0e38f)) + 1u);
  const float limit = best;
  int best_i = -1;
  float3 best_p = make_float3(0.f, 0.f, 0.f);
  int l = top, n = 0;
  unsigned long long masks = 0;
  if (top == 0) {
    masks = 1ull << 8;
    l = 1;
  } else {
    masks = (unsigned long long)cloud_children_mask<true>(C, l - 1, 0, p, best) << (8 * l);
  }
  const int vtop = l;
  for (;;) {
    const unsigned m = (unsigned)(masks >> (8 * l)) & 0xffu;
    if (!m) {
      if (l == vtop) break;
      n >>= 3;
      l++;
      continue;
    }
    const int j = __ffs(m) - 1;
    masks &= ~(1ull << (8 * l + j));
    const int child = n * 8 + j;
    if (l == 1) {
      const float4* q = C.pts + 8 * (size_t)child;
#pragma unroll
      for (int k = 0; k < 8; k++) {
        const float4 x = __ldg(q + k);
        const float dx = x.x - p.x, dy = x.y - p.y, dz = x.z - p.z;
        const float d2 = dx * dx + dy * dy + dz * dz;
        if (d2 < best && (cls < 0 || cloud_point_class(x.w) == cls)) {
          best = d2;
          best_i = __float_as_int(x.w);
          best_p = make_float3(x.x, x.y, x.z);
        }
      }
    } else {
      l--;
      n = child;
      masks = (masks & ~(0xffull << (8 * l))) |
              ((unsigned long long)cloud_children_mask<true>(C, l - 1, n * 8, p, best) << (8 * l));
    }
  }
  if (idx) *idx = best_i;
  if (nearest) *nearest = best_p;
  return (best < limit) ? best : INFINITY;
}

// Lower / upper bounds of NN(p) from the distance field.  Outside the grid the lower bound is the
// distance to the point bounding box and the upper bound is +inf.
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
  const int ix = (int)fx, iy = (int)fy, iz = (int)fz;
  const int cell = (iz * C.ny + iy) * C.nx + ix;  // (the build keeps the field below 64 M cells)
  const float d = __ldg(C.dist + cell);
  // NN is 1-Lipschitz: |NN(p) - NN(centre)| <= |p - centre| (<= e, the half diagonal); the guard
  // covers the FP32 rounding of the offset
  const float ux = (fx - (float)ix - 0.5f), uy = (fy - (float)iy - 0.5f), uz = (fz - (float)iz - 0.5f);
  const float off = fminf(C.h * sqrtf(ux * ux + uy * uy + uz * uz) + 2e-6f, C.e);
  lo = d - off;
  hi = (d >= C.dist_cap) ? INFINITY : d + off;
}

// CloudSDF::is_outside(p, r) preceded by is_outside_aabb(p, r), exact:
// outside  <=>  aabb_out  ||  sqrt(NN^2) - rho > r.
__device__ inline bool cloud_is_outside(const CloudDev& C, const float3 p, const float r) {
  if (p.x < C.lbx - r || p.x > C.ubx + r || p.y < C.lby - r || p.y > C.uby + r || p.z < C.lbz - r ||
      p.z > C.ubz + r)
    return true;
  float lo, hi;
  cloud_bounds(C, p, lo, hi);
  // keep a small guard so that FP32 rounding of the field never decides a borderline query
  if (lo > r + C.rho + 1e-5f) return true;       // farther than the largest radius allows
  if (hi < r + C.rho_min - 1e-5f) return false;  // some point closer than even the smallest radius needs
  return !cloud_any_within(C, p, r);
}

// ---- 8-lane cooperative traversal ------------------------------------------------------------------
// The tree is 8-ary and a query is worked by 8 lanes, so one sub-warp visits one node per step: lane j tests
// child j (two LDG.128 + ~12 instructions per lane instead of 8 x that in one lane), a ballot gives the child
// mask, and a leaf is one point per lane.  The 8 lanes of a sub-warp (lanes 8k .. 8k+7) share p and the bound;
// the traversal state (level, node, pending masks) is sub-warp uniform, so there is no divergence inside a
// sub-warp.
//
// "Any point within r + rho(x) of p", four queries at once, one per 8-lane sub-warp, in lockstep: every
// collective is a full-warp ballot of which each sub-warp reads its own byte, and every step of the loop is one
// node (or leaf) visit of every sub-warp that is still walking.  All 32 lanes must call; `valid` is sub-warp
// uniform (false = this sub-warp has no query); the result is sub-warp uniform.  (Letting each sub-warp walk on
// its own with collectives on its 8-lane mask measured the same in the validity kernels, 8.0e8 vs 8.0e8 PR2
// configurations/s; for the exact nearest-point search below, where walks are long and of very different
// lengths, lockstep measured 8-9 % SLOWER, so cloud_nn8 keeps independent sub-warps.)
__device__ inline bool cloud_any_within8x4(const CloudDev& C, const float3 p, const float r, const bool valid, const int sub,
                                           const int kshift) {
  constexpr unsigned full = 0xffffffffu;
  const float bound2 = (r + C.rho) * (r + C.rho);
  const int top = C.n_levels - 1;
  int l = top, n = 0;
  unsigned long long masks = 0;
  bool done = !valid, hit = false;
  if (top == 0) {
    masks = 1ull << 8;
    l = 1;
  } else {
    const float4* b = C.nodes + 2 * (size_t)(C.level_off[top - 1] + sub);
    const float d2 = done ? INFINITY : box_dist2(__ldg(b), __ldg(b + 1), p);
    masks = (unsigned long long)((__ballot_sync(full, d2 <= bound2) >> kshift) & 0xffu) << (8 * l);
  }
  const int vtop = l;
  for (;;) {
    unsigned m = 0;
    if (!done) {
      m = (unsigned)(masks >> (8 * l)) & 0xffu;
      while (!m && l != vtop) {  // this level is exhausted: back to the parent's pending children
        n >>= 3;
        l++;
        m = (unsigned)(masks >> (8 * l)) & 0xffu;
      }
      done = !m;
    }
    if (__all_sync(full, done)) break;
    float d2 = INFINITY, thr = bound2;
    int child = 0;
    const bool leaf = l == 1;
    if (!done) {
      const int j = __ffs(m) - 1;
      masks &= ~(1ull << (8 * l + j));
      child = n * 8 + j;
      if (leaf) {  // leaf box: one point per lane (padding points are +inf)
        const float4 x = __ldg(C.pts + 8 * (size_t)child + sub);
        const float dx = x.x - p.x, dy = x.y - p.y, dz = x.z - p.z;
        d2 = dx * dx + dy * dy + dz * dz;
        thr = cloud_point_bound2(C, x.w, r, bound2);
      } else {  // inner node: one child box per lane (padding boxes are at +inf)
        const float4* b = C.nodes + 2 * (size_t)(C.level_off[l - 2] + 8 * child + sub);
        d2 = box_dist2(__ldg(b), __ldg(b + 1), p);
      }
    }
    const unsigned within = (__ballot_sync(full, !done && d2 <= thr) >> kshift) & 0xffu;
    if (!done) {
      if (leaf) {
        if (within) hit = done = true;
      } else {
        l--;
        n = child;
        masks = (masks & ~(0xffull << (8 * l))) | ((unsigned long long)within << (8 * l));
      }
    }
  }
  return hit;
}

// 8-lane cooperative exact nearest neighbour: the sub-warp shares one query p; lane j tests child box j
// of the current node (strict test against the running best), a leaf is one point per lane followed by
// a sub-warp arg-min.  `best` enters as an exclusive upper bound on the squared distance and is sub-warp
// uniform throughout.  Returns the squared NN distance (or the bound if nothing beat it) and the
// original index of the nearest point through *idx (-1 if none).
__device__ inline float cloud_nn8(const CloudDev& C, const float3 p, float best, const unsigned smask, const int sub,
                                  const int kshift, int* idx, int* pos = nullptr, const int cls = -1) {
  const int top = C.n_levels - 1;
  int best_i = -1, best_pos = -1;
  int l = top, n = 0;
  unsigned long long masks = 0;
  auto child_mask = [&](const int level, const int first) -> unsigned {
    const float4* b = C.nodes + 2 * (size_t)(C.level_off[level] + first + sub);
    const float d2 = box_dist2(__ldg(b), __ldg(b + 1), p);
    return (__ballot_sync(smask, d2 < best) >> kshift) & 0xffu;
  };
  if (top == 0) {
    masks = 1ull << 8;
    l = 1;
  } else {
    masks = (unsigned long long)child_mask(l - 1, 0) << (8 * l);
  }
  const int vtop = l;
  for (;;) {
    const unsigned m = (unsigned)(masks >> (8 * l)) & 0xffu;
    if (!m) {
      if (l == vtop) break;
      n >>= 3;
      l++;
      continue;
    }
    const int j = __ffs(m) - 1;
    masks &= ~(1ull << (8 * l + j));
    const int child = n * 8 + j;
    if (l == 1) {
      const float4 x = __ldg(C.pts + 8 * (size_t)child + sub);  // padding points are +inf
      const float dx = x.x - p.x, dy = x.y - p.y, dz = x.z - p.z;
      float d2 = dx * dx + dy * dy + dz * dz;
      if (cls >= 0 && cloud_point_class(x.w) != cls) d2 = INFINITY;  // another radius class
      int di = __float_as_int(x.w);
      int dp = 8 * child + sub;  // position in the Morton-sorted array
#pragma unroll
      for (int off = 4; off; off >>= 1) {
        const float od = __shfl_xor_sync(smask, d2, off);
        const int oi = __shfl_xor_sync(smask, di, off);
        const int op = __shfl_xor_sync(smask, dp, off);
        if (od < d2 || (od == d2 && oi < di)) {  // ties -> smaller original index, the same on every lane
          d2 = od;
          di = oi;
          dp = op;
        }
      }
      if (d2 < best) {
        best = d2;
        best_i = di;
        best_pos = dp;
      }
    } else {
      l--;
      n = child;
      masks = (masks & ~(0xffull << (8 * l))) | ((unsigned long long)child_mask(l - 1, n * 8) << (8 * l));
    }
  }
  *idx = best_i;
  if (pos) *pos = best_pos;
  return best;
}

// cooperative form of cloud_greedy_bound2: lane j looks at child j, the sub-warp steps into the nearest
__device__ inline float cloud_greedy_bound2_8(const CloudDev& C, const float3 p, const unsigned smask, const int sub) {
  int n = 0;
  for (int l = C.n_levels - 1; l > 0; l--) {
    const float4* b = C.nodes + 2 * (size_t)(C.level_off[l - 1] + 8 * n + sub);
    float d2 = box_dist2(__ldg(b), __ldg(b + 1), p);
    int dj = sub;
#pragma unroll
    for (int off = 4; off; off >>= 1) {
      const float od = __shfl_xor_sync(smask, d2, off);
      const int oj = __shfl_xor_sync(smask, dj, off);
      if (od < d2 || (od == d2 && oj < dj)) {
        d2 = od;
        dj = oj;
      }
    }
    n = 8 * n + dj;
  }
  const float4 x = __ldg(C.pts + 8 * (size_t)n + sub);
  const float dx = x.x - p.x, dy = x.y - p.y, dz = x.z - p.z;
  float d2 = dx * dx + dy * dy + dz * dz;
#pragma unroll
  for (int off = 4; off; off >>= 1) d2 = fminf(d2, __shfl_xor_sync(smask, d2, off));
  return d2;
}

// The O(1) part of cloud_is_outside: 0 = certainly outside, 1 = certainly not outside, 2 = the
// distance field cannot decide (p lies within +-e of the r + rho iso-surface): ask the tree with
// bound (r + rho)^2.
__device__ __forceinline__ int cloud_classify(const CloudDev& C, const float3 p, const float r) {
  if (p.x < C.lbx - r || p.x > C.ubx + r || p.y < C.lby - r || p.y > C.uby + r || p.z < C.lbz - r ||
      p.z > C.ubz + r)
    return 0;
  float lo, hi;
  cloud_bounds(C, p, lo, hi);
  if (lo > r + C.rho + 1e-5f) return 0;
  if (hi < r + C.rho_min - 1e-5f) return 1;
  return 2;
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

// Exact NN squared distance, unbounded (KDTree::sqdist).  The field's upper bound (or, far from
// every point, the distance to the far corner of the bounding box) starts the search.
// Upper bound of NN(p) from one greedy descent: at every level step into the child box nearest to p
// and, at the leaf, take the nearest of its 8 points.  n_levels x 8 box tests; for a query far from
// the cloud (outside the distance field) this is what makes the bounded search below prune.
__device__ inline float cloud_greedy_bound2(const CloudDev& C, const float3 p) {
  int n = 0;
  for (int l = C.n_levels - 1; l > 0; l--) {
    const float4* b = C.nodes + 2 * (size_t)(C.level_off[l - 1] + 8 * n);
    float bd = INFINITY;
    int bj = 0;
#pragma unroll
    for (int j = 0; j < 8; j++) {
      const float d2 = box_dist2(__ldg(b + 2 * j), __ldg(b + 2 * j + 1), p);  // padding boxes give +inf
      if (d2 < bd) {
        bd = d2;
        bj = j;
      }
    }
    n = 8 * n + bj;
  }
  const float4* q = C.pts + 8 * (size_t)n;
  float best = INFINITY;
#pragma unroll
  for (int k = 0; k < 8; k++) {
    const float4 x = __ldg(q + k);  // padding points are +inf
    const float dx = x.x - p.x, dy = x.y - p.y, dz = x.z - p.z;
    best = fminf(best, dx * dx + dy * dy + dz * dz);
  }
  return best;
}

__device__ inline float cloud_nn_exact(const CloudDev& C, const float3 p, int* idx) {
  float lo, hi;
  cloud_bounds(C, p, lo, hi);
  {
    // one greedy descent: outside the field it is the only bound, inside it usually beats the field's
    // +-cell-diagonal bound by orders of magnitude and turns the search into a verification
    const float g2 = cloud_greedy_bound2(C, p);
    if (g2 < INFINITY) {
      hi = fminf(hi, sqrtf(g2));
    } else if (!(hi < INFINITY)) {
      const float bx = fmaxf(fabsf(p.x - (C.lbx + C.rho)), fabsf(p.x - (C.ubx - C.rho)));
      const float by = fmaxf(fabsf(p.y - (C.lby + C.rho)), fabsf(p.y - (C.uby - C.rho)));
      const float bz = fmaxf(fabsf(p.z - (C.lbz + C.rho)), fabsf(p.z - (C.ubz - C.rho)));
      hi = sqrtf(bx * bx + by * by + bz * bz);
    }
  }
  const float b = hi * (1.f + 1e-6f) + 1e-6f;
  return cloud_nn_within(C, p, b * b, idx);
}

}  // namespace pmp
