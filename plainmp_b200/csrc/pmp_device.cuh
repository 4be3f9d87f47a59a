// pmp_device.cuh — device-side arithmetic of the plainmp hot path in FP32 (sm_100a).
//
// What each block restates (reference paths relative to cpp/plainmp/):
//   approx_sincos      kinematics/algorithm.cpp:49-86      (Taylor 9/8 polynomial, same powers)
//   joint_local / fk   kinematics/algorithm.cpp:109-158, :181-210, transform.hpp:64-74
//   rotmat             Eigen Quaternion::toRotationMatrix (used at
//                      constraints/primitive_sphere_collision.cpp:32)
//   prim_is_outside    collision/primitive_sdf.hpp:180-184 (AABB), :246-248 (Ground),
//                      :313-398 (Box), :464-498 (Cylinder), :546-552 (Sphere)
//   prim_evaluate      collision/primitive_sdf.hpp:245, :299-311, :444-462, :540-544
// Quaternions are never renormalised along the chain.
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/plainmp_b200.h"

namespace pmp {

struct Frame {
  float qx, qy, qz, qw;
  float tx, ty, tz;
};

struct Rot {
  float m00, m01, m02, m10, m11, m12, m20, m21, m22;
};

__device__ __forceinline__ float3 make3(float x, float y, float z) { return make_float3(x, y, z); }

// Eigen quaternion product a (x) b, (x,y,z,w) storage.
__device__ __forceinline__ void quat_mul(float ax, float ay, float az, float aw, float bx, float by,
                                         float bz, float bw, float& rx, float& ry, float& rz,
                                         float& rw) {
  rw = aw * bw - ax * bx - ay * by - az * bz;
  rx = aw * bx + ax * bw + ay * bz - az * by;
  ry = aw * by + ay * bw + az * bx - ax * bz;
  rz = aw * bz + az * bw + ax * by - ay * bx;
}

// Eigen q * v: uv = 2 (q.vec x v); v + w uv + q.vec x uv
__device__ __forceinline__ float3 quat_rot(float qx, float qy, float qz, float qw, float3 v) {
  float ux = qy * v.z - qz * v.y;
  float uy = qz * v.x - qx * v.z;
  float uz = qx * v.y - qy * v.x;
  ux += ux;
  uy += uy;
  uz += uz;
  return make3(v.x + qw * ux + (qy * uz - qz * uy), v.y + qw * uy + (qz * ux - qx * uz),
               v.z + qw * uz + (qx * uy - qy * ux));
}

__device__ __forceinline__ Rot rotmat(const Frame& f) {
  const float tx = 2.f * f.qx, ty = 2.f * f.qy, tz = 2.f * f.qz;
  const float twx = tx * f.qw, twy = ty * f.qw, twz = tz * f.qw;
  const float txx = tx * f.qx, txy = ty * f.qx, txz = tz * f.qx;
  const float tyy = ty * f.qy, tyz = tz * f.qy, tzz = tz * f.qz;
  Rot r;
  r.m00 = 1.f - (tyy + tzz);
  r.m01 = txy - twz;
  r.m02 = txz + twy;
  r.m10 = txy + twz;
  r.m11 = 1.f - (txx + tzz);
  r.m12 = tyz - twx;
  r.m20 = txz - twy;
  r.m21 = tyz + twx;
  r.m22 = 1.f - (txx + tyy);
  return r;
}

__device__ __forceinline__ float3 xform(const Rot& r, const Frame& f, float px, float py, float pz) {
  return make3(fmaf(r.m00, px, fmaf(r.m01, py, fmaf(r.m02, pz, f.tx))),
               fmaf(r.m10, px, fmaf(r.m11, py, fmaf(r.m12, pz, f.ty))),
               fmaf(r.m20, px, fmaf(r.m21, py, fmaf(r.m22, pz, f.tz))));
}

// kinematics/algorithm.cpp:49-86, FP32.  x is the HALF joint angle.
__device__ __forceinline__ void approx_sincos(float x, float& s, float& c) {
  constexpr float kPi = 3.14159265358979323846f;
  constexpr float kHalfPi = 1.57079632679489661923f;
  constexpr float kInv2Pi = 0.15915494309189533577f;
  float sign = 1.f;
  if (x > kHalfPi || x < -kHalfPi) {
    if (x > kPi || x < -kPi) x = x - 2.f * kPi * floorf(x * kInv2Pi + 0.5f);
    if (x < -kHalfPi) {
      x = -x - kPi;
      sign = -1.f;
    } else if (x > kHalfPi) {
      x = -x + kPi;
      sign = -1.f;
    }
  }
  const float xx = x * x;
  const float x4 = xx * xx;
  const float x6 = x4 * xx;
  const float x8 = x4 * x4;
  s = x * (1.f - xx * (1.f / 6.f) + x4 * (1.f / 120.f) - x6 * (1.f / 5040.f) + x8 * (1.f / 362880.f));
  c = sign * (1.f - xx * 0.5f + x4 * (1.f / 24.f) - x6 * (1.f / 720.f) + x8 * (1.f / 40320.f));
}

template <bool APPROX>
__device__ __forceinline__ void half_sincos(float theta, float& s, float& c) {
  const float x = theta * 0.5f;
  if (APPROX) {
    approx_sincos(x, s, c);
  } else {
    sincosf(x, &s, &c);
  }
}

// Device copy of one link op: 5 x 16 bytes.
//   w0 = origin quat (x,y,z,w); exactly (0,0,0,1) when flagged identity
//   w1 = origin xyz, kind
//   w2 = axis xyz, q_index
//   w3 = parent, flags, slot (stash slot of this op's frame or -1), parent_slot
//   w4 = revolute: origin (x) (axis, 0), the sine basis of the local quaternion (v3)
#define PMP_OPFLAG_ZERO_XYZ 256 /* device-internal: origin_xyz == 0, the translation term is skipped */
struct OpWords {
  float4 w0, w1, w2;
  int4 w3;
};

// frame = parent (x) local(q)
template <bool APPROX>
__device__ __forceinline__ Frame apply_op(const Frame& p, const float4 w0, const float4 w1,
                                          const float4 w2, const int kind, const int flags,
                                          const float qv) {
  Frame f;
  float ltx = w1.x, lty = w1.y, ltz = w1.z;
  if (kind == PMP_OP_REVOLUTE) {
    float s, c;
    half_sincos<APPROX>(qv, s, c);
    float lx = w2.x * s, ly = w2.y * s, lz = w2.z * s, lw = c;
    if (!(flags & PMP_OPFLAG_ORIGIN_IDENTITY)) {
      float ox, oy, oz, ow;
      quat_mul(w0.x, w0.y, w0.z, w0.w, lx, ly, lz, lw, ox, oy, oz, ow);
      lx = ox;
      ly = oy;
      lz = oz;
      lw = ow;
    }
    quat_mul(p.qx, p.qy, p.qz, p.qw, lx, ly, lz, lw, f.qx, f.qy, f.qz, f.qw);
  } else {
    // prismatic: translation origin + axis * q.  The joint itself carries no rotation
    // (kinematics/algorithm.cpp:149-155), but in the folded tables the constant rotation of the fixed
    // links above it lives in this op's origin quaternion, so it is applied like a fixed op's.
    if (kind == PMP_OP_PRISMATIC) {
      ltx = fmaf(w2.x, qv, ltx);
      lty = fmaf(w2.y, qv, lty);
      ltz = fmaf(w2.z, qv, ltz);
    }
    if (flags & PMP_OPFLAG_ORIGIN_IDENTITY) {
      f.qx = p.qx;
      f.qy = p.qy;
      f.qz = p.qz;
      f.qw = p.qw;
    } else {
      quat_mul(p.qx, p.qy, p.qz, p.qw, w0.x, w0.y, w0.z, w0.w, f.qx, f.qy, f.qz, f.qw);
    }
  }
  const float3 r = quat_rot(p.qx, p.qy, p.qz, p.qw, make3(ltx, lty, ltz));
  f.tx = p.tx + r.x;
  f.ty = p.ty + r.y;
  f.tz = p.tz + r.z;
  return f;
}

// QuatTrans::setQuaternionFromRPY (kinematics/transform.hpp:122-130), exact sin/cos.
__device__ __forceinline__ void quat_from_rpy(float roll, float pitch, float yaw, float& x, float& y,
                                              float& z, float& w) {
  float sp, cp, st, ct, ss, cs;
  sincosf(roll * 0.5f, &sp, &cp);
  sincosf(pitch * 0.5f, &st, &ct);
  sincosf(yaw * 0.5f, &ss, &cs);
  x = sp * ct * cs - cp * st * ss;
  y = cp * st * cs + sp * ct * ss;
  z = cp * ct * ss - sp * st * cs;
  w = cp * ct * cs + sp * st * ss;
}

// ---------------------------------------------------------------------------------------------
// primitives: kPrimWords x float4 per primitive, repacked from pmp_prim_t by pmp_scene_create
//   [0] position, type | flags << 8 | kind << 16   (kind: the test variant resolved at scene build)
//   [1] params: box half widths x,y,z | cylinder radius, half height | sphere radius | ground height
//   [2..4] rotation columns (read only when the pose is not axis- / z-aligned)
//   [5] AABB centre      [6] AABB half extent (+inf = unbounded)      [7] unused
// The AABB words are the reference's lb/ub (primitive_sdf.hpp:218-219; Ground pinned at z = 0,
// :250-255) in centre/half form, inflated by a few ulps so that the FP32 cull stays conservative.
// ---------------------------------------------------------------------------------------------
constexpr int kPrimWords = 8;
// test variants: the aligned ones need no rotation columns at all
enum { PMP_KIND_BOX = 1, PMP_KIND_CYL = 2, PMP_KIND_SPHERE = 3, PMP_KIND_GROUND = 4, PMP_KIND_BOX_AA = 5, PMP_KIND_CYL_AA = 6 };

struct Prim {
  const float4* w;  // the primitive's record (rotation columns are fetched on demand)
  float4 p0, pr;
  int type, flags, kind;
};

__device__ __forceinline__ Prim load_prim(const float4* __restrict__ prims, int i) {
  Prim P;
  P.w = prims + kPrimWords * i;
  P.p0 = P.w[0];
  P.pr = P.w[1];
  const int tf = __float_as_int(P.p0.w);
  P.type = tf & 0xff;
  P.flags = (tf >> 8) & 0xff;
  P.kind = tf >> 16;
  return P;
}

// |R^T d| per axis, following the reference's alignment shortcuts: axis-aligned poses skip the
// rotation (primitive_sdf.hpp:328-340), z-aligned ones use world z for the third coordinate
// (:341-345, :466-472).  `box` selects the Box rule (world x, y only when fully axis-aligned); for
// cylinders the host stores identity columns whenever the pose is z-aligned, so the same rule holds.
__device__ __forceinline__ void prim_local_abs(const Prim& P, const float dx, const float dy, const float dz,
                                               float& ax, float& ay, float& az) {
  if (P.flags & PMP_PRIMFLAG_AXIS_ALIGNED) {
    ax = fabsf(dx); ay = fabsf(dy); az = fabsf(dz);
    return;
  }
  const float4 c0 = P.w[2], c1 = P.w[3];
  ax = fabsf(dx * c0.x + dy * c0.y + dz * c0.z);
  ay = fabsf(dx * c1.x + dy * c1.y + dz * c1.z);
  if (P.flags & PMP_PRIMFLAG_Z_ALIGNED) {
    az = fabsf(dz);
  } else {
    const float4 c2 = P.w[4];
    az = fabsf(dx * c2.x + dy * c2.y + dz * c2.z);
  }
}

// !is_outside_aabb && !is_outside of the reference collapsed into one predicate: for Box /
// Cylinder / Sphere the AABB is conservative and therefore implied; for Ground it is
// semantically active (AABB pinned at z = 0).  Contact (sdf == r) is NOT
// outside (strict >).  One straight-line case per test variant, no early returns inside a case.
__device__ __forceinline__ bool box_outside(const float sx, const float sy, const float sz, const float r) {
  const float px = fmaxf(sx, 0.f), py = fmaxf(sy, 0.f), pz = fmaxf(sz, 0.f);
  // edge / corner case (primitive_sdf.hpp:374-397): with fewer than two positive axis distances, each
  // <= r, the sum below cannot exceed r^2, so the reference's out_count branches reduce to this;
  // r < 1e-6 skips it (:371-373).
  return (fmaxf(fmaxf(sx, sy), sz) > r) | ((r >= 1e-6f) & (px * px + py * py + pz * pz > r * r));
}
__device__ __forceinline__ bool cyl_outside(const float d2, const float sz, const float R, const float r) {
  const bool tiny = r < 1e-6f;  // :483-485
  const float rr = tiny ? R : R + r;
  bool out = (sz > r) | (d2 > rr * rr);
  if (!out & !tiny & (sz > 0.f) & (d2 > R * R)) {  // rim case (rare): the only place that needs a square root
    const float sr = sqrtf(d2) - R;
    out = sz * sz + sr * sr > r * r;
  }
  return out;
}
__device__ __forceinline__ bool prim_is_outside(const Prim& P, const float3 c, const float r) {
  if (P.kind == PMP_KIND_GROUND) return (fabsf(c.z) > r) | (c.z + P.pr.x > r);
  const float dx = c.x - P.p0.x, dy = c.y - P.p0.y, dz = c.z - P.p0.z;
  switch (P.kind) {
    case PMP_KIND_CYL_AA:
      return cyl_outside(dx * dx + dy * dy, fabsf(dz) - P.pr.y, P.pr.x, r);
    case PMP_KIND_BOX_AA:
      return box_outside(fabsf(dx) - P.pr.x, fabsf(dy) - P.pr.y, fabsf(dz) - P.pr.z, r);
    case PMP_KIND_SPHERE: {
      const float rr = (r < 1e-6f) ? P.pr.x : P.pr.x + r;  // :547-549
      return dx * dx + dy * dy + dz * dz > rr * rr;
    }
    default: {
      float ax, ay, az;
      prim_local_abs(P, dx, dy, dz, ax, ay, az);
      if (P.kind == PMP_KIND_BOX) return box_outside(ax - P.pr.x, ay - P.pr.y, az - P.pr.z, r);
      return cyl_outside(ax * ax + ay * ay, az - P.pr.y, P.pr.x, r);
    }
  }
}

// SDF value.  Ground has no AABB here; callers apply the reference's AABB pre-test where the
// reference does (evaluate_ext_collision, primitive_sphere_collision.cpp:272-283).
__device__ __forceinline__ float prim_evaluate(const Prim& P, const float3 c) {
  if (P.type == PMP_PRIM_GROUND) return c.z + P.pr.x;
  const float dx = c.x - P.p0.x, dy = c.y - P.p0.y, dz = c.z - P.p0.z;
  if (P.type == PMP_PRIM_SPHERE) return sqrtf(dx * dx + dy * dy + dz * dz) - P.pr.x;
  if (P.type == PMP_PRIM_BOX) {
    // BoxSDF::evaluate uses rot_inv for every non-axis-aligned pose (no z-aligned shortcut)
    float ax, ay, az;
    if (P.flags & PMP_PRIMFLAG_AXIS_ALIGNED) {
      ax = fabsf(dx); ay = fabsf(dy); az = fabsf(dz);
    } else {
      const float4 c0 = P.w[2], c1 = P.w[3], c2 = P.w[4];
      ax = fabsf(dx * c0.x + dy * c0.y + dz * c0.z);
      ay = fabsf(dx * c1.x + dy * c1.y + dz * c1.z);
      az = fabsf(dx * c2.x + dy * c2.y + dz * c2.z);
    }
    const float sx = ax - P.pr.x, sy = ay - P.pr.y, sz = az - P.pr.z;
    const float px = fmaxf(sx, 0.f), py = fmaxf(sy, 0.f), pz = fmaxf(sz, 0.f);
    const float outside = sqrtf(px * px + py * py + pz * pz);
    const float inside = fmaxf(fmaxf(fminf(sx, 0.f), fminf(sy, 0.f)), fminf(sz, 0.f));
    return outside + inside;
  }
  float ax, ay, az;
  prim_local_abs(P, dx, dy, dz, ax, ay, az);
  const float sz = az - P.pr.y;
  const float sr = sqrtf(ax * ax + ay * ay) - P.pr.x;
  const float pr = fmaxf(sr, 0.f), pz = fmaxf(sz, 0.f);
  return sqrtf(pr * pr + pz * pz) + fmaxf(fminf(sr, 0.f), fminf(sz, 0.f));
}

// Analytic gradient of prim_evaluate w.r.t. the query point (world frame, unit length away from the
// medial axis).  The reference differentiates evaluate() numerically (forward difference, step 1e-6,
// constraints/primitive_sphere_collision.cpp:307-314); the closed forms below are the limits of
// that quotient wherever the SDF is differentiable.
__device__ __forceinline__ float3 prim_gradient(const Prim& P, const float3 c) {
  if (P.type == PMP_PRIM_GROUND) return make3(0.f, 0.f, 1.f);
  const float dx = c.x - P.p0.x, dy = c.y - P.p0.y, dz = c.z - P.p0.z;
  if (P.type == PMP_PRIM_SPHERE) {
    const float inv = rsqrtf(fmaxf(dx * dx + dy * dy + dz * dz, 1e-30f));
    return make3(dx * inv, dy * inv, dz * inv);
  }
  const bool aligned = P.flags & PMP_PRIMFLAG_AXIS_ALIGNED;
  float4 c0 = make_float4(1.f, 0.f, 0.f, 0.f), c1 = make_float4(0.f, 1.f, 0.f, 0.f), c2 = make_float4(0.f, 0.f, 1.f, 0.f);
  if (!aligned) {
    c0 = P.w[2]; c1 = P.w[3]; c2 = P.w[4];
  }
  const float lx = dx * c0.x + dy * c0.y + dz * c0.z, ly = dx * c1.x + dy * c1.y + dz * c1.z,
              lz = dx * c2.x + dy * c2.y + dz * c2.z;
  float gx, gy, gz;  // gradient in the primitive's frame
  if (P.type == PMP_PRIM_BOX) {
    const float sx = fabsf(lx) - P.pr.x, sy = fabsf(ly) - P.pr.y, sz = fabsf(lz) - P.pr.z;
    if (fmaxf(fmaxf(sx, sy), sz) > 0.f) {
      const float px = fmaxf(sx, 0.f), py = fmaxf(sy, 0.f), pz = fmaxf(sz, 0.f);
      const float inv = rsqrtf(px * px + py * py + pz * pz);
      gx = copysignf(px * inv, lx); gy = copysignf(py * inv, ly); gz = copysignf(pz * inv, lz);
    } else {  // inside: the face with the largest (least negative) distance
      gx = (sx >= sy && sx >= sz) ? copysignf(1.f, lx) : 0.f;
      gy = (gx == 0.f && sy >= sz) ? copysignf(1.f, ly) : 0.f;
      gz = (gx == 0.f && gy == 0.f) ? copysignf(1.f, lz) : 0.f;
    }
  } else {  // cylinder: axis = local z
    const float rho = sqrtf(lx * lx + ly * ly);
    const float irho = rho > 0.f ? 1.f / rho : 0.f;
    const float sr = rho - P.pr.x, sz = fabsf(lz) - P.pr.y;
    float wr, wz;  // weights of the radial and axial directions
    if (fmaxf(sr, sz) > 0.f) {
      const float pr = fmaxf(sr, 0.f), pz = fmaxf(sz, 0.f);
      const float inv = rsqrtf(pr * pr + pz * pz);
      wr = pr * inv; wz = pz * inv;
    } else {
      wr = (sr > sz) ? 1.f : 0.f;
      wz = 1.f - wr;
    }
    gx = wr * lx * irho; gy = wr * ly * irho; gz = copysignf(wz, lz);
  }
  return make3(c0.x * gx + c1.x * gy + c2.x * gz, c0.y * gx + c1.y * gy + c2.y * gz, c0.z * gx + c1.z * gy + c2.z * gz);
}

// is_outside_aabb(c, r) (primitive_sdf.hpp:180-184) against the AABB in centre / half-extent form,
// branch-free: outside <=> max_k(|c_k - ctr_k| - half_k) > r.
__device__ __forceinline__ bool prim_aabb_outside(const float4 ctr, const float4 half, const float3 c,
                                                  const float r) {
  const float ex = fabsf(c.x - ctr.x) - half.x, ey = fabsf(c.y - ctr.y) - half.y, ez = fabsf(c.z - ctr.z) - half.z;
  return fmaxf(fmaxf(ex, ey), ez) > r;
}

}  // namespace pmp
