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

// Device copy of one link op: 4 x 16 bytes.
//   w0 = origin quat (x,y,z,w)
//   w1 = origin xyz, kind
//   w2 = axis xyz, q_index
//   w3 = parent, flags, slot (stash slot of this op's frame or -1), parent_slot
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
  } else if (kind == PMP_OP_PRISMATIC) {
    ltx = fmaf(w2.x, qv, ltx);
    lty = fmaf(w2.y, qv, lty);
    ltz = fmaf(w2.z, qv, ltz);
    f.qx = p.qx;
    f.qy = p.qy;
    f.qz = p.qz;
    f.qw = p.qw;
  } else {  // fixed
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
// primitives: 6 x float4 per primitive (same layout as pmp_prim_t)
//   p0 = position, type      p1 = col0, param0     p2 = col1, param1
//   p3 = col2, param2        p4 = lb, flags        p5 = ub, -
// ---------------------------------------------------------------------------------------------
struct Prim {
  float4 p0, p1, p2, p3;
  int type, flags;
};

__device__ __forceinline__ Prim load_prim(const float4* __restrict__ prims, int i) {
  Prim P;
  P.p0 = prims[6 * i + 0];
  P.p1 = prims[6 * i + 1];
  P.p2 = prims[6 * i + 2];
  P.p3 = prims[6 * i + 3];
  P.type = __float_as_int(P.p0.w);
  P.flags = __float_as_int(prims[6 * i + 4].w);
  return P;
}

// !is_outside_aabb && !is_outside of the reference collapsed into one predicate: for Box /
// Cylinder / Sphere the AABB is conservative and therefore implied; for Ground it is
// semantically active (AABB pinned at z = 0).  Contact (sdf == r) is NOT
// outside (strict >).
__device__ __forceinline__ bool prim_is_outside(const Prim& P, const float3 c, const float r) {
  if (P.type == PMP_PRIM_GROUND) {
    const bool aabb_out = (c.z < -r) || (c.z > r);
    return aabb_out || (c.z + P.p1.w > r);
  }
  const float dx = c.x - P.p0.x, dy = c.y - P.p0.y, dz = c.z - P.p0.z;
  if (P.type == PMP_PRIM_SPHERE) {
    const float d2 = dx * dx + dy * dy + dz * dz;
    const float R = P.p1.w;
    if (r < 1e-6f) return d2 > R * R;
    return d2 > (R + r) * (R + r);
  }
  const float ax = fabsf(dx * P.p1.x + dy * P.p1.y + dz * P.p1.z);
  const float ay = fabsf(dx * P.p2.x + dy * P.p2.y + dz * P.p2.z);
  // z test in world axes when the pose is z-aligned (primitive_sdf.hpp:341-345, :466-472)
  const float az = (P.flags & PMP_PRIMFLAG_Z_ALIGNED) ? fabsf(dz)
                                                      : fabsf(dx * P.p3.x + dy * P.p3.y + dz * P.p3.z);
  if (P.type == PMP_PRIM_BOX) {
    const float sx = ax - P.p1.w, sy = ay - P.p2.w, sz = az - P.p3.w;
    if (sx > r || sy > r || sz > r) return true;
    if (r < 1e-6f) return false;
    // edge / corner case: with fewer than two positive axis distances the sum below cannot
    // exceed r^2 (each is <= r), so the reference's out_count branches reduce to this.
    const float px = fmaxf(sx, 0.f), py = fmaxf(sy, 0.f), pz = fmaxf(sz, 0.f);
    const int cnt = (sx > 0.f) + (sy > 0.f) + (sz > 0.f);
    if (cnt < 2) return false;
    return px * px + py * py + pz * pz > r * r;
  }
  // cylinder: p1.w = radius, p2.w = half height
  const float sz = az - P.p2.w;
  if (sz > r) return true;
  const float R = P.p1.w;
  const float d2 = ax * ax + ay * ay;
  if (r < 1e-6f) return d2 > R * R;
  if (d2 > (R + r) * (R + r)) return true;
  if (sz > 0.f && d2 > R * R) {
    const float sr = sqrtf(d2) - R;
    return sz * sz + sr * sr > r * r;
  }
  return false;
}

// SDF value.  Ground has no AABB here; callers apply the reference's AABB pre-test where the
// reference does (evaluate_ext_collision, primitive_sphere_collision.cpp:272-283).
__device__ __forceinline__ float prim_evaluate(const Prim& P, const float3 c) {
  if (P.type == PMP_PRIM_GROUND) return c.z + P.p1.w;
  const float dx = c.x - P.p0.x, dy = c.y - P.p0.y, dz = c.z - P.p0.z;
  if (P.type == PMP_PRIM_SPHERE) return sqrtf(dx * dx + dy * dy + dz * dz) - P.p1.w;
  const float ax = fabsf(dx * P.p1.x + dy * P.p1.y + dz * P.p1.z);
  const float ay = fabsf(dx * P.p2.x + dy * P.p2.y + dz * P.p2.z);
  if (P.type == PMP_PRIM_BOX) {
    // BoxSDF::evaluate uses rot_inv for every non-axis-aligned pose (no z-aligned shortcut)
    const float az = fabsf(dx * P.p3.x + dy * P.p3.y + dz * P.p3.z);
    const float sx = ax - P.p1.w, sy = ay - P.p2.w, sz = az - P.p3.w;
    const float px = fmaxf(sx, 0.f), py = fmaxf(sy, 0.f), pz = fmaxf(sz, 0.f);
    const float outside = sqrtf(px * px + py * py + pz * pz);
    const float inside = fmaxf(fmaxf(fminf(sx, 0.f), fminf(sy, 0.f)), fminf(sz, 0.f));
    return outside + inside;
  }
  const float az = (P.flags & PMP_PRIMFLAG_Z_ALIGNED) ? fabsf(dz)
                                                      : fabsf(dx * P.p3.x + dy * P.p3.y + dz * P.p3.z);
  const float sz = az - P.p2.w;
  const float sr = sqrtf(ax * ax + ay * ay) - P.p1.w;
  const float pr = fmaxf(sr, 0.f), pz = fmaxf(sz, 0.f);
  return sqrtf(pr * pr + pz * pz) + fmaxf(fminf(sr, 0.f), fminf(sz, 0.f));
}

// is_outside_aabb(c, r) against the stored AABB (primitive_sdf.hpp:180-184).
__device__ __forceinline__ bool prim_aabb_outside(const float4 lb, const float4 ub, const float3 c,
                                                  const float r) {
  return c.x < lb.x - r || c.x > ub.x + r || c.y < lb.y - r || c.y > ub.y + r || c.z < lb.z - r ||
         c.z > ub.z + r;
}

}  // namespace pmp
