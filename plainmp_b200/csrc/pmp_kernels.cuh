// pmp_kernels.cuh — the hot-path kernels.
//
// Reference call stack restated per item (paths relative to cpp/plainmp/):
//   IneqConstraintBase::is_valid            constraints/primitive.hpp:113-117
//   ConstraintBase::update_kintree          constraints/primitive.hpp:38-62
//   SphereCollisionCst::check_ext_collision constraints/primitive_sphere_collision.cpp:191-220
//   SphereCollisionCst::check_self_collision                                      .cpp:222-254
//   evaluate_ext_collision / evaluate_self_collision (values)      .cpp:256-295, .cpp:324-361
//   CustomValidatorBase::checkMotion        ompl/motion_validator.cpp:17-79, :89-126
//
// Data layout.  Robot and scene tables live in HBM as 16-byte words (a few KB) and are staged
// into shared memory once per CTA; every table read in the hot loops is a warp-uniform LDS.128
// broadcast.  Per-thread link frames that a later op or the self-collision phase needs are
// kept in a shared-memory stash laid out [slot][component][thread] (bank-conflict free).
#pragma once
#include "pmp_cloud.cuh"
#include "pmp_device.cuh"

namespace pmp {

struct RobotDev {
  const float4* blob;  // [ops: 5 words][groups: 2 words][spheres: 1 word][pairs: 1 word per 2 pairs][pairs4: 1 word per pair][gmask: 1 word per 4 groups, optional]
  int blob_words;
  int n_ops, n_groups, n_spheres, n_pairs;
  int off_groups, off_spheres, off_pairs, off_pairs4;  // word offsets into blob
  int off_gmask;  // per-group scene mask (one unsigned per group), -1 = every group sees every scene entry
  int dof, q_dim, base_type;
  int n_slots;  // frames kept in the stash
  float base_pose[7];
};

struct SceneDev {
  const float4* prims;  // kPrimWords words per primitive
  int n_prims;
  CloudDev cloud;
};

// shared-memory carve-up common to all kernels
struct Smem {
  const float4* ops;
  const float4* groups;
  const float4* spheres;
  const int2* pairs;
  const float4* pairs4;  // per pair: group a, group b (int bits), (Ra + Rb)^2, -
  const float4* prims;
  const unsigned* gmask;  // nullptr unless the robot was created with per-group scene masks
  float* stash;
};

constexpr int kOpWords = 5;  // 16-byte words per link op (see OpWords in pmp_device.cuh + the v3 basis word)

__device__ __forceinline__ Smem stage_tables(const RobotDev& R, const SceneDev& S, float4* sm) {
  for (int i = threadIdx.x; i < R.blob_words; i += blockDim.x) sm[i] = __ldg(R.blob + i);
  float4* pr = sm + R.blob_words;
  for (int i = threadIdx.x; i < S.n_prims * kPrimWords; i += blockDim.x) pr[i] = __ldg(S.prims + i);
  __syncthreads();
  Smem t;
  t.ops = sm;
  t.groups = sm + R.off_groups;
  t.spheres = sm + R.off_spheres;
  t.pairs = reinterpret_cast<const int2*>(sm + R.off_pairs);
  t.pairs4 = sm + R.off_pairs4;
  t.gmask = R.off_gmask >= 0 ? reinterpret_cast<const unsigned*>(sm + R.off_gmask) : nullptr;
  t.prims = pr;
  t.stash = reinterpret_cast<float*>(pr + S.n_prims * kPrimWords);
  return t;
}

// slot code of a frame that is not stashed because it is the same for every configuration: the root
// op of a FIXED-base robot in the folded (`fast`) tables (frame = base pose)
constexpr int kSlotBase = 0xffff;

__device__ __forceinline__ void stash_put(float* stash, int slot, const Frame& f) {
  float* p = stash + (slot * 7) * blockDim.x + threadIdx.x;
  const int n = blockDim.x;
  p[0] = f.qx; p[n] = f.qy; p[2 * n] = f.qz; p[3 * n] = f.qw;
  p[4 * n] = f.tx; p[5 * n] = f.ty; p[6 * n] = f.tz;
}
__device__ __forceinline__ Frame stash_get(const float* stash, int slot) {
  const float* p = stash + (slot * 7) * blockDim.x + threadIdx.x;
  const int n = blockDim.x;
  Frame f;
  f.qx = p[0]; f.qy = p[n]; f.qz = p[2 * n]; f.qw = p[3 * n];
  f.tx = p[4 * n]; f.ty = p[5 * n]; f.tz = p[6 * n];
  return f;
}

// Reads of one configuration's coordinates.
struct QDirect {  // q[item][j]
  const float* q;
  __device__ __forceinline__ float operator()(int j) const { return __ldg(q + j); }
};
struct QLerp {  // s1 + (s2 - s1) * t  (OMPL RealVectorStateSpace::interpolate)
  const float* a;
  const float* b;
  float t;
  __device__ __forceinline__ float operator()(int j) const {
    const float x = __ldg(a + j), y = __ldg(b + j);
    return x + (y - x) * t;
  }
};

// (bt = R.base_type; the validity kernels of a FIXED base pass the constant, which drops the other two cases
// and their three sincosf from the kernel)
template <class QAcc>
__device__ __forceinline__ Frame base_frame_bt(const RobotDev& R, const QAcc& q, const int bt) {
  Frame f;
  if (bt == PMP_BASE_FIXED) {
    f.tx = R.base_pose[0]; f.ty = R.base_pose[1]; f.tz = R.base_pose[2];
    f.qx = R.base_pose[3]; f.qy = R.base_pose[4]; f.qz = R.base_pose[5]; f.qw = R.base_pose[6];
  } else if (bt == PMP_BASE_PLANAR) {  // primitive.hpp:52-57
    f.tx = q(R.dof); f.ty = q(R.dof + 1); f.tz = 0.f;
    quat_from_rpy(0.f, 0.f, q(R.dof + 2), f.qx, f.qy, f.qz, f.qw);
  } else {  // primitive.hpp:44-51
    f.tx = q(R.dof); f.ty = q(R.dof + 1); f.tz = q(R.dof + 2);
    quat_from_rpy(q(R.dof + 3), q(R.dof + 4), q(R.dof + 5), f.qx, f.qy, f.qz, f.qw);
  }
  return f;
}
template <class QAcc>
__device__ __forceinline__ Frame base_frame(const RobotDev& R, const QAcc& q) {
  return base_frame_bt(R, q, R.base_type);
}

// Walk the link program; call visit(op_index, frame) after each frame is known.
template <bool APPROX, class QAcc, class Visit>
__device__ __forceinline__ void walk_links(const RobotDev& R, const Smem& T, const QAcc& q, Visit&& visit) {
  const Frame base = base_frame(R, q);
  Frame cur = base;
  for (int i = 0; i < R.n_ops; i++) {
    const float4 w0 = T.ops[kOpWords * i], w1 = T.ops[kOpWords * i + 1], w2 = T.ops[kOpWords * i + 2];
    const int4 w3 = *reinterpret_cast<const int4*>(&T.ops[kOpWords * i + 3]);
    const int kind = __float_as_int(w1.w), qi = __float_as_int(w2.w);
    Frame parent = cur;
    if (w3.x != i - 1) parent = (w3.x < 0 || w3.w == kSlotBase) ? base : stash_get(T.stash, w3.w);
    const float qv = (kind != PMP_OP_FIXED) ? q(qi) : 0.f;
    cur = apply_op<APPROX>(parent, w0, w1, w2, kind, w3.y, qv);
    if (w3.z >= 0) stash_put(T.stash, w3.z, cur);
    if (!visit(i, cur)) return;
  }
}

}  // namespace pmp
#include "pmp_v3.cuh"
namespace pmp {

// waypoint of an edge: s2 itself (use_b) or s1 + (s2 - s1) * t
struct QEdge {
  const float* a;
  const float* b;
  float t;
  bool use_b;
  __device__ __forceinline__ float operator()(int j) const {
    const float y = __ldg(b + j);
    if (use_b) return y;
    const float x = __ldg(a + j);
    return x + (y - x) * t;
  }
};

// ---- kernels ------------------------------------------------------------------------------------
// Resident threads per SM the register allocation of the validity kernel is bounded for (A/B knob of the build).
#ifndef PMP_CFG_MIN_THREADS
#define PMP_CFG_MIN_THREADS 512  // the FIXED-base, primitives-only kernels come out at 76 registers (6 CTAs/SM) on their own; bounding them for 768 threads (same register count!) measured 3 % slower, for 640 4 % (profiles/r02_s15_*, r02_s5_ab_register_bound.txt)
#endif
// v3: lean phase 1, environment narrow phase in place per sphere group, self queue (pmp_v3.cuh).  MODE: see warp_check_interp.  B = blockDim.x (compile-time so that the
// stash offsets fold into immediates); CLOUD = the scene has a CloudSDF (keeps the tree traversal
// out of the primitives-only kernel).
template <int B, int MODE, bool CLOUD>
__global__ void __launch_bounds__(B, PMP_CFG_MIN_THREADS / B)
k_check_configs_v3(const RobotDev R, const SceneDev S, const float* __restrict__ q, const long long n,
                   uint32_t* __restrict__ bits, unsigned long long* __restrict__ n_valid) {
  extern __shared__ float4 sm[];
  const Smem T = stage_tables(R, S, sm);
  unsigned cnt = 0;
  const long long n_tiles = (n + B - 1) / B;
  for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const long long item = tile * B + threadIdx.x;
    const bool active = item < n;
    QDirect acc{q + (active ? item : 0) * R.q_dim};
    const bool ok = warp_check_v3<B, MODE, CLOUD>(R, S, T, acc, active);
    const unsigned b = __ballot_sync(0xffffffffu, ok);
    if ((threadIdx.x & 31) == 0 && active) {
      bits[item >> 5] = b;
      cnt += __popc(b);
    }
    __syncwarp();
  }
  if (n_valid && cnt) atomicAdd(n_valid, (unsigned long long)cnt);
}

// evaluate(): exact FK, min signed distances clipped at 0.1, optionally with the Jacobian rows
// (primitive_sphere_collision.cpp:256-383).  GRAD: the arg-min sphere (ext) and sphere pair (self)
// are remembered in pass 1; pass 2 walks the links once more and, for every controlled joint that is
// an ancestor of the arg-min sphere's link, adds the column of get_attached_point_jacobian
// (kinematics/algorithm.cpp:346-375: world_axis x (c - joint position), or world_axis for a
// prismatic joint) contracted with the distance gradient.  Two deliberate differences from the
// reference, both below its own test tolerance (tests/test_constraints.py:36-44, 4 decimals):
// the SDF gradient is the analytic one (the reference takes a forward difference with a 1e-6 step
// in FP64, .cpp:307-314, which FP32 cannot resolve), and the base-rotation columns are the analytic
// derivatives w.r.t. the roll/pitch/yaw entries of q (reference: forward difference, eps 1e-7,
// algorithm.cpp:388-430).
struct EvalArgmin {
  int ext_op, ext_prim;  // op carrying the arg-min sphere (-1 = cutoff case), primitive (-1 = cloud)
  float3 ext_c, ext_nn;  // its world centre; nearest cloud point when ext_prim < 0
  int self_op_a, self_op_b;
  float3 self_ca, self_cb;
};

// ancestor-or-self test on the topologically ordered op table (parent < index)
__device__ __forceinline__ bool op_is_ancestor(const Smem& T, const int anc, int x) {
  while (x > anc) x = __float_as_int(T.ops[kOpWords * x + 3].x);
  return x == anc;
}

// Pass 2 of evaluate(): the Jacobian rows at the arg-min sphere / sphere pair recorded in A.
template <class QAcc>
__device__ __forceinline__ void eval_jacobian_rows(const RobotDev& R, const SceneDev& S, const Smem& T, const QAcc& acc, const EvalArgmin& A,
                                                   const bool ext, float* __restrict__ je, float* __restrict__ js) {
  const float nanv = __int_as_float(0x7fc00000);
  for (int c = 0; c < R.q_dim; c++) {
    je[c] = ext ? 0.f : nanv;
    js[c] = (R.n_pairs > 0) ? 0.f : nanv;
  }
  float3 ge = make3(0.f, 0.f, 0.f);  // d(sdf)/d(centre) at the arg-min sphere
  if (A.ext_op >= 0) {
    if (A.ext_prim >= 0) {
      ge = prim_gradient(load_prim(T.prims, A.ext_prim), A.ext_c);
    } else {
      const float dx = A.ext_c.x - A.ext_nn.x, dy = A.ext_c.y - A.ext_nn.y, dz = A.ext_c.z - A.ext_nn.z;
      const float inv = rsqrtf(fmaxf(dx * dx + dy * dy + dz * dz, 1e-30f));
      ge = make3(dx * inv, dy * inv, dz * inv);
    }
  }
  float3 ns = make3(0.f, 0.f, 0.f);  // (c1 - c2) / |c1 - c2|
  if (A.self_op_a >= 0) {
    const float dx = A.self_ca.x - A.self_cb.x, dy = A.self_ca.y - A.self_cb.y, dz = A.self_ca.z - A.self_cb.z;
    const float inv = rsqrtf(fmaxf(dx * dx + dy * dy + dz * dz, 1e-30f));
    ns = make3(dx * inv, dy * inv, dz * inv);
  }
  if (A.ext_op < 0 && A.self_op_a < 0) return;
  walk_links<false>(R, T, acc, [&](int op, const Frame& f) -> bool {
    const int kind = __float_as_int(T.ops[kOpWords * op + 1].w);
    if (kind == PMP_OP_FIXED) return true;
    const float4 w2 = T.ops[kOpWords * op + 2];
    const int qi = __float_as_int(w2.w);
    // joint axis in the op's own frame: word 2 for a revolute joint; a prismatic op keeps it in word 4
    // (word 2 holds the axis folded into the parent's frame, which is what the translation needs)
    const float4 wj = (kind == PMP_OP_PRISMATIC) ? T.ops[kOpWords * op + 4] : w2;
    const float3 wa = quat_rot(f.qx, f.qy, f.qz, f.qw, make3(wj.x, wj.y, wj.z));  // axis w.r.t. the root
    auto column = [&](const float3 c) -> float3 {
      if (kind == PMP_OP_PRISMATIC) return wa;
      const float rx = c.x - f.tx, ry = c.y - f.ty, rz = c.z - f.tz;
      return make3(wa.y * rz - wa.z * ry, wa.z * rx - wa.x * rz, wa.x * ry - wa.y * rx);
    };
    if (A.ext_op >= 0 && op_is_ancestor(T, op, A.ext_op)) {
      const float3 col = column(A.ext_c);
      je[qi] = col.x * ge.x + col.y * ge.y + col.z * ge.z;
    }
    if (A.self_op_a >= 0) {
      float v = 0.f;
      if (op_is_ancestor(T, op, A.self_op_a)) {
        const float3 col = column(A.self_ca);
        v += col.x * ns.x + col.y * ns.y + col.z * ns.z;
      }
      if (op_is_ancestor(T, op, A.self_op_b)) {
        const float3 col = column(A.self_cb);
        v -= col.x * ns.x + col.y * ns.y + col.z * ns.z;
      }
      js[qi] = v;
    }
    return true;
  });
  // moving base (algorithm.cpp:378-430): translation columns are unit vectors, rotation columns
  // axis_k x (c - t_base) with the roll / pitch / yaw axes of R = Rz(yaw) Ry(pitch) Rx(roll).
  // For the self row every base column cancels (both points move rigidly with the base).
  if (R.base_type != PMP_BASE_FIXED && A.ext_op >= 0) {
    const Frame b = base_frame(R, acc);
    const float rx = A.ext_c.x - b.tx, ry = A.ext_c.y - b.ty, rz = A.ext_c.z - b.tz;
    auto rot_col = [&](const float ax, const float ay, const float az) -> float {
      return (ay * rz - az * ry) * ge.x + (az * rx - ax * rz) * ge.y + (ax * ry - ay * rx) * ge.z;
    };
    je[R.dof] = ge.x;
    je[R.dof + 1] = ge.y;
    if (R.base_type == PMP_BASE_PLANAR) {
      je[R.dof + 2] = rot_col(0.f, 0.f, 1.f);
    } else {
      je[R.dof + 2] = ge.z;
      float sy, cy, sp, cp;
      sincosf(acc(R.dof + 5), &sy, &cy);
      sincosf(acc(R.dof + 4), &sp, &cp);
      je[R.dof + 3] = rot_col(cy * cp, sy * cp, -sp);
      je[R.dof + 4] = rot_col(-sy, cy, 0.f);
      je[R.dof + 5] = rot_col(0.f, 0.f, 1.f);
    }
  }
}

// Scratch of the cooperative cloud stage of evaluate() (COOP): for every configuration the env spheres
// whose group passes the cloud's group-level cull, placed with exact FK — [config][slot] float4 (x, y, z, r)
// and, for the Jacobian rows, the op that carries each of them.
struct CloudCand {
  float4* sph;         // [n][stride]
  int* op;             // [n][stride] (GRAD only)
  int* count;          // [n]
  EvalArgmin* argmin;  // [n] (GRAD only): pass 1's arg-min record, completed by the cloud stage
  int stride;          // env spheres of the robot
};

// COOP: the scene's CloudSDF is NOT queried here; the candidate spheres are handed to k_eval_cloud_coop,
// which works each configuration with a whole warp (8 lanes per nearest-point query), and for GRAD the
// Jacobian rows are left to k_eval_jacobian.  Without a cloud the kernel does everything itself.
template <bool GRAD, bool COOP>
__global__ void __launch_bounds__(128)
k_eval_configs(const RobotDev R, const SceneDev S, const float* __restrict__ q, const long long n,
               float* __restrict__ out, float* __restrict__ jac, const CloudCand CC) {
  extern __shared__ float4 sm[];
  const Smem T = stage_tables(R, S, sm);
  const float cutoff = 0.1f;  // primitive_sphere_collision.hpp:115
  float* gc_stash = T.stash + R.n_slots * 7 * blockDim.x;
  const bool ext = (S.n_prims > 0) || S.cloud.enabled;
  const long long n_tiles = (n + blockDim.x - 1) / blockDim.x;
  for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const long long item = tile * blockDim.x + threadIdx.x;
    if (item >= n) continue;
    QDirect acc{q + item * R.q_dim};
    float min_ext = cutoff, min_self = cutoff;
    EvalArgmin A;
    A.ext_op = -1; A.ext_prim = -1; A.self_op_a = -1; A.self_op_b = -1;
    A.ext_c = A.ext_nn = A.self_ca = A.self_cb = make3(0.f, 0.f, 0.f);
    int gi = 0, n_cand = 0;
    auto visit = [&](int op, const Frame& f) -> bool {
      while (gi < R.n_groups) {
        const int4 gw = *reinterpret_cast<const int4*>(&T.groups[2 * gi + 1]);
        if (gw.x != op) break;
        const float4 g = T.groups[2 * gi];
        const Rot rm = rotmat(f);
        const float3 gc = xform(rm, f, g.x, g.y, g.z);
        if (R.n_pairs > 0) {
          float* p = gc_stash + (gi * 3) * blockDim.x + threadIdx.x;
          p[0] = gc.x; p[blockDim.x] = gc.y; p[2 * blockDim.x] = gc.z;
        }
        if (ext && !(gw.w & 1)) {
          for (int k = 0; k < S.n_prims; k++) {
            const Prim P = load_prim(T.prims, k);
            // group-level cull with the cutoff margin (.cpp:272-275)
            if (prim_is_outside(P, gc, g.w + cutoff)) continue;
            for (int s = 0; s < gw.z; s++) {
              const float4 sp = T.spheres[gw.y + s];
              const float3 c = xform(rm, f, sp.x, sp.y, sp.z);
              // sphere-level AABB pre-test (.cpp:280-283): result-relevant for Ground only
              if (P.type == PMP_PRIM_GROUND && (c.z < -(sp.w + cutoff) || c.z > sp.w + cutoff)) continue;
              const float val = prim_evaluate(P, c) - sp.w;
              if (val < min_ext) {
                min_ext = val;
                if (GRAD) {
                  A.ext_op = op; A.ext_prim = k; A.ext_c = c;
                }
              }
            }
          }
          // the scene's cloud: eval_launch always takes the COOP form when there is one (the candidates go to
          // k_eval_cloud_coop), so the primitives-only instantiation carries no tree traversal at all
          if (COOP && !cloud_surely_outside(S.cloud, gc, g.w + cutoff)) {
            for (int s = 0; s < gw.z; s++) {
              const float4 sp = T.spheres[gw.y + s];
              const float3 c = xform(rm, f, sp.x, sp.y, sp.z);
              const float rr = sp.w + cutoff;
              if (c.x < S.cloud.lbx - rr || c.x > S.cloud.ubx + rr || c.y < S.cloud.lby - rr ||
                  c.y > S.cloud.uby + rr || c.z < S.cloud.lbz - rr || c.z > S.cloud.ubz + rr)
                continue;
              const long long slot = item * CC.stride + n_cand;
              CC.sph[slot] = make_float4(c.x, c.y, c.z, sp.w);
              if (GRAD) CC.op[slot] = op;
              n_cand++;
            }
          }
        }
        gi++;
      }
      return true;
    };
    walk_links<false>(R, T, acc, visit);
    for (int p = 0; p < R.n_pairs; p++) {
      const int2 pr = T.pairs[p];
      const float4 ga = T.groups[2 * pr.x], gb = T.groups[2 * pr.y];
      const float* pa = gc_stash + (pr.x * 3) * blockDim.x + threadIdx.x;
      const float* pb = gc_stash + (pr.y * 3) * blockDim.x + threadIdx.x;
      const float dx = pa[0] - pb[0], dy = pa[blockDim.x] - pb[blockDim.x], dz = pa[2 * blockDim.x] - pb[2 * blockDim.x];
      const float rs = ga.w + gb.w + cutoff;
      if (dx * dx + dy * dy + dz * dz > rs * rs) continue;
      const int4 wa = *reinterpret_cast<const int4*>(&T.groups[2 * pr.x + 1]);
      const int4 wb = *reinterpret_cast<const int4*>(&T.groups[2 * pr.y + 1]);
      const Frame fbase = base_frame(R, acc);
      const Frame fa = (wa.w >> 8) == kSlotBase ? fbase : stash_get(T.stash, wa.w >> 8);
      const Frame fb = (wb.w >> 8) == kSlotBase ? fbase : stash_get(T.stash, wb.w >> 8);
      // The inner loop works in the link frame of one group ("inner"): a sphere of the other ("outer", the SMALLER
      // group: the carry costs a dozen instructions per outer sphere) is carried there once, x = Ri^-1 (co - ti), and
      // compared with the inner group's spheres as the table holds them — no placement per (i, j), and a square
      // root only for a pair that lowers the running minimum.  Ri is the matrix of a quaternion that is never
      // renormalised, i.e. |q|^2 times a rotation: Ri^-1 = Ri^T / |q|^4 and a world distance is |q|^2 times the
      // distance in that frame.  (The value differs from the world-frame formula by FP32 rounding, ~1e-7 m.)
      const bool sw = wa.z > wb.z;
      const int4 wo = sw ? wb : wa, wi = sw ? wa : wb;
      const Frame fo = sw ? fb : fa, fi = sw ? fa : fb;
      const Rot ro = rotmat(fo), ri = rotmat(fi);
      const float n2 = fi.qx * fi.qx + fi.qy * fi.qy + fi.qz * fi.qz + fi.qw * fi.qw;  // |q_i|^2
      const float n4 = n2 * n2, inv_n4 = 1.f / n4;
      for (int i = 0; i < wo.z; i++) {
        const float4 so = T.spheres[wo.y + i];
        const float3 co = xform(ro, fo, so.x, so.y, so.z);
        const float ux = co.x - fi.tx, uy = co.y - fi.ty, uz = co.z - fi.tz;
        const float xx = (ri.m00 * ux + ri.m10 * uy + ri.m20 * uz) * inv_n4, xy = (ri.m01 * ux + ri.m11 * uy + ri.m21 * uz) * inv_n4,
                    xz = (ri.m02 * ux + ri.m12 * uy + ri.m22 * uz) * inv_n4;
        for (int j = 0; j < wi.z; j++) {
          const float4 si = T.spheres[wi.y + j];
          const float lx = xx - si.x, ly = xy - si.y, lz = xz - si.z;
          const float d2 = n4 * (lx * lx + ly * ly + lz * lz);  // squared world distance of the centres
          const float reach = fmaxf(min_self + so.w + si.w, 0.f);  // a centre distance below this lowers the minimum
          if (d2 < reach * reach) {
            const float val = sqrtf(d2) - (so.w + si.w);
            if (val < min_self) {
              min_self = val;
              if (GRAD) {
                const float3 ci = xform(ri, fi, si.x, si.y, si.z);
                A.self_op_a = wa.x; A.self_op_b = wb.x;
                A.self_ca = sw ? ci : co; A.self_cb = sw ? co : ci;
              }
            }
          }
        }
      }
    }
    const float nanv = __int_as_float(0x7fc00000);
    out[2 * item] = ext ? min_ext : nanv;
    out[2 * item + 1] = (R.n_pairs > 0) ? min_self : nanv;
    if (COOP) {
      CC.count[item] = n_cand;
      if (GRAD) CC.argmin[item] = A;
      continue;  // k_eval_cloud_coop finishes the ext value, k_eval_jacobian writes the rows
    }
    if (!GRAD) continue;
    eval_jacobian_rows(R, S, T, acc, A, ext, jac + (2 * item) * R.q_dim, jac + (2 * item + 1) * R.q_dim);
  }
}

// The cloud stage of evaluate(), one WARP per configuration: the candidate spheres (those of the groups
// that passed the group-level cull in k_eval_configs<., true>) are taken 32 at a time — every lane reads
// the distance field's lower / upper bound for one of them (the upper bounds tighten the running minimum
// before any tree is touched), then the candidates the field cannot rule out go to the tree four at a
// time, 8 lanes per exact nearest-point query (cloud_nn8) bounded by the running minimum, which every
// answered query tightens for the ones that follow.  Same value as evaluate_ext_collision's loop
// (primitive_sphere_collision.cpp:256-295): min over env spheres of sqrt(NN^2) - rho - r, cut off at 0.1.
template <bool GRAD>
__global__ void __launch_bounds__(128)
k_eval_cloud_coop(const SceneDev S, const CloudCand CC, const long long n, float* __restrict__ out) {
  const CloudDev& C = S.cloud;
  constexpr unsigned full = 0xffffffffu;
  const int lane = threadIdx.x & 31, sub = lane & 7, k = lane >> 3;
  const unsigned smask = 0xffu << (8 * k);
  const long long n_warps = (long long)gridDim.x * (blockDim.x >> 5);
  for (long long cfg = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); cfg < n; cfg += n_warps) {
    const int m = CC.count[cfg];
    if (m == 0) continue;
    float cur = out[2 * cfg];  // min over the primitives (or the cutoff); warp-uniform from here on
    int best_slot = -1, best_pos = -1;
    for (int c0 = 0; c0 < m; c0 += 32) {
      const int slot = c0 + lane;
      float4 sp = make_float4(0.f, 0.f, 0.f, 0.f);
      float lo = INFINITY, hi = INFINITY;
      if (slot < m) {
        sp = CC.sph[cfg * CC.stride + slot];
        cloud_bounds(C, make3(sp.x, sp.y, sp.z), lo, hi);
      }
      // the field's upper bounds: an upper bound of the answer (used only to prune, never reported)
      float ub = (slot < m && hi < INFINITY) ? (hi * (1.f + 1e-6f) + 1e-6f) - C.rho_min - sp.w : INFINITY;  // (a point that near has AT LEAST the smallest radius)
#pragma unroll
      for (int off = 16; off; off >>= 1) ub = fminf(ub, __shfl_xor_sync(full, ub, off));
      const float prune = fminf(cur, ub + 2e-6f);
      unsigned cand = __ballot_sync(full, slot < m && lo - C.rho - sp.w < prune);
      while (cand) {
        // the next four candidates, one per 8-lane sub-warp
        int owner = -1;
        unsigned rest = cand;
        for (int j = 0; j <= k && rest; j++) {
          owner = __ffs(rest) - 1;
          rest &= rest - 1u;
          if (j < k) owner = -1;
        }
        for (int j = 0; j < 4 && cand; j++) cand &= cand - 1u;
        const int src = owner < 0 ? 0 : owner;
        const float4 qs = make_float4(__shfl_sync(full, sp.x, src), __shfl_sync(full, sp.y, src), __shfl_sync(full, sp.z, src), __shfl_sync(full, sp.w, src));
        const float qlo = __shfl_sync(full, lo, src);
        float val = INFINITY;
        int pos = -1;
        // one bounded search per radius class (one class, no filter, unless clouds of different radii were merged)
        for (int kc = 0; kc < C.n_classes; kc++) {
          const bool one = C.n_classes == 1;
          const float rho = one ? C.rho : C.rho_cls[kc];
          const float need = fminf(fminf(cur, prune), val) + qs.w + rho;  // only a nearest point closer than this lowers the minimum
          if (owner >= 0 && qlo < need && need > 0.f) {
            // strict search against a bound nudged up by one ulp, so that a point AT the bound is found
            const float b2 = __uint_as_float(__float_as_uint(need * need) + 1u);
            int idx, ps;
            const float d2 = cloud_nn8(C, make3(qs.x, qs.y, qs.z), b2, smask, sub, 8 * k, &idx, &ps, one ? -1 : kc);
            if (idx >= 0) {
              val = sqrtf(d2) - rho - qs.w;
              pos = ps;
            }
          }
        }
        // warp-wide minimum over the four answers (ties -> the lower sub-warp, deterministic)
        float bv = val;
        int bk = k;
#pragma unroll
        for (int off = 16; off >= 8; off >>= 1) {
          const float ov = __shfl_xor_sync(full, bv, off);
          const int ok = __shfl_xor_sync(full, bk, off);
          if (ov < bv || (ov == bv && ok < bk)) {
            bv = ov;
            bk = ok;
          }
        }
        if (bv < cur) {
          cur = bv;
          if (GRAD) {
            best_slot = c0 + __shfl_sync(full, src, bk * 8);
            best_pos = __shfl_sync(full, pos, bk * 8);
          }
        }
      }
    }
    if (lane == 0) {
      out[2 * cfg] = cur;
      if (GRAD && best_slot >= 0) {
        EvalArgmin& A = CC.argmin[cfg];
        const float4 sp = CC.sph[cfg * CC.stride + best_slot];
        const float4 x = C.pts[best_pos];
        A.ext_op = CC.op[cfg * CC.stride + best_slot];
        A.ext_prim = -1;
        A.ext_c = make3(sp.x, sp.y, sp.z);
        A.ext_nn = make3(x.x, x.y, x.z);
      }
    }
  }
}

// Pass 2 of evaluate() as its own kernel (COOP scenes): Jacobian rows from the completed arg-min records.
__global__ void __launch_bounds__(128)
k_eval_jacobian(const RobotDev R, const SceneDev S, const float* __restrict__ q, const long long n, const CloudCand CC,
                float* __restrict__ jac) {
  extern __shared__ float4 sm[];
  const Smem T = stage_tables(R, S, sm);
  const bool ext = (S.n_prims > 0) || S.cloud.enabled;
  const long long n_tiles = (n + blockDim.x - 1) / blockDim.x;
  for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const long long item = tile * blockDim.x + threadIdx.x;
    if (item >= n) continue;
    QDirect acc{q + item * R.q_dim};
    const EvalArgmin A = CC.argmin[item];
    eval_jacobian_rows(R, S, T, acc, A, ext, jac + (2 * item) * R.q_dim, jac + (2 * item + 1) * R.q_dim);
  }
}

// get_all_spheres for a batch: out[item][sphere][4], caller's sphere order
template <bool APPROX>
__global__ void __launch_bounds__(128)
k_fk_spheres(const RobotDev R, const SceneDev S, const float* __restrict__ q, const long long n,
             float4* __restrict__ out) {
  extern __shared__ float4 sm[];
  const Smem T = stage_tables(R, S, sm);
  const long long n_tiles = (n + blockDim.x - 1) / blockDim.x;
  for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const long long item = tile * blockDim.x + threadIdx.x;
    if (item >= n) continue;
    QDirect acc{q + item * R.q_dim};
    int gi = 0;
    walk_links<APPROX>(R, T, acc, [&](int op, const Frame& f) -> bool {
      while (gi < R.n_groups) {
        const int4 gw = *reinterpret_cast<const int4*>(&T.groups[2 * gi + 1]);
        if (gw.x != op) break;
        const Rot rm = rotmat(f);
        for (int s = 0; s < gw.z; s++) {
          const float4 sp = T.spheres[gw.y + s];
          const float3 c = xform(rm, f, sp.x, sp.y, sp.z);
          out[item * R.n_spheres + gw.y + s] = make_float4(c.x, c.y, c.z, sp.w);
        }
        gi++;
      }
      return true;
    });
  }
}

// link frames for a batch: out[item][op][7] = x,y,z,qx,qy,qz,qw
template <bool APPROX>
__global__ void __launch_bounds__(128)
k_fk_frames(const RobotDev R, const SceneDev S, const float* __restrict__ q, const long long n,
            float* __restrict__ out) {
  extern __shared__ float4 sm[];
  const Smem T = stage_tables(R, S, sm);
  const long long n_tiles = (n + blockDim.x - 1) / blockDim.x;
  for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const long long item = tile * blockDim.x + threadIdx.x;
    if (item >= n) continue;
    QDirect acc{q + item * R.q_dim};
    walk_links<APPROX>(R, T, acc, [&](int op, const Frame& f) -> bool {
      float* o = out + (item * R.n_ops + op) * 7;
      o[0] = f.tx; o[1] = f.ty; o[2] = f.tz; o[3] = f.qx; o[4] = f.qy; o[5] = f.qz; o[6] = f.qw;
      return true;
    });
  }
}

// ---- motion edges ---------------------------------------------------------------------------------
struct ValidatorDev {
  int type;           // PMP_VALIDATOR_*
  double resolution;  // euclidean
  const double* width;      // [q_dim] device
  const double* inv_width;  // [q_dim] device (1.0 / width, motion_validator.cpp:81-87)
  const uint8_t* seq;       // [128][128] SEQUENCE_TABLE (generate_sequence.py), device
};

// determine_step_ratio in FP64 (the waypoint count must not depend on FP32 rounding).  T = float for
// FP32 endpoints (every float is a double, so this is what the reference computes for those states) or
// double for the reference's own state type.
template <class T>
__device__ __forceinline__ double edge_step_ratio(const ValidatorDev& V, const T* a, const T* b, int dim) {
  if (V.type == PMP_VALIDATOR_BOX) {  // motion_validator.cpp:89-111
    double diff_longest = 0.0, max_diff = -INFINITY;
    int longest = 0;
    for (int i = 0; i < dim; i++) {
      const double diff = (double)__ldg(b + i) - (double)__ldg(a + i);
      const double sc = fabs(diff) * V.inv_width[i];
      if (sc > max_diff) {
        max_diff = sc;
        longest = i;
        diff_longest = diff;
      }
    }
    if (fabs(diff_longest) < 1e-6) return INFINITY;
    return V.width[longest] / fabs(diff_longest);
  }
  double d2 = 0.0;  // motion_validator.cpp:113-126
  for (int i = 0; i < dim; i++) {
    const double d = (double)__ldg(a + i) - (double)__ldg(b + i);
    d2 += d * d;
  }
  return V.resolution / sqrt(d2);
}

// n_test of motion_validator.cpp:45 from the step ratio: 0 = "no check at all" (ratio == inf, :41-43).
// The reference converts floor(1 / ratio) + 2 to size_t; anything beyond 2^31 waypoints is clamped (such
// an edge never finishes there either), NaN (non-finite endpoints) becomes 2 = "s2 only".
__host__ __device__ __forceinline__ int edge_n_test(const double ratio) {
  if (ratio == INFINITY) return 0;
  const double nt = floor(1.0 / ratio) + 2.0;
  if (!(nt == nt)) return 2;
  return nt > 2147483000.0 ? 2147483000 : (nt < 2.0 ? 2 : (int)nt);
}

// FP64 -> FP32 of a resident batch (the pinned-FP64 host path: the doubles cross PCIe as they are and are
// narrowed here, instead of by host threads that the host's memory bandwidth limits)
__global__ void __launch_bounds__(256)
k_f64_to_f32(const double* __restrict__ src, float* __restrict__ dst, const long long n) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) dst[i] = (float)src[i];
}

// Pre-pass of the edge kernel: everything FP64 about an edge happens here, once, instead of inside
// the persistent loop below (where it cost registers for the whole kernel): step ratio and n_test per
// edge, the FP64 -> FP32 conversion of the endpoints when the caller holds doubles, and the zeroing
// of the result words and of the chunk dispenser, and the keys of the longest-first order.
template <class T>
__global__ void __launch_bounds__(256)
k_edge_prepare(const ValidatorDev V, const T* __restrict__ q0, const T* __restrict__ q1, const long long n, const int dim,
               float* __restrict__ q0f, float* __restrict__ q1f, double* __restrict__ ratio, int* __restrict__ ntest,
               uint32_t* __restrict__ bits, unsigned long long* __restrict__ next_chunk, uint8_t* __restrict__ sort_key,
               uint32_t* __restrict__ sort_idx) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  const long long t0 = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t0 == 0) *next_chunk = 0ull;
  for (long long w = t0; w < (n + 31) / 32; w += stride) bits[w] = 0u;
  if (q0f) {  // T = double: the kernels read FP32 endpoints
    for (long long i = t0; i < n * dim; i += stride) {
      q0f[i] = (float)q0[i];
      q1f[i] = (float)q1[i];
    }
  }
  for (long long e = t0; e < n; e += stride) {
    const double r = edge_step_ratio<T>(V, q0 + e * dim, q1 + e * dim, dim);
    ratio[e] = r;
    const int nt = edge_n_test(r);
    ntest[e] = nt;
    if (sort_key) {  // longest edges first (see EdgeArgs::order)
      sort_key[e] = (uint8_t)(255 - min(nt, 255));
      sort_idx[e] = (uint32_t)e;
    }
  }
}

// ---- motion edges: lane = edge with dynamic refill ------------------------------------------------
// Every lane walks ITS OWN edge through the reference's visiting order (s2 first, then the
// farthest-first table order, motion_validator.cpp:45-78) one waypoint per round, exactly like the
// serial reference; the 32 waypoints of a round (from 32 different edges) go through one
// warp_check_v3.  A lane whose edge is decided (first invalid waypoint, or all waypoints valid)
// immediately takes the next edge of the warp's chunk, and the warp takes a new chunk of
// kEdgeChunk edges from a global counter when its chunk runs dry, so no check is ever spent on a
// waypoint the reference would not have tested and the lanes stay full until the batch ends.
// Valid edges set their bit with one atomicOr (k_edge_prepare zeroes the words first).
// calls (optional): per edge, the number of is_valid calls = the reference's own count for that edge.
constexpr int kCoopAfter = 12, kCoopMin = 20;  // see the cooperative round in k_check_edges_v3
constexpr int kEdgeChunk = 32;  // one dispenser atomic per 32 edges: with 2960 resident warps, 256-edge chunks left a 1 M-edge batch with 1-2 chunks per warp (no balancing at all)

struct EdgeArgs {
  const float* q0;
  const float* q1;
  const double* ratio;  // [n] k_edge_prepare / the session's host pass
  const int* ntest;     // [n]
  // [n] or nullptr: the order in which the dispenser hands the edges out — by decreasing waypoint count
  // (one 8-bit radix pass over n_test).  A lane walks its edge serially, one waypoint per round, so a
  // 64-waypoint edge taken last would keep its warp (and the kernel) alive for 64 more rounds with
  // nothing left to overlap; taken first, it finishes while the short edges fill the other lanes.
  const uint32_t* order;
  long long n;
  uint32_t* bits;
  uint32_t* calls;  // optional
  unsigned long long* n_valid;
  unsigned long long* n_checked;
  unsigned long long* next_chunk;
};

template <int B, int MODE, bool CLOUD>
__global__ void __launch_bounds__(B, 512 / B)  // (78-80 registers for a FIXED base: 6 CTAs of 128 threads per SM where the stash allows)
k_check_edges_v3(const RobotDev R, const SceneDev S, const uint8_t* __restrict__ seq, const EdgeArgs A) {
  extern __shared__ float4 sm[];
  const Smem T = stage_tables(R, S, sm);
  constexpr unsigned full = 0xffffffffu;
  const int lane = threadIdx.x & 31;
  const unsigned lt = (1u << lane) - 1u;
  int next = 0, chunk_end = 0;  // warp-uniform: the unassigned part of the warp's chunk
  bool more_chunks = true;      // warp-uniform
  // per-lane edge state: edge index within this launch (the host splits batches of 2^31 edges or more), -1 = none
  int e = -1, w = 0, n_test = 0;
  unsigned cnt_valid = 0, cnt_checked = 0;
  auto decide = [&](const bool valid) {
    if (valid) {
      atomicOr(A.bits + (e >> 5), 1u << (e & 31));
      cnt_valid++;
    }
    if (A.calls) A.calls[e] = (unsigned)w;
    e = -1;
  };
  for (;;) {
    // ---- refill: every lane without an edge takes the next one ----
    for (;;) {
      const unsigned need = __ballot_sync(full, e < 0);
      if (!need) break;
      if (next >= chunk_end) {
        if (!more_chunks) break;
        unsigned long long c = 0;
        if (lane == 0) c = atomicAdd(A.next_chunk, (unsigned long long)kEdgeChunk);
        c = __shfl_sync(full, c, 0);
        if (c >= (unsigned long long)A.n) {
          more_chunks = false;
          break;
        }
        next = (int)c;
        chunk_end = (int)min((long long)c + kEdgeChunk, A.n);
        if (cnt_checked > (1u << 30)) {  // keep the 32-bit per-lane counter far from wrapping
          if (A.n_checked) atomicAdd(A.n_checked, (unsigned long long)cnt_checked);
          cnt_checked = 0;
        }
      }
      const int cand = next + __popc(need & lt);
      next = min(next + __popc(need), chunk_end);
      if (e < 0 && cand < chunk_end) {
        e = A.order ? (int)__ldg(A.order + cand) : cand;
        w = 0;
        n_test = __ldg(A.ntest + e);
        if (n_test == 0) decide(true);  // zero-length edge: valid without any check (motion_validator.cpp:41-43)
      }
    }
    const bool work = e >= 0;
    if (!__any_sync(full, work)) break;
    // ---- the round's mode (warp-uniform).  Normally every lane tests the next waypoint of ITS edge.  An edge
    // that has survived kCoopAfter waypoints and still has kCoopMin or more to go is worked by the whole warp
    // instead: 32 consecutive positions of the reference's visiting order in one round, the first failing
    // position decides (and gives the reference's call count; the tests behind it are the only ones the
    // reference would not have made, and an edge that got this far rarely fails).  A 116-waypoint valid
    // edge then takes 12 + 4 rounds instead of 116 — it was the serial tail of every launch. ----
    const unsigned coop = __ballot_sync(full, work && w >= kCoopAfter && n_test - 1 - w >= kCoopMin);
    const int owner = coop ? __ffs(coop) - 1 : 0;
    int ce = e, cn = n_test, ci = w;  // edge, its n_test, position in the visiting order (0 = s2)
    bool act = work;
    if (coop) {
      ce = __shfl_sync(full, e, owner);
      cn = __shfl_sync(full, n_test, owner);
      ci = __shfl_sync(full, w, owner) + lane;
      act = ci < cn - 1;
    }
    // ---- one waypoint per lane: s2 (position 0), then the interior points 1 .. n_test-2 ----
    const long long ge = act ? ce : 0;
    float t = 1.f;
    if (act && ci > 0) {
      const int idx = (cn < 129) ? (int)seq[(cn - 1) * 128 + (ci + 1)] : ci;  // sequence[i], i = position + 1
      t = (float)((double)idx * __ldg(A.ratio + ge));
    }
    QEdge acc{A.q0 + ge * R.q_dim, A.q1 + ge * R.q_dim, t, ci == 0};
    const bool ok = warp_check_v3<B, MODE, CLOUD>(R, S, T, acc, act);
    if (coop) {
      const unsigned fail = __ballot_sync(full, act && !ok);
      if (lane == owner) {
        const int consumed = fail ? __ffs(fail) : min(32, n_test - 1 - w);  // what the reference's walk would have called
        cnt_checked += consumed;
        w += consumed;
        if (fail) decide(false);
        else if (w >= n_test - 1) decide(true);
      }
    } else if (work) {
      cnt_checked++;
      w++;
      if (!ok) decide(false);  // first invalid waypoint decides the edge
      else if (w >= n_test - 1) decide(true);
    }
  }
  unsigned long long tot_valid = cnt_valid, tot_checked = cnt_checked;
  for (int off = 16; off; off >>= 1) {
    tot_valid += __shfl_down_sync(full, tot_valid, off);
    tot_checked += __shfl_down_sync(full, tot_checked, off);
  }
  if (lane == 0) {
    if (A.n_valid && tot_valid) atomicAdd(A.n_valid, tot_valid);
    if (A.n_checked && tot_checked) atomicAdd(A.n_checked, tot_checked);
  }
}

// ---- motion edges, waypoint-parallel: the latency path of the validator session --------------------
// The early-exit kernel above is the throughput path; a planner that asks about ONE edge (or a few
// hundred) wants the answer after one configuration-check latency instead.  Here lane = one
// (edge, waypoint): every waypoint the reference could test (s2 and t = j * ratio, j = 1 .. n_test-2;
// the farthest-first table visits the same set in another order, motion_validator.cpp:45-78) is
// tested in the same wave and an edge is valid iff none of its waypoints reports a collision —
// checkMotion's result is an AND, so the answer is the reference's.  The step ratio and the waypoint
// counts come from the host in FP64 (computed from the caller's doubles exactly as
// determine_step_ratio does); off[e] .. off[e + 1] are edge e's waypoints.
// Results are per WAYPOINT, so that the host can also report the reference's own is_valid call count
// (the position of the first failing waypoint in the reference's visiting order):
//   way_ok   (zero-copy mode, tiny batches): one byte per waypoint in host-mapped memory, plain stores;
//   way_bits (staged mode): one bit per waypoint, one word per warp (ballot), copied back by the host.
// n_edges <= kFlatSmemEdges: the offset table is staged in shared memory so the search does not walk
// host memory.
constexpr int kFlatSmemEdges = 1024;

template <int B, int MODE, bool CLOUD>
__global__ void __launch_bounds__(B, 512 / B)
k_check_waypoints_flat(const RobotDev R, const SceneDev S, const float* __restrict__ q0, const float* __restrict__ q1,
                       const double* __restrict__ ratio, const int* __restrict__ off, const int n_edges,
                       uint32_t* __restrict__ way_bits, uint8_t* __restrict__ way_ok) {
  extern __shared__ float4 sm[];
  __shared__ int s_off[kFlatSmemEdges + 1];
  const bool off_in_smem = n_edges <= kFlatSmemEdges;
  if (off_in_smem)
    for (int i = threadIdx.x; i <= n_edges; i += B) s_off[i] = off[i];
  const Smem T = stage_tables(R, S, sm);  // ends with a CTA barrier
  const int* o = off_in_smem ? s_off : off;
  const long long total = o[n_edges];
  const long long n_tiles = (total + B - 1) / B;
  for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const long long t = tile * B + threadIdx.x;
    const bool active = t < total;
    int e = 0;
    if (active) {  // largest e with off[e] <= t
      int lo = 0, hi = n_edges;
      while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if ((long long)o[mid] <= t) lo = mid; else hi = mid;
      }
      e = lo;
    }
    const int j = active ? (int)(t - o[e]) : 0;
    const float rate = j == 0 ? 1.f : (float)((double)j * __ldg(ratio + e));
    QEdge acc{q0 + (long long)e * R.q_dim, q1 + (long long)e * R.q_dim, rate, j == 0};
    const bool ok = warp_check_v3<B, MODE, CLOUD>(R, S, T, acc, active);
    if (way_ok) {
      if (active) way_ok[t] = ok ? 1 : 0;
    } else {
      const unsigned b = __ballot_sync(0xffffffffu, ok);
      if ((threadIdx.x & 31) == 0 && active) way_bits[t >> 5] = b;
    }
    __syncwarp();
  }
}

// ---- psdf / kdtree surface ------------------------------------------------------------------------
__global__ void k_sdf_evaluate(const SceneDev S, const float* __restrict__ pts, const long long n,
                               float* __restrict__ out) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float3 p = make3(pts[3 * i], pts[3 * i + 1], pts[3 * i + 2]);
    float v = 3.402823466e+38f;  // UnionSDF::evaluate starts from numeric_limits::max()
    for (int k = 0; k < S.n_prims; k++) v = fminf(v, prim_evaluate(load_prim(S.prims, k), p));
    if (S.cloud.enabled) {
      // min over the CloudSDF leaves of sqrt(NN_k^2) - rho_k: the nearest point of any class first, then
      // every other class only within the distance at which it could still win
      int bi;
      const float d_any = sqrtf(cloud_nn_exact(S.cloud, p, &bi));
      const int c0 = (S.cloud.n_classes == 1 || bi < 0) ? 0 : cloud_point_class(__int_as_float(bi));
      float vc = d_any - (S.cloud.n_classes == 1 ? S.cloud.rho : S.cloud.rho_cls[c0]);
      for (int kc = 0; kc < S.cloud.n_classes && S.cloud.n_classes > 1; kc++) {
        if (kc == c0) continue;
        const float b = vc + S.cloud.rho_cls[kc];
        if (!(b > 0.f)) continue;
        const float d2 = cloud_nn_within(S.cloud, p, b * b, nullptr, nullptr, kc);
        if (d2 <= b * b) vc = fminf(vc, sqrtf(d2) - S.cloud.rho_cls[kc]);
      }
      v = fminf(v, vc);
    }
    out[i] = v;
  }
}

// UnionSDF::is_outside / PrimitiveSDF::is_outside — NO AABB pre-test here (the binding calls
// is_outside directly, bindings/primitive.cpp:52), so Ground uses the plain plane test.
__global__ void k_sdf_is_outside(const SceneDev S, const float* __restrict__ pts, const float* __restrict__ radii,
                                 const long long n, uint8_t* __restrict__ out) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float3 p = make3(pts[3 * i], pts[3 * i + 1], pts[3 * i + 2]);
    const float r = radii ? radii[i] : 0.f;
    bool outside = true;
    for (int k = 0; k < S.n_prims && outside; k++) {
      const Prim P = load_prim(S.prims, k);
      if (P.type == PMP_PRIM_GROUND)
        outside = (p.z + P.pr.x > r);
      else
        outside = prim_is_outside(P, p, r);
    }
    if (outside && S.cloud.enabled) {
      // CloudSDF::is_outside(p, r) = evaluate(p) > r, no AABB
      float lo, hi;
      cloud_bounds(S.cloud, p, lo, hi);
      if (lo > r + S.cloud.rho + 1e-5f) outside = true;
      else if (hi < r + S.cloud.rho_min - 1e-5f) outside = false;
      else if (r + S.cloud.rho < 0.f) outside = true;
      else outside = !cloud_any_within(S.cloud, p, r);
    }
    out[i] = outside ? 1 : 0;
  }
}

// KDTree::sqdist / query for a batch: 8 lanes per query (4 queries per warp in flight), see cloud_nn8.
__global__ void __launch_bounds__(128)
k_cloud_nearest(const SceneDev S, const float* __restrict__ pts, const long long n, float* __restrict__ sqd,
                int* __restrict__ idx) {
  const CloudDev& C = S.cloud;
  const int lane = threadIdx.x & 31, sub = lane & 7, k = lane >> 3;
  const unsigned smask = 0xffu << (8 * k);
  const long long n_sub = (long long)gridDim.x * (blockDim.x / 8);
  for (long long i = (long long)blockIdx.x * (blockDim.x / 8) + (threadIdx.x >> 3); i < n; i += n_sub) {
    const float3 p = make3(__ldg(pts + 3 * i), __ldg(pts + 3 * i + 1), __ldg(pts + 3 * i + 2));
    float lo, hi;
    cloud_bounds(C, p, lo, hi);  // the distance field's upper bound starts the search ...
    // ... tightened by one greedy descent (n_levels node visits): the field only knows NN to within
    // the cell diagonal, and on flat geometry a shell that thick holds hundreds of leaf boxes; the
    // greedy leaf is usually the nearest point itself, which makes the search below a verification
    float bound2 = cloud_greedy_bound2_8(C, p, smask, sub);
    if (hi < INFINITY) {
      const float b = hi * (1.f + 1e-6f) + 1e-6f;
      bound2 = fminf(bound2, b * b);
    }
    // strict search against a bound nudged up by one ulp, so that a point AT the bound is found
    bound2 = __uint_as_float(__float_as_uint(fminf(fmaxf(bound2, 0.f), 3.0e38f)) + 1u);
    int bi;
    const float d2 = cloud_nn8(C, p, bound2, smask, sub, 8 * k, &bi);
    if (sub == 0) {
      sqd[i] = d2;
      if (idx) idx[i] = bi;
    }
  }
}

}  // namespace pmp
