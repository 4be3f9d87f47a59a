// pmp_v3.cuh — v3 validity check: lean per-lane phase 1 + flat, warp-local work queues for the
// narrow phases.
//
// ncu on v2 (profiles/r01_v2b_check_configs_ncu.md): 259 warp-instructions per configuration, of
// which ~60 % sit in the narrow phases although only ~1.7 (configuration, group | pair) items per
// configuration survive the broad phase.  v2 walks every group and every listed pair with a
// ballot + publish + partially filled rounds; v3 instead
//
//   phase 1  (lane = configuration, uniform): FK with the joint quaternion taken from a
//            host-precomputed basis (local = s * A + c * B, A = origin (x) (axis, 0), B = origin:
//            8 instructions instead of 19), points placed with q * v directly (no rotation
//            matrices), group bounding spheres against the primitives' AABBs (branch-free
//            centre / half-extent form) -> every surviving (configuration, group, primitive)
//            triple is PUSHED into a warp-local queue in shared memory (the warp is converged, so
//            the queue tail is a warp-uniform register: ballot + popc + one STS, no atomics);
//   drain    one flat loop over the queue, four independent 8-lane sub-warps per round: an env
//            item is one group against ONE primitive (one sphere per lane, the primitive's test
//            variant resolved at scene build), a self item is one listed group pair.  A
//            warp-uniform `bad` mask retires configurations (REDUX.OR); later items of a retired
//            configuration are skipped.  If a push would overflow a queue, phase 1 stops, the
//            queue is drained and phase 1 resumes where it stopped, so any table size works with
//            fixed 512 + 128 byte queues.
//
// The result is unchanged: validity depends neither on the broad-phase hierarchy nor on the order
// of the tests (SURVEY finding 7, App. A.4; reference: check_ext_collision / check_self_collision,
// constraints/primitive_sphere_collision.cpp:191-254).
#pragma once

namespace pmp {

constexpr int kQE = 128;  // env items per warp:  lane | group << 5 | primitive << 16
constexpr int kQS = 32;   // self items per warp: lane | pair << 5
// The self drain's scratch row (float4[32] per warp) lives in the bytes of the env queue: every env
// item has been drained before the first self item is looked at, so the two are never live together.
static_assert(kQE * 4 == 32 * 16, "the env queue and the scratch row share their bytes");
constexpr int kV3WordsPerThread = (kQE + kQS) / 32;  // env queue (= scratch row) + self queue, per warp
// CLOUD kernels: the cloud narrow phase does not go through the queue (v3_cloud_run below).  The upper half of
// the env queue's bytes holds the spheres the distance field could not decide until enough of them have
// gathered for tree walks four at a time, and — during phase 1 only — the bytes of the self queue hold two
// small byte tables (rank -> lane of the run being worked, lane of each pending sphere).
constexpr int kQEC = 64;    // env (primitive) items per warp in a CLOUD kernel
constexpr int kPend = 16;   // pending spheres per warp: float4 (centre, radius) in the words kQEC .. kQE-1
static_assert(kQEC * 4 + kPend * 16 == kQE * 4 && 32 + kPend <= kQS * 4, "the pending list lives in the queues' bytes");
// (kSlotBase, pmp_kernels.cuh: the root frame of a FIXED-base robot is taken from the handle)
//
// Group centres in the stash.  They are read back only by the SELF broad phase (bounding spheres of a
// listed pair; culling of member spheres against the other group's bounding sphere), which merely
// has to be conservative: everything that survives goes through the exact FP32 narrow phase.  For a
// FIXED base (GC16) they are therefore kept as FP16 offsets from the base position — half the bytes,
// which buys one more resident CTA per SM — and every test that uses them is inflated by a bound on
// the rounding: |offset component| < kGc16Range  =>  component error <= 2^-12 * 4 m, centre error
// <= 1.7 mm, pair distance error <= 3.4 mm < kGc16Pair.  A configuration with a group centre farther
// than kGc16Range from the base (none of the stock robots can reach that) is flagged `far` and skips
// these culls altogether.  Moving bases keep FP32 centres.
constexpr float kGc16Range = 4.0f;
constexpr float kGc16Pair = 4.0e-3f;    // added to Ra + Rb of a pair test (host: pairs4.w)
constexpr float kGc16Sphere = 2.0e-3f;  // added to r + R of a sphere-vs-bounding-sphere cull
constexpr unsigned kSelfItemFar = 1u << 31;
constexpr unsigned kItemCloud = 0xffffu;     // primitive field: the scene's point cloud
constexpr unsigned kItemOverflow = 0xfffeu;  // primitive field: every primitive with index >= 30
constexpr int kMaxGroupsV3 = 2048;           // the group field of an env item has 11 bits
// bits of the per-(lane, group) cull mask: 0..29 primitive k, 30 cloud, 31 some primitive >= 30
constexpr unsigned kMaskCloud = 1u << 30;
constexpr unsigned kMaskOverflow = 1u << 31;

// shared-memory views with the CTA size B known at compile time (offsets fold into immediates)
template <int B>
struct V3View {
  float* frames;    // [slot][7][B]
  float* gc;        // [group][3][B] floats, or (GC16) the same shape in halves
  float4* scratch;  // [warp][32]   (aliases qenv)
  unsigned* qenv;   // [warp][kQE]
  unsigned* qself;  // [warp][kQS]
};

// floats per thread taken by the group centres
__host__ __device__ constexpr int v3_gc_floats(int n_groups, bool gc16) { return gc16 ? (3 * n_groups + 1) / 2 : 3 * n_groups; }

template <int B>
__device__ __forceinline__ V3View<B> v3_view(const RobotDev& R, const Smem& T, const bool gc16) {
  V3View<B> v;
  v.frames = T.stash;
  v.gc = T.stash + R.n_slots * 7 * B;
  v.scratch = reinterpret_cast<float4*>(v.gc + v3_gc_floats(R.n_groups, gc16) * B);
  v.qenv = reinterpret_cast<unsigned*>(v.scratch);
  v.qself = v.qenv + kQE * (B / 32);
  return v;
}

template <int B>
__device__ __forceinline__ Frame v3_frame_at(const V3View<B>& v, int slot, int tid) {
  const float* p = v.frames + slot * (7 * B) + tid;
  Frame f;
  f.qx = p[0]; f.qy = p[B]; f.qz = p[2 * B]; f.qw = p[3 * B];
  f.tx = p[4 * B]; f.ty = p[5 * B]; f.tz = p[6 * B];
  return f;
}

// frame of a group's op: the stash, or the constant base frame for kSlotBase
template <int B>
__device__ __forceinline__ Frame v3_group_frame(const V3View<B>& v, const Frame& basef, int slot, int tid) {
  if (slot == kSlotBase) return basef;
  return v3_frame_at(v, slot, tid);
}

// group centre of thread `stid` as stored: absolute (FP32) or relative to the base position (GC16)
template <int B, bool GC16>
__device__ __forceinline__ float3 v3_load_gc(const V3View<B>& V, const int g, const int stid) {
  if (GC16) {
    const __half* p = reinterpret_cast<const __half*>(V.gc) + g * (3 * B) + stid;
    return make3(__half2float(p[0]), __half2float(p[B]), __half2float(p[2 * B]));
  }
  const float* p = V.gc + g * (3 * B) + stid;
  return make3(p[0], p[B], p[2 * B]);
}

// t + q * v  (Eigen's q * v, algebraically equal to toRotationMatrix() * v for any q)
__device__ __forceinline__ float3 place(const Frame& f, const float vx, const float vy, const float vz) {
  const float3 r = quat_rot(f.qx, f.qy, f.qz, f.qw, make3(vx, vy, vz));
  return make3(f.tx + r.x, f.ty + r.y, f.tz + r.z);
}

// frame = parent (x) local(q).  Revolute: local quaternion = s * A + c * B with
// A = origin (x) (axis, 0) and B = origin (the product origin (x) (axis s, c) of
// kinematics/algorithm.cpp:140-147 is linear in (s, c)); for identity-flagged origins A = (axis, 0),
// B = (0, 0, 0, 1), which reproduces the reference's skipped product exactly.
template <bool APPROX>
__device__ __forceinline__ Frame apply_op_v3(const Frame& p, const float4 wB, const float4 w1, const float4 w2,
                                             const float4 wA, const int kind, const int flags, const float qv) {
  Frame f;
  float ltx = w1.x, lty = w1.y, ltz = w1.z;
  if (kind == PMP_OP_REVOLUTE) {
    float s, c;
    half_sincos<APPROX>(qv, s, c);
    const float lx = fmaf(s, wA.x, c * wB.x), ly = fmaf(s, wA.y, c * wB.y), lz = fmaf(s, wA.z, c * wB.z),
                lw = fmaf(s, wA.w, c * wB.w);
    quat_mul(p.qx, p.qy, p.qz, p.qw, lx, ly, lz, lw, f.qx, f.qy, f.qz, f.qw);
  } else {
    if (kind == PMP_OP_PRISMATIC) {
      ltx = fmaf(w2.x, qv, ltx);
      lty = fmaf(w2.y, qv, lty);
      ltz = fmaf(w2.z, qv, ltz);
    }
    if (flags & PMP_OPFLAG_ORIGIN_IDENTITY) {
      f.qx = p.qx; f.qy = p.qy; f.qz = p.qz; f.qw = p.qw;
    } else {
      quat_mul(p.qx, p.qy, p.qz, p.qw, wB.x, wB.y, wB.z, wB.w, f.qx, f.qy, f.qz, f.qw);
    }
  }
  if ((flags & PMP_OPFLAG_ZERO_XYZ) && kind != PMP_OP_PRISMATIC) {
    f.tx = p.tx; f.ty = p.ty; f.tz = p.tz;
  } else {
    const float3 r = quat_rot(p.qx, p.qy, p.qz, p.qw, make3(ltx, lty, ltz));
    f.tx = p.tx + r.x; f.ty = p.ty + r.y; f.tz = p.tz + r.z;
  }
  return f;
}

// ---- drains ---------------------------------------------------------------------------------------
// env items: 8 lanes per (configuration, group, primitive) item, one sphere per lane per round.
// Sub-warps are independent; they meet again at the REDUX that retires configurations.  (The scene's point
// cloud never appears here: see v3_cloud_run.)
template <int B>
__device__ __forceinline__ unsigned drain_env(const SceneDev& S, const Smem& T, const V3View<B>& V, const Frame& basef,
                                              const unsigned* qenv, const int n_env, unsigned bad, const int lane,
                                              const int wbase) {
  constexpr unsigned full = 0xffffffffu;
  const int sub = lane & 7, k = lane >> 3;
  for (int b0 = 0; b0 < n_env; b0 += 4) {
    const int it = b0 + k;
    unsigned hitbit = 0;
    if (it < n_env) {
      const unsigned h = qenv[it];
      const int src = h & 31;
      if (!((bad >> src) & 1u)) {
        const int4 gw = *reinterpret_cast<const int4*>(&T.groups[2 * ((h >> 5) & (kMaxGroupsV3 - 1)) + 1]);
        const Frame f = v3_group_frame(V, basef, gw.w >> 8, wbase + src);
        const unsigned kp = h >> 16;
        bool hit = false;
        if (kp < 30u) {
          const Prim P = load_prim(T.prims, kp);
          for (int s = sub; s < gw.z; s += 8) {
            const float4 sp = T.spheres[gw.y + s];
            hit |= !prim_is_outside(P, place(f, sp.x, sp.y, sp.z), sp.w);
          }
        } else {
          for (int s = sub; s < gw.z; s += 8) {
            const float4 sp = T.spheres[gw.y + s];
            const float3 c = place(f, sp.x, sp.y, sp.z);
            for (int kk = 30; kk < S.n_prims; kk++) hit |= !prim_is_outside(load_prim(T.prims, kk), c, sp.w);
          }
        }
        hitbit = hit ? (1u << src) : 0u;
      }
    }
    bad |= __reduce_or_sync(full, hitbit);
  }
  return bad;
}

// self items: 8 lanes per (configuration, pair) item, four independent sub-warps per round (their
// ballots / barriers use the sub-warp's own lane mask).  The lanes place 8 spheres of A at a time,
// cull them against B's bounding sphere and compact the survivors into the sub-warp's scratch row;
// then every lane owns one sphere of B (culled against A's bounding sphere) and tests it against the
// compacted A spheres (strict <, .cpp:240-248).
template <int B, bool GC16>
__device__ __forceinline__ unsigned drain_self(const Smem& T, const V3View<B>& V, const Frame& basef, const unsigned* qself,
                                               float4* scratch, const int n_self, unsigned bad, const int lane,
                                               const int wbase) {
  constexpr unsigned full = 0xffffffffu;
  const int sub = lane & 7, k = lane >> 3;
  const unsigned smask = 0xffu << (k * 8);
  float4* my = scratch + k * 8;
  for (int b0 = 0; b0 < n_self; b0 += 4) {
    const int it = b0 + k;
    unsigned hitbit = 0;
    if (it < n_self) {
      const unsigned h = qself[it];
      const int src = h & 31;
      if (!((bad >> src) & 1u)) {
        const float4 pw = T.pairs4[(h & ~kSelfItemFar) >> 5];
        const bool far = (h & kSelfItemFar) != 0u;  // GC16 only: the stored centres are not usable
        const int ga = __float_as_int(pw.x), gb = __float_as_int(pw.y);
        const float Ra = T.groups[2 * ga].w, Rb = T.groups[2 * gb].w;
        const int4 wa = *reinterpret_cast<const int4*>(&T.groups[2 * ga + 1]);
        const int4 wb = *reinterpret_cast<const int4*>(&T.groups[2 * gb + 1]);
        const int stid = wbase + src;
        const Frame fa = v3_group_frame(V, basef, wa.w >> 8, stid), fb = v3_group_frame(V, basef, wb.w >> 8, stid);
        float3 gca = v3_load_gc<B, GC16>(V, ga, stid), gcb = v3_load_gc<B, GC16>(V, gb, stid);
        const float slack = GC16 ? kGc16Sphere : 0.f;
        if (GC16) {
          gca = make3(gca.x + basef.tx, gca.y + basef.ty, gca.z + basef.tz);
          gcb = make3(gcb.x + basef.tx, gcb.y + basef.ty, gcb.z + basef.tz);
        }
        bool hit = false;
        for (int i0 = 0; i0 < wa.z; i0 += 8) {
          const int i = i0 + sub;
          const float4 sa = T.spheres[wa.y + i];  // reads past the group stay inside the CTA's tables and are masked
          const float3 ca = place(fa, sa.x, sa.y, sa.z);
          const float ex = ca.x - gcb.x, ey = ca.y - gcb.y, ez = ca.z - gcb.z, er = sa.w + Rb + slack;
          const bool keep_a = (i < wa.z) & (far | !(ex * ex + ey * ey + ez * ez > er * er));
          const unsigned MA = (__ballot_sync(smask, keep_a) >> (k * 8)) & 0xffu;
          if (!MA) continue;
          __syncwarp(smask);  // readers of the previous chunk are done
          if (keep_a) my[__popc(MA & ((1u << sub) - 1u))] = make_float4(ca.x, ca.y, ca.z, sa.w);
          __syncwarp(smask);
          const int ka = __popc(MA);
          for (int j0 = 0; j0 < wb.z; j0 += 8) {
            const int j = j0 + sub;
            const float4 sb = T.spheres[wb.y + j];
            const float3 cb = place(fb, sb.x, sb.y, sb.z);
            const float fx = cb.x - gca.x, fy = cb.y - gca.y, fz = cb.z - gca.z, fr = sb.w + Ra + slack;
            if ((j < wb.z) & (far | !(fx * fx + fy * fy + fz * fz > fr * fr))) {
              for (int t = 0; t < ka; t++) {
                const float4 a = my[t];
                const float hx = a.x - cb.x, hy = a.y - cb.y, hz = a.z - cb.z, hr = a.w + sb.w;
                hit |= hx * hx + hy * hy + hz * hz < hr * hr;
              }
            }
          }
        }
        hitbit = hit ? (1u << src) : 0u;
      }
    }
    bad |= __reduce_or_sync(full, hitbit);
  }
  return bad;
}

// ---- per-warp context and the building blocks of phase 1 -------------------------------------------
// Used by the table interpreter below.
template <int B>
struct V3Ctx {
  V3View<B> V;
  unsigned* qenv;
  unsigned* qself;
  float4* scratch;
  int tid, lane, wbase;
  unsigned lt;        // lanes below this one
  unsigned bad;       // warp-uniform: bit l set = lane l's configuration is known to be invalid
  int n_env, n_self;  // warp-uniform queue fills
  bool active;
  // CLOUD kernels, phase 1 only (see kQEC / kPend)
  float4* pend;             // [kPend] centre + radius of the spheres waiting for a tree walk
  unsigned char* sel;       // [32] rank -> lane of the run being worked
  unsigned char* pend_src;  // [kPend] lane (= configuration) of each pending sphere
  int n_pend;               // warp-uniform
};

template <int B>
__device__ __forceinline__ V3Ctx<B> v3_begin(const RobotDev& R, const Smem& T, const bool active, const bool gc16) {
  V3Ctx<B> X;
  X.tid = threadIdx.x;
  X.lane = X.tid & 31;
  X.wbase = X.tid & ~31;
  X.lt = (1u << X.lane) - 1u;
  X.V = v3_view<B>(R, T, gc16);
  const int warp = X.tid >> 5;
  X.qenv = X.V.qenv + warp * kQE;
  X.qself = X.V.qself + warp * kQS;
  X.scratch = X.V.scratch + warp * 32;
  X.bad = 0;
  X.n_env = 0;
  X.n_self = 0;
  X.active = active;
  X.pend = reinterpret_cast<float4*>(X.qenv + kQEC);
  X.sel = reinterpret_cast<unsigned char*>(X.qself);
  X.pend_src = X.sel + 32;
  X.n_pend = 0;
  return X;
}

template <int B>
__device__ __forceinline__ void v3_stash_frame(const V3Ctx<B>& X, const int slot, const Frame& f) {
  float* p = X.V.frames + slot * (7 * B) + X.tid;
  p[0] = f.qx; p[B] = f.qy; p[2 * B] = f.qz; p[3 * B] = f.qw;
  p[4 * B] = f.tx; p[5 * B] = f.ty; p[6 * B] = f.tz;
}

// GC16: offset from the base position in FP16; returns true when the offset leaves the range the
// inflation constants cover (the configuration is then `far`)
template <int B, bool GC16>
__device__ __forceinline__ bool v3_store_gc(const V3Ctx<B>& X, const Frame& basef, const int gi, const float3 gc) {
  if (GC16) {
    __half* p = reinterpret_cast<__half*>(X.V.gc) + gi * (3 * B) + X.tid;
    const float rx = gc.x - basef.tx, ry = gc.y - basef.ty, rz = gc.z - basef.tz;
    p[0] = __float2half_rn(rx); p[B] = __float2half_rn(ry); p[2 * B] = __float2half_rn(rz);
    return !(fmaxf(fmaxf(fabsf(rx), fabsf(ry)), fabsf(rz)) < kGc16Range);
  }
  float* p = X.V.gc + gi * (3 * B) + X.tid;
  p[0] = gc.x; p[B] = gc.y; p[2 * B] = gc.z;
  return false;
}

// Group-level cull of a bounding sphere against every primitive's AABB and the cloud's distance
// field.  Bits 0..29: primitive k passes, 30: cloud, 31: some primitive >= 30.
template <bool CLOUD>
__device__ __forceinline__ unsigned v3_cull_mask(const SceneDev& S, const Smem& T, const float3 gc, const float radius) {
  unsigned pm = 0;
  const int n30 = min(S.n_prims, 30);
  for (int kp = 0; kp < n30; kp++) {
    const float4* pw = T.prims + kPrimWords * kp;
    pm |= prim_aabb_outside(pw[5], pw[6], gc, radius) ? 0u : (1u << kp);
  }
  for (int kp = 30; kp < S.n_prims; kp++) {
    const float4* pw = T.prims + kPrimWords * kp;
    pm |= prim_aabb_outside(pw[5], pw[6], gc, radius) ? 0u : kMaskOverflow;
  }
  if (CLOUD && !cloud_surely_outside(S.cloud, gc, radius)) pm |= kMaskCloud;
  return pm;
}

// Push one queue item per set bit of this lane's cull mask `pm` for group gi, consuming the bits as
// they are queued.  Returns false when the items of the next round do not fit; `pm` then holds what
// is still to be queued.  The caller drains and either restarts the group from its full mask (items
// queued before are then only re-tested) or, for a group that cannot fit even into an empty queue,
// calls again with the remaining bits.
template <int B, int CAP = kQE>
__device__ __forceinline__ bool v3_push_env(V3Ctx<B>& X, const int gi, unsigned& pm) {
  constexpr unsigned full = 0xffffffffu;
  if (!(X.active && !((X.bad >> X.lane) & 1u))) pm = 0u;
  unsigned ball = __ballot_sync(full, pm != 0u);
  while (ball) {
    const int cnt = __popc(ball);
    if (X.n_env + cnt > CAP) return false;
    if (pm) {
      const unsigned b = __ffs(pm) - 1;
      pm &= pm - 1u;
      const unsigned kp = (b < 30u) ? b : (b == 30u ? kItemCloud : kItemOverflow);
      X.qenv[X.n_env + __popc(ball & X.lt)] = (unsigned)X.lane | ((unsigned)gi << 5) | (kp << 16);
    }
    X.n_env += cnt;
    ball = __ballot_sync(full, pm != 0u);
  }
  return true;
}

// ---- cloud narrow phase (CLOUD kernels) ---------------------------------------------------------------
// Tree walks of the pending spheres, four at a time (one per 8-lane sub-warp, in lockstep); a sphere whose
// configuration has been retired in the meantime is dropped.
template <int B>
__device__ __forceinline__ void v3_cloud_flush(const SceneDev& S, V3Ctx<B>& X) {
  constexpr unsigned full = 0xffffffffu;
  const int sub = X.lane & 7, k = X.lane >> 3;
  __syncwarp();
  for (int i0 = 0; i0 < X.n_pend; i0 += 4) {
    const int it = i0 + k;
    int src = 0;
    float4 pr = make_float4(0.f, 0.f, 0.f, 0.f);
    bool valid = it < X.n_pend;
    if (valid) {
      src = X.pend_src[it];
      pr = X.pend[it];
      valid = !((X.bad >> src) & 1u);
    }
    const bool hit = cloud_any_within8x4(S.cloud, make3(pr.x, pr.y, pr.z), pr.w, valid, sub, k * 8);
    X.bad |= __reduce_or_sync(full, hit ? (1u << src) : 0u);
  }
  X.n_pend = 0;
  __syncwarp();
}

// The scene's point cloud against ONE sphere group of every configuration of the warp that passed the
// group-level cull (`ball`, warp-uniform).  The group is the same for all of them, so the work is a dense
// list of (configuration, member sphere) units: unit u belongs to the (u / count)-th set lane of `ball`
// and member u % count, 32 units per round with one unit per lane whatever the size of the group (a
// sub-warp per item would leave most of its 8 lanes idle on the 2-5 sphere groups of a PR2).  The group's
// link frame is still in the registers of the lane that owns the configuration (phase 1 is at that link),
// so a unit fetches it with seven shuffles — no stash read, no queue.  A unit places its sphere and asks the
// distance field (cloud_classify): a certain hit retires the configuration, a sphere the field cannot
// decide joins the pending list.
template <int B>
__device__ __forceinline__ void v3_cloud_run(const SceneDev& S, const Smem& T, V3Ctx<B>& X, const Frame& cur, unsigned ball,
                                             const int sphere_begin, const int count) {
  constexpr unsigned full = 0xffffffffu;
  ball &= ~X.bad;
  if (!ball || count <= 0) return;
  if ((ball >> X.lane) & 1u) X.sel[__popc(ball & X.lt)] = (unsigned char)X.lane;
  __syncwarp();
  const int units = __popc(ball) * count;
  const float inv_count = 1.f / (float)count;
  for (int u0 = 0; u0 < units; u0 += 32) {
    const int u = u0 + X.lane;
    const bool have = u < units;
    const int j = have ? (int)(((float)u + 0.5f) * inv_count) : 0;  // u / count (exact: u < 32 * count)
    const int s = u - j * count;
    const int src = have ? X.sel[j] : X.lane;
    Frame f;
    f.qx = __shfl_sync(full, cur.qx, src); f.qy = __shfl_sync(full, cur.qy, src);
    f.qz = __shfl_sync(full, cur.qz, src); f.qw = __shfl_sync(full, cur.qw, src);
    f.tx = __shfl_sync(full, cur.tx, src); f.ty = __shfl_sync(full, cur.ty, src);
    f.tz = __shfl_sync(full, cur.tz, src);
    int st = 0;
    float4 c = make_float4(0.f, 0.f, 0.f, 0.f);
    if (have && !((X.bad >> src) & 1u)) {
      const float4 sp = T.spheres[sphere_begin + s];
      const float3 w = place(f, sp.x, sp.y, sp.z);
      c = make_float4(w.x, w.y, w.z, sp.w);
      st = cloud_classify(S.cloud, w, sp.w);
    }
    X.bad |= __reduce_or_sync(full, st == 1 ? (1u << src) : 0u);
    unsigned und = __ballot_sync(full, st == 2 && !((X.bad >> src) & 1u));
    while (und) {  // (a round can leave more undecided spheres than the list has room for)
      const int rank = __popc(und & X.lt);
      const bool mine = ((und >> X.lane) & 1u) && X.n_pend + rank < kPend;
      if (mine) {
        X.pend[X.n_pend + rank] = c;
        X.pend_src[X.n_pend + rank] = (unsigned char)src;
      }
      const unsigned took = __ballot_sync(full, mine);
      und &= ~took;
      X.n_pend += __popc(took);
      if (X.n_pend == kPend) v3_cloud_flush<B>(S, X);
    }
  }
  __syncwarp();  // the rank table is rewritten by the next run
}

// One listed group pair p; `near` = this lane's bounding spheres are within reach and the
// configuration is still alive.  Returns false when the items do not fit (drain, call again).
template <int B>
__device__ __forceinline__ bool v3_push_self(V3Ctx<B>& X, const int p, const bool near, const bool far = false) {
  const unsigned ball = __ballot_sync(0xffffffffu, near);
  if (ball) {
    const int cnt = __popc(ball);
    if (X.n_self + cnt > kQS) return false;
    if (near) X.qself[X.n_self + __popc(ball & X.lt)] = (unsigned)X.lane | ((unsigned)p << 5) | (far ? kSelfItemFar : 0u);
    X.n_self += cnt;
  }
  return true;
}

// Table interpreter: FK over the link program, groups in op order, single drain site per queue
// (phase 1 is resumable: when a push does not fit, the queue is drained and the group redone).
// All 32 lanes of the warp must call this (lanes without a configuration pass active = false).
// MODE 0: exact sin/cos (moving base), FP32 group centres; 1: polynomial sin/cos (FIXED base), FP32
// group centres; 2: FIXED base with FP16 group centres (chosen by the host when it buys a CTA per SM)
template <int B, int MODE, bool CLOUD, class QAcc>
__device__ __forceinline__ bool warp_check_interp(const RobotDev& R, const SceneDev& S, const Smem& T, const QAcc& q,
                                                  const bool active) {
  constexpr bool APPROX = MODE != 0;  // FIXED base (constraints/primitive.hpp:39-60)
  constexpr bool GC16 = MODE == 2;
  V3Ctx<B> X = v3_begin<B>(R, T, active, GC16);
  const bool ext = (S.n_prims > 0) || CLOUD;
  const Frame base = base_frame(R, q);
  bool far = false;
  unsigned pm = 0u;     // cull mask of the group being queued (what is left of it when resuming)
  int stall_gi = -1;    // group at which the queue last overflowed (warp-uniform)
  int cloud_gi = -1;    // last group whose cloud narrow phase has run (warp-uniform)
  bool resume = false;  // warp-uniform: the group stalled twice in a row, keep pm across the drain
  Frame cur = base;
  int i = 0, gi = 0;
  bool ready = false;
  for (;;) {
    bool stalled = false;
    for (; i < R.n_ops; i++) {
      if (!ready) {
        const float4* ow = T.ops + kOpWords * i;
        const float4 wB = ow[0], w1 = ow[1], w2 = ow[2];
        const int4 w3 = *reinterpret_cast<const int4*>(ow + 3);
        const float4 wA = ow[4];
        const int kind = __float_as_int(w1.w), qi = __float_as_int(w2.w);
        if (w3.x != i - 1) cur = (w3.x < 0 || w3.w == kSlotBase) ? base : v3_frame_at(X.V, w3.w, X.tid);
        const float qv = (kind != PMP_OP_FIXED) ? q(qi) : 0.f;
        cur = apply_op_v3<APPROX>(cur, wB, w1, w2, wA, kind, w3.y, qv);
        if (w3.z >= 0) v3_stash_frame(X, w3.z, cur);
        ready = true;
      }
      while (gi < R.n_groups) {
        const int4 gw = *reinterpret_cast<const int4*>(&T.groups[2 * gi + 1]);
        if (gw.x != i) break;
        if (!resume) {
          const float4 g = T.groups[2 * gi];
          const float3 gc = place(cur, g.x, g.y, g.z);
          far |= v3_store_gc<B, GC16>(X, base, gi, gc);
          pm = (ext && !(gw.w & 1)) ? v3_cull_mask<CLOUD>(S, T, gc, g.w) : 0u;
          if (T.gmask) pm &= T.gmask[gi];  // composite constraints: a group sees its own member's scene entries only
          if (CLOUD && cloud_gi != gi) {  // (a group restarted after a queue overflow has been through this already)
            const bool alive = X.active && !((X.bad >> X.lane) & 1u);
            v3_cloud_run<B>(S, T, X, cur, __ballot_sync(0xffffffffu, alive && (pm & kMaskCloud)), gw.y, gw.z);
            cloud_gi = gi;
          }
          if (CLOUD) pm &= ~kMaskCloud;
        }
        if (!v3_push_env<B, CLOUD ? kQEC : kQE>(X, gi, pm)) {
          // After the drain the group is normally restarted from its full mask.  A group that yields
          // more items than the whole queue holds (e.g. 5+ primitives passing the cull on all 32
          // lanes) would stall again in exactly the same place for ever if no lane died in between:
          // from its second consecutive stall on, continue with the bits that are still to be queued.
          resume = (stall_gi == gi);
          stall_gi = gi;
          stalled = true;
          break;
        }
        resume = false;
        gi++;
      }
      if (stalled) break;
      ready = false;
    }
    __syncwarp();
    X.bad = drain_env<B>(S, T, X.V, base, X.qenv, X.n_env, X.bad, X.lane, X.wbase);
    X.n_env = 0;
    if (!stalled) break;
    __syncwarp();
  }
  if (CLOUD && X.n_pend) v3_cloud_flush<B>(S, X);  // the list's bytes become the self queue / scratch row now
  // listed group pairs: bounding-sphere test per lane, survivors queued
  for (int p = 0; R.n_pairs > 0;) {
    const bool alive = active && !((X.bad >> X.lane) & 1u);
    for (; p < R.n_pairs; p++) {
      const float4 pw = T.pairs4[p];
      const float3 ca = v3_load_gc<B, GC16>(X.V, __float_as_int(pw.x), X.tid);
      const float3 cb = v3_load_gc<B, GC16>(X.V, __float_as_int(pw.y), X.tid);
      const float dx = ca.x - cb.x, dy = ca.y - cb.y, dz = ca.z - cb.z;
      // pw.z = (Ra + Rb)^2, pw.w = (Ra + Rb + kGc16Pair)^2
      const bool near = far | !(dx * dx + dy * dy + dz * dz > (GC16 ? pw.w : pw.z));
      if (!v3_push_self(X, p, alive & near, GC16 && far)) break;
    }
    __syncwarp();
    X.bad = drain_self<B, GC16>(T, X.V, base, X.qself, X.scratch, X.n_self, X.bad, X.lane, X.wbase);
    X.n_self = 0;
    if (p >= R.n_pairs) break;
    __syncwarp();
  }
  return active && !((X.bad >> X.lane) & 1u);
}

template <int B, int MODE, bool CLOUD, class QAcc>
__device__ __forceinline__ bool warp_check_v3(const RobotDev& R, const SceneDev& S, const Smem& T, const QAcc& q,
                                              const bool active) {
  return warp_check_interp<B, MODE, CLOUD>(R, S, T, q, active);
}

}  // namespace pmp
