// pmp_v3.cuh — validity check of one configuration per lane: lean phase 1, environment narrow phase in
// place per sphere group, self narrow phase from a warp-local queue.
//
//   phase 1  (lane = configuration, the warp converged): FK with the joint quaternion taken from a
//            host-precomputed basis (local = s * A + c * B, A = origin (x) (axis, 0), B = origin:
//            8 instructions instead of 19), points placed with q * v directly (no rotation
//            matrices), group bounding spheres against the primitives' AABBs (branch-free
//            centre / half-extent form) and the cloud's distance field;
//   env      right after a group's cull, while phase 1 is still at the group's link: the surviving
//            (configuration, member sphere) units of that ONE group, densely packed 32 per round, the
//            link frame fetched from the owning lane's registers by shuffles, the primitives visited in
//            a warp-uniform loop (v3_group_run).  Round 1 / early round 2 pushed (configuration, group,
//            primitive) items into a warp-local queue and drained them with four 8-lane sub-warps per
//            round; ncu had that drain at 32 % of the panda kernel's and 68 % of the PR2-vs-cloud
//            kernel's instructions at 16-19 of 32 lanes (sub-warps half empty on groups of 2-5 spheres,
//            the four sub-warps on different primitives and trip counts);
//   self     listed group pairs: bounding-sphere test per lane, survivors PUSHED into a warp-local queue
//            in shared memory (the warp is converged, so the queue tail is a warp-uniform register:
//            ballot + popc + one STS, no atomics), drained by four independent 8-lane sub-warps per
//            round.  If a push would overflow the queue, the pair loop stops, the queue is drained and
//            the loop resumes where it stopped.
//   A warp-uniform `bad` mask retires configurations (REDUX.OR); work of a retired configuration is
//   skipped from then on.
//
// The result is unchanged: validity depends neither on the broad-phase hierarchy nor on the order
// of the tests (SURVEY finding 7, App. A.4; reference: check_ext_collision / check_self_collision,
// constraints/primitive_sphere_collision.cpp:191-254).
#pragma once

namespace pmp {

constexpr int kQS = 32;   // self items per warp: lane | pair << 5
// Per-warp scratch, 512 + 128 bytes.  The 512-byte row is, during phase 1, the list of spheres the cloud's
// distance field could not decide (float4 centre + radius, gathered for tree walks four at a time) and,
// afterwards, the self drain's row of compacted spheres (float4[32]); the 128 bytes are, during phase 1,
// two byte tables (rank -> lane of the group run being worked, lane of each pending sphere) and,
// afterwards, the self queue.  The two uses of each are never live together.
constexpr int kRowWords = 128;
constexpr int kPend = 32;   // pending spheres per warp
static_assert(kPend * 16 == kRowWords * 4 && 32 + kPend <= kQS * 4, "the pending list and its byte tables live in the row / the self queue");
constexpr int kV3WordsPerThread = (kRowWords + kQS) / 32;  // row + self queue, per warp
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
constexpr int kMaxGroupsV3 = 2048;
// bits of the per-(lane, group) cull mask: 0..29 primitive k, 30 cloud, 31 some primitive >= 30
constexpr unsigned kMaskCloud = 1u << 30;
constexpr unsigned kMaskOverflow = 1u << 31;

// shared-memory views with the CTA size B known at compile time (offsets fold into immediates)
template <int B>
struct V3View {
  float* frames;    // [slot][7][B]
  float* gc;        // [group][3][B] floats, or (GC16) the same shape in halves
  float4* scratch;  // [warp][32]   the 512-byte row
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
  v.qself = reinterpret_cast<unsigned*>(v.scratch) + kRowWords * (B / 32);
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
  unsigned* qself;
  float4* scratch;
  int tid, lane, wbase;
  unsigned lt;        // lanes below this one
  unsigned bad;       // warp-uniform: bit l set = lane l's configuration is known to be invalid
  int n_self;  // warp-uniform fill of the self queue
  bool active;
  // phase 1 only (see kPend)
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
  X.qself = X.V.qself + warp * kQS;
  X.scratch = X.V.scratch + warp * 32;
  X.bad = 0;
  X.n_self = 0;
  X.active = active;
  X.pend = X.scratch;
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

// Scenes with more than 30 primitives: the sphere against every primitive from index 30 on (bit 31 of the cull
// mask).  Out of line — a second inlined copy of the primitive tests made every validity kernel larger for a
// case the stock scenes never reach, and these kernels are sensitive to their instruction footprint.
__device__ __noinline__ bool v3_overflow_prims_hit(const float4* prims, const int n_prims, const bool have, const float cx,
                                                   const float cy, const float cz, const float r) {
  bool hit = false;
  for (int kp = 30; kp < n_prims; kp++) {
    const Prim P = load_prim(prims, kp);
    if (have) hit |= !prim_is_outside(P, make3(cx, cy, cz), r);
  }
  return hit;
}

// One (configuration, member sphere) unit of v3_group_run: the sphere at world centre c against the scene
// entries of the owner's mask spm.  The primitive loop is warp-uniform (um = union of the masks of the run).
template <int B, bool CLOUD>
__device__ __forceinline__ void v3_unit_test(const SceneDev& S, const Smem& T, V3Ctx<B>& X, const unsigned um, const bool have,
                                             const unsigned spm, const int src, const float3 c, const float r) {
  constexpr unsigned full = 0xffffffffu;
  bool hit = false;
  for (unsigned m = um & (kMaskCloud - 1u); m; m &= m - 1u) {
    const int kp = __ffs(m) - 1;
    const Prim P = load_prim(T.prims, kp);
    if (have && ((spm >> kp) & 1u)) hit |= !prim_is_outside(P, c, r);
  }
  if (um & kMaskOverflow) hit |= v3_overflow_prims_hit(T.prims, S.n_prims, have && (spm & kMaskOverflow), c.x, c.y, c.z, r);
  int st = 0;
  if (CLOUD && have && !hit && (spm & kMaskCloud)) st = cloud_classify(S.cloud, c, r);
  X.bad |= __reduce_or_sync(full, (hit || st == 1) ? (1u << src) : 0u);
  if (CLOUD) {
    unsigned und = __ballot_sync(full, st == 2 && !((X.bad >> src) & 1u));
    while (und) {  // (a round can leave more undecided spheres than the list has room for)
      const int rank = __popc(und & X.lt);
      const bool mine = ((und >> X.lane) & 1u) && X.n_pend + rank < kPend;
      if (mine) {
        X.pend[X.n_pend + rank] = make_float4(c.x, c.y, c.z, r);
        X.pend_src[X.n_pend + rank] = (unsigned char)src;
      }
      const unsigned took = __ballot_sync(full, mine);
      und &= ~took;
      X.n_pend += __popc(took);
      if (X.n_pend == kPend) v3_cloud_flush<B>(S, X);
    }
  }
}

// Narrow phase of ONE sphere group against the scene, for every configuration of the warp whose group
// bounding sphere passed the cull (`pm` = this lane's cull mask: bits 0..29 primitive k, 30 the point cloud,
// 31 some primitive >= 30; 0 for a lane that is retired or has no configuration).  It runs in place, right
// after the group's cull, while phase 1 is at the group's link:
//   * the group — member spheres, count — is the same for all those configurations, so the work is a dense
//     list of (configuration, member sphere) units: unit u belongs to the (u / count)-th lane with a non-zero
//     mask and member u % count; 32 units per round, one per lane, whatever the size of the group (an 8-lane
//     sub-warp per item leaves most lanes idle on groups of 2-5 spheres);
//   * the link frame is still in the registers of the lane that owns the configuration: a unit fetches it,
//     and the owner's mask, with eight shuffles — no frame stash read, no queue, no queue overflow.  (Letting
//     every lane place its OWN spheres when most lanes have work — no shuffles, no unit map — was measured
//     twice and lost both times: as a second copy of the unit test, which made the kernel larger, and folded
//     into this loop body, where a round per member at 60 - 80 % fill loses to dense rounds, profiles/r02_s15_*);
//   * primitives are visited in a warp-uniform loop over the UNION of the masks: the primitive's record and
//     its test variant are the same for all 32 units (one broadcast load, no divergence on the variant); a
//     unit whose owner's mask lacks the bit just sits the test out;
//   * cloud: the unit asks the distance field (cloud_classify); a sphere the field cannot decide joins the
//     pending list for a tree walk (v3_cloud_flush).
// A hit retires the configuration (bit in X.bad): its remaining units, and its later groups, are skipped.
// Same tests on the same spheres as check_ext_collision (constraints/primitive_sphere_collision.cpp:191-220);
// the order does not matter for the result.
template <int B, bool CLOUD>
__device__ __forceinline__ void v3_group_run(const SceneDev& S, const Smem& T, V3Ctx<B>& X, const Frame& cur, const unsigned pm,
                                             const int sphere_begin, const int count) {
  constexpr unsigned full = 0xffffffffu;
  const unsigned ball = __ballot_sync(full, pm != 0u);
  if (!ball || count <= 0) return;
  const unsigned um = __reduce_or_sync(full, pm);  // union of the masks
  if ((ball >> X.lane) & 1u) X.sel[__popc(ball & X.lt)] = (unsigned char)X.lane;
  __syncwarp();
  const int units = __popc(ball) * count;
  const float inv_count = 1.f / (float)count;
  for (int u0 = 0; u0 < units; u0 += 32) {
    const int u = u0 + X.lane;
    const int j = u < units ? (int)(((float)u + 0.5f) * inv_count) : 0;  // u / count (exact: u < 32 * count)
    const int s = u - j * count;
    const int src = u < units ? X.sel[j] : X.lane;
    Frame f;
    f.qx = __shfl_sync(full, cur.qx, src); f.qy = __shfl_sync(full, cur.qy, src);
    f.qz = __shfl_sync(full, cur.qz, src); f.qw = __shfl_sync(full, cur.qw, src);
    f.tx = __shfl_sync(full, cur.tx, src); f.ty = __shfl_sync(full, cur.ty, src);
    f.tz = __shfl_sync(full, cur.tz, src);
    const unsigned spm = __shfl_sync(full, pm, src);  // the owner's mask
    const bool have = u < units && !((X.bad >> src) & 1u);
    const float4 sp = T.spheres[sphere_begin + (have ? s : 0)];
    const float3 c = place(f, sp.x, sp.y, sp.z);
    v3_unit_test<B, CLOUD>(S, T, X, um, have, spm, src, c, sp.w);
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

// Table interpreter: FK over the link program, the sphere groups of a link right after its frame (cull + narrow
// phase in place, v3_group_run), then the listed group pairs (bounding-sphere test per lane, survivors queued
// for the sub-warp drain).  All 32 lanes of the warp must call this (lanes without a configuration pass
// active = false).
// MODE 0: exact sin/cos (moving base), FP32 group centres; 1: polynomial sin/cos (FIXED base), FP32
// group centres; 2: FIXED base with FP16 group centres (chosen by the host when it buys a CTA per SM)
template <int B, int MODE, bool CLOUD, class QAcc>
__device__ __forceinline__ bool warp_check_interp(const RobotDev& R, const SceneDev& S, const Smem& T, const QAcc& q,
                                                  const bool active) {
  constexpr bool APPROX = MODE != 0;  // FIXED base (constraints/primitive.hpp:39-60)
  constexpr bool GC16 = MODE == 2;
  V3Ctx<B> X = v3_begin<B>(R, T, active, GC16);
  const bool ext = (S.n_prims > 0) || CLOUD;
  const Frame base = base_frame_bt(R, q, APPROX ? (int)PMP_BASE_FIXED : R.base_type);  // MODE 1, 2: FIXED base by construction
  bool far = false;
  Frame cur = base;
  int gi = 0;
  for (int i = 0; i < R.n_ops; i++) {
    const float4* ow = T.ops + kOpWords * i;
    const float4 wB = ow[0], w1 = ow[1], w2 = ow[2];
    const int4 w3 = *reinterpret_cast<const int4*>(ow + 3);
    const float4 wA = ow[4];
    const int kind = __float_as_int(w1.w), qi = __float_as_int(w2.w);
    if (w3.x != i - 1) cur = (w3.x < 0 || w3.w == kSlotBase) ? base : v3_frame_at(X.V, w3.w, X.tid);
    const float qv = (kind != PMP_OP_FIXED) ? q(qi) : 0.f;
    cur = apply_op_v3<APPROX>(cur, wB, w1, w2, wA, kind, w3.y, qv);
    if (w3.z >= 0) v3_stash_frame(X, w3.z, cur);
    while (gi < R.n_groups) {
      const int4 gw = *reinterpret_cast<const int4*>(&T.groups[2 * gi + 1]);
      if (gw.x != i) break;
      const float4 g = T.groups[2 * gi];
      const float3 gc = place(cur, g.x, g.y, g.z);
      far |= v3_store_gc<B, GC16>(X, base, gi, gc);
      if (ext && !(gw.w & 1)) {
        unsigned pm = v3_cull_mask<CLOUD>(S, T, gc, g.w);
        if (T.gmask) pm &= T.gmask[gi];  // composite constraints: a group sees its own member's scene entries only
        if (!(X.active && !((X.bad >> X.lane) & 1u))) pm = 0u;
        v3_group_run<B, CLOUD>(S, T, X, cur, pm, gw.y, gw.z);
      }
      gi++;
    }
  }
  if (CLOUD && X.n_pend) v3_cloud_flush<B>(S, X);  // the list's bytes become the self queue / scratch row now
  __syncwarp();
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
