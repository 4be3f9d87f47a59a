// pmp_warp.cuh — v2 validity check: one configuration per lane for the uniform part, the whole
// warp for the irregular part.
//
// ncu on the v1 kernel (profiles/r01_v1_check_configs_ncu.md) showed 8.07 of 32 lanes active per
// issued instruction: the narrow phases (per-sphere tests of the few groups / group pairs that
// survive the broad phase) run under heavy divergence.  v2 splits the reference's
// check_ext_collision / check_self_collision (constraints/primitive_sphere_collision.cpp:191-254)
// differently — the result is unchanged because validity does not depend on the broad-phase
// hierarchy or on the order of the tests (SURVEY finding 7, App. A.4):
//
//   phase 1  (lane = configuration, no divergence): FK over the link program, group bounding
//            spheres, broad phase against every primitive -> per-(lane, group) primitive mask;
//            frames, group centres and masks go to the per-thread shared-memory stash.
//   phase 2a (warp cooperative): for each group, the lanes that need its narrow phase are
//            compacted: sub-warps of LPI lanes (LPI = sphere count rounded up to a power of two)
//            each take one (configuration, group) item, one sphere per lane.
//   phase 2b (warp cooperative): the same for the listed group pairs that pass the bounding-
//            sphere test, with a sphere-vs-other-group cull before the sphere x sphere tests.
#pragma once

namespace pmp {

struct Stash {
  float* frames;    // [slot][7][thread]
  float* gc;        // [group][3][thread]
  unsigned* mask;   // [group][thread]
  unsigned* srctab; // [thread]: per-warp scratch, rank -> source lane
  int stride;       // blockDim.x
};

__device__ __forceinline__ Stash stash_view(const RobotDev& R, const Smem& T) {
  Stash s;
  s.stride = blockDim.x;
  s.frames = T.stash;
  s.gc = T.stash + R.n_slots * 7 * s.stride;
  s.mask = reinterpret_cast<unsigned*>(s.gc + R.n_groups * 3 * s.stride);
  s.srctab = s.mask + R.n_groups * s.stride;
  return s;
}

__device__ __forceinline__ Frame frame_at(const Stash& st, int slot, int tid) {
  const float* p = st.frames + (slot * 7) * st.stride + tid;
  const int n = st.stride;
  Frame f;
  f.qx = p[0]; f.qy = p[n]; f.qz = p[2 * n]; f.qw = p[3 * n];
  f.tx = p[4 * n]; f.ty = p[5 * n]; f.tz = p[6 * n];
  return f;
}

// mask bits: 0..29 primitive k needs a narrow phase, 30 cloud, 31 some primitive >= 30
constexpr unsigned kMaskCloud = 1u << 30;
constexpr unsigned kMaskOverflow = 1u << 31;

__device__ __forceinline__ bool sphere_hits_scene(const SceneDev& S, const Smem& T, const float3 c, const float r,
                                                  const unsigned pm) {
  unsigned mm = pm & 0x3fffffffu;
  while (mm) {
    const int k = __ffs(mm) - 1;
    mm &= mm - 1;
    if (!prim_is_outside(load_prim(T.prims, k), c, r)) return true;
  }
  if (pm & kMaskOverflow)
    for (int k = 30; k < S.n_prims; k++)
      if (!prim_is_outside(load_prim(T.prims, k), c, r)) return true;
  if (pm & kMaskCloud)
    if (!cloud_is_outside(S.cloud, c, r)) return true;
  return false;
}

// Compaction helper: the lanes whose bit is set in `ball` publish their lane id at their rank in
// the warp's scratch row, so that sub-warp k of round t can pick up item t * ipr + k with one LDS.
__device__ __forceinline__ int publish_sources(const Stash& st, unsigned ball, bool mine, int lane, int wbase) {
  __syncwarp();  // previous readers of the scratch row are done
  if (mine) st.srctab[wbase + __popc(ball & ((1u << lane) - 1u))] = (unsigned)lane;
  __syncwarp();
  return __popc(ball);
}

// All 32 lanes of the warp must call this (lanes without a configuration pass active = false).
template <bool APPROX, class QAcc>
__device__ bool warp_check(const RobotDev& R, const SceneDev& S, const Smem& T, const QAcc& q, const bool active) {
  constexpr unsigned full = 0xffffffffu;
  const int tid = threadIdx.x, lane = tid & 31, wbase = tid & ~31;
  const Stash st = stash_view(R, T);
  const bool ext = (S.n_prims > 0) || S.cloud.enabled;
  bool valid = active;

  // ---- phase 1: per-lane FK + broad phase ----
  if (active) {
    int gi = 0;
    walk_links<APPROX>(R, T, q, [&](int op, const Frame& f) -> bool {
      while (gi < R.n_groups) {
        const int4 gw = *reinterpret_cast<const int4*>(&T.groups[2 * gi + 1]);
        if (gw.x != op) break;
        const float4 g = T.groups[2 * gi];
        const Rot rm = rotmat(f);
        const float3 gc = xform(rm, f, g.x, g.y, g.z);
        float* p = st.gc + (gi * 3) * st.stride + tid;
        p[0] = gc.x; p[st.stride] = gc.y; p[2 * st.stride] = gc.z;
        unsigned pm = 0;
        if (ext && !(gw.w & 1)) {
          for (int k = 0; k < S.n_prims; k++)
            if (!prim_is_outside(load_prim(T.prims, k), gc, g.w)) pm |= (k < 30) ? (1u << k) : kMaskOverflow;
          if (S.cloud.enabled && !cloud_surely_outside(S.cloud, gc, g.w)) pm |= kMaskCloud;
        }
        st.mask[gi * st.stride + tid] = pm;
        gi++;
      }
      return true;
    });
  }
  __syncwarp();

  // ---- phase 2a: env narrow phase, items compacted over the warp ----
  if (ext) {
    for (int g = 0; g < R.n_groups; g++) {
      const int4 gw = *reinterpret_cast<const int4*>(&T.groups[2 * g + 1]);
      if (gw.w & 1) continue;
      const unsigned mine = valid ? st.mask[g * st.stride + tid] : 0u;
      unsigned ball = __ballot_sync(full, mine != 0u);
      if (!ball) continue;
      const int n = gw.z;
      const int lg = n <= 2 ? 1 : n <= 4 ? 2 : n <= 8 ? 3 : n <= 16 ? 4 : 5;
      const int lpi = 1 << lg, ipr = 32 >> lg;
      const int sub = lane & (lpi - 1), k = lane >> lg;
      const int fslot = gw.w >> 8;
      const int cnt = publish_sources(st, ball, mine != 0u, lane, wbase);
      for (int it = k; it - k < cnt; it += ipr) {
        const int src = it < cnt ? (int)st.srctab[wbase + it] : 32;
        bool hit = false;
        if (src < 32) {
          const int stid = wbase + src;
          const Frame f = frame_at(st, fslot, stid);
          const unsigned pm = st.mask[g * st.stride + stid];
          const Rot rm = rotmat(f);
          for (int s = sub; s < n && !hit; s += lpi) {
            const float4 sp = T.spheres[gw.y + s];
            hit = sphere_hits_scene(S, T, xform(rm, f, sp.x, sp.y, sp.z), sp.w, pm);
          }
        }
        const unsigned bad = __reduce_or_sync(full, hit ? (1u << src) : 0u);
        if ((bad >> lane) & 1u) valid = false;
      }
    }
  }

  // ---- phase 2b: self collision (strict <), 32 listed pairs at a time ----
  for (int p0 = 0; p0 < R.n_pairs; p0 += 32) {
    const int np = min(32, R.n_pairs - p0);
    unsigned pmask = 0;
    if (valid) {
      for (int j = 0; j < np; j++) {
        const int2 pr = T.pairs[p0 + j];
        const float ra = T.groups[2 * pr.x].w, rb = T.groups[2 * pr.y].w;
        const float* pa = st.gc + (pr.x * 3) * st.stride + tid;
        const float* pb = st.gc + (pr.y * 3) * st.stride + tid;
        const float dx = pa[0] - pb[0], dy = pa[st.stride] - pb[st.stride], dz = pa[2 * st.stride] - pb[2 * st.stride];
        const float rs = ra + rb;
        if (!(dx * dx + dy * dy + dz * dz > rs * rs)) pmask |= 1u << j;
      }
    }
    if (!__ballot_sync(full, pmask != 0u)) continue;
    for (int j = 0; j < np; j++) {
      unsigned ball = __ballot_sync(full, ((pmask >> j) & 1u) && valid);
      if (!ball) continue;
      const int2 pr = T.pairs[p0 + j];
      const float4 ga = T.groups[2 * pr.x], gb = T.groups[2 * pr.y];
      const int4 wa = *reinterpret_cast<const int4*>(&T.groups[2 * pr.x + 1]);
      const int4 wb = *reinterpret_cast<const int4*>(&T.groups[2 * pr.y + 1]);
      const int na = wa.z, nb = wb.z;
      const int lg = (na <= 16 && nb <= 16) ? 4 : 5;
      const int lpi = 1 << lg, ipr = 32 >> lg;
      const int sub = lane & (lpi - 1), k = lane >> lg;
      const unsigned submask = (lpi == 32) ? full : ((1u << lpi) - 1u);
      const int cnt = publish_sources(st, ball, ((pmask >> j) & 1u) && valid, lane, wbase);
      for (int it = k; it - k < cnt; it += ipr) {
        const int src = it < cnt ? (int)st.srctab[wbase + it] : 32;
        const bool have = src < 32;
        const int stid = wbase + (have ? src : lane);
        const Frame fa = frame_at(st, wa.w >> 8, stid), fb = frame_at(st, wb.w >> 8, stid);
        const Rot rma = rotmat(fa), rmb = rotmat(fb);
        const float* pa = st.gc + (pr.x * 3) * st.stride + stid;
        const float* pb = st.gc + (pr.y * 3) * st.stride + stid;
        const float3 gca = make3(pa[0], pa[st.stride], pa[2 * st.stride]);
        const float3 gcb = make3(pb[0], pb[st.stride], pb[2 * st.stride]);
        bool hit = false;
        for (int i0 = 0; i0 < na; i0 += lpi) {
          const int i = i0 + sub;
          const float4 sa = T.spheres[wa.y + min(i, na - 1)];
          const float3 ca = xform(rma, fa, sa.x, sa.y, sa.z);
          // cull: sphere i of A can only touch a sphere of B if it touches B's bounding sphere
          const float ex = ca.x - gcb.x, ey = ca.y - gcb.y, ez = ca.z - gcb.z, er = sa.w + gb.w;
          const bool keep_a = have && i < na && !(ex * ex + ey * ey + ez * ez > er * er);
          const unsigned MA = (__ballot_sync(full, keep_a) >> (k * lpi)) & submask;
          if (!__ballot_sync(full, MA != 0u)) continue;
          for (int j0 = 0; j0 < nb; j0 += lpi) {
            const int jj = j0 + sub;
            const float4 sb = T.spheres[wb.y + min(jj, nb - 1)];
            const float3 cb = xform(rmb, fb, sb.x, sb.y, sb.z);
            const float fx = cb.x - gca.x, fy = cb.y - gca.y, fz = cb.z - gca.z, fr = sb.w + ga.w;
            const bool keep_b = have && jj < nb && !(fx * fx + fy * fy + fz * fz > fr * fr);
            unsigned ma = MA;
            while (__ballot_sync(full, ma != 0u)) {
              const bool on = ma != 0u;
              const int ii = on ? (__ffs(ma) - 1) : 0;
              ma &= ma - 1u;
              const float ax = __shfl_sync(full, ca.x, ii, lpi);
              const float ay = __shfl_sync(full, ca.y, ii, lpi);
              const float az = __shfl_sync(full, ca.z, ii, lpi);
              const float ar = __shfl_sync(full, sa.w, ii, lpi);
              const float hx = ax - cb.x, hy = ay - cb.y, hz = az - cb.z, hr = ar + sb.w;
              if (on && keep_b && (hx * hx + hy * hy + hz * hz < hr * hr)) hit = true;
            }
          }
        }
        const unsigned bad = __reduce_or_sync(full, (have && hit) ? (1u << src) : 0u);
        if ((bad >> lane) & 1u) valid = false;
      }
    }
  }
  return valid;
}

}  // namespace pmp
