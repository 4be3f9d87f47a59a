// plainmp_b200.cu — C ABI (include/plainmp_b200.h) over the sm_100a kernels.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -shared -Xcompiler -fPIC
// No torch, no CPU fallback: every compute entry point needs a CUDA device.
#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <condition_variable>
#include <cstdio>
#include <cstring>
#include <functional>
#include <mutex>
#include <string>
#include <thread>
#include <unordered_map>
#include <vector>

#include <cub/device/device_radix_sort.cuh>

#include "pmp_cloud_build.h"
#include "pmp_kernels.cuh"

using namespace pmp;

namespace {

thread_local std::string g_err;
std::atomic<uint64_t> g_launches{0};

int fail(int code, const std::string& msg) {
  g_err = msg;
  return code;
}
#define CUDA_OK(expr)                                                                          \
  do {                                                                                         \
    cudaError_t _e = (expr);                                                                   \
    if (_e != cudaSuccess)                                                                     \
      return fail(PMP_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e));           \
  } while (0)

constexpr int kBlock = 128;
constexpr int kMaxSmem = 200 * 1024;

// Everything the library caches on a device lives in a per-device record: the SM count, the occupancy
// cache (cudaFuncSetAttribute is per device), the sequence table, the box-width copies and the host-call
// workspaces.  The record of the calling thread's CURRENT device is used, so one process may drive
// several GPUs (pmp_set_device).
struct GridCacheEntry {
  int smem_attr = 0;                                  // MaxDynamicSharedMemorySize currently set
  std::vector<std::pair<std::pair<size_t, int>, int>> per_sm;  // ((smem, block), CTAs per SM)
};
struct ValidatorCache {  // device copies of box widths, keyed by content
  std::vector<double> width;
  double* dev = nullptr;  // [2][dim]: width, inv_width
  uint64_t last_use = 0;
};
struct HostWorkspace;
struct DeviceState {
  int ordinal = -1;
  int num_sms = 0;
  std::mutex grid_mu;
  std::unordered_map<const void*, GridCacheEntry> grid_cache;
  std::mutex seq_mu;
  uint8_t* seq_dev = nullptr;
  std::mutex val_mu;
  std::vector<ValidatorCache> val_cache;
  uint64_t val_clock = 0;
  std::mutex ws_mu;
  std::vector<HostWorkspace*> ws_free;  // idle host-call workspaces (one is checked out per *_host call)
};
std::mutex g_dev_mu;
std::vector<DeviceState*> g_devs;  // index = device ordinal; entries are never freed

int device_state(DeviceState** out) {
  int dev = 0;
  cudaError_t err = cudaGetDevice(&dev);
  if (err == cudaSuccess && dev < 0) err = cudaErrorNoDevice;
  if (err != cudaSuccess) return fail(PMP_ERR_CUDA, std::string("no usable CUDA device: ") + cudaGetErrorString(err));
  std::lock_guard<std::mutex> lk(g_dev_mu);
  if ((int)g_devs.size() <= dev) g_devs.resize(dev + 1, nullptr);
  if (!g_devs[dev]) {
    int n = 0;
    err = cudaGetDeviceCount(&n);
    if (err == cudaSuccess && n == 0) err = cudaErrorNoDevice;
    if (err != cudaSuccess) return fail(PMP_ERR_CUDA, std::string("no usable CUDA device: ") + cudaGetErrorString(err));
    auto* d = new DeviceState();
    d->ordinal = dev;
    cudaDeviceGetAttribute(&d->num_sms, cudaDevAttrMultiProcessorCount, dev);
    // stream-ordered scratch (cudaMallocAsync) is handed back to the pool, not to the driver, between calls
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
      uint64_t keep = ~0ull;
      cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
    }
    g_devs[dev] = d;
  }
  *out = g_devs[dev];
  return PMP_OK;
}
int ensure_device() {
  DeviceState* d;
  return device_state(&d);
}
int num_sms() {
  DeviceState* d;
  return device_state(&d) == PMP_OK ? d->num_sms : 1;
}

// Occupancy per (kernel, dynamic smem, block) is queried once per device and remembered: the two runtime
// calls cost several microseconds, which is most of a small launch on the session's latency path.
template <class K>
int grid_for(K kernel, size_t smem, long long n_tiles, int* grid, int block = kBlock) {
  DeviceState* D;
  if (int rc = device_state(&D)) return rc;
  int per_sm = 0;
  {
    std::lock_guard<std::mutex> lk(D->grid_mu);
    GridCacheEntry& e = D->grid_cache[reinterpret_cast<const void*>(kernel)];
    if ((int)smem > e.smem_attr) {
      CUDA_OK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      e.smem_attr = (int)smem;
    }
    for (auto& kv : e.per_sm)
      if (kv.first.first == smem && kv.first.second == block) per_sm = kv.second;
    if (!per_sm) {
      CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, block, smem));
      if (per_sm < 1) return fail(PMP_ERR_LIMIT, "kernel does not fit on an SM with these tables");
      e.per_sm.push_back({{smem, block}, per_sm});
    }
  }
  long long g = (long long)D->num_sms * per_sm;
  *grid = (int)std::max<long long>(1, std::min<long long>(g, n_tiles));
  return PMP_OK;
}

}  // namespace

// One flattened program on the device: the table blob + what the kernels need to carve it up.
struct RobotTables {
  RobotDev dev{};
  float4* blob_dev = nullptr;
  int v3_mode = 0;                          // MODE of the validity kernels (pmp_v3.cuh: warp_check_interp)
  size_t stash_floats_per_thread = 0;       // validity kernels (v3 layout) / FK kernels
  size_t stash_floats_per_thread_eval = 0;  // evaluate kernel: one word per group (+1) after the group centres
  int n_env_spheres = 0;                    // spheres of the groups that take part in environment checks
};

struct pmp_robot {
  // `full` is the program exactly as the caller described it (one op per link): it serves the
  // calls that report per-link or per-sphere quantities (evaluate, get_all_spheres, link poses).
  // `fast` is the same kinematics with every fixed op folded away on the host in FP64 (see
  // fold_fixed_ops): the validity kernels walk only the controlled joints.
  RobotTables full, fast;
};

struct pmp_scene {
  SceneDev dev{};
  float4* prims_dev = nullptr;
  // two grow-only buffer sets: a new cloud is indexed into the idle set while kernels in flight may
  // still walk the current one, then the sets swap; the build temporaries are kept too
  CloudBuffers cloud[2];
  int cloud_cur = 0;
  CloudScratch cloud_scratch;
};

static SceneDev empty_scene() {
  SceneDev s{};
  s.prims = nullptr;
  s.n_prims = 0;
  s.cloud.enabled = 0;
  return s;
}

static size_t extra_smem() {  // PMP_EXTRA_SMEM=<bytes>: occupancy experiments only
  static const size_t v = [] {
    const char* e = getenv("PMP_EXTRA_SMEM");
    return e ? (size_t)atol(e) : (size_t)0;
  }();
  return v;
}
static size_t smem_bytes(const RobotTables& r, const SceneDev& s, int block = kBlock, bool eval_kernel = false) {
  const size_t stash = eval_kernel ? r.stash_floats_per_thread_eval : r.stash_floats_per_thread;
  return (size_t)(r.dev.blob_words + s.n_prims * kPrimWords) * 16 + stash * 4 * block + extra_smem();
}
// Per-thread kernels (evaluate, FK dumps) read blockDim.x at run time: take the largest CTA whose
// stash fits the budget, so that very large robots (dozens of links with sphere groups) still run.
static int pick_block_per_thread(const RobotTables& r, const SceneDev& s, bool eval_kernel) {
  for (int b : {128, 64, 32})
    if (smem_bytes(r, s, b, eval_kernel) <= (size_t)kMaxSmem) return b;
  return 0;
}
// a robot with per-group scene masks addresses primitives through the 30 direct bits of the cull mask
static bool mask_limit_ok(const RobotTables& r, const SceneDev& s) { return r.dev.off_gmask < 0 || s.n_prims <= 30; }
// 128 threads per CTA unless the per-thread stash would leave fewer than 3 CTAs per SM
static int pick_block(const RobotTables& r, const SceneDev& s) {
  return smem_bytes(r, s, 128) * 3 <= (size_t)227 * 1024 ? 128 : 64;
}

namespace {
// Host-side (FP64) description of a program, the common input of build_tables.
struct OpD {
  double oq[4], oxyz[3], axis[3];
  double axis_child[3];  // the joint axis in the op's OWN frame (= axis unless folding moved a prismatic axis)
  int kind, q_index, parent, flags;
};
struct GroupD {
  double center[3], radius;
  int op, sphere_begin, sphere_count, only_self;
  unsigned prim_mask = 0xffffffffu;  // which scene entries the group is tested against (pmp_robot_create_ex)
};
struct ProgramD {
  std::vector<OpD> ops;
  std::vector<GroupD> groups;
  std::vector<double> spheres;  // [n][4]
  std::vector<int> pairs;       // [n][2]
};

inline void qmul(const double* a, const double* b, double* r) {  // Eigen product, (x, y, z, w)
  const double w = a[3] * b[3] - a[0] * b[0] - a[1] * b[1] - a[2] * b[2];
  const double x = a[3] * b[0] + a[0] * b[3] + a[1] * b[2] - a[2] * b[1];
  const double y = a[3] * b[1] + a[1] * b[3] + a[2] * b[0] - a[0] * b[2];
  const double z = a[3] * b[2] + a[2] * b[3] + a[0] * b[1] - a[1] * b[0];
  r[0] = x; r[1] = y; r[2] = z; r[3] = w;
}
inline void qrot(const double* q, const double* v, double* r) {  // Eigen q * v
  double u[3] = {2 * (q[1] * v[2] - q[2] * v[1]), 2 * (q[2] * v[0] - q[0] * v[2]), 2 * (q[0] * v[1] - q[1] * v[0])};
  const double x = v[0] + q[3] * u[0] + (q[1] * u[2] - q[2] * u[1]);
  const double y = v[1] + q[3] * u[1] + (q[2] * u[0] - q[0] * u[2]);
  const double z = v[2] + q[3] * u[2] + (q[0] * u[1] - q[1] * u[0]);
  r[0] = x; r[1] = y; r[2] = z;
}

// Fold every fixed op into what hangs off it.  For op i let anchor(i) be the nearest movable
// ancestor-or-self (or the base) and F_i the constant transform with frame_i = frame_anchor (x) F_i
// (kinematics/algorithm.cpp:181-210 composes exactly these products, in FP64, at run time; the
// identity-flag shortcut of transform.hpp:64-74 is kept by treating flagged origins as the exact
// identity).  Then
//   * a movable op keeps its joint and takes origin' = F_parent (x) origin,
//   * a sphere group on op i moves to anchor(i) with centre and members mapped through F_i,
// and the folded program has one op per controlled joint plus a root op that stands for the base.
// World-frame results are the same up to FP64 rounding of the constant products.
ProgramD fold_fixed_ops(const ProgramD& in) {
  const int n = (int)in.ops.size();
  struct Fix { double q[4], t[3]; int anchor; };  // anchor = folded op index
  std::vector<Fix> F(n);
  ProgramD out;
  OpD root{};
  root.oq[3] = 1.0;
  root.kind = PMP_OP_FIXED;
  root.q_index = -1;
  root.parent = -1;
  root.flags = PMP_OPFLAG_ORIGIN_IDENTITY;
  out.ops.push_back(root);
  for (int i = 0; i < n; i++) {
    const OpD& o = in.ops[i];
    Fix P;  // the parent's (anchor, F)
    if (o.parent < 0) {
      P.q[0] = P.q[1] = P.q[2] = 0.0; P.q[3] = 1.0;
      P.t[0] = P.t[1] = P.t[2] = 0.0;
      P.anchor = 0;
    } else {
      P = F[o.parent];
    }
    const bool ident = o.flags & PMP_OPFLAG_ORIGIN_IDENTITY;
    const double oq[4] = {ident ? 0.0 : o.oq[0], ident ? 0.0 : o.oq[1], ident ? 0.0 : o.oq[2], ident ? 1.0 : o.oq[3]};
    Fix L;  // P (x) (oq, oxyz)
    qmul(P.q, oq, L.q);
    double rt[3];
    qrot(P.q, o.oxyz, rt);
    for (int a = 0; a < 3; a++) L.t[a] = P.t[a] + rt[a];
    if (o.kind == PMP_OP_FIXED) {
      L.anchor = P.anchor;
      F[i] = L;
      continue;
    }
    OpD c = o;
    for (int a = 0; a < 4; a++) c.oq[a] = L.q[a];
    for (int a = 0; a < 3; a++) c.oxyz[a] = L.t[a];
    // a prismatic joint translates along its axis in the PARENT link's frame (the joint has no rotation
    // of its own): fold the parent's constant rotation into the axis used for the translation; the
    // Jacobian pass wants the axis in the op's own frame, which is the unrotated one
    if (o.kind == PMP_OP_PRISMATIC) qrot(P.q, o.axis, c.axis);
    c.parent = P.anchor;
    c.flags = (L.q[0] == 0.0 && L.q[1] == 0.0 && L.q[2] == 0.0 && L.q[3] == 1.0) ? PMP_OPFLAG_ORIGIN_IDENTITY : 0;
    out.ops.push_back(c);
    F[i].q[0] = F[i].q[1] = F[i].q[2] = 0.0; F[i].q[3] = 1.0;
    F[i].t[0] = F[i].t[1] = F[i].t[2] = 0.0;
    F[i].anchor = (int)out.ops.size() - 1;
  }
  out.spheres = in.spheres;
  out.pairs = in.pairs;
  for (const GroupD& g : in.groups) {
    GroupD c = g;
    const Fix& f = F[g.op];
    c.op = f.anchor;
    double r[3];
    qrot(f.q, g.center, r);
    for (int a = 0; a < 3; a++) c.center[a] = f.t[a] + r[a];
    for (int s = g.sphere_begin; s < g.sphere_begin + g.sphere_count; s++) {
      qrot(f.q, &in.spheres[4 * s], r);
      for (int a = 0; a < 3; a++) out.spheres[4 * s + a] = f.t[a] + r[a];
    }
    out.groups.push_back(c);
  }
  return out;
}

// v3: the tables feed the validity kernels (pmp_v3.cuh), whose stash has the v3 layout and which
// understand kSlotBase; otherwise the evaluate / FK kernels of pmp_kernels.cuh.
int build_tables(const ProgramD& P, const pmp_robot_desc_t* d, int q_dim, bool v3, RobotTables* out) {
  const int n_ops = (int)P.ops.size(), n_groups = (int)P.groups.size(), n_spheres = (int)P.spheres.size() / 4,
            n_pairs = (int)P.pairs.size() / 2;
  // groups sorted by op (stable), pairs remapped
  std::vector<int> order(n_groups);
  for (int g = 0; g < n_groups; g++) order[g] = g;
  std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return P.groups[a].op < P.groups[b].op; });
  std::vector<int> pos(n_groups);
  for (int k = 0; k < n_groups; k++) pos[order[k]] = k;

  // stash slots: a frame is kept if a later, non-adjacent op hangs off it or if it carries a group
  // (the cooperative narrow phases read the frames of the groups they work on)
  std::vector<int> slot(n_ops, -1);
  std::vector<char> need(n_ops, 0);
  for (int i = 0; i < n_ops; i++)
    if (P.ops[i].parent >= 0 && P.ops[i].parent != i - 1) need[P.ops[i].parent] = 1;
  for (int g = 0; g < n_groups; g++) need[P.groups[g].op] = 1;
  int n_slots = 0;
  for (int i = 0; i < n_ops; i++) {
    if (!need[i]) continue;
    const OpD& o = P.ops[i];
    // the root op of a FIXED-base robot carries the base pose itself, the same for every
    // configuration: the v3 kernels take it from the handle instead of stashing it per thread
    const bool is_base = v3 && d->base_type == PMP_BASE_FIXED && o.parent < 0 && o.kind == PMP_OP_FIXED &&
                         (o.flags & PMP_OPFLAG_ORIGIN_IDENTITY) && o.oxyz[0] == 0.0 && o.oxyz[1] == 0.0 && o.oxyz[2] == 0.0;
    slot[i] = is_base ? kSlotBase : n_slots++;
  }

  const int off_groups = n_ops * kOpWords;
  const int off_spheres = off_groups + n_groups * 2;
  const int off_pairs = off_spheres + n_spheres;
  const int off_pairs4 = off_pairs + (n_pairs + 1) / 2;
  const int off_gmask = off_pairs4 + n_pairs;
  bool has_gmask = false;
  for (const GroupD& G : P.groups) has_gmask |= G.prim_mask != 0xffffffffu;
  const int words = off_gmask + (has_gmask ? (n_groups + 3) / 4 : 0);
  std::vector<float4> blob(std::max(words, 1));
  auto fi = [](int v) {
    float f;
    memcpy(&f, &v, 4);
    return f;
  };
  for (int i = 0; i < n_ops; i++) {
    const OpD& o = P.ops[i];
    const bool ident = o.flags & PMP_OPFLAG_ORIGIN_IDENTITY;
    // identity-flagged origins are never multiplied in (kinematics/transform.hpp:64-74): store the exact identity
    const double oq[4] = {ident ? 0.0 : o.oq[0], ident ? 0.0 : o.oq[1], ident ? 0.0 : o.oq[2], ident ? 1.0 : o.oq[3]};
    blob[kOpWords * i + 0] = make_float4((float)oq[0], (float)oq[1], (float)oq[2], (float)oq[3]);
    blob[kOpWords * i + 1] = make_float4((float)o.oxyz[0], (float)o.oxyz[1], (float)o.oxyz[2], fi(o.kind));
    blob[kOpWords * i + 2] = make_float4((float)o.axis[0], (float)o.axis[1], (float)o.axis[2], fi(o.q_index));
    const int pslot = (o.parent >= 0 && o.parent != i - 1) ? slot[o.parent] : -1;
    int flags = o.flags & PMP_OPFLAG_ORIGIN_IDENTITY;
    if ((float)o.oxyz[0] == 0.f && (float)o.oxyz[1] == 0.f && (float)o.oxyz[2] == 0.f) flags |= PMP_OPFLAG_ZERO_XYZ;
    // an op whose frame is the constant base pose is never stashed (own slot -1); readers get kSlotBase
    blob[kOpWords * i + 3] = make_float4(fi(o.parent), fi(flags), fi(slot[i] == kSlotBase ? -1 : slot[i]), fi(pslot));
    // v3 sine basis: origin (x) (axis, 0) in FP64 (Eigen product, w = -o.a, vec = ow a + o x a)
    const double ax = o.axis[0], ay = o.axis[1], az = o.axis[2];
    blob[kOpWords * i + 4] = make_float4((float)(oq[3] * ax + oq[1] * az - oq[2] * ay), (float)(oq[3] * ay + oq[2] * ax - oq[0] * az),
                                         (float)(oq[3] * az + oq[0] * ay - oq[1] * ax), (float)(-oq[0] * ax - oq[1] * ay - oq[2] * az));
    // prismatic ops have no sine basis: word 4 carries the axis in the op's own frame (Jacobian pass)
    if (o.kind == PMP_OP_PRISMATIC)
      blob[kOpWords * i + 4] = make_float4((float)o.axis_child[0], (float)o.axis_child[1], (float)o.axis_child[2], 0.f);
  }
  for (int k = 0; k < n_groups; k++) {
    const GroupD& G = P.groups[order[k]];
    blob[off_groups + 2 * k] = make_float4((float)G.center[0], (float)G.center[1], (float)G.center[2], (float)G.radius);
    const int sl = slot[G.op] < 0 ? 0 : slot[G.op];  // kSlotBase = 0xffff fits the 16-bit field
    const int fl = (G.only_self ? 1 : 0) | (sl << 8);
    blob[off_groups + 2 * k + 1] = make_float4(fi(G.op), fi(G.sphere_begin), fi(G.sphere_count), fi(fl));
  }
  for (int s = 0; s < n_spheres; s++)
    blob[off_spheres + s] = make_float4((float)P.spheres[4 * s], (float)P.spheres[4 * s + 1], (float)P.spheres[4 * s + 2], (float)P.spheres[4 * s + 3]);
  {
    int* pw = reinterpret_cast<int*>(blob.data() + off_pairs);
    for (int p = 0; p < n_pairs; p++) {
      pw[2 * p] = pos[P.pairs[2 * p]];
      pw[2 * p + 1] = pos[P.pairs[2 * p + 1]];
      const float rs = (float)P.groups[P.pairs[2 * p]].radius + (float)P.groups[P.pairs[2 * p + 1]].radius;
      blob[off_pairs4 + p] = make_float4(fi(pw[2 * p]), fi(pw[2 * p + 1]), rs * rs, (rs + kGc16Pair) * (rs + kGc16Pair));
    }
  }
  if (has_gmask) {
    unsigned* gm = reinterpret_cast<unsigned*>(blob.data() + off_gmask);
    for (int k = 0; k < n_groups; k++) gm[k] = P.groups[order[k]].prim_mask;
  }
  cudaError_t e = cudaMalloc(&out->blob_dev, blob.size() * sizeof(float4));
  if (e == cudaSuccess) e = cudaMemcpy(out->blob_dev, blob.data(), blob.size() * sizeof(float4), cudaMemcpyHostToDevice);
  if (e != cudaSuccess) return fail(PMP_ERR_CUDA, std::string("robot table upload: ") + cudaGetErrorString(e));
  RobotDev& D = out->dev;
  D.blob = out->blob_dev;
  D.blob_words = words;
  D.n_ops = n_ops;
  D.n_groups = n_groups;
  D.n_spheres = n_spheres;
  D.n_pairs = n_pairs;
  D.off_groups = off_groups;
  D.off_spheres = off_spheres;
  D.off_pairs = off_pairs;
  D.off_pairs4 = off_pairs4;
  D.off_gmask = has_gmask ? off_gmask : -1;
  out->n_env_spheres = 0;
  for (const GroupD& G : P.groups)
    if (!G.only_self) out->n_env_spheres += G.sphere_count;
  D.dof = d->dof;
  D.q_dim = q_dim;
  D.base_type = d->base_type;
  D.n_slots = n_slots;
  memcpy(D.base_pose, d->base_pose, sizeof(D.base_pose));
  // per thread: frames [slot][7], group centres [group][3], + the per-warp scratch row and self queue
  // (the evaluate kernel keeps one word per group instead)
  out->stash_floats_per_thread_eval = (size_t)n_slots * 7 + (size_t)n_groups * 3 + (size_t)n_groups + 1;
  out->stash_floats_per_thread = out->stash_floats_per_thread_eval;
  if (v3) {
    // FP16 group centres (MODE 2) cost a few instructions per pair; they are used only where the
    // bytes they save buy another resident CTA per SM (estimated with a 4-primitive scene and the
    // register bound of these kernels: the FIXED-base validity kernels need 76-80 registers since the environment
    // narrow phase runs in place and the moving-base code is compiled out of them, i.e. up to 6 CTAs of 128 threads)
    const bool fixed = d->base_type == PMP_BASE_FIXED;
    auto ctas = [&](bool gc16) {
      const size_t stash = (size_t)n_slots * 7 + (size_t)v3_gc_floats(n_groups, gc16) + (size_t)kV3WordsPerThread;
      const size_t cta = (size_t)words * 16 + 4 * kPrimWords * 16 + stash * 4 * kBlock + 1024;
      return std::min<size_t>(6, (size_t)233472 / cta);
    };
    bool gc16 = fixed && ctas(true) > ctas(false);
    if (const char* e = getenv("PMP_V3_MODE")) gc16 = fixed && atoi(e) == 2;  // A/B experiments and tests only
    out->v3_mode = !fixed ? 0 : (gc16 ? 2 : 1);
    out->stash_floats_per_thread = (size_t)n_slots * 7 + (size_t)v3_gc_floats(n_groups, gc16) + (size_t)kV3WordsPerThread;
  }
  if (pick_block_per_thread(*out, empty_scene(), true) == 0 || smem_bytes(*out, empty_scene(), 64) > (size_t)kMaxSmem)
    return fail(PMP_ERR_LIMIT, "robot tables exceed the shared-memory budget of one CTA");
  return PMP_OK;
}
}  // namespace

extern "C" {

const char* pmp_last_error(void) { return g_err.c_str(); }
int pmp_abi_version(void) { return PMP_ABI_VERSION; }
uint64_t pmp_launch_count(void) { return g_launches.load(); }

int pmp_host_alloc(size_t bytes, int write_combined, void** out) {
  if (!out) return fail(PMP_ERR_INVALID, "out is null");
  *out = nullptr;
  if (bytes == 0) return PMP_OK;
  if (int rc = ensure_device()) return rc;
  CUDA_OK(cudaHostAlloc(out, bytes, cudaHostAllocPortable | (write_combined ? cudaHostAllocWriteCombined : 0)));
  return PMP_OK;
}

int pmp_host_free(void* p) {
  if (p) CUDA_OK(cudaFreeHost(p));
  return PMP_OK;
}
int pmp_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
  return n;
}

int pmp_set_device(int device) {
  CUDA_OK(cudaSetDevice(device));
  return ensure_device();  // creates this device's record (SM count, caches) on first use
}

// ---------------------------------------------------------------------------------------------
int pmp_robot_create(const pmp_robot_desc_t* d, pmp_robot_t** out) { return pmp_robot_create_ex(d, nullptr, out); }

int pmp_robot_create_ex(const pmp_robot_desc_t* d, const uint32_t* group_prim_mask, pmp_robot_t** out) {
  if (!d || !out) return fail(PMP_ERR_INVALID, "null argument");
  if (d->n_ops < 0 || d->n_groups < 0 || d->n_spheres < 0 || d->n_pairs < 0 || d->dof < 0)
    return fail(PMP_ERR_INVALID, "negative table size");
  if (d->n_groups > kMaxGroupsV3) return fail(PMP_ERR_LIMIT, "too many sphere groups (limit 2048)");
  if (d->base_type < PMP_BASE_FIXED || d->base_type > PMP_BASE_FLOATING) return fail(PMP_ERR_INVALID, "bad base_type");
  const int q_dim = d->dof + (d->base_type == PMP_BASE_PLANAR ? 3 : 0) + (d->base_type == PMP_BASE_FLOATING ? 6 : 0);
  for (int i = 0; i < d->n_ops; i++) {
    const pmp_link_op_t& o = d->ops[i];
    if (o.parent < -1 || o.parent >= i) return fail(PMP_ERR_INVALID, "ops must be in topological order (parent < index)");
    if (o.kind < PMP_OP_FIXED || o.kind > PMP_OP_PRISMATIC) return fail(PMP_ERR_INVALID, "bad op kind");
    if (o.kind != PMP_OP_FIXED && (o.q_index < 0 || o.q_index >= d->dof)) return fail(PMP_ERR_INVALID, "q_index out of range");
  }
  for (int p = 0; p < d->n_pairs; p++) {
    const int a = d->pairs[2 * p], b = d->pairs[2 * p + 1];
    if (a < 0 || a >= d->n_groups || b < 0 || b >= d->n_groups) return fail(PMP_ERR_INVALID, "pair refers to an unknown group");
  }
  for (int g = 0; g < d->n_groups; g++) {
    const pmp_group_t& G = d->groups[g];
    if (G.op < 0 || G.op >= d->n_ops) return fail(PMP_ERR_INVALID, "group refers to an unknown op");
    if (G.sphere_begin < 0 || G.sphere_count < 0 || G.sphere_begin + G.sphere_count > d->n_spheres)
      return fail(PMP_ERR_INVALID, "group sphere range out of bounds");
  }
  if (int rc = ensure_device()) return rc;

  ProgramD P;
  P.ops.resize(d->n_ops);
  for (int i = 0; i < d->n_ops; i++) {
    const pmp_link_op_t& o = d->ops[i];
    OpD& c = P.ops[i];
    for (int a = 0; a < 4; a++) c.oq[a] = o.origin_quat[a];
    for (int a = 0; a < 3; a++) c.oxyz[a] = o.origin_xyz[a];
    for (int a = 0; a < 3; a++) c.axis[a] = c.axis_child[a] = o.axis[a];
    c.kind = o.kind;
    c.q_index = o.q_index;
    c.parent = o.parent;
    c.flags = o.flags & PMP_OPFLAG_ORIGIN_IDENTITY;
  }
  P.groups.resize(d->n_groups);
  for (int g = 0; g < d->n_groups; g++) {
    const pmp_group_t& G = d->groups[g];
    GroupD& c = P.groups[g];
    for (int a = 0; a < 3; a++) c.center[a] = G.center[a];
    c.radius = G.radius;
    c.op = G.op;
    c.sphere_begin = G.sphere_begin;
    c.sphere_count = G.sphere_count;
    c.only_self = G.only_self_collision ? 1 : 0;
    if (group_prim_mask) c.prim_mask = group_prim_mask[g];
  }
  P.spheres.assign(d->spheres, d->spheres + (size_t)4 * d->n_spheres);
  P.pairs.assign(d->pairs, d->pairs + (size_t)2 * d->n_pairs);

  auto* r = new pmp_robot();
  int rc = build_tables(P, d, q_dim, false, &r->full);
  if (rc == PMP_OK) rc = build_tables(fold_fixed_ops(P), d, q_dim, true, &r->fast);
  if (rc != PMP_OK) {
    pmp_robot_destroy(r);
    return rc;
  }
  *out = r;
  return PMP_OK;
}

void pmp_robot_destroy(pmp_robot_t* r) {
  if (!r) return;
  if (r->full.blob_dev) cudaFree(r->full.blob_dev);
  if (r->fast.blob_dev) cudaFree(r->fast.blob_dev);
  delete r;
}
int pmp_robot_q_dim(const pmp_robot_t* r) { return r ? r->full.dev.q_dim : -1; }
int pmp_robot_n_spheres(const pmp_robot_t* r) { return r ? r->full.dev.n_spheres : -1; }

// ---------------------------------------------------------------------------------------------
// repack pmp_prim_t into the device record (pmp_device.cuh): type | flags | test variant in p0.w,
// AABB as centre / half extent
static void pack_prims(const pmp_prim_t* prims, int n, std::vector<float4>* out) {
  std::vector<float4>& rec = *out;
  rec.assign((size_t)n * kPrimWords, make_float4(0.f, 0.f, 0.f, 0.f));
  for (int i = 0; i < n; i++) {
    const pmp_prim_t& P = prims[i];
    int flags = P.flags;
    // a z-aligned cylinder is evaluated in world axes for all three coordinates (primitive_sdf.hpp:446-449,466-472)
    if (P.type == PMP_PRIM_CYLINDER && (flags & PMP_PRIMFLAG_Z_ALIGNED)) flags |= PMP_PRIMFLAG_AXIS_ALIGNED;
    int kind = P.type;
    if (P.type == PMP_PRIM_BOX && (flags & PMP_PRIMFLAG_AXIS_ALIGNED)) kind = PMP_KIND_BOX_AA;
    if (P.type == PMP_PRIM_CYLINDER && (flags & PMP_PRIMFLAG_AXIS_ALIGNED)) kind = PMP_KIND_CYL_AA;
    const int tf = P.type | (flags << 8) | (kind << 16);
    float tfw;
    memcpy(&tfw, &tf, 4);
    float4* w = rec.data() + (size_t)kPrimWords * i;
    w[0] = make_float4(P.position[0], P.position[1], P.position[2], tfw);
    w[1] = make_float4(P.param0, P.param1, P.param2, 0.f);
    w[2] = make_float4(P.rot_col0[0], P.rot_col0[1], P.rot_col0[2], 0.f);
    w[3] = make_float4(P.rot_col1[0], P.rot_col1[1], P.rot_col1[2], 0.f);
    w[4] = make_float4(P.rot_col2[0], P.rot_col2[1], P.rot_col2[2], 0.f);
    float ctr[3], half[3];
    for (int a = 0; a < 3; a++) {
      const double lo = P.lb[a], hi = P.ub[a];
      if (!(lo > -1e37) || !(hi < 1e37)) {  // unbounded axis (Ground x, y)
        ctr[a] = 0.f;
        half[a] = INFINITY;
      } else {
        ctr[a] = (float)(0.5 * (lo + hi));
        half[a] = (float)(0.5 * (hi - lo) + 1e-6 + 1e-6 * std::max(std::fabs(lo), std::fabs(hi)));
      }
    }
    w[5] = make_float4(ctr[0], ctr[1], ctr[2], 0.f);
    w[6] = make_float4(half[0], half[1], half[2], 0.f);
  }
}

static int cloud_build_status(const std::string& err) {
  if (err.find("non-finite") != std::string::npos) return PMP_ERR_INVALID;
  if (err.find("distinct cloud radii") != std::string::npos || err.find("too many cloud points") != std::string::npos) return PMP_ERR_LIMIT;
  return PMP_ERR_CUDA;
}

int pmp_scene_create(const pmp_scene_desc_t* d, pmp_scene_t** out) {
  if (!d || !out) return fail(PMP_ERR_INVALID, "null argument");
  if (d->n_prims < 0 || d->n_cloud_points < 0) return fail(PMP_ERR_INVALID, "negative table size");
  for (int i = 0; i < d->n_prims; i++)
    if (d->prims[i].type < PMP_PRIM_BOX || d->prims[i].type > PMP_PRIM_GROUND) return fail(PMP_ERR_INVALID, "bad primitive type");
  if (int rc = ensure_device()) return rc;
  auto* s = new pmp_scene();
  s->dev = empty_scene();
  static_assert(sizeof(pmp_prim_t) == 96, "pmp_prim_t layout");
  if (d->n_prims > 0) {
    std::vector<float4> rec;
    pack_prims(d->prims, d->n_prims, &rec);
    cudaError_t e = cudaMalloc(&s->prims_dev, rec.size() * sizeof(float4));
    if (e == cudaSuccess) e = cudaMemcpy(s->prims_dev, rec.data(), rec.size() * sizeof(float4), cudaMemcpyHostToDevice);
    if (e != cudaSuccess) {
      pmp_scene_destroy(s);
      return fail(PMP_ERR_CUDA, std::string("scene table upload: ") + cudaGetErrorString(e));
    }
    s->dev.prims = s->prims_dev;
    s->dev.n_prims = d->n_prims;
  }
  if (d->n_cloud_points > 0) {
    if (d->n_cloud_points >= (int64_t)1 << 31) {
      pmp_scene_destroy(s);
      return fail(PMP_ERR_LIMIT, "cloud too large");
    }
    if (!d->cloud_points) {
      pmp_scene_destroy(s);
      return fail(PMP_ERR_INVALID, "cloud_points is null");
    }
    std::string err;
    int launches = 0;
    if (cloud_build(d->cloud_points, d->n_cloud_points, d->cloud_radius, d->cloud_point_radii, d->cloud_cell, num_sms(), &s->dev.cloud,
                    &s->cloud[0], &s->cloud_scratch, &launches, &err)) {
      pmp_scene_destroy(s);
      return fail(cloud_build_status(err), "cloud build: " + err);
    }
    g_launches += launches;
  }
  *out = s;
  return PMP_OK;
}

void pmp_scene_destroy(pmp_scene_t* s) {
  if (!s) return;
  if (s->prims_dev) cudaFree(s->prims_dev);
  cloud_free(&s->cloud[0]);
  cloud_free(&s->cloud[1]);
  cloud_free(&s->cloud_scratch);
  delete s;
}

// Scene mutation without re-creating the handle: translate / rotate_z on the leaves (and a new depth
// frame for the cloud) are visible to the reference's constraint immediately because it holds the
// SDF objects by pointer (primitive_sdf.hpp:135-145,265-273; primitive_sphere_collision.cpp:426-448).
int pmp_scene_update_prims(pmp_scene_t* s, const pmp_prim_t* prims, int32_t n_prims, void* stream) {
  if (!s) return fail(PMP_ERR_INVALID, "scene handle is null");
  if (n_prims != s->dev.n_prims) return fail(PMP_ERR_INVALID, "primitive count differs from the scene's: create a new scene");
  if (n_prims == 0) return PMP_OK;
  if (!prims) return fail(PMP_ERR_INVALID, "prims is null");
  for (int i = 0; i < n_prims; i++)
    if (prims[i].type < PMP_PRIM_BOX || prims[i].type > PMP_PRIM_GROUND) return fail(PMP_ERR_INVALID, "bad primitive type");
  std::vector<float4> rec;
  pack_prims(prims, n_prims, &rec);
  // pageable source: the runtime stages the bytes before returning, so `rec` may die here
  CUDA_OK(cudaMemcpyAsync(s->prims_dev, rec.data(), rec.size() * sizeof(float4), cudaMemcpyHostToDevice, (cudaStream_t)stream));
  return PMP_OK;
}

int pmp_scene_update_cloud(pmp_scene_t* s, const float* points, int64_t n, float radius, float cell) {
  return pmp_scene_update_cloud_ex(s, points, n, radius, nullptr, cell);
}

int pmp_scene_update_cloud_ex(pmp_scene_t* s, const float* points, int64_t n, float radius, const float* point_radii, float cell) {
  if (!s) return fail(PMP_ERR_INVALID, "scene handle is null");
  if (n < 0 || n >= (int64_t)1 << 31) return fail(PMP_ERR_INVALID, "bad cloud size");
  if (n > 0 && !points) return fail(PMP_ERR_INVALID, "cloud_points is null");
  if (int rc = ensure_device()) return rc;
  CloudDev fresh{};
  const int idle = 1 - s->cloud_cur;
  if (n > 0) {
    std::string err;
    int launches = 0;
    if (cloud_build(points, n, radius, point_radii, cell, num_sms(), &fresh, &s->cloud[idle], &s->cloud_scratch, &launches, &err))
      return fail(cloud_build_status(err), "cloud build: " + err);
    g_launches += launches;
  } else {
    fresh.enabled = 0;
  }
  CUDA_OK(cudaDeviceSynchronize());  // nothing in flight may still walk the old tree when it becomes the idle set
  s->cloud_cur = idle;
  s->dev.cloud = fresh;
  return PMP_OK;
}

// ---------------------------------------------------------------------------------------------
#define CHECK_COMMON(robot, n)                                                       \
  if (!(robot)) return fail(PMP_ERR_INVALID, "robot handle is null");               \
  if ((n) < 0) return fail(PMP_ERR_INVALID, "negative item count");                 \
  if (int rc_ = ensure_device()) return rc_;

int pmp_check_configs(const pmp_robot_t* robot, const pmp_scene_t* scene, const float* q_dev, int64_t n,
                      uint32_t* valid_bits_dev, unsigned long long* n_valid_dev, void* stream) {
  CHECK_COMMON(robot, n);
  if (n == 0) return PMP_OK;
  if ((!q_dev && robot->fast.dev.q_dim > 0) || !valid_bits_dev) return fail(PMP_ERR_INVALID, "null device pointer");
  const SceneDev S = scene ? scene->dev : empty_scene();
  if (!mask_limit_ok(robot->fast, S)) return fail(PMP_ERR_LIMIT, "a robot with per-group scene masks needs a scene of at most 30 primitives");
  const int block = pick_block(robot->fast, S);
  const size_t smem = smem_bytes(robot->fast, S, block);
  if (smem > (size_t)kMaxSmem) return fail(PMP_ERR_LIMIT, "robot + scene tables exceed the shared-memory budget");
  const long long n_tiles = (n + block - 1) / block;
  int grid = 0;
  cudaStream_t st = (cudaStream_t)stream;
  // is_valid uses the polynomial sin/cos only for a FIXED base (constraints/primitive.hpp:39-60)
  const int mode = robot->fast.v3_mode;  // 0 moving base (exact sin/cos), 1 FIXED base, 2 FIXED base + FP16 group centres
#define PMP_LAUNCH_CHECK(KERNEL)                                                                  \
  do {                                                                                            \
    if (int rc = grid_for(KERNEL, smem, n_tiles, &grid, block)) return rc;                        \
    KERNEL<<<grid, block, smem, st>>>(robot->fast.dev, S, q_dev, n, valid_bits_dev, n_valid_dev);      \
  } while (0)
  {
      const bool cloud = S.cloud.enabled != 0;
#define PMP_V3(BB, AP, CL) PMP_LAUNCH_CHECK((k_check_configs_v3<BB, AP, CL>))
#define PMP_V3_B(BB)                                                                      \
  do {                                                                                    \
    if (mode == 2) { if (cloud) PMP_V3(BB, 2, true); else PMP_V3(BB, 2, false); }         \
    else if (mode == 1) { if (cloud) PMP_V3(BB, 1, true); else PMP_V3(BB, 1, false); }    \
    else { if (cloud) PMP_V3(BB, 0, true); else PMP_V3(BB, 0, false); }                   \
  } while (0)
      if (block == 128) PMP_V3_B(128); else PMP_V3_B(64);
#undef PMP_V3_B
#undef PMP_V3
  }
#undef PMP_LAUNCH_CHECK
  g_launches++;
  CUDA_OK(cudaGetLastError());
  return PMP_OK;
}

static inline size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

// evaluate(): one kernel, or — when the scene has a CloudSDF — three: pass 1 without the cloud (candidate
// spheres written to stream-ordered scratch), the warp-per-configuration cloud stage, and the Jacobian rows.
constexpr int64_t kEvalCoopChunk = 1 << 17;  // configurations per pass of the cooperative path (bounds the scratch)

static int eval_launch(const pmp_robot_t* robot, const pmp_scene_t* scene, const float* q_dev, int64_t n, float* dist_dev,
                       float* jac_dev, void* stream) {
  CHECK_COMMON(robot, n);
  if (n == 0) return PMP_OK;
  if ((!q_dev && robot->fast.dev.q_dim > 0) || !dist_dev) return fail(PMP_ERR_INVALID, "null device pointer");
  const SceneDev S = scene ? scene->dev : empty_scene();
  const int block = pick_block_per_thread(robot->fast, S, true);
  if (!block) return fail(PMP_ERR_LIMIT, "robot + scene tables exceed the shared-memory budget");
  const size_t smem = smem_bytes(robot->fast, S, block, true);
  cudaStream_t st = (cudaStream_t)stream;
  int grid = 0;
  const bool coop = S.cloud.enabled != 0 && robot->fast.n_env_spheres > 0;
  if (!coop) {
    CloudCand none{};
    if (jac_dev) {
      if (int rc = grid_for(k_eval_configs<true, false>, smem, (n + block - 1) / block, &grid, block)) return rc;
      k_eval_configs<true, false><<<grid, block, smem, st>>>(robot->fast.dev, S, q_dev, n, dist_dev, jac_dev, none);
    } else {
      if (int rc = grid_for(k_eval_configs<false, false>, smem, (n + block - 1) / block, &grid, block)) return rc;
      k_eval_configs<false, false><<<grid, block, smem, st>>>(robot->fast.dev, S, q_dev, n, dist_dev, nullptr, none);
    }
    g_launches++;
    CUDA_OK(cudaGetLastError());
    return PMP_OK;
  }
  const int qd = robot->fast.dev.q_dim;
  const int stride = robot->fast.n_env_spheres;
  for (int64_t off = 0; off < n; off += kEvalCoopChunk) {
    const int64_t m = std::min<int64_t>(kEvalCoopChunk, n - off);
    const size_t o_op = align256((size_t)m * stride * 16), o_cnt = o_op + (jac_dev ? align256((size_t)m * stride * 4) : 0);
    const size_t o_arg = o_cnt + align256((size_t)m * 4), total = o_arg + (jac_dev ? align256((size_t)m * sizeof(EvalArgmin)) : 0);
    char* scratch = nullptr;
    CUDA_OK(cudaMallocAsync(&scratch, total, st));
    CloudCand CC{};
    CC.sph = reinterpret_cast<float4*>(scratch);
    CC.op = jac_dev ? reinterpret_cast<int*>(scratch + o_op) : nullptr;
    CC.count = reinterpret_cast<int*>(scratch + o_cnt);
    CC.argmin = jac_dev ? reinterpret_cast<EvalArgmin*>(scratch + o_arg) : nullptr;
    CC.stride = stride;
    const float* qc = q_dev + off * qd;
    float* dc = dist_dev + off * 2;
    int rc = PMP_OK;
    const int warps_grid = (int)std::min<int64_t>((m + 3) / 4, (int64_t)num_sms() * 16);  // 4 warps (configurations) per CTA
    if (jac_dev) {
      float* jc = jac_dev + off * 2 * qd;
      rc = grid_for(k_eval_configs<true, true>, smem, (m + block - 1) / block, &grid, block);
      if (rc == PMP_OK) {
        k_eval_configs<true, true><<<grid, block, smem, st>>>(robot->fast.dev, S, qc, m, dc, jc, CC);
        k_eval_cloud_coop<true><<<warps_grid, 128, 0, st>>>(S, CC, m, dc);
        rc = grid_for(k_eval_jacobian, smem, (m + block - 1) / block, &grid, block);
      }
      if (rc == PMP_OK) k_eval_jacobian<<<grid, block, smem, st>>>(robot->fast.dev, S, qc, m, CC, jc);
      g_launches += 3;
    } else {
      rc = grid_for(k_eval_configs<false, true>, smem, (m + block - 1) / block, &grid, block);
      if (rc == PMP_OK) {
        k_eval_configs<false, true><<<grid, block, smem, st>>>(robot->fast.dev, S, qc, m, dc, nullptr, CC);
        k_eval_cloud_coop<false><<<warps_grid, 128, 0, st>>>(S, CC, m, dc);
      }
      g_launches += 2;
    }
    cudaError_t e = cudaGetLastError();
    cudaFreeAsync(scratch, st);
    if (rc != PMP_OK) return rc;
    if (e != cudaSuccess) return fail(PMP_ERR_CUDA, cudaGetErrorString(e));
  }
  return PMP_OK;
}

int pmp_eval_configs(const pmp_robot_t* robot, const pmp_scene_t* scene, const float* q_dev, int64_t n,
                     float* dist_dev, void* stream) {
  return eval_launch(robot, scene, q_dev, n, dist_dev, nullptr, stream);
}

int pmp_eval_grad_configs(const pmp_robot_t* robot, const pmp_scene_t* scene, const float* q_dev, int64_t n,
                          float* dist_dev, float* jac_dev, void* stream) {
  if (!jac_dev && n > 0 && robot && robot->fast.dev.q_dim > 0) return fail(PMP_ERR_INVALID, "null device pointer");
  if (!jac_dev) return eval_launch(robot, scene, q_dev, n, dist_dev, nullptr, stream);  // q_dim == 0: no columns
  return eval_launch(robot, scene, q_dev, n, dist_dev, jac_dev, stream);
}

int pmp_fk_spheres(const pmp_robot_t* robot, const float* q_dev, int64_t n, int accurate, float* spheres_dev, void* stream) {
  CHECK_COMMON(robot, n);
  if (n == 0) return PMP_OK;
  if ((!q_dev && robot->full.dev.q_dim > 0) || !spheres_dev) return fail(PMP_ERR_INVALID, "null device pointer");
  const SceneDev S = empty_scene();
  const int block = pick_block_per_thread(robot->full, S, false);
  if (!block) return fail(PMP_ERR_LIMIT, "robot tables exceed the shared-memory budget");
  const size_t smem = smem_bytes(robot->full, S, block);
  int grid = 0;
  // a moving base always uses exact sin/cos (constraints/primitive.hpp:42)
  const bool approx = !accurate && robot->full.dev.base_type == PMP_BASE_FIXED;
  if (approx) {
    if (int rc = grid_for(k_fk_spheres<true>, smem, (n + block - 1) / block, &grid, block)) return rc;
    k_fk_spheres<true><<<grid, block, smem, (cudaStream_t)stream>>>(robot->full.dev, S, q_dev, n, (float4*)spheres_dev);
  } else {
    if (int rc = grid_for(k_fk_spheres<false>, smem, (n + block - 1) / block, &grid, block)) return rc;
    k_fk_spheres<false><<<grid, block, smem, (cudaStream_t)stream>>>(robot->full.dev, S, q_dev, n, (float4*)spheres_dev);
  }
  g_launches++;
  CUDA_OK(cudaGetLastError());
  return PMP_OK;
}

int pmp_fk_frames(const pmp_robot_t* robot, const float* q_dev, int64_t n, int accurate, float* frames_dev, void* stream) {
  CHECK_COMMON(robot, n);
  if (n == 0) return PMP_OK;
  if ((!q_dev && robot->full.dev.q_dim > 0) || !frames_dev) return fail(PMP_ERR_INVALID, "null device pointer");
  const SceneDev S = empty_scene();
  const int block = pick_block_per_thread(robot->full, S, false);
  if (!block) return fail(PMP_ERR_LIMIT, "robot tables exceed the shared-memory budget");
  const size_t smem = smem_bytes(robot->full, S, block);
  int grid = 0;
  const bool approx = !accurate && robot->full.dev.base_type == PMP_BASE_FIXED;
  if (approx) {
    if (int rc = grid_for(k_fk_frames<true>, smem, (n + block - 1) / block, &grid, block)) return rc;
    k_fk_frames<true><<<grid, block, smem, (cudaStream_t)stream>>>(robot->full.dev, S, q_dev, n, frames_dev);
  } else {
    if (int rc = grid_for(k_fk_frames<false>, smem, (n + block - 1) / block, &grid, block)) return rc;
    k_fk_frames<false><<<grid, block, smem, (cudaStream_t)stream>>>(robot->full.dev, S, q_dev, n, frames_dev);
  }
  g_launches++;
  CUDA_OK(cudaGetLastError());
  return PMP_OK;
}

// ---- edges ---------------------------------------------------------------------------------------
extern "C++" {
namespace {
// ompl/generate_sequence.py:11-63: greedy farthest-first permutation, ties -> larger number
std::vector<int> farthest_sequence(int N) {
  std::vector<int> result(N, 0);
  std::vector<char> used(N, 0);
  used[0] = 1;
  for (int i = 1; i < N; i++) {
    int best_d = -1, best = -1;
    for (int num = 0; num < N; num++) {
      if (used[num]) continue;
      int md = 1 << 30;
      for (int j = 0; j < i; j++) md = std::min(md, std::abs(num - result[j]));
      if (md > best_d || (md == best_d && num > best)) {
        best_d = md;
        best = num;
      }
    }
    result[i] = best;
    used[best] = 1;
  }
  return result;
}
const std::vector<uint8_t>& sequence_table_host() {  // SEQUENCE_TABLE rows zero-padded to 128 (generate_sequence.py:76-98)
  static const std::vector<uint8_t> t = [] {
    std::vector<uint8_t> v(128 * 128, 0);
    for (int n = 1; n <= 128; n++) {
      auto row = farthest_sequence(n);
      for (int i = 0; i < n; i++) v[(n - 1) * 128 + i] = (uint8_t)row[i];
    }
    return v;
  }();
  return t;
}
int ensure_sequence_table(DeviceState* D) {
  std::lock_guard<std::mutex> lk(D->seq_mu);
  if (D->seq_dev) return PMP_OK;
  const auto& t = sequence_table_host();
  CUDA_OK(cudaMalloc(&D->seq_dev, t.size()));
  CUDA_OK(cudaMemcpy(D->seq_dev, t.data(), t.size(), cudaMemcpyHostToDevice));
  return PMP_OK;
}

constexpr size_t kMaxValidatorCache = 64;  // distinct box-width vectors kept per device (least recently used goes first)

int validator_to_dev(const pmp_validator_t* v, int q_dim, ValidatorDev* out) {
  if (!v) return fail(PMP_ERR_INVALID, "validator is null");
  if (v->dim != q_dim) return fail(PMP_ERR_INVALID, "box dimension and space dimension mismatch");  // ompl_thin_wrap.hpp:84-86
  DeviceState* D;
  if (int rc = device_state(&D)) return rc;
  if (int rc = ensure_sequence_table(D)) return rc;
  out->seq = D->seq_dev;
  out->type = v->type;
  out->resolution = v->resolution;
  out->width = out->inv_width = nullptr;
  if (v->type == PMP_VALIDATOR_BOX) {
    if (!v->box_width) return fail(PMP_ERR_INVALID, "box_width is null");
    std::lock_guard<std::mutex> lk(D->val_mu);
    std::vector<double> w(v->box_width, v->box_width + q_dim);
    for (auto& c : D->val_cache)
      if (c.width == w) {
        c.last_use = ++D->val_clock;
        out->width = c.dev;
        out->inv_width = c.dev + q_dim;
        return PMP_OK;
      }
    std::vector<double> both(2 * q_dim);
    for (int i = 0; i < q_dim; i++) {
      if (!(w[i] > 0)) return fail(PMP_ERR_INVALID, "box width must be positive");
      both[i] = w[i];
      both[q_dim + i] = 1.0 / w[i];
    }
    ValidatorCache c;
    c.width = w;
    if (D->val_cache.size() >= kMaxValidatorCache) {
      // evict the least recently used entry; kernels already queued may still read it, so its memory is
      // released in stream order behind everything submitted so far (legacy stream = device-wide order)
      size_t lru = 0;
      for (size_t i = 1; i < D->val_cache.size(); i++)
        if (D->val_cache[i].last_use < D->val_cache[lru].last_use) lru = i;
      CUDA_OK(cudaFreeAsync(D->val_cache[lru].dev, (cudaStream_t)0));
      D->val_cache.erase(D->val_cache.begin() + (long)lru);
    }
    CUDA_OK(cudaMalloc(&c.dev, both.size() * 8));
    CUDA_OK(cudaMemcpy(c.dev, both.data(), both.size() * 8, cudaMemcpyHostToDevice));
    c.last_use = ++D->val_clock;
    D->val_cache.push_back(c);
    out->width = c.dev;
    out->inv_width = c.dev + q_dim;
  } else if (v->type == PMP_VALIDATOR_EUCLIDEAN) {
    if (!(v->resolution > 0)) return fail(PMP_ERR_INVALID, "resolution must be positive");
  } else {
    return fail(PMP_ERR_INVALID, "unknown validator type");  // ompl_thin_wrap.hpp:89-91
  }
  return PMP_OK;
}

// the persistent edge kernel on endpoints, ratios and n_test that are already on the device
// Launch hint of the chunked host path (experiments): edges per lane the grid of the NEXT edge launch of this
// thread is sized for (0 = fill the GPU, the default).
thread_local int tl_edges_per_lane = 0;

int launch_edges(const pmp_robot_t* robot, const SceneDev& S, const uint8_t* seq, const EdgeArgs& A, cudaStream_t st) {
  if (!mask_limit_ok(robot->fast, S)) return fail(PMP_ERR_LIMIT, "a robot with per-group scene masks needs a scene of at most 30 primitives");
  const int block = pick_block(robot->fast, S);
  const size_t smem = smem_bytes(robot->fast, S, block);
  if (smem > (size_t)kMaxSmem) return fail(PMP_ERR_LIMIT, "robot + scene tables exceed the shared-memory budget");
  const long long n_words = (A.n + 31) / 32;
  long long tiles = (n_words + (block / 32) - 1) / (block / 32);  // one warp per 32 edges at most
  if (tl_edges_per_lane > 0) tiles = std::max<long long>(1, (A.n + (long long)block * tl_edges_per_lane - 1) / ((long long)block * tl_edges_per_lane));
  int grid = 0;
  const int mode = robot->fast.v3_mode;  // 0 moving base (exact sin/cos), 1 FIXED base, 2 FIXED base + FP16 group centres
  const bool cloud = S.cloud.enabled != 0;
#define PMP_E3(BB, AP, CL)                                                                                   \
  do {                                                                                                       \
    if (int rc = grid_for(k_check_edges_v3<BB, AP, CL>, smem, tiles, &grid, block)) return rc;              \
    k_check_edges_v3<BB, AP, CL><<<grid, block, smem, st>>>(robot->fast.dev, S, seq, A);                    \
  } while (0)
#define PMP_E3_B(BB)                                                                      \
  do {                                                                                    \
    if (mode == 2) { if (cloud) PMP_E3(BB, 2, true); else PMP_E3(BB, 2, false); }         \
    else if (mode == 1) { if (cloud) PMP_E3(BB, 1, true); else PMP_E3(BB, 1, false); }    \
    else { if (cloud) PMP_E3(BB, 0, true); else PMP_E3(BB, 0, false); }                   \
  } while (0)
  if (block == 128) PMP_E3_B(128); else PMP_E3_B(64);
#undef PMP_E3_B
#undef PMP_E3
  g_launches++;
  CUDA_OK(cudaGetLastError());
  return PMP_OK;
}

constexpr int64_t kMaxEdgesPerLaunch = (int64_t)1 << 30;  // the kernel indexes edges with 32 bits
constexpr int64_t kSortEdgesFrom = 4096;                   // smaller batches are not worth the radix pass

// prepare pass + persistent kernel for one launch-sized batch.  Scratch (ratio, n_test, chunk counter and,
// for FP64 endpoints, their FP32 copies) comes from the stream-ordered pool and goes back to it in stream
// order, so concurrent callers on different streams never share it and nothing is allocated in steady state.
template <class T>
int check_edges_batch(const pmp_robot_t* robot, const SceneDev& S, const ValidatorDev& V, const T* q0, const T* q1, int64_t n,
                      uint32_t* bits, uint32_t* calls, unsigned long long* n_valid, unsigned long long* n_checked, cudaStream_t st) {
  constexpr bool f64 = sizeof(T) == 8;
  const int qd = robot->fast.dev.q_dim;
  static const int64_t sort_from = getenv("PMP_EDGE_SORT_FROM") ? atoll(getenv("PMP_EDGE_SORT_FROM")) : kSortEdgesFrom;  // (A/B runs)
  const bool sorted = n >= sort_from;
  size_t sort_tmp = 0;
  if (sorted)
    CUDA_OK(cub::DeviceRadixSort::SortPairs(nullptr, sort_tmp, (const uint8_t*)nullptr, (uint8_t*)nullptr, (const uint32_t*)nullptr,
                                            (uint32_t*)nullptr, (int)n, 0, 8, st));
  const size_t o_ntest = align256((size_t)n * 8), o_ctr = o_ntest + align256((size_t)n * 4), o_q0 = o_ctr + 256;
  const size_t qbytes = f64 ? align256((size_t)n * std::max(qd, 1) * 4) : 0;
  // longest-first order: keys in / out (u8), indices in / out (u32), CUB's temporary storage
  const size_t o_key = o_q0 + 2 * qbytes, o_idx = o_key + (sorted ? 2 * align256((size_t)n) : 0);
  const size_t o_tmp = o_idx + (sorted ? 2 * align256((size_t)n * 4) : 0), total = o_tmp + (sorted ? align256(sort_tmp) : 0);
  char* scratch = nullptr;
  CUDA_OK(cudaMallocAsync(&scratch, total, st));
  EdgeArgs A{};
  A.ratio = reinterpret_cast<double*>(scratch);
  A.ntest = reinterpret_cast<int*>(scratch + o_ntest);
  A.next_chunk = reinterpret_cast<unsigned long long*>(scratch + o_ctr);
  float* q0f = f64 ? reinterpret_cast<float*>(scratch + o_q0) : nullptr;
  float* q1f = f64 ? reinterpret_cast<float*>(scratch + o_q0 + qbytes) : nullptr;
  uint8_t* key = sorted ? reinterpret_cast<uint8_t*>(scratch + o_key) : nullptr;
  uint32_t* idx = sorted ? reinterpret_cast<uint32_t*>(scratch + o_idx) : nullptr;
  A.q0 = f64 ? q0f : reinterpret_cast<const float*>(q0);
  A.q1 = f64 ? q1f : reinterpret_cast<const float*>(q1);
  A.n = n;
  A.bits = bits;
  A.calls = calls;
  A.n_valid = n_valid;
  A.n_checked = n_checked;
  const int grid = (int)std::min<int64_t>((n * (f64 ? std::max(qd, 1) : 1) + 255) / 256, (int64_t)num_sms() * 8);
  k_edge_prepare<T><<<grid, 256, 0, st>>>(V, q0, q1, n, qd, q0f, q1f, const_cast<double*>(A.ratio), const_cast<int*>(A.ntest), bits,
                                          A.next_chunk, key, idx);
  g_launches++;
  int rc = PMP_OK;
  if (sorted) {
    uint32_t* order = idx + align256((size_t)n * 4) / 4;
    cudaError_t e = cub::DeviceRadixSort::SortPairs(scratch + o_tmp, sort_tmp, (const uint8_t*)key, key + align256((size_t)n),
                                                    (const uint32_t*)idx, order, (int)n, 0, 8, st);
    if (e != cudaSuccess) rc = fail(PMP_ERR_CUDA, std::string("edge order sort: ") + cudaGetErrorString(e));
    A.order = order;
    g_launches++;
  }
  if (rc == PMP_OK) rc = launch_edges(robot, S, V.seq, A, st);
  cudaError_t e = cudaFreeAsync(scratch, st);
  if (rc == PMP_OK && e != cudaSuccess) return fail(PMP_ERR_CUDA, std::string("cudaFreeAsync: ") + cudaGetErrorString(e));
  return rc;
}

template <class T>
int check_edges_any(const pmp_robot_t* robot, const pmp_scene_t* scene, const pmp_validator_t* validator, const T* q0_dev,
                    const T* q1_dev, int64_t n, uint32_t* valid_bits_dev, uint32_t* calls_dev, unsigned long long* n_valid_dev,
                    unsigned long long* n_checked_dev, void* stream) {
  CHECK_COMMON(robot, n);
  ValidatorDev V;
  if (int rc = validator_to_dev(validator, robot->fast.dev.q_dim, &V)) return rc;
  if (n == 0) return PMP_OK;
  if (!q0_dev || !q1_dev || !valid_bits_dev) return fail(PMP_ERR_INVALID, "null device pointer");
  const SceneDev S = scene ? scene->dev : empty_scene();
  const int qd = robot->fast.dev.q_dim;
  for (int64_t off = 0; off < n; off += kMaxEdgesPerLaunch) {  // kMaxEdgesPerLaunch is a multiple of 32: words do not straddle
    const int64_t m = std::min<int64_t>(kMaxEdgesPerLaunch, n - off);
    if (int rc = check_edges_batch<T>(robot, S, V, q0_dev + off * qd, q1_dev + off * qd, m, valid_bits_dev + off / 32,
                                      calls_dev ? calls_dev + off : nullptr, n_valid_dev, n_checked_dev, (cudaStream_t)stream))
      return rc;
  }
  return PMP_OK;
}
}  // namespace
}  // extern "C++"

int pmp_check_edges(const pmp_robot_t* robot, const pmp_scene_t* scene, const pmp_validator_t* validator,
                    const float* q0_dev, const float* q1_dev, int64_t n, uint32_t* valid_bits_dev,
                    unsigned long long* n_valid_dev, unsigned long long* n_checked_dev, void* stream) {
  return check_edges_any<float>(robot, scene, validator, q0_dev, q1_dev, n, valid_bits_dev, nullptr, n_valid_dev, n_checked_dev, stream);
}

int pmp_check_edges_ex(const pmp_robot_t* robot, const pmp_scene_t* scene, const pmp_validator_t* validator,
                       const void* q0_dev, const void* q1_dev, int endpoints_f64, int64_t n, uint32_t* valid_bits_dev,
                       uint32_t* calls_per_edge_dev, unsigned long long* n_valid_dev, unsigned long long* n_checked_dev,
                       void* stream) {
  if (endpoints_f64)
    return check_edges_any<double>(robot, scene, validator, (const double*)q0_dev, (const double*)q1_dev, n, valid_bits_dev,
                                   calls_per_edge_dev, n_valid_dev, n_checked_dev, stream);
  return check_edges_any<float>(robot, scene, validator, (const float*)q0_dev, (const float*)q1_dev, n, valid_bits_dev,
                                calls_per_edge_dev, n_valid_dev, n_checked_dev, stream);
}

// ---- host-buffer entry points ----------------------------------------------------------------------
extern "C++" {
namespace {
// Device scratch + streams of ONE *_host call in flight.  A call checks a workspace out of its device's
// free list (creating one when none is idle) and hands it back when it returns, so concurrent callers
// never wait on each other; buffers are grow-only and reused by later calls.
constexpr int kWsStreams = 3;
struct HostWorkspace {
  cudaStream_t stream[kWsStreams] = {};
  void* in[kWsStreams][2] = {};
  size_t in_cap[kWsStreams][2] = {};
  void* outb[kWsStreams] = {};
  size_t out_cap[kWsStreams] = {};
  unsigned long long* counters = nullptr;       // device [2]: n_valid, n_checked
  unsigned long long* counters_host = nullptr;  // pinned [2]
  // edge batches (check_edges_host_any): ONE copy stream that never waits for a kernel, and a ring of
  // compute slots (stream + input buffers + "inputs arrived" / "slot free again" events)
  static constexpr int kEdgeSlots = 6;
  cudaStream_t ecopy = nullptr;
  cudaStream_t ecomp[kEdgeSlots] = {};
  cudaEvent_t e_in[kEdgeSlots] = {}, e_done[kEdgeSlots] = {};
  void* ein[kEdgeSlots][2] = {};
  size_t ein_cap[kEdgeSlots][2] = {};
  void* eout[kEdgeSlots] = {};
  size_t eout_cap[kEdgeSlots] = {};
  int init_edges() {
    if (ecopy) return PMP_OK;
    CUDA_OK(cudaStreamCreateWithFlags(&ecopy, cudaStreamNonBlocking));
    for (int i = 0; i < kEdgeSlots; i++) {
      CUDA_OK(cudaStreamCreateWithFlags(&ecomp[i], cudaStreamNonBlocking));
      CUDA_OK(cudaEventCreateWithFlags(&e_in[i], cudaEventDisableTiming));
      CUDA_OK(cudaEventCreateWithFlags(&e_done[i], cudaEventDisableTiming));
    }
    return PMP_OK;
  }
  int init() {
    for (int i = 0; i < kWsStreams; i++) CUDA_OK(cudaStreamCreateWithFlags(&stream[i], cudaStreamNonBlocking));
    CUDA_OK(cudaMalloc(&counters, 2 * sizeof(unsigned long long)));
    CUDA_OK(cudaMallocHost(&counters_host, 2 * sizeof(unsigned long long)));
    return PMP_OK;
  }
  int reserve(void** p, size_t* cap, size_t bytes) {
    if (*cap >= bytes) return PMP_OK;
    if (*p) cudaFree(*p);
    *p = nullptr;
    *cap = 0;
    CUDA_OK(cudaMalloc(p, bytes));
    *cap = bytes;
    return PMP_OK;
  }
  int sync_all() {
    for (int i = 0; i < kWsStreams; i++) CUDA_OK(cudaStreamSynchronize(stream[i]));
    if (ecopy) {
      CUDA_OK(cudaStreamSynchronize(ecopy));
      for (int i = 0; i < kEdgeSlots; i++) CUDA_OK(cudaStreamSynchronize(ecomp[i]));
    }
    return PMP_OK;
  }
};
struct WorkspaceLease {  // RAII check-out of a workspace of the current device
  DeviceState* D = nullptr;
  HostWorkspace* ws = nullptr;
  int acquire() {
    if (int rc = device_state(&D)) return rc;
    {
      std::lock_guard<std::mutex> lk(D->ws_mu);
      if (!D->ws_free.empty()) {
        ws = D->ws_free.back();
        D->ws_free.pop_back();
        return PMP_OK;
      }
    }
    ws = new HostWorkspace();
    return ws->init();
  }
  ~WorkspaceLease() {
    if (!ws) return;
    ws->sync_all();  // nothing of this call may still be in flight when another caller takes the buffers
    std::lock_guard<std::mutex> lk(D->ws_mu);
    D->ws_free.push_back(ws);
  }
};
constexpr int64_t kHostChunk = 1 << 19;  // items per pipelined chunk (multiple of 32): the tail that cannot overlap is one chunk's kernel
// Edge batches: the first chunks are small so that the first kernel starts after a short copy, later ones
// larger so that per-chunk costs (prepare pass, sort, launch) stay small — but not so large that the last
// chunk, whose inputs arrive last, holds much of the work.
inline int64_t edge_chunk_size(int k) {
  // log2 of the first / largest chunk; the environment overrides are for A/B runs (profiles/r02_s10_edges_e2e_pipeline.txt)
  static const int first = getenv("PMP_EDGE_CHUNK_FIRST") ? atoi(getenv("PMP_EDGE_CHUNK_FIRST")) : 17;
  static const int last = getenv("PMP_EDGE_CHUNK_LAST") ? atoi(getenv("PMP_EDGE_CHUNK_LAST")) : 19;
  return (int64_t)1 << std::min(last, first + k);
}

template <class T>
int check_edges_host_any(const pmp_robot_t* robot, const pmp_scene_t* scene, const pmp_validator_t* validator, const T* q0_host,
                         const T* q1_host, int64_t n, uint32_t* valid_bits_host, uint64_t* n_valid, uint64_t* n_checked) {
  CHECK_COMMON(robot, n);
  if (n_valid) *n_valid = 0;
  if (n_checked) *n_checked = 0;
  if (n == 0) {
    ValidatorDev V;
    return validator_to_dev(validator, robot->full.dev.q_dim, &V);
  }
  if (!q0_host || !q1_host || !valid_bits_host) return fail(PMP_ERR_INVALID, "null host pointer");
  WorkspaceLease L;
  if (int rc = L.acquire()) return rc;
  HostWorkspace& W = *L.ws;
  const int qd = robot->full.dev.q_dim;
  CUDA_OK(cudaMemsetAsync(W.counters, 0, 2 * sizeof(unsigned long long), W.stream[0]));
  CUDA_OK(cudaStreamSynchronize(W.stream[0]));
  if (int rc = W.init_edges()) return rc;
  static const bool trace = getenv("PMP_TRACE_EDGES") != nullptr;
  std::vector<cudaEvent_t> marks;
  auto trace_mark = [&](cudaStream_t s) {
    cudaEvent_t e;
    cudaEventCreate(&e);
    cudaEventRecord(e, s);
    marks.push_back(e);
  };
  const auto t_host0 = std::chrono::steady_clock::now();
  // The pipeline (device timelines: PMP_TRACE_EDGES=1).  Copies run back to back on their own stream at the
  // link rate and never wait for a kernel (only for a ring slot, kEdgeSlots chunks back); every chunk's
  // prepare pass / sort / persistent kernel / result copy run on the slot's stream behind the chunk's
  // "inputs arrived" event; as a persistent CTA retires when its warps run dry, the tail of one chunk's kernel
  // runs under the head of the next.  (Grids sized for a few edges per lane — PMP_EDGE_EPL, so that neighbouring
  // chunks share the SMs from the start — were measured and lost: profiles/r02_s10_edges_e2e_pipeline.txt.)
  static const int epl = getenv("PMP_EDGE_EPL") ? atoi(getenv("PMP_EDGE_EPL")) : 0;
  constexpr int kSlots = HostWorkspace::kEdgeSlots;
  const int64_t max_chunk = std::min<int64_t>(n, 3 * edge_chunk_size(64) / 2);
  for (int b = 0; b < kSlots; b++) {  // all growth happens before anything is in flight (cudaMalloc synchronises)
    for (int a = 0; a < 2; a++)
      if (int rc = W.reserve(&W.ein[b][a], &W.ein_cap[b][a], (size_t)max_chunk * qd * sizeof(T))) return rc;
    if (int rc = W.reserve(&W.eout[b], &W.eout_cap[b], (size_t)((max_chunk + 31) / 32) * 4)) return rc;
  }
  int64_t off = 0;
  for (int k = 0; off < n; k++) {
    int64_t m = std::min<int64_t>(edge_chunk_size(k), n - off);
    if (n - off - m < m / 2) m = n - off;  // a short remainder rides with this chunk: a small last chunk would be all tail
    const int b = k % kSlots;
    cudaStream_t st = W.ecomp[b];
    if (k >= kSlots) CUDA_OK(cudaStreamWaitEvent(W.ecopy, W.e_done[b], 0));  // the slot's previous chunk has been consumed
    if (trace) trace_mark(W.ecopy);
    CUDA_OK(cudaMemcpyAsync(W.ein[b][0], q0_host + off * qd, (size_t)m * qd * sizeof(T), cudaMemcpyHostToDevice, W.ecopy));
    CUDA_OK(cudaMemcpyAsync(W.ein[b][1], q1_host + off * qd, (size_t)m * qd * sizeof(T), cudaMemcpyHostToDevice, W.ecopy));
    CUDA_OK(cudaEventRecord(W.e_in[b], W.ecopy));
    if (trace) trace_mark(W.ecopy);
    CUDA_OK(cudaStreamWaitEvent(st, W.e_in[b], 0));
    tl_edges_per_lane = epl;
    const int rc = check_edges_any<T>(robot, scene, validator, (const T*)W.ein[b][0], (const T*)W.ein[b][1], m, (uint32_t*)W.eout[b], nullptr,
                                      W.counters, W.counters + 1, st);
    tl_edges_per_lane = 0;
    if (rc) return rc;
    if (trace) trace_mark(st);
    CUDA_OK(cudaMemcpyAsync(valid_bits_host + off / 32, W.eout[b], (size_t)((m + 31) / 32) * 4, cudaMemcpyDeviceToHost, st));
    CUDA_OK(cudaEventRecord(W.e_done[b], st));
    if (trace) trace_mark(st);
    off += m;
  }
  if (int rc = W.sync_all()) return rc;
  if (trace) {  // PMP_TRACE_EDGES=1 (diagnostics): device timeline of the chunks, ms from the first mark
    fprintf(stderr, "[pmp trace] check_edges_host n=%lld host enqueue %.3f ms\n", (long long)n,
            std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_host0).count());
    for (size_t i = 0; i + 3 < marks.size(); i += 4) {
      float a = 0, b2 = 0, c = 0, d = 0;
      cudaEventElapsedTime(&a, marks[0], marks[i]);
      cudaEventElapsedTime(&b2, marks[0], marks[i + 1]);
      cudaEventElapsedTime(&c, marks[0], marks[i + 2]);
      cudaEventElapsedTime(&d, marks[0], marks[i + 3]);
      fprintf(stderr, "[pmp trace]   chunk %zu: h2d %.3f -> %.3f, kernels -> %.3f, d2h -> %.3f\n", i / 4, a, b2, c, d);
    }
    for (cudaEvent_t e : marks) cudaEventDestroy(e);
  }
  CUDA_OK(cudaMemcpyAsync(W.counters_host, W.counters, 2 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, W.stream[0]));
  CUDA_OK(cudaStreamSynchronize(W.stream[0]));
  if (n_valid) *n_valid = W.counters_host[0];
  if (n_checked) *n_checked = W.counters_host[1];
  return PMP_OK;
}
}  // namespace
}  // extern "C++"

int pmp_check_configs_host(const pmp_robot_t* robot, const pmp_scene_t* scene, const float* q_host, int64_t n,
                           uint32_t* valid_bits_host, uint64_t* n_valid) {
  CHECK_COMMON(robot, n);
  if (n_valid) *n_valid = 0;
  if (n == 0) return PMP_OK;
  if (!q_host || !valid_bits_host) return fail(PMP_ERR_INVALID, "null host pointer");
  WorkspaceLease L;
  if (int rc = L.acquire()) return rc;
  HostWorkspace& W = *L.ws;
  const int qd = robot->full.dev.q_dim;
  CUDA_OK(cudaMemsetAsync(W.counters, 0, 2 * sizeof(unsigned long long), W.stream[0]));
  CUDA_OK(cudaStreamSynchronize(W.stream[0]));
  int k = 0;
  for (int64_t off = 0; off < n; off += kHostChunk, k = (k + 1) % kWsStreams) {
    const int64_t m = std::min<int64_t>(kHostChunk, n - off);
    if (int rc = W.reserve(&W.in[k][0], &W.in_cap[k][0], (size_t)m * qd * 4)) return rc;
    if (int rc = W.reserve(&W.outb[k], &W.out_cap[k], (size_t)((m + 31) / 32) * 4)) return rc;
    cudaStream_t st = W.stream[k];
    CUDA_OK(cudaMemcpyAsync(W.in[k][0], q_host + off * qd, (size_t)m * qd * 4, cudaMemcpyHostToDevice, st));
    if (int rc = pmp_check_configs(robot, scene, (const float*)W.in[k][0], m, (uint32_t*)W.outb[k], W.counters, st)) return rc;
    CUDA_OK(cudaMemcpyAsync(valid_bits_host + off / 32, W.outb[k], (size_t)((m + 31) / 32) * 4, cudaMemcpyDeviceToHost, st));
  }
  if (int rc = W.sync_all()) return rc;
  CUDA_OK(cudaMemcpyAsync(W.counters_host, W.counters, sizeof(unsigned long long), cudaMemcpyDeviceToHost, W.stream[0]));
  CUDA_OK(cudaStreamSynchronize(W.stream[0]));
  if (n_valid) *n_valid = W.counters_host[0];
  return PMP_OK;
}

int pmp_eval_configs_host(const pmp_robot_t* robot, const pmp_scene_t* scene, const float* q_host, int64_t n, float* dist_host) {
  CHECK_COMMON(robot, n);
  if (n == 0) return PMP_OK;
  if (!q_host || !dist_host) return fail(PMP_ERR_INVALID, "null host pointer");
  WorkspaceLease L;
  if (int rc = L.acquire()) return rc;
  HostWorkspace& W = *L.ws;
  const int qd = robot->full.dev.q_dim;
  int k = 0;
  for (int64_t off = 0; off < n; off += kHostChunk, k = (k + 1) % kWsStreams) {
    const int64_t m = std::min<int64_t>(kHostChunk, n - off);
    if (int rc = W.reserve(&W.in[k][0], &W.in_cap[k][0], (size_t)m * qd * 4)) return rc;
    if (int rc = W.reserve(&W.outb[k], &W.out_cap[k], (size_t)m * 8)) return rc;
    cudaStream_t st = W.stream[k];
    CUDA_OK(cudaMemcpyAsync(W.in[k][0], q_host + off * qd, (size_t)m * qd * 4, cudaMemcpyHostToDevice, st));
    if (int rc = pmp_eval_configs(robot, scene, (const float*)W.in[k][0], m, (float*)W.outb[k], st)) return rc;
    CUDA_OK(cudaMemcpyAsync(dist_host + off * 2, W.outb[k], (size_t)m * 8, cudaMemcpyDeviceToHost, st));
  }
  return W.sync_all();
}

// ---- FP64 host states (what the reference's bindings hold: numpy float64 -> Eigen::VectorXd) ---------------
extern "C++" {
namespace {
// A few persistent host threads for the FP64 -> FP32 packing of large host batches (a pageable double
// array cannot be DMA'd as it is: the bytes have to be touched once anyway, so they are converted while
// being staged, which also halves what crosses PCIe).
// FP64 -> FP32 over a contiguous range; cloned for AVX2 where the host compiler supports function
// multi-versioning (the scalar loop converts ~1.3 GB/s per core, the vector one is memory-bound)
#if defined(__GNUC__) && defined(__x86_64__)
__attribute__((target_clones("avx2", "default")))
#endif
void pack_f64_f32(float* __restrict__ dst, const double* __restrict__ src, int64_t n) {
  for (int64_t i = 0; i < n; i++) dst[i] = (float)src[i];
}

class PackPool {
 public:
  static PackPool& get() {
    // never destroyed: its detached threads wait on the condition variable for the life of the process
    // (destroying a condition variable that has waiters blocks at exit)
    static PackPool* p = new PackPool();
    return *p;
  }
  int size() const { return (int)th_.size() + 1; }
  // runs fn(t) for t in [0, n_tasks) on the pool's threads and the caller
  void run(int n_tasks, const std::function<void(int)>& fn) {
    std::lock_guard<std::mutex> serial(run_mu_);
    {
      std::lock_guard<std::mutex> lk(mu_);
      fn_ = &fn;
      n_tasks_ = n_tasks;
      next_ = 0;
      pending_ = n_tasks;
      gen_++;
    }
    cv_.notify_all();
    work();
    std::unique_lock<std::mutex> lk(mu_);
    done_.wait(lk, [&] { return pending_ == 0; });
    fn_ = nullptr;
  }

 private:
  PackPool() {
    unsigned hw = std::thread::hardware_concurrency();
    const int n = (int)std::min<unsigned>(hw ? hw : 4u, 16u) / 2;  // half of the cores, at most 8
    for (int i = 0; i + 1 < std::max(n, 1); i++) th_.emplace_back([this] { loop(); });
    for (auto& t : th_) t.detach();
  }
  void work() {
    for (;;) {
      int t;
      {
        std::lock_guard<std::mutex> lk(mu_);
        if (!fn_ || next_ >= n_tasks_) return;
        t = next_++;
      }
      (*fn_)(t);
      std::lock_guard<std::mutex> lk(mu_);
      if (--pending_ == 0) done_.notify_all();
    }
  }
  void loop() {
    uint64_t seen = 0;
    for (;;) {
      {
        std::unique_lock<std::mutex> lk(mu_);
        cv_.wait(lk, [&] { return gen_ != seen; });
        seen = gen_;
      }
      work();
    }
  }
  std::mutex mu_, run_mu_;
  std::condition_variable cv_, done_;
  std::vector<std::thread> th_;
  const std::function<void(int)>* fn_ = nullptr;
  int n_tasks_ = 0, next_ = 0, pending_ = 0;
  uint64_t gen_ = 0;
};

struct PinnedRing {  // pinned FP32 staging of the FP64 entry points: grow-only, one ring per workspace stream
  float* buf[kWsStreams] = {};
  size_t cap[kWsStreams] = {};
  cudaEvent_t copied[kWsStreams] = {};
  int reserve(int k, size_t bytes) {
    if (!copied[k]) CUDA_OK(cudaEventCreateWithFlags(&copied[k], cudaEventDisableTiming));
    if (cap[k] >= bytes) return PMP_OK;
    if (buf[k]) cudaFreeHost(buf[k]);
    buf[k] = nullptr;
    cap[k] = 0;
    CUDA_OK(cudaMallocHost(&buf[k], bytes));
    cap[k] = bytes;
    return PMP_OK;
  }
};
std::mutex g_ring_mu;
std::unordered_map<HostWorkspace*, PinnedRing*> g_rings;
PinnedRing* ring_of(HostWorkspace* w) {
  std::lock_guard<std::mutex> lk(g_ring_mu);
  PinnedRing*& r = g_rings[w];
  if (!r) r = new PinnedRing();
  return r;
}
}  // namespace
}  // extern "C++"

int pmp_check_configs_host_f64(const pmp_robot_t* robot, const pmp_scene_t* scene, const double* q_host, int64_t n,
                               uint32_t* valid_bits_host, uint64_t* n_valid) {
  CHECK_COMMON(robot, n);
  if (n_valid) *n_valid = 0;
  if (n == 0) return PMP_OK;
  if (!q_host || !valid_bits_host) return fail(PMP_ERR_INVALID, "null host pointer");
  WorkspaceLease L;
  if (int rc = L.acquire()) return rc;
  HostWorkspace& W = *L.ws;
  PinnedRing& R = *ring_of(L.ws);
  const int qd = robot->full.dev.q_dim;
  CUDA_OK(cudaMemsetAsync(W.counters, 0, 2 * sizeof(unsigned long long), W.stream[0]));
  CUDA_OK(cudaStreamSynchronize(W.stream[0]));
  // Page-locked source (cudaHostAlloc / cudaHostRegister / pmp_host_alloc, e.g. a planner's reused state
  // array): the doubles are DMA'd as they are and narrowed on the device — 56 B per panda configuration at the
  // link rate instead of a host pack pass that the host's own memory bandwidth limits.
  {
    cudaPointerAttributes attr{};
    const bool pinned = cudaPointerGetAttributes(&attr, q_host) == cudaSuccess && attr.type == cudaMemoryTypeHost;
    cudaGetLastError();  // (an unregistered pointer sets the sticky-free "invalid value" on older drivers)
    if (pinned) {
      int k = 0;
      for (int64_t off = 0; off < n; off += kHostChunk, k = (k + 1) % kWsStreams) {
        const int64_t m = std::min<int64_t>(kHostChunk, n - off);
        const size_t elems = (size_t)m * qd;
        if (int rc = W.reserve(&W.in[k][0], &W.in_cap[k][0], elems * 4)) return rc;
        if (int rc = W.reserve(&W.in[k][1], &W.in_cap[k][1], elems * 8)) return rc;
        if (int rc = W.reserve(&W.outb[k], &W.out_cap[k], (size_t)((m + 31) / 32) * 4)) return rc;
        cudaStream_t st = W.stream[k];
        CUDA_OK(cudaMemcpyAsync(W.in[k][1], q_host + off * qd, elems * 8, cudaMemcpyHostToDevice, st));
        k_f64_to_f32<<<(int)std::min<int64_t>(((int64_t)elems + 255) / 256, (int64_t)num_sms() * 8), 256, 0, st>>>(
            (const double*)W.in[k][1], (float*)W.in[k][0], (long long)elems);
        g_launches++;
        if (int rc = pmp_check_configs(robot, scene, (const float*)W.in[k][0], m, (uint32_t*)W.outb[k], W.counters, st)) return rc;
        CUDA_OK(cudaMemcpyAsync(valid_bits_host + off / 32, W.outb[k], (size_t)((m + 31) / 32) * 4, cudaMemcpyDeviceToHost, st));
      }
      if (int rc = W.sync_all()) return rc;
      CUDA_OK(cudaMemcpyAsync(W.counters_host, W.counters, sizeof(unsigned long long), cudaMemcpyDeviceToHost, W.stream[0]));
      CUDA_OK(cudaStreamSynchronize(W.stream[0]));
      if (n_valid) *n_valid = W.counters_host[0];
      return PMP_OK;
    }
  }
  constexpr int64_t kChunk = 1 << 18;  // small enough that packing chunk k + 1 overlaps the copy and kernel of chunk k
  PackPool& pool = PackPool::get();
  int k = 0;
  for (int64_t off = 0; off < n; off += kChunk, k = (k + 1) % kWsStreams) {
    const int64_t m = std::min<int64_t>(kChunk, n - off);
    const size_t bytes = (size_t)m * qd * 4;
    if (int rc = R.reserve(k, bytes)) return rc;
    if (int rc = W.reserve(&W.in[k][0], &W.in_cap[k][0], bytes)) return rc;
    if (int rc = W.reserve(&W.outb[k], &W.out_cap[k], (size_t)((m + 31) / 32) * 4)) return rc;
    CUDA_OK(cudaEventSynchronize(R.copied[k]));  // the previous copy out of this staging buffer is done
    const double* src = q_host + off * qd;
    float* dst = R.buf[k];
    const int64_t total = m * qd;
    const int tasks = (int)std::min<int64_t>(pool.size(), std::max<int64_t>(1, total / 16384));
    pool.run(tasks, [&](int t) {
      const int64_t lo = total * t / tasks, hi = total * (t + 1) / tasks;
      pack_f64_f32(dst + lo, src + lo, hi - lo);
    });
    cudaStream_t st = W.stream[k];
    CUDA_OK(cudaMemcpyAsync(W.in[k][0], dst, bytes, cudaMemcpyHostToDevice, st));
    CUDA_OK(cudaEventRecord(R.copied[k], st));
    if (int rc = pmp_check_configs(robot, scene, (const float*)W.in[k][0], m, (uint32_t*)W.outb[k], W.counters, st)) return rc;
    CUDA_OK(cudaMemcpyAsync(valid_bits_host + off / 32, W.outb[k], (size_t)((m + 31) / 32) * 4, cudaMemcpyDeviceToHost, st));
  }
  if (int rc = W.sync_all()) return rc;
  CUDA_OK(cudaMemcpyAsync(W.counters_host, W.counters, sizeof(unsigned long long), cudaMemcpyDeviceToHost, W.stream[0]));
  CUDA_OK(cudaStreamSynchronize(W.stream[0]));
  if (n_valid) *n_valid = W.counters_host[0];
  return PMP_OK;
}

int pmp_check_edges_host(const pmp_robot_t* robot, const pmp_scene_t* scene, const pmp_validator_t* validator,
                         const float* q0_host, const float* q1_host, int64_t n, uint32_t* valid_bits_host,
                         uint64_t* n_valid, uint64_t* n_checked) {
  return check_edges_host_any<float>(robot, scene, validator, q0_host, q1_host, n, valid_bits_host, n_valid, n_checked);
}

int pmp_check_edges_host_f64(const pmp_robot_t* robot, const pmp_scene_t* scene, const pmp_validator_t* validator,
                             const double* q0_host, const double* q1_host, int64_t n, uint32_t* valid_bits_host,
                             uint64_t* n_valid, uint64_t* n_checked) {
  return check_edges_host_any<double>(robot, scene, validator, q0_host, q1_host, n, valid_bits_host, n_valid, n_checked);
}

// ---- validator sessions (the OMPL-facing adaptor's back end) ----------------------------------------
// A session owns a stream, one pinned + device-mapped host arena and one device arena sized at
// creation.  A call packs FP64 -> FP32 into the arena and then either
//   zero-copy (tiny batches): ONE kernel that reads its inputs from, and writes its results to, the
//                             mapped host arena, and a stream synchronise;
//   staged:                   ONE H2D copy of the compactly packed inputs, ONE kernel, ONE D2H copy;
// no allocation, no memset, no shared workspace lock, occupancy queries cached (grid_for).
struct pmp_session {
  const pmp_robot_t* robot = nullptr;
  const pmp_scene_t* scene = nullptr;
  bool has_validator = false;
  pmp_validator_t vdesc{};
  std::vector<double> width, inv_width;
  int64_t cap = 0;  // states / edges per launch; larger calls are chunked
  int qd = 0;
  cudaStream_t stream = nullptr;
  char* host = nullptr;      // pinned, mapped
  char* host_dev = nullptr;  // the device's view of `host`
  char* dev = nullptr;
  size_t arena = 0;
  unsigned long long* counters = nullptr;       // device [3]: n_valid, n_checked and the chunk dispenser of the early-exit path
  unsigned long long* counters_host = nullptr;  // pinned [2]
  std::mutex mu;
};

namespace {
constexpr long long kFlatMaxWaypoints = 1 << 17;  // beyond this the early-exit kernel wins
constexpr long long kZeroCopyWaypoints = 4096;    // flat path: results come back as bytes in mapped memory
constexpr int64_t kZeroCopyStates = 2048;
// Completion of a small session call without cudaStreamSynchronize: the flat kernel writes one byte (0 / 1) per
// waypoint into mapped host memory, the host presets those bytes to 0xff and spins until none is left.  One kernel
// round trip costs 11.4 us with a stream synchronise and 7.2 us this way (profiles/r02_s21_latency_floor.txt).
// Returns false when the bytes did not all arrive within the limit (the caller then synchronises, which also
// surfaces a failed launch).
bool spin_until_written(const uint8_t* bytes, long long n) {
  const volatile uint8_t* p = bytes;
  const auto t0 = std::chrono::steady_clock::now();
  unsigned polls = 0;
  for (long long i = 0; i < n; i++) {
    while (p[i] == 0xffu) {
#if defined(__x86_64__)
      __builtin_ia32_pause();
#endif
      if ((++polls & 0x3ffu) == 0 && std::chrono::steady_clock::now() - t0 > std::chrono::microseconds(2000)) return false;
    }
  }
  std::atomic_thread_fence(std::memory_order_acquire);
  return true;
}

// every waypoint of every edge in one wave (k_check_waypoints_flat); results as bits (way_bits) or bytes (way_ok)
int launch_waypoints_flat(const pmp_robot_t* robot, const SceneDev& S, const float* d0, const float* d1, const double* dr, const int* dof,
                          int m, long long total, uint32_t* dwb, uint8_t* dway, cudaStream_t st) {
  const int block = pick_block(robot->fast, S);
  const size_t smem = smem_bytes(robot->fast, S, block);  // the static offset table is accounted for by the occupancy query
  if (smem > (size_t)kMaxSmem) return fail(PMP_ERR_LIMIT, "robot + scene tables exceed the shared-memory budget");
  const int mode = robot->fast.v3_mode;  // 0 moving base (exact sin/cos), 1 FIXED base, 2 FIXED base + FP16 group centres
  const bool cloud = S.cloud.enabled != 0;
  const long long tiles = (total + block - 1) / block;
  int grid = 0;
#define PMP_F3(BB, AP, CL)                                                                                      \
  do {                                                                                                          \
    if (int rc = grid_for(k_check_waypoints_flat<BB, AP, CL>, smem, tiles, &grid, block)) return rc;            \
    k_check_waypoints_flat<BB, AP, CL><<<grid, block, smem, st>>>(robot->fast.dev, S, d0, d1, dr, dof, m, dwb, dway); \
  } while (0)
#define PMP_F3_B(BB)                                                                      \
  do {                                                                                    \
    if (mode == 2) { if (cloud) PMP_F3(BB, 2, true); else PMP_F3(BB, 2, false); }         \
    else if (mode == 1) { if (cloud) PMP_F3(BB, 1, true); else PMP_F3(BB, 1, false); }    \
    else { if (cloud) PMP_F3(BB, 0, true); else PMP_F3(BB, 0, false); }                   \
  } while (0)
  if (block == 128) PMP_F3_B(128); else PMP_F3_B(64);
#undef PMP_F3_B
#undef PMP_F3
  g_launches++;
  CUDA_OK(cudaGetLastError());
  return PMP_OK;
}

// BoxMotionValidator / EuclideanMotionValidator::determine_step_ratio in FP64 on the caller's
// doubles (ompl/motion_validator.cpp:89-126)
double host_step_ratio(const pmp_session* s, const double* a, const double* b) {
  if (s->vdesc.type == PMP_VALIDATOR_BOX) {
    double diff_longest = 0.0, max_diff = -std::numeric_limits<double>::infinity();
    int longest = 0;
    for (int i = 0; i < s->qd; i++) {  // first maximum wins (strict >)
      const double diff = b[i] - a[i];
      const double sc = std::abs(diff) * s->inv_width[i];
      if (sc > max_diff) {
        max_diff = sc;
        longest = i;
        diff_longest = diff;
      }
    }
    if (std::abs(diff_longest) < 1e-6) return std::numeric_limits<double>::infinity();
    return s->width[longest] / std::abs(diff_longest);
  }
  double d2 = 0.0;
  for (int i = 0; i < s->qd; i++) {
    const double d = a[i] - b[i];
    d2 += d * d;
  }
  return s->vdesc.resolution / std::sqrt(d2);
}
}  // namespace

int pmp_session_create(const pmp_robot_t* robot, const pmp_scene_t* scene, const pmp_validator_t* validator,
                       int64_t capacity, pmp_session_t** out) {
  if (!out) return fail(PMP_ERR_INVALID, "out is null");
  *out = nullptr;
  CHECK_COMMON(robot, capacity);
  if (capacity < 1) return fail(PMP_ERR_INVALID, "capacity must be positive");
  auto* s = new pmp_session();
  s->robot = robot;
  s->scene = scene;
  s->qd = robot->fast.dev.q_dim;
  s->cap = (capacity + 31) / 32 * 32;
  if (validator) {
    ValidatorDev V;  // validates the description (dimension, widths, type) like pmp_check_edges does
    if (int rc = validator_to_dev(validator, s->qd, &V)) {
      delete s;
      return rc;
    }
    s->has_validator = true;
    s->vdesc = *validator;
    if (validator->type == PMP_VALIDATOR_BOX) {
      s->width.assign(validator->box_width, validator->box_width + s->qd);
      for (double w : s->width) s->inv_width.push_back(1.0 / w);
      s->vdesc.box_width = s->width.data();
    }
  }
  // worst-case compact layout: q0, q1, ratio, off, bits, and the zero-copy result bytes
  const size_t qbytes = align256((size_t)s->cap * std::max(1, s->qd) * 4);
  s->arena = 2 * qbytes + align256((size_t)s->cap * 8) + align256((size_t)(s->cap + 1) * 4) +
             align256(std::max<size_t>((size_t)(s->cap / 32) * 4, (size_t)kFlatMaxWaypoints / 8)) + align256((size_t)kZeroCopyWaypoints);
  cudaError_t e = cudaStreamCreateWithFlags(&s->stream, cudaStreamNonBlocking);
  if (e == cudaSuccess) e = cudaHostAlloc(&s->host, s->arena, cudaHostAllocMapped);
  if (e == cudaSuccess) e = cudaHostGetDevicePointer(&s->host_dev, s->host, 0);
  if (e == cudaSuccess) e = cudaMalloc(&s->dev, s->arena);
  if (e == cudaSuccess) e = cudaMalloc(&s->counters, 3 * sizeof(unsigned long long));
  if (e == cudaSuccess) e = cudaMallocHost(&s->counters_host, 2 * sizeof(unsigned long long));
  if (e != cudaSuccess) {
    pmp_session_destroy(s);
    return fail(PMP_ERR_CUDA, cudaGetErrorString(e));
  }
  *out = s;
  return PMP_OK;
}

void pmp_session_destroy(pmp_session_t* s) {
  if (!s) return;
  if (s->stream) cudaStreamDestroy(s->stream);
  if (s->host) cudaFreeHost(s->host);
  if (s->dev) cudaFree(s->dev);
  if (s->counters) cudaFree(s->counters);
  if (s->counters_host) cudaFreeHost(s->counters_host);
  delete s;
}

int pmp_session_set_scene(pmp_session_t* s, const pmp_scene_t* scene) {
  if (!s) return fail(PMP_ERR_INVALID, "session handle is null");
  std::lock_guard<std::mutex> lk(s->mu);
  s->scene = scene;
  return PMP_OK;
}

int pmp_session_is_valid(pmp_session_t* s, const double* states, int64_t n, uint8_t* valid) {
  if (!s) return fail(PMP_ERR_INVALID, "session handle is null");
  if (n < 0) return fail(PMP_ERR_INVALID, "negative item count");
  if (n == 0) return PMP_OK;
  if ((!states && s->qd > 0) || !valid) return fail(PMP_ERR_INVALID, "null host pointer");
  std::lock_guard<std::mutex> lk(s->mu);
  if (int rc = ensure_device()) return rc;
  float* hq = reinterpret_cast<float*>(s->host);
  const SceneDev S = s->scene ? s->scene->dev : empty_scene();
  if (!mask_limit_ok(s->robot->fast, S)) return fail(PMP_ERR_LIMIT, "a robot with per-group scene masks needs a scene of at most 30 primitives");
  for (int64_t off = 0; off < n; off += s->cap) {
    const int64_t m = std::min<int64_t>(s->cap, n - off);
    const size_t nq = (size_t)m * s->qd;
    const size_t o_bits = align256(std::max<size_t>(nq, 1) * 4);
    const size_t words = (size_t)((m + 31) / 32);
    pack_f64_f32(hq, states + off * s->qd, (int64_t)nq);
    if (m <= kFlatSmemEdges) {
      // latency path: the states go through the flat waypoint kernel as m "edges" of one waypoint (s2 itself, the
      // ratio is never read), read from the mapped staging; one result byte per state comes back the same way and
      // the host sees the call complete by spinning on those bytes instead of synchronising the stream
      const size_t o_off = o_bits, o_way = o_off + align256((size_t)(m + 1) * 4);
      int* ho = reinterpret_cast<int*>(s->host + o_off);
      for (int64_t i = 0; i <= m; i++) ho[i] = (int)i;
      uint8_t* hway = reinterpret_cast<uint8_t*>(s->host + o_way);
      std::memset(hway, 0xff, (size_t)m);
      const float* dq = reinterpret_cast<const float*>(s->host_dev);
      if (int rc = launch_waypoints_flat(s->robot, S, dq, dq, reinterpret_cast<const double*>(s->host_dev), reinterpret_cast<const int*>(s->host_dev + o_off),
                                         (int)m, m, nullptr, reinterpret_cast<uint8_t*>(s->host_dev + o_way), s->stream))
        return rc;
      if (!spin_until_written(hway, m)) CUDA_OK(cudaStreamSynchronize(s->stream));
      for (int64_t i = 0; i < m; i++) valid[off + i] = hway[i] != 0;
      continue;
    }
    const uint32_t* hbits = reinterpret_cast<const uint32_t*>(s->host + o_bits);
    if (m <= kZeroCopyStates) {
      if (int rc = pmp_check_configs(s->robot, s->scene, reinterpret_cast<const float*>(s->host_dev), m,
                                     reinterpret_cast<uint32_t*>(s->host_dev + o_bits), nullptr, s->stream))
        return rc;
    } else {
      if (nq) CUDA_OK(cudaMemcpyAsync(s->dev, hq, nq * 4, cudaMemcpyHostToDevice, s->stream));
      if (int rc = pmp_check_configs(s->robot, s->scene, reinterpret_cast<const float*>(s->dev), m,
                                     reinterpret_cast<uint32_t*>(s->dev + o_bits), nullptr, s->stream))
        return rc;
      CUDA_OK(cudaMemcpyAsync(s->host + o_bits, s->dev + o_bits, words * 4, cudaMemcpyDeviceToHost, s->stream));
    }
    CUDA_OK(cudaStreamSynchronize(s->stream));
    for (int64_t i = 0; i < m; i++) valid[off + i] = (hbits[i >> 5] >> (i & 31)) & 1u;
  }
  return PMP_OK;
}

int pmp_session_check_motion(pmp_session_t* s, const double* s1, const double* s2, int64_t n, uint8_t* valid,
                             uint64_t* n_checked) {
  if (!s) return fail(PMP_ERR_INVALID, "session handle is null");
  if (n < 0) return fail(PMP_ERR_INVALID, "negative item count");
  if (!s->has_validator) return fail(PMP_ERR_INVALID, "the session was created without a validator");
  if (n_checked) *n_checked = 0;
  if (n == 0) return PMP_OK;
  if (((!s1 || !s2) && s->qd > 0) || !valid) return fail(PMP_ERR_INVALID, "null host pointer");
  std::lock_guard<std::mutex> lk(s->mu);
  DeviceState* D;
  if (int rc = device_state(&D)) return rc;
  if (int rc = ensure_sequence_table(D)) return rc;
  const pmp_robot_t* robot = s->robot;
  const SceneDev S = s->scene ? s->scene->dev : empty_scene();
  if (!mask_limit_ok(robot->fast, S)) return fail(PMP_ERR_LIMIT, "a robot with per-group scene masks needs a scene of at most 30 primitives");
  const int block = pick_block(robot->fast, S);
  const size_t smem = smem_bytes(robot->fast, S, block);
  if (smem > (size_t)kMaxSmem) return fail(PMP_ERR_LIMIT, "robot + scene tables exceed the shared-memory budget");
  const uint8_t* seq = sequence_table_host().data();
  uint64_t checked = 0;
  for (int64_t off = 0; off < n; off += s->cap) {
    const int64_t m = std::min<int64_t>(s->cap, n - off);
    // compact layout for this chunk: q0 | q1 | ratio | off (flat) or n_test (early exit) | result words | zero-copy bytes
    const size_t qb = align256((size_t)m * std::max(1, s->qd) * 4);
    const size_t o_q1 = qb, o_ratio = 2 * qb, o_off = o_ratio + align256((size_t)m * 8);
    const size_t o_bits = o_off + align256((size_t)(m + 1) * 4);
    float* h0 = reinterpret_cast<float*>(s->host);
    float* h1 = reinterpret_cast<float*>(s->host + o_q1);
    double* hr = reinterpret_cast<double*>(s->host + o_ratio);
    int* ho = reinterpret_cast<int*>(s->host + o_off);
    uint32_t* hbits = reinterpret_cast<uint32_t*>(s->host + o_bits);
    // the host pass: FP64 -> FP32 endpoints, step ratio and waypoint count from the caller's doubles
    // (motion_validator.cpp:89-126, :45); edges are independent, so large chunks are split over threads
    auto pack = [&](int64_t lo, int64_t hi) {
      for (int64_t e = lo; e < hi; e++) {
        const double* a = s1 + (off + e) * s->qd;
        const double* b = s2 + (off + e) * s->qd;
        pack_f64_f32(h0 + e * s->qd, a, s->qd);
        pack_f64_f32(h1 + e * s->qd, b, s->qd);
        const double r = host_step_ratio(s, a, b);
        hr[e] = r;
        ho[e + 1] = edge_n_test(r);  // n_test; turned into offsets below when the flat path is taken
      }
    };
    // (the persistent pool of the FP64 host entry points: spawning threads per call cost more than the pass itself)
    const int n_thr = (int)std::min<int64_t>(PackPool::get().size(), m / 2048);
    if (n_thr > 1) {
      PackPool::get().run(n_thr, [&](int t) { pack(m * t / n_thr, m * (t + 1) / n_thr); });
    } else {
      pack(0, m);
    }
    long long total = 0;  // s2 + the n_test - 2 interior points per edge
    for (int64_t e = 0; e < m; e++) total += std::max(0, ho[e + 1] - 1);
    if (total <= kFlatMaxWaypoints) {
      // latency path: every waypoint in one wave
      ho[0] = 0;
      for (int64_t e = 0; e < m; e++) ho[e + 1] = ho[e] + std::max(0, ho[e + 1] - 1);
      const bool zero_copy = total <= kZeroCopyWaypoints && m <= kFlatSmemEdges;
      const size_t way_words = (size_t)((total + 31) / 32);
      char* base = zero_copy ? s->host_dev : s->dev;
      if (total > 0) {
        if (zero_copy) std::memset(hbits, 0xff, (size_t)total);  // completion is seen byte by byte (spin_until_written)
        else CUDA_OK(cudaMemcpyAsync(s->dev, s->host, o_bits, cudaMemcpyHostToDevice, s->stream));
        const float* d0 = reinterpret_cast<const float*>(base);
        const float* d1 = reinterpret_cast<const float*>(base + o_q1);
        const double* dr = reinterpret_cast<const double*>(base + o_ratio);
        const int* dof = reinterpret_cast<const int*>(base + o_off);
        uint32_t* dwb = zero_copy ? nullptr : reinterpret_cast<uint32_t*>(base + o_bits);
        uint8_t* dway = zero_copy ? reinterpret_cast<uint8_t*>(base + o_bits) : nullptr;
        if (int rc = launch_waypoints_flat(robot, S, d0, d1, dr, dof, (int)m, total, dwb, dway, s->stream)) return rc;
        if (!zero_copy) CUDA_OK(cudaMemcpyAsync(hbits, s->dev + o_bits, way_words * 4, cudaMemcpyDeviceToHost, s->stream));
        if (!zero_copy || !spin_until_written(reinterpret_cast<const uint8_t*>(hbits), total)) CUDA_OK(cudaStreamSynchronize(s->stream));
      }
      // per edge: AND over its waypoints, and the reference's call count = position of the first failing
      // waypoint in ITS visiting order (s2, then sequence[2..] of the farthest-first table; natural order
      // from n_test = 129 on, motion_validator.cpp:45-78), else the waypoint count
      const uint8_t* way = reinterpret_cast<const uint8_t*>(hbits);
      auto ok_at = [&](long long t) -> bool { return zero_copy ? way[t] != 0 : ((hbits[t >> 5] >> (t & 31)) & 1u) != 0; };
      for (int64_t e = 0; e < m; e++) {
        const int n_way = ho[e + 1] - ho[e], n_test = n_way + 1;
        bool ok = true;
        for (int t = 0; t < n_way && ok; t++) ok = ok_at(ho[e] + t);
        valid[off + e] = ok ? 1 : 0;
        if (ok) {
          checked += (uint64_t)n_way;
        } else {
          int k = 0;
          for (;; k++) {
            const int j = k == 0 ? 0 : (n_test < 129 ? (int)seq[(n_test - 1) * 128 + (k + 1)] : k);
            if (!ok_at(ho[e] + j)) break;
          }
          checked += (uint64_t)k + 1;
        }
      }
    } else {
      // throughput path: lane = edge with early exit (k_check_edges_v3) on the host-computed ratios / n_test
      for (int64_t e = 0; e < m; e++) ho[e] = ho[e + 1];
      const size_t words = (size_t)((m + 31) / 32);
      std::memset(hbits, 0, words * 4);
      CUDA_OK(cudaMemcpyAsync(s->dev, s->host, o_bits + words * 4, cudaMemcpyHostToDevice, s->stream));
      CUDA_OK(cudaMemsetAsync(s->counters, 0, 3 * sizeof(unsigned long long), s->stream));
      EdgeArgs A{};
      A.q0 = reinterpret_cast<const float*>(s->dev);
      A.q1 = reinterpret_cast<const float*>(s->dev + o_q1);
      A.ratio = reinterpret_cast<const double*>(s->dev + o_ratio);
      A.ntest = reinterpret_cast<const int*>(s->dev + o_off);
      A.n = m;
      A.bits = reinterpret_cast<uint32_t*>(s->dev + o_bits);
      A.n_valid = s->counters;
      A.n_checked = s->counters + 1;
      A.next_chunk = s->counters + 2;
      if (int rc = launch_edges(robot, S, D->seq_dev, A, s->stream)) return rc;
      CUDA_OK(cudaMemcpyAsync(hbits, s->dev + o_bits, words * 4, cudaMemcpyDeviceToHost, s->stream));
      CUDA_OK(cudaMemcpyAsync(s->counters_host, s->counters, 2 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, s->stream));
      CUDA_OK(cudaStreamSynchronize(s->stream));
      for (int64_t e = 0; e < m; e++) valid[off + e] = (hbits[e >> 5] >> (e & 31)) & 1u;
      checked += s->counters_host[1];
    }
  }
  if (n_checked) *n_checked = checked;
  return PMP_OK;
}

// ---- psdf / kdtree ---------------------------------------------------------------------------------
int pmp_sdf_evaluate(const pmp_scene_t* scene, const float* points_dev, int64_t n, float* out_dev, void* stream) {
  if (!scene) return fail(PMP_ERR_INVALID, "scene handle is null");
  if (n < 0) return fail(PMP_ERR_INVALID, "negative item count");
  if (int rc = ensure_device()) return rc;
  if (n == 0) return PMP_OK;
  if (!points_dev || !out_dev) return fail(PMP_ERR_INVALID, "null device pointer");
  const int grid = (int)std::min<int64_t>((n + 255) / 256, (int64_t)num_sms() * 8);
  k_sdf_evaluate<<<grid, 256, 0, (cudaStream_t)stream>>>(scene->dev, points_dev, n, out_dev);
  g_launches++;
  CUDA_OK(cudaGetLastError());
  return PMP_OK;
}

int pmp_sdf_is_outside(const pmp_scene_t* scene, const float* points_dev, const float* radii_dev, int64_t n,
                       uint8_t* out_dev, void* stream) {
  if (!scene) return fail(PMP_ERR_INVALID, "scene handle is null");
  if (n < 0) return fail(PMP_ERR_INVALID, "negative item count");
  if (int rc = ensure_device()) return rc;
  if (n == 0) return PMP_OK;
  if (!points_dev || !out_dev) return fail(PMP_ERR_INVALID, "null device pointer");
  const int grid = (int)std::min<int64_t>((n + 255) / 256, (int64_t)num_sms() * 8);
  k_sdf_is_outside<<<grid, 256, 0, (cudaStream_t)stream>>>(scene->dev, points_dev, radii_dev, n, out_dev);
  g_launches++;
  CUDA_OK(cudaGetLastError());
  return PMP_OK;
}

int pmp_cloud_nearest(const pmp_scene_t* scene, const float* points_dev, int64_t n, float* sqdist_dev,
                      int32_t* nearest_index_dev, void* stream) {
  if (!scene) return fail(PMP_ERR_INVALID, "scene handle is null");
  if (!scene->dev.cloud.enabled) return fail(PMP_ERR_INVALID, "scene has no point cloud");
  if (n < 0) return fail(PMP_ERR_INVALID, "negative item count");
  if (int rc = ensure_device()) return rc;
  if (n == 0) return PMP_OK;
  if (!points_dev || !sqdist_dev) return fail(PMP_ERR_INVALID, "null device pointer");
  const int grid = (int)std::min<int64_t>((n + 15) / 16, (int64_t)num_sms() * 16);  // 16 queries per CTA of 128 threads
  k_cloud_nearest<<<grid, 128, 0, (cudaStream_t)stream>>>(scene->dev, points_dev, n, sqdist_dev, nearest_index_dev);
  g_launches++;
  CUDA_OK(cudaGetLastError());
  return PMP_OK;
}

}  // extern "C"

// ---- FP32 roofline denominator ---------------------------------------------------------------------
namespace {
__global__ void k_ffma_peak(float* out, int iters, float a, float b) {
  float x0 = threadIdx.x * 1e-3f, x1 = x0 + 1.f, x2 = x0 + 2.f, x3 = x0 + 3.f, x4 = x0 + 4.f, x5 = x0 + 5.f,
        x6 = x0 + 6.f, x7 = x0 + 7.f;
#pragma unroll 1
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int k = 0; k < 16; k++) {
      x0 = fmaf(x0, a, b); x1 = fmaf(x1, a, b); x2 = fmaf(x2, a, b); x3 = fmaf(x3, a, b);
      x4 = fmaf(x4, a, b); x5 = fmaf(x5, a, b); x6 = fmaf(x6, a, b); x7 = fmaf(x7, a, b);
    }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
}
}  // namespace

extern "C" int pmp_measure_fp32_peak(double* tflops) {
  if (!tflops) return fail(PMP_ERR_INVALID, "null argument");
  if (int rc = ensure_device()) return rc;
  const int block = 256, grid = num_sms() * 8, iters = 4096;
  float* out = nullptr;
  CUDA_OK(cudaMalloc(&out, (size_t)grid * block * 4));
  cudaEvent_t e0, e1;
  CUDA_OK(cudaEventCreate(&e0));
  CUDA_OK(cudaEventCreate(&e1));
  double best = 0;
  for (int rep = 0; rep < 5; rep++) {
    CUDA_OK(cudaEventRecord(e0));
    k_ffma_peak<<<grid, block>>>(out, iters, 0.999f, 1e-3f);
    CUDA_OK(cudaEventRecord(e1));
    CUDA_OK(cudaEventSynchronize(e1));
    float ms = 0;
    CUDA_OK(cudaEventElapsedTime(&ms, e0, e1));
    const double flop = 2.0 * 8 * 16 * (double)iters * grid * block;
    best = std::max(best, flop / (ms * 1e-3) / 1e12);
  }
  g_launches += 5;
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(out);
  *tflops = best;
  return PMP_OK;
}
