// TEST INFRASTRUCTURE — never part of the product, never loaded by plainmp_b200/.
//
// A small SIMT emulator: enough of the CUDA C++ device language and runtime API to compile the
// kernel sources of plainmp_b200/csrc/ UNCHANGED with g++ and run them on CPU threads, one OS thread
// per CUDA thread, warp collectives implemented as rendezvous between the lanes named in the mask.
// Purpose: exercise the kernels' CONTROL FLOW (warp-local queues, overflow handling, sub-warp drains,
// dynamic edge refill) in the CPU test tier and on machines without a GPU — e.g. to reproduce a
// livelock without spending GPU time.  It says nothing about performance and is far too slow for
// anything but a few hundred items.  tests/cpu_emul/build.py builds it; tests/test_simt_emul.py uses it.
//
// Semantics implemented
//   * __ballot_sync / __any_sync / __all_sync / __shfl_sync / __shfl_xor_sync / __shfl_down_sync / __reduce_or_sync /
//     __syncwarp: every lane named in `mask` must arrive at a collective with the same mask; lanes of
//     different (disjoint) masks proceed independently, like independent thread scheduling on sm_70+.
//   * __syncthreads: barrier over the CTA.  One CTA runs at a time (grids are persistent / grid-stride
//     in this code base, so any grid size computes the same result).
//   * shared memory: `extern __shared__ float4 sm[]` binds to one global arena; function-scope
//     `__shared__` arrays become function-local statics (one CTA at a time).
//   * cudaMalloc / cudaMemcpy* / streams / events: synchronous host implementations.
#pragma once
#ifndef _GNU_SOURCE
#define _GNU_SOURCE
#endif
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <cmath>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <thread>
#include <tuple>
#include <utility>
#include <vector>

#define __device__
#define __host__
#define __global__
#define __forceinline__ inline
#define __noinline__ inline
#define __launch_bounds__(...)
#define __shared__ static

// ---- vector types --------------------------------------------------------------------------------
struct float2 { float x, y; };
struct float3 { float x, y, z; };
struct alignas(16) float4 { float x, y, z, w; };
struct int2 { int x, y; };
struct alignas(16) int4 { int x, y, z, w; };
struct uint3 { unsigned x, y, z; };
struct dim3 { unsigned x = 1, y = 1, z = 1; dim3(unsigned a = 1, unsigned b = 1, unsigned c = 1) : x(a), y(b), z(c) {} };
inline float3 make_float3(float x, float y, float z) { return {x, y, z}; }
inline float4 make_float4(float x, float y, float z, float w) { return {x, y, z, w}; }
inline int2 make_int2(int x, int y) { return {x, y}; }

// ---- scalar intrinsics ---------------------------------------------------------------------------
template <class T> inline T __ldg(const T* p) { return *p; }
inline int __float_as_int(float f) { int i; memcpy(&i, &f, 4); return i; }
inline unsigned __float_as_uint(float f) { unsigned i; memcpy(&i, &f, 4); return i; }
inline float __int_as_float(int i) { float f; memcpy(&f, &i, 4); return f; }
inline float __uint_as_float(unsigned i) { float f; memcpy(&f, &i, 4); return f; }
inline int __popc(unsigned v) { return __builtin_popcount(v); }
inline int __ffs(unsigned v) { return __builtin_ffs((int)v); }
inline float rsqrtf(float x) { return 1.0f / sqrtf(x); }
inline int min(int a, int b) { return a < b ? a : b; }
inline int max(int a, int b) { return a > b ? a : b; }
inline long long min(long long a, long long b) { return a < b ? a : b; }
inline long long max(long long a, long long b) { return a > b ? a : b; }
typedef _Float16 __half;
inline __half __float2half_rn(float f) { return (__half)f; }
inline float __half2float(__half h) { return (float)h; }

template <class T> inline T atomicAdd(T* p, T v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
template <class T> inline T atomicOr(T* p, T v) { return __atomic_fetch_or(p, v, __ATOMIC_SEQ_CST); }

// ---- thread identity -------------------------------------------------------------------------------
namespace emu {
struct Warp;
struct Cta;
struct ThreadCtx {
  uint3 tid{0, 0, 0}, bid{0, 0, 0}, bdim{1, 1, 1}, gdim{1, 1, 1};
  Warp* warp = nullptr;
  Cta* cta = nullptr;
};
inline ThreadCtx& ctx() {
  static thread_local ThreadCtx c;
  return c;
}
}  // namespace emu
#define threadIdx (emu::ctx().tid)
#define blockIdx (emu::ctx().bid)
#define blockDim (emu::ctx().bdim)
#define gridDim (emu::ctx().gdim)

// ---- warp collectives --------------------------------------------------------------------------------
namespace emu {
struct Meeting {
  unsigned mask = 0, arrived = 0;
  unsigned long long gen = 0;
  uint64_t vals[32];
  uint64_t res[32];
};
struct Warp {
  std::mutex m;
  std::condition_variable cv;
  std::vector<Meeting> meetings;  // one per mask currently in use (a handful)
  Meeting& meeting(unsigned mask) {
    for (auto& x : meetings)
      if (x.mask == mask) return x;
    meetings.emplace_back();
    meetings.back().mask = mask;
    return meetings.back();
  }
};
// every lane of `mask` contributes v; returns a snapshot of all contributions
inline void exchange(unsigned mask, uint64_t v, uint64_t* out) {
  ThreadCtx& c = ctx();
  Warp& w = *c.warp;
  const int lane = c.tid.x & 31;
  std::unique_lock<std::mutex> lk(w.m);
  // (std::vector may reallocate when a new mask appears: re-find the meeting after every wait)
  {
    Meeting& M = w.meeting(mask);
    M.vals[lane] = v;
    M.arrived |= 1u << lane;
    if (M.arrived == mask) {
      memcpy(M.res, M.vals, sizeof(M.vals));
      M.arrived = 0;
      M.gen++;
      memcpy(out, M.res, sizeof(M.res));
      w.cv.notify_all();
      return;
    }
  }
  const unsigned long long g = w.meeting(mask).gen;
  w.cv.wait(lk, [&] { return w.meeting(mask).gen != g; });
  memcpy(out, w.meeting(mask).res, sizeof(uint64_t) * 32);
}
template <class T> inline uint64_t to_bits(T v) { uint64_t b = 0; memcpy(&b, &v, sizeof(T)); return b; }
template <class T> inline T from_bits(uint64_t b) { T v; memcpy(&v, &b, sizeof(T)); return v; }
}  // namespace emu

inline unsigned __ballot_sync(unsigned mask, bool pred) {
  uint64_t all[32];
  emu::exchange(mask, pred ? 1 : 0, all);
  unsigned r = 0;
  for (int l = 0; l < 32; l++)
    if (((mask >> l) & 1u) && all[l]) r |= 1u << l;
  return r;
}
inline bool __any_sync(unsigned mask, bool pred) { return __ballot_sync(mask, pred) != 0u; }
inline bool __all_sync(unsigned mask, bool pred) { return __ballot_sync(mask, !pred) == 0u; }
inline void __syncwarp(unsigned mask = 0xffffffffu) {
  uint64_t all[32];
  emu::exchange(mask, 0, all);
}
template <class T> inline T __shfl_sync(unsigned mask, T v, int src) {
  uint64_t all[32];
  emu::exchange(mask, emu::to_bits(v), all);
  return emu::from_bits<T>(all[src & 31]);
}
template <class T> inline T __shfl_xor_sync(unsigned mask, T v, int off) {
  uint64_t all[32];
  emu::exchange(mask, emu::to_bits(v), all);
  return emu::from_bits<T>(all[((int)(threadIdx.x & 31)) ^ off]);
}
template <class T> inline T __shfl_down_sync(unsigned mask, T v, int off) {
  uint64_t all[32];
  emu::exchange(mask, emu::to_bits(v), all);
  const int lane = (int)(threadIdx.x & 31);
  return emu::from_bits<T>(all[lane + off < 32 ? lane + off : lane]);
}
inline unsigned __reduce_or_sync(unsigned mask, unsigned v) {
  uint64_t all[32];
  emu::exchange(mask, v, all);
  unsigned r = 0;
  for (int l = 0; l < 32; l++)
    if ((mask >> l) & 1u) r |= (unsigned)all[l];
  return r;
}

// ---- CTA barrier ---------------------------------------------------------------------------------------
namespace emu {
struct Cta {
  std::mutex m;
  std::condition_variable cv;
  int n = 0, arrived = 0;
  unsigned long long gen = 0;
  std::vector<Warp> warps;
};
}  // namespace emu
inline void __syncthreads() {
  emu::Cta& c = *emu::ctx().cta;
  std::unique_lock<std::mutex> lk(c.m);
  const unsigned long long g = c.gen;
  if (++c.arrived == c.n) {
    c.arrived = 0;
    c.gen++;
    c.cv.notify_all();
  } else {
    c.cv.wait(lk, [&] { return c.gen != g; });
  }
}

// ---- shared memory arena (one CTA at a time): defined in emul_extra.cpp as pmp::sm ------------------------
namespace pmp { extern float4 sm[1 << 16]; }
#ifdef PMP_EMUL_ASAN
#include <sanitizer/asan_interface.h>
#endif

// ---- kernel launch: K % emu::Cfg(grid, block, smem, stream)(args...) ------------------------------------
// (tests/cpu_emul/build.py rewrites `K<<<g, b, s, st>>>(args)` into that form)
namespace emu {
template <class... A>
struct Bound {
  unsigned grid, block;
  size_t smem;
  std::tuple<A...> args;
};
struct Cfg {
  unsigned grid, block;
  size_t smem;
  template <class G, class B> Cfg(G g, B b, size_t s = 0, void* = nullptr) : grid((unsigned)g), block((unsigned)b), smem(s) {}
  template <class... A> Bound<std::decay_t<A>...> operator()(A&&... a) const {
    return Bound<std::decay_t<A>...>{grid, block, smem, std::tuple<std::decay_t<A>...>(std::forward<A>(a)...)};
  }
};
#ifdef PMP_EMUL_ASAN
#ifndef PMP_EMUL_ASAN_SELFTEST
#define PMP_EMUL_ASAN_SELFTEST 0  // 64: poison the last 64 requested bytes too (checks that the checker fires)
#endif
struct SmemGuard {
  size_t used;
  explicit SmemGuard(size_t n) : used((n + 7) / 8 * 8 - (n >= 64 ? PMP_EMUL_ASAN_SELFTEST : 0)) {
    ASAN_POISON_MEMORY_REGION((char*)pmp::sm + used, sizeof(float4) * (1 << 16) - used);
  }
  ~SmemGuard() { ASAN_UNPOISON_MEMORY_REGION((char*)pmp::sm + used, sizeof(float4) * (1 << 16) - used); }
};
#endif
constexpr unsigned kMaxGrid = 2;  // persistent / grid-stride kernels: a couple of CTAs cover every path
template <class... P, class... A>
inline void launch(void (*k)(P...), const Bound<A...>& b) {
  const unsigned grid = std::min(b.grid, kMaxGrid);
#ifdef PMP_EMUL_ASAN
  // memcheck mode: the dynamic shared memory the launch did not ask for is poisoned, device
  // allocations carry no slack (cudaMalloc below), so AddressSanitizer reports what
  // compute-sanitizer's memcheck would
  emu::SmemGuard guard(b.smem);
#endif
  for (unsigned bid = 0; bid < grid; bid++) {
    Cta cta;
    cta.n = (int)b.block;
    cta.warps = std::vector<Warp>((b.block + 31) / 32);
    std::vector<std::thread> th;
    for (unsigned t = 0; t < b.block; t++)
      th.emplace_back([&, t] {
        ThreadCtx& c = ctx();
        c.tid = {t, 0, 0};
        c.bid = {bid, 0, 0};
        c.bdim = {b.block, 1, 1};
        c.gdim = {grid, 1, 1};
        c.cta = &cta;
        c.warp = &cta.warps[t >> 5];
        std::apply([&](const A&... a) { k(static_cast<P>(a)...); }, b.args);
      });
    for (auto& x : th) x.join();
  }
}
}  // namespace emu
template <class... P, class... A>
inline void operator%(void (*k)(P...), const emu::Bound<A...>& b) { emu::launch(k, b); }

// ---- runtime API (synchronous host implementations) -----------------------------------------------------
typedef int cudaError_t;
typedef void* cudaStream_t;
typedef void* cudaEvent_t;
enum { cudaSuccess = 0, cudaErrorNoDevice = 100, cudaErrorMemoryAllocation = 2 };
enum cudaMemcpyKind { cudaMemcpyHostToDevice = 1, cudaMemcpyDeviceToHost = 2, cudaMemcpyDeviceToDevice = 3 };
enum { cudaStreamNonBlocking = 1, cudaHostAllocMapped = 2, cudaHostAllocPortable = 1, cudaHostAllocWriteCombined = 4 };
enum { cudaDevAttrMultiProcessorCount = 16, cudaFuncAttributeMaxDynamicSharedMemorySize = 8 };
inline const char* cudaGetErrorString(cudaError_t e) { return e == cudaSuccess ? "no error" : "emulated CUDA error"; }
#ifdef PMP_EMUL_ASAN
template <class T> inline cudaError_t cudaMalloc(T** p, size_t n) { void* q = nullptr; if (posix_memalign(&q, 256, n ? n : 1)) q = nullptr; *p = (T*)q; return q ? cudaSuccess : cudaErrorMemoryAllocation; }
#else
template <class T> inline cudaError_t cudaMalloc(T** p, size_t n) { *p = (T*)aligned_alloc(256, (n + 255) / 256 * 256 + 256); return *p ? cudaSuccess : cudaErrorMemoryAllocation; }
#endif
inline cudaError_t cudaFree(void* p) { free(p); return cudaSuccess; }
// stream-ordered allocation: everything is synchronous here, so these are plain malloc / free
template <class T> inline cudaError_t cudaMallocAsync(T** p, size_t n, cudaStream_t) { return cudaMalloc(p, n); }
inline cudaError_t cudaFreeAsync(void* p, cudaStream_t) { free(p); return cudaSuccess; }
typedef void* cudaMemPool_t;
enum cudaMemPoolAttr { cudaMemPoolAttrReleaseThreshold = 4 };
inline cudaError_t cudaDeviceGetDefaultMemPool(cudaMemPool_t* p, int) { *p = nullptr; return cudaSuccess; }
inline cudaError_t cudaMemPoolSetAttribute(cudaMemPool_t, cudaMemPoolAttr, void*) { return cudaSuccess; }
template <class T> inline cudaError_t cudaMallocHost(T** p, size_t n) { return cudaMalloc(p, n); }
template <class T> inline cudaError_t cudaHostAlloc(T** p, size_t n, unsigned) { return cudaMalloc(p, n); }
inline cudaError_t cudaFreeHost(void* p) { free(p); return cudaSuccess; }
// pointer attributes: every pointer is "pageable host" unless PMP_EMUL_PINNED is set (then every pointer counts as
// page-locked, which sends the FP64 host entry point down its DMA-the-doubles path)
enum cudaMemoryType { cudaMemoryTypeUnregistered = 0, cudaMemoryTypeHost = 1, cudaMemoryTypeDevice = 2 };
struct cudaPointerAttributes { cudaMemoryType type; };
inline cudaError_t cudaPointerGetAttributes(cudaPointerAttributes* a, const void*) {
  a->type = getenv("PMP_EMUL_PINNED") ? cudaMemoryTypeHost : cudaMemoryTypeUnregistered;
  return cudaSuccess;
}
template <class T, class U> inline cudaError_t cudaHostGetDevicePointer(T** d, U* h, unsigned) { *d = (T*)h; return cudaSuccess; }
inline cudaError_t cudaMemcpy(void* d, const void* s, size_t n, cudaMemcpyKind) { memcpy(d, s, n); return cudaSuccess; }
inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind, cudaStream_t = nullptr) { memcpy(d, s, n); return cudaSuccess; }
inline cudaError_t cudaMemsetAsync(void* d, int v, size_t n, cudaStream_t = nullptr) { memset(d, v, n); return cudaSuccess; }
inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t* s, unsigned) { *s = (void*)1; return cudaSuccess; }
inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
inline cudaError_t cudaStreamDestroy(cudaStream_t) { return cudaSuccess; }
inline cudaError_t cudaDeviceSynchronize() { return cudaSuccess; }
inline cudaError_t cudaGetLastError() { return cudaSuccess; }
inline cudaError_t cudaGetDeviceCount(int* n) { *n = 1; return cudaSuccess; }
inline cudaError_t cudaGetDevice(int* d) { *d = 0; return cudaSuccess; }
inline cudaError_t cudaSetDevice(int) { return cudaSuccess; }
inline cudaError_t cudaDeviceGetAttribute(int* v, int, int) { *v = 1; return cudaSuccess; }  // one "SM": grids of a CTA or two
template <class K> inline cudaError_t cudaFuncSetAttribute(K, int, int) { return cudaSuccess; }
template <class K> inline cudaError_t cudaOccupancyMaxActiveBlocksPerMultiprocessor(int* n, K, int, size_t) { *n = 1; return cudaSuccess; }
inline cudaError_t cudaEventCreate(cudaEvent_t* e) { *e = (void*)1; return cudaSuccess; }
enum { cudaEventDisableTiming = 2 };
inline cudaError_t cudaEventCreateWithFlags(cudaEvent_t* e, unsigned) { *e = (void*)1; return cudaSuccess; }
inline cudaError_t cudaEventRecord(cudaEvent_t, cudaStream_t = nullptr) { return cudaSuccess; }
inline cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned = 0) { return cudaSuccess; }
inline cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }
inline cudaError_t cudaEventElapsedTime(float* ms, cudaEvent_t, cudaEvent_t) { *ms = 1.0f; return cudaSuccess; }
inline cudaError_t cudaEventDestroy(cudaEvent_t) { return cudaSuccess; }
