// gm_common.cuh -- shared host/device helpers for libgeosmie_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/geosmie_b200.h"

// ------------------------------------------------------------------------------------------------ errors
void gm_set_error(const char* fmt, ...);

#define GM_CUDA_TRY(expr)                                                                       \
  do {                                                                                          \
    cudaError_t _e = (expr);                                                                    \
    if (_e != cudaSuccess) {                                                                    \
      gm_set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
      return GM_ECUDA;                                                                          \
    }                                                                                           \
  } while (0)

#define GM_REQUIRE(cond, msg)                                     \
  do {                                                            \
    if (!(cond)) {                                                \
      gm_set_error("invalid argument: %s (%s)", msg, #cond);      \
      return GM_EINVAL;                                           \
    }                                                             \
  } while (0)

// ------------------------------------------------------------------------------------------------ device buffers
// Process-wide cache of released device buffers (per device, best fit).  A table build walks through its size bins with one
// gm_table after the other; each owns ~30 small device arrays, and cudaFree / cudaMalloc of those -- normally 20 us each -- were seen
// to take milliseconds each in a third of the optics_SS builds on some boxes (gm_table_destroy 0.34 s instead of 3 ms,
// tools/diag_lut_outliers.py).  Released buffers up to 64 MB are therefore kept (at most 1 GB / 1024 buffers per process) and handed to
// the next allocation of a similar size; the large per-table outputs go through the handle's pool (gm_pool_*).  All work of the library
// is ordered on the handle's stream, so a recycled buffer cannot still be in use.  GEOSMIE_NO_ALLOC_CACHE=1 turns the cache off.
struct GmCacheEntry {
  void* p;
  size_t cap;
  int dev;
};
inline std::mutex g_gm_cache_mu;
inline std::vector<GmCacheEntry> g_gm_cache;
inline size_t g_gm_cache_bytes = 0;
inline bool gm_cache_enabled() {
  static const bool on = getenv("GEOSMIE_NO_ALLOC_CACHE") == nullptr;
  return on;
}
inline void* gm_cache_take(size_t want, size_t* cap) {
  if (!gm_cache_enabled()) return nullptr;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return nullptr;
  std::lock_guard<std::mutex> lk(g_gm_cache_mu);
  int best = -1;
  for (int i = 0; i < (int)g_gm_cache.size(); ++i) {
    const GmCacheEntry& e = g_gm_cache[i];
    if (e.dev == dev && e.cap >= want && e.cap <= want + want / 2 + 4096 && (best < 0 || e.cap < g_gm_cache[best].cap)) best = i;
  }
  if (best < 0) return nullptr;
  void* p = g_gm_cache[best].p;
  *cap = g_gm_cache[best].cap;
  g_gm_cache_bytes -= *cap;
  g_gm_cache.erase(g_gm_cache.begin() + best);
  return p;
}
inline void gm_cache_put(void* p, size_t cap) {
  if (!p) return;
  int dev = 0;
  if (gm_cache_enabled() && cap <= ((size_t)64 << 20) && cudaGetDevice(&dev) == cudaSuccess) {
    std::lock_guard<std::mutex> lk(g_gm_cache_mu);
    if (g_gm_cache.size() < 1024 && g_gm_cache_bytes + cap <= ((size_t)1 << 30)) {
      g_gm_cache.push_back({p, cap, dev});
      g_gm_cache_bytes += cap;
      return;
    }
  }
  cudaFree(p);
}
// frees the cached buffers of the current device (gm_destroy)
inline void gm_cache_flush() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return;
  std::lock_guard<std::mutex> lk(g_gm_cache_mu);
  for (size_t i = 0; i < g_gm_cache.size();) {
    if (g_gm_cache[i].dev == dev) {
      cudaFree(g_gm_cache[i].p);
      g_gm_cache_bytes -= g_gm_cache[i].cap;
      g_gm_cache.erase(g_gm_cache.begin() + i);
    } else {
      ++i;
    }
  }
}

// Grow-only device allocation; keeps repeated API calls free of cudaMalloc.
struct DevBuf {
  void* p = nullptr;
  size_t cap = 0;
  int ensure(size_t bytes) {
    if (bytes <= cap) return GM_OK;
    gm_cache_put(p, cap);
    p = nullptr;
    cap = 0;
    size_t want = bytes + bytes / 8 + 256;
    if ((p = gm_cache_take(want, &cap)) != nullptr) return GM_OK;
    cudaError_t e = cudaMalloc(&p, want);
    if (e == cudaErrorMemoryAllocation) {      // give the cached buffers back to the driver and try once more
      cudaGetLastError();
      gm_cache_flush();
      e = cudaMalloc(&p, want);
    }
    if (e != cudaSuccess) {
      gm_set_error("cudaMalloc(%zu) failed: %s", want, cudaGetErrorString(e));
      p = nullptr;
      return GM_ENOMEM;
    }
    cap = want;
    return GM_OK;
  }
  void release() {
    gm_cache_put(p, cap);
    p = nullptr;
    cap = 0;
  }
  template <typename T>
  T* as() const { return reinterpret_cast<T*>(p); }
};

struct gm_handle_s {
  int device = 0;
  int sm_count = 148;
  long long l2_persist_max = 0, l2_window_max = 0;   // persisting-L2 carve-out set at gm_init, largest access-policy window
  cudaStream_t stream = nullptr;
  int64_t launches = 0;
  // scratch for the per-particle API and GSF/bands
  DevBuf ws[16];
  size_t coef_budget_bytes = (size_t)4 << 30;  // coefficient staging buffer per gm_table_run batch (optics_SU dense: 2.9 GB = one batch)
  // per-run scratch of gm_table_run* (coefficient stream, partial sums, weights): grow-only and shared by all tables of the
  // handle, so that building one table per size bin does not pay a multi-GB cudaMalloc / cudaFree per bin
  // buffers handed back by destroyed tables (phase sums, normalised planes, GSF moments: up to ~1 GB each on fine spectral grids):
  // a table build walks through its size bins with one table after the other, and cudaFree + cudaMalloc of these buffers was 0.34 s of
  // the 2.3 s optics_SS 2048-wavelength build
  std::vector<DevBuf> pool;
  DevBuf scratch_coef, scratch_gact, scratch_scal_part, scratch_part, scratch_g_hpart, scratch_g_hsum, scratch_wphase, scratch_wscal, scratch_taskc, scratch_nodepart;
  // GSF constants (Gauss nodes, interpolation brackets, generalized spherical functions) cached per angle grid
  DevBuf gsf_nodes, gsf_table, gsf_alt, gsf_raw;
  std::vector<double> gsf_key;
  bool gsf_alt_valid = false;
  // per-order constants of k_coeff: [n] = ((2n+1)/(n(n+1)), n(n+2)/(n+1)) (mie_props.py:58-64), grown on demand
  DevBuf ntab;
  int ntab_n = 0;
  // multi-GPU exchange (gm_peer.cu): copy-engine transfers run on their own stream, ordered against `stream` by events
  cudaStream_t peer_stream = nullptr;
  cudaEvent_t peer_ev_compute = nullptr, peer_ev_done = nullptr;
  cudaEvent_t peer_marks[4] = {nullptr, nullptr, nullptr, nullptr};   // gm_peer_mark / gm_peer_wait
  bool peer_pending = false;
};

// take a buffer of at least `bytes` for `b` from the handle's pool (smallest fit) or allocate it; give one back (bounded pool)
int gm_pool_take(gm_handle_s* h, DevBuf& b, size_t bytes);
void gm_pool_give(gm_handle_s* h, DevBuf& b);

// GSF expansion of gm_table_run's phase layout on device pointers (gm_gsf.cu); asynchronous on the handle's stream
int gm_gsf_phase4_async(gm_handle_s* h, int ncell, int nang, const double* h_ang_deg, const double* d_P4, int ng, double* d_coef,
                        double* d_cnorm, int quantize10);

// ------------------------------------------------------------------------------------------------ table geometry
// Layout constants shared by the coefficient kernel (producer) and the DMMA contraction kernel (consumer).
constexpr int GM_GROUP = 32;          // particles per group = one warp of the coefficient kernel = DMMA N extent / 2
constexpr int GM_NHALF = 2;           // the angle axis is split in two halves of GM_HALF_ANG angles
constexpr int GM_HALF_ANG = 192;      // angles per CTA of the contraction kernel (12 warps x 16)
constexpr int GM_NANG_PAD = GM_NHALF * GM_HALF_ANG;
constexpr int GM_LAH = 194;           // doubles per table row per half (192 + 2 pad: row stride 388 = 4 mod 16 -> conflict-free A fragments)
constexpr int GM_TROW = 2 * GM_LAH;   // p_n row then q_n row
constexpr int GM_SB = 132;            // doubles per coefficient row: 64 (c+) + 64 (c-) + 4 pad (stride 4 mod 16 -> conflict-free B fragments)
constexpr int GM_KSTEP = 4;           // DMMA k extent
constexpr int GM_BESSEL_SLACK_ROWS = 8;   // rows of 32 doubles in front of and behind the psi and chi tables: k_coeff's ring reads them unchecked
constexpr int GM_STAGE_DBL = GM_KSTEP * GM_TROW + GM_KSTEP * GM_SB;
#ifndef GM_PRODUCER_WARP
#define GM_PRODUCER_WARP 1   // 1: dedicated TMA producer warp (4th warpgroup, setmaxnreg); 0: thread 0 of warp 0 produces
#endif
#ifndef GM_WANT_ITEMS_CFG
#define GM_WANT_ITEMS_CFG 32
#endif
constexpr int GM_STAGES = 8;              // pipeline stages (one k4 step each)
constexpr int GM_CONTRACT_WARPS = 12;     // consumer (MMA) warps
constexpr int GM_CONTRACT_THREADS = (GM_CONTRACT_WARPS + (GM_PRODUCER_WARP ? 4 : 0)) * 32;
constexpr int GM_MAX_CHUNK_GROUPS = 1024;   // group metadata staged in shared memory per contraction CTA
constexpr int GM_CONTRACT_SMEM = GM_STAGES * GM_STAGE_DBL * 8 + 2 * GM_STAGES * 8 + GM_MAX_CHUNK_GROUPS * 8;

// ------------------------------------------------------------------------------------------------ complex helpers
__device__ __forceinline__ double2 cmul(double2 a, double2 b) {
  return make_double2(fma(a.x, b.x, -a.y * b.y), fma(a.x, b.y, a.y * b.x));
}
// reciprocal of a normal, finite double: MUFU.RCP64H seed (~20 bits) + two Newton steps (<= 1 ulp), no slow path
__device__ __forceinline__ double fast_rcp(double x) {
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
  r = fma(r, fma(-x, r, 1.0), r);
  r = fma(r, fma(-x, r, 1.0), r);
  return r;
}
__device__ __forceinline__ double2 crcp(double2 b) {
  double inv = fast_rcp(fma(b.x, b.x, b.y * b.y));
  return make_double2(b.x * inv, -b.y * inv);
}
__device__ __forceinline__ double2 cdiv(double2 a, double2 b) {
  double inv = fast_rcp(fma(b.x, b.x, b.y * b.y));
  return make_double2(fma(a.x, b.x, a.y * b.y) * inv, fma(a.y, b.x, -a.x * b.y) * inv);
}
// Sums 16 per-lane values over the warp with 16 shuffles instead of 80: each round halves the number of values a lane
// carries (the half it gives away goes to its partner), so after four rounds lane L holds the sum over 16 lanes of value
// idx(L) = bits 4,3,2,1 of L, and the last round adds the two halves.  Fixed order: deterministic.  Returns the total of
// value warp_reduce16_index(lane) in every lane.
__device__ __forceinline__ int warp_reduce16_index(int lane) {
  return ((lane >> 4) & 1) * 8 + ((lane >> 3) & 1) * 4 + ((lane >> 2) & 1) * 2 + ((lane >> 1) & 1);
}
__device__ __forceinline__ double warp_reduce16(const double (&v)[16], int lane) {
  double a[8], b[4], c[2];
  const bool u16 = lane & 16, u8 = lane & 8, u4 = lane & 4, u2 = lane & 2;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const double keep = u16 ? v[8 + i] : v[i], give = u16 ? v[i] : v[8 + i];
    a[i] = keep + __shfl_xor_sync(0xffffffffu, give, 16);
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const double keep = u8 ? a[4 + i] : a[i], give = u8 ? a[i] : a[4 + i];
    b[i] = keep + __shfl_xor_sync(0xffffffffu, give, 8);
  }
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const double keep = u4 ? b[2 + i] : b[i], give = u4 ? b[i] : b[2 + i];
    c[i] = keep + __shfl_xor_sync(0xffffffffu, give, 4);
  }
  const double keep = u2 ? c[1] : c[0], give = u2 ? c[0] : c[1];
  double d = keep + __shfl_xor_sync(0xffffffffu, give, 2);
  d += __shfl_xor_sync(0xffffffffu, d, 1);
  return d;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
