// gm_api.cu -- host side of the C ABI declared in include/geosmie_b200.h (lifetime, per-particle Mie, table cells).
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <cmath>

#include "gm_mie_kernels.cuh"
#include "gm_coated.cuh"
#include "gm_psd.cuh"
#include "gm_gram.cuh"
#include "gm_small.cuh"

#ifndef GM_HIO_MIN_BATCHES
#define GM_HIO_MIN_BATCHES 3     // host-buffer calls: batches whose H2D / D2H copies are pipelined against the kernels (optics_SU e2e, equal batches: 2 -> 3.92, 3 -> 3.72, 4 -> 3.87, 6 -> 4.1 ms)
#define GM_HIO_SPLIT {42, 42, 16}   // per cent of the tasks in each batch (optics_SU e2e: 34/33/33 3.74 ms, 40/36/24 3.67, 50/35/15 3.64, 42/42/16 3.60)
#endif
#ifndef GM_GRAM_INTERLEAVE
#define GM_GRAM_INTERLEAVE 1     // k_gram launch order: HBM-bound (class 0/1) work items interleaved with the pipe-bound ones
#endif
#ifndef GM_GRAM_SLOW0
#define GM_GRAM_SLOW0 1.0        // cost weight of class 0 / class 1 work items relative to their DMMA count.  Alone they run at 0.42 / 0.56 of
#define GM_GRAM_SLOW1 1.0        // the DMMA peak against 0.76-0.84 for the others (tools/gram_class_probe.py), but weighting them 1.85 / 1.4
                                 // only made their task ranges shorter: optics_SU k_gram 1.249 ms against 1.197 ms (1.0 / 1.0)
#endif
#ifndef GM_SMALL_SEG0
#define GM_SMALL_SEG0 16         // k_small: particle groups per work item (warp), class 0 (max nmax <= 4) / class 1 (max nmax <= 8)
#define GM_SMALL_SEG1 8
#endif
#ifndef GM_GRAM_ITEMS_PER_SM
#define GM_GRAM_ITEMS_PER_SM 6   // k_gram: work items (class x task range) per SM, launched longest first
#endif
#ifndef GM_EVAL_CTAS_PER_SM
#define GM_EVAL_CTAS_PER_SM 2    // k_gram_eval: CTAs (angle block x task range) per SM
#endif

// ------------------------------------------------------------------------------------------------ errors / lifetime
static thread_local char g_err[512] = "";

void gm_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

extern "C" const char* gm_last_error(void) { return g_err; }
extern "C" int gm_version(void) { return 100; }

extern "C" int gm_init(int device, gm_handle_t* out) {
  GM_REQUIRE(out != nullptr, "out handle pointer is NULL");
  int ndev = 0;
  GM_CUDA_TRY(cudaGetDeviceCount(&ndev));
  GM_REQUIRE(device >= 0 && device < ndev, "device index out of range");
  GM_CUDA_TRY(cudaSetDevice(device));
  cudaDeviceProp prop;
  GM_CUDA_TRY(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10) {
    gm_set_error("libgeosmie_b200 is built for sm_100a only; device %d is sm_%d%d (%s)", device, prop.major, prop.minor, prop.name);
    return GM_ECUDA;
  }
  gm_handle_s* h = new gm_handle_s();
  h->device = device;
  h->sm_count = prop.multiProcessorCount;
  h->l2_window_max = prop.accessPolicyMaxWindowSize;
  h->l2_persist_max = 0;
  if (getenv("GEOSMIE_L2_PERSIST") && atoi(getenv("GEOSMIE_L2_PERSIST")) > 0 && prop.persistingL2CacheMaxSize > 0) {
    // experiment (off by default): a slice of the 126 MB L2 set aside for the Riccati-Bessel tables (see table_run_core)
    const size_t want = std::min<size_t>((size_t)prop.persistingL2CacheMaxSize, (size_t)atoi(getenv("GEOSMIE_L2_PERSIST")) << 20);
    if (cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, want) == cudaSuccess) h->l2_persist_max = (long long)want;
    else cudaGetLastError();
  }
  // opt in to the large dynamic shared memory of the contraction kernels once
  const int smem = GM_CONTRACT_SMEM;
  GM_CUDA_TRY(cudaFuncSetAttribute(k_contract<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  GM_CUDA_TRY(cudaFuncSetAttribute(k_contract<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  GM_CUDA_TRY(cudaFuncSetAttribute(k_gram, cudaFuncAttributeMaxDynamicSharedMemorySize, GM_GRAM_SMEM));
  GM_CUDA_TRY(cudaFuncSetAttribute(k_gram_eval<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, GM_GRAM_EVAL_SMEM_MAX));
  GM_CUDA_TRY(cudaFuncSetAttribute(k_gram_eval<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, GM_GRAM_EVAL_SMEM_MAX));
  GM_CUDA_TRY(cudaFuncSetAttribute(k_gram_eval<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, GM_GRAM_EVAL_SMEM_MAX));
  GM_CUDA_TRY(cudaFuncSetAttribute(k_gram_eval<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, GM_GRAM_EVAL_SMEM_MAX));
  GM_CUDA_TRY(cudaFuncSetAttribute(k_gram_eval<5>, cudaFuncAttributeMaxDynamicSharedMemorySize, GM_GRAM_EVAL_SMEM_MAX));
  GM_CUDA_TRY(cudaFuncSetAttribute(k_gram_eval<6>, cudaFuncAttributeMaxDynamicSharedMemorySize, GM_GRAM_EVAL_SMEM_MAX));
  GM_CUDA_TRY(cudaFuncSetAttribute(k_gram_eval<7>, cudaFuncAttributeMaxDynamicSharedMemorySize, GM_GRAM_EVAL_SMEM_MAX));
  GM_CUDA_TRY(cudaFuncSetAttribute(k_gram_eval<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, GM_GRAM_EVAL_SMEM_MAX));
  // Experiment switch (GEOSMIE_COEFF_CARVEOUT=maxl1): ask for the largest L1 for k_coeff, which uses no shared memory.  Tried
  // against the rank-dependent slow mode of k_coeff in multi-GPU jobs (1.6-1.86 ms instead of 1.03 ms):
  // no effect (4 GPUs, same box: 3.77 ms/step without, 3.75 ms with), so the driver's default stays.
  const char* cv = getenv("GEOSMIE_COEFF_CARVEOUT");
  if (cv && !strcmp(cv, "maxl1")) {
    GM_CUDA_TRY(cudaFuncSetAttribute(k_coeff<0>, cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxL1));
    GM_CUDA_TRY(cudaFuncSetAttribute((k_coeff<0, GM_COEFF_MINB_LONG>), cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxL1));
    GM_CUDA_TRY(cudaFuncSetAttribute(k_coeff<1>, cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxL1));
    GM_CUDA_TRY(cudaFuncSetAttribute(k_coeff<2>, cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxL1));
  }
  *out = h;
  return GM_OK;
}

int gm_pool_take(gm_handle_s* h, DevBuf& b, size_t bytes) {
  if (bytes <= b.cap) return GM_OK;
  int best = -1;
  for (int i = 0; i < (int)h->pool.size(); ++i)
    if (h->pool[i].cap >= bytes && (best < 0 || h->pool[i].cap < h->pool[best].cap)) best = i;      // smallest fit
  if (best < 0) return b.ensure(bytes);
  const DevBuf got = h->pool[best];
  h->pool.erase(h->pool.begin() + best);
  gm_pool_give(h, b);                   // the too-small buffer, if any
  b = got;
  return GM_OK;
}

void gm_pool_give(gm_handle_s* h, DevBuf& b) {
  if (!b.p) return;
  size_t held = 0;
  for (const auto& q : h->pool) held += q.cap;
  if (b.cap <= ((size_t)64 << 20) || h->pool.size() >= 32 || held + b.cap > ((size_t)8 << 30)) {
    b.release();                        // buffers up to 64 MB go to the process-wide cache (gm_common.cuh); the pool stays bounded (32 buffers, 8 GB)
    return;
  }
  h->pool.push_back(b);
  b.p = nullptr;
  b.cap = 0;
}

extern "C" int gm_destroy(gm_handle_t h) {
  if (!h) return GM_OK;
  cudaSetDevice(h->device);
  for (auto& b : h->pool) b.release();
  h->pool.clear();
  struct Flush { ~Flush() { gm_cache_flush(); } } flush_cache_at_return;   // after every buffer of the handle has been released
  for (auto& b : h->ws) b.release();
  for (DevBuf* b : {&h->scratch_coef, &h->scratch_gact, &h->scratch_scal_part, &h->scratch_part, &h->scratch_g_hpart, &h->scratch_g_hsum,
                    &h->scratch_wphase, &h->scratch_wscal, &h->scratch_taskc, &h->scratch_nodepart})
    b->release();
  h->gsf_nodes.release();
  h->gsf_table.release();
  h->gsf_alt.release();
  h->ntab.release();
  h->gsf_raw.release();
  if (h->peer_stream) cudaStreamDestroy(h->peer_stream);
  if (h->peer_ev_compute) cudaEventDestroy(h->peer_ev_compute);
  if (h->peer_ev_done) cudaEventDestroy(h->peer_ev_done);
  for (cudaEvent_t e : h->peer_marks)
    if (e) cudaEventDestroy(e);
  delete h;
  return GM_OK;
}

extern "C" int gm_set_stream(gm_handle_t h, void* s) {
  GM_REQUIRE(h != nullptr, "handle is NULL");
  h->stream = reinterpret_cast<cudaStream_t>(s);
  return GM_OK;
}

extern "C" int gm_sync(gm_handle_t h) {
  GM_REQUIRE(h != nullptr, "handle is NULL");
  GM_CUDA_TRY(cudaSetDevice(h->device));
  GM_CUDA_TRY(cudaStreamSynchronize(h->stream));
  return GM_OK;
}

extern "C" int64_t gm_launch_count(gm_handle_t h) { return h ? h->launches : 0; }

#define GM_LAUNCH_CHECK(h)               \
  do {                                   \
    (h)->launches++;                     \
    GM_CUDA_TRY(cudaGetLastError());     \
  } while (0)

// order-dependent factors of the efficiency sums (mie_props.py:58-64) for every order up to nmaxmax, evaluated once on the host
static int ensure_ntab(gm_handle_s* h, int nmaxmax) {
  if (nmaxmax + 2 <= h->ntab_n) return GM_OK;
  int n = 4096;
  while (n < nmaxmax + 2) n *= 2;
  std::vector<double> v((size_t)2 * n, 0.0);
  for (int k = 1; k < n; ++k) {
    const double dn = (double)k;
    v[2 * (size_t)k] = (2.0 * dn + 1.0) / (dn * (dn + 1.0));
    v[2 * (size_t)k + 1] = dn * (dn + 2.0) / (dn + 1.0);
  }
  GM_CUDA_TRY(cudaStreamSynchronize(h->stream));   // kernels still reading the old table
  int rc = h->ntab.ensure(sizeof(double) * v.size());
  if (rc) return rc;
  GM_CUDA_TRY(cudaMemcpy(h->ntab.p, v.data(), sizeof(double) * v.size(), cudaMemcpyHostToDevice));
  h->ntab_n = n;
  return GM_OK;
}

// ------------------------------------------------------------------------------------------------ grouping
// Host description of how a particle list is cut into groups of 32 (one warp of k_coeff, one N-extent of k_contract).
struct Groups {
  int nx = 0, ngroup = 0, nmaxmax = 0;
  std::vector<long long> gboff;  // Bessel-table offset (doubles) of group g
  std::vector<int> gk4, grow;    // DMMA k4 steps and first coefficient row of group g
  long long bessel_len = 0;      // doubles per Bessel table
  long long task_rows = 0;       // coefficient rows per task
  // Gram path (gm_gram.cuh): groups with max nmax <= 64, listed class by class (class = number of 8-row tiles; 0 = max nmax <= 4)
  std::vector<unsigned char> gskip;   // 1 = Gram group
  std::vector<int> glist;             // Gram groups ordered by class, ascending group index inside a class
  int cls_begin[GM_GRAM_MAX_TG + 2] = {0};   // glist range of class c: [cls_begin[c], cls_begin[c + 1])
  int gram_nmax = 0;                  // largest nmax among Gram groups
  int gram_tgmax = 0;                 // largest class present
  int ndirect = 0;                    // groups left to the per-angle contraction
  static int gram_class(int gm) {
    if (gm <= 4) return 0;      // stacked c+/c- tile (gm_gram.cuh, GramCfg<0>)
    return (gm + 7) / 8;
  }
  void build_gram(const int32_t* nmax) {
    gskip.assign(ngroup, 0);
    glist.clear();
    gram_nmax = gram_tgmax = 0;
    std::vector<int> gm(ngroup, 0);
    for (int g = 0; g < ngroup; ++g)
      for (int i = g * GM_GROUP; i < std::min(nx, (g + 1) * GM_GROUP); ++i) gm[g] = std::max(gm[g], (int)nmax[i]);
    for (int c = 0; c <= GM_GRAM_MAX_TG; ++c) {
      cls_begin[c] = (int)glist.size();
      for (int g = 0; g < ngroup; ++g)
        if (gm[g] <= 8 * GM_GRAM_MAX_TG && gram_class(gm[g]) == c) {
          glist.push_back(g);
          gskip[g] = 1;
          gram_nmax = std::max(gram_nmax, gm[g]);
          gram_tgmax = std::max(gram_tgmax, std::max(c, 1));
        }
    }
    cls_begin[GM_GRAM_MAX_TG + 1] = (int)glist.size();
    ndirect = ngroup - (int)glist.size();
  }
  void build(int n, const int32_t* nmax) {
    nx = n;
    ngroup = (n + GM_GROUP - 1) / GM_GROUP;
    gboff.resize(ngroup);
    gk4.resize(ngroup);
    grow.resize(ngroup);
    bessel_len = 0;
    task_rows = 0;
    nmaxmax = 0;
    for (int g = 0; g < ngroup; ++g) {
      int gm = 0;
      for (int i = g * GM_GROUP; i < std::min(n, (g + 1) * GM_GROUP); ++i) gm = std::max(gm, (int)nmax[i]);
      nmaxmax = std::max(nmaxmax, gm);
      gboff[g] = bessel_len;
      bessel_len += (long long)(gm + 1) * GM_GROUP;
      gk4[g] = (gm + GM_KSTEP - 1) / GM_KSTEP;
      grow[g] = (int)task_rows;
      task_rows += (long long)gk4[g] * GM_KSTEP;
    }
    build_gram(nmax);
  }
};

struct DevGroups {
  DevBuf x, xinv, nmax, gboff, gk4, grow, psichi;
  double* psi_p = nullptr;
  double* chi_p = nullptr;
  size_t psichi_bytes = 0;
  int upload(const Groups& G, const double* hx, const int32_t* hnmax, cudaStream_t st, gm_handle_s* pool = nullptr) {
    int rc;
    if ((rc = x.ensure(sizeof(double) * G.nx))) return rc;
    if ((rc = xinv.ensure(sizeof(double) * G.nx))) return rc;
    if ((rc = nmax.ensure(sizeof(int) * G.nx))) return rc;
    if ((rc = gboff.ensure(sizeof(long long) * G.ngroup))) return rc;
    if ((rc = gk4.ensure(sizeof(int) * G.ngroup))) return rc;
    if ((rc = grow.ensure(sizeof(int) * G.ngroup))) return rc;
    const size_t slack = (size_t)GM_BESSEL_SLACK_ROWS * GM_GROUP;                // doubles of slack around each table (k_coeff's unchecked ring)
    psichi_bytes = sizeof(double) * 2 * ((size_t)G.bessel_len + 2 * slack);
    if ((rc = pool ? gm_pool_take(pool, psichi, psichi_bytes) : psichi.ensure(psichi_bytes))) return rc;   // psi | chi in ONE allocation
    psi_p = psichi.as<double>() + slack;
    chi_p = psi_p + G.bessel_len + 2 * slack;
    GM_CUDA_TRY(cudaMemcpyAsync(x.p, hx, sizeof(double) * G.nx, cudaMemcpyHostToDevice, st));
    GM_CUDA_TRY(cudaMemcpyAsync(nmax.p, hnmax, sizeof(int) * G.nx, cudaMemcpyHostToDevice, st));
    GM_CUDA_TRY(cudaMemcpyAsync(gboff.p, G.gboff.data(), sizeof(long long) * G.ngroup, cudaMemcpyHostToDevice, st));
    GM_CUDA_TRY(cudaMemcpyAsync(gk4.p, G.gk4.data(), sizeof(int) * G.ngroup, cudaMemcpyHostToDevice, st));
    GM_CUDA_TRY(cudaMemcpyAsync(grow.p, G.grow.data(), sizeof(int) * G.ngroup, cudaMemcpyHostToDevice, st));
    GM_CUDA_TRY(cudaMemsetAsync(psichi.p, 0, psichi_bytes, st));
    return GM_OK;
  }
  void release() {
    x.release(); xinv.release(); nmax.release(); gboff.release(); gk4.release(); grow.release(); psichi.release();
  }
};

static int check_particles(int n, const double* x, const int32_t* nmax) {
  GM_REQUIRE(n > 0, "need at least one particle");
  GM_REQUIRE(x && nmax, "x / nmax is NULL");
  for (int i = 0; i < n; ++i) {
    GM_REQUIRE(x[i] > 0.0 && std::isfinite(x[i]), "size parameters must be finite and > 0");
    GM_REQUIRE(nmax[i] >= 1 && nmax[i] < (1 << 20), "nmax out of range");
  }
  return GM_OK;
}

// ------------------------------------------------------------------------------------------------ gm_mie_eval
extern "C" int gm_mie_eval(gm_handle_t h, int n, const double* x, const double* xcore, const double* mz, const double* mrel,
                           int mat_stride, const int32_t* nmax, const int64_t* bes_off, const double* ajv, const double* ayv,
                           int nang, const double* u, double* q, double* s12, double* ab) {
  GM_REQUIRE(h != nullptr, "handle is NULL");
  GM_REQUIRE(mz && mrel, "material arrays are NULL");
  GM_REQUIRE(mat_stride == 0 || mat_stride == 1, "mat_stride must be 0 or 1");
  GM_REQUIRE(nang >= 0 && (nang == 0 || u), "u is NULL");
  GM_REQUIRE(!s12 || nang > 0, "s12 requested without angles");
  int rc = check_particles(n, x, nmax);
  if (rc) return rc;
  GM_CUDA_TRY(cudaSetDevice(h->device));
  cudaStream_t st = h->stream;

  Groups G;
  G.build(n, nmax);
  std::vector<long long> aboff(n + 1, 0);
  for (int i = 0; i < n; ++i) aboff[i + 1] = aboff[i] + nmax[i];
  const long long nab = aboff[n];

  // workspace slots: 0 x,1 nmax,2 gboff,3 psi,4 chi,5 mz,6 mrel,7 aboff,8 ab,9 q,10 u,11 s12,12 jv,13 yv,14 besoff,15 xcore
  DevBuf* W = h->ws;
  const int nmat = mat_stride ? n : 1;
  if ((rc = W[0].ensure(sizeof(double) * n)) || (rc = W[1].ensure(sizeof(int) * n)) ||
      (rc = W[2].ensure(sizeof(long long) * G.ngroup)) || (rc = W[3].ensure(sizeof(double) * G.bessel_len)) ||
      (rc = W[4].ensure(sizeof(double) * G.bessel_len)) || (rc = W[5].ensure(sizeof(double2) * nmat)) ||
      (rc = W[6].ensure(sizeof(double2) * nmat)) || (rc = W[7].ensure(sizeof(long long) * (n + 1))) ||
      (rc = W[8].ensure(sizeof(double4) * nab)) || (rc = W[9].ensure(sizeof(double) * 6 * n)))
    return rc;
  GM_CUDA_TRY(cudaMemcpyAsync(W[0].p, x, sizeof(double) * n, cudaMemcpyHostToDevice, st));
  GM_CUDA_TRY(cudaMemcpyAsync(W[1].p, nmax, sizeof(int) * n, cudaMemcpyHostToDevice, st));
  GM_CUDA_TRY(cudaMemcpyAsync(W[2].p, G.gboff.data(), sizeof(long long) * G.ngroup, cudaMemcpyHostToDevice, st));
  GM_CUDA_TRY(cudaMemcpyAsync(W[5].p, mz, sizeof(double2) * nmat, cudaMemcpyHostToDevice, st));
  GM_CUDA_TRY(cudaMemcpyAsync(W[6].p, mrel, sizeof(double2) * nmat, cudaMemcpyHostToDevice, st));
  GM_CUDA_TRY(cudaMemcpyAsync(W[7].p, aboff.data(), sizeof(long long) * (n + 1), cudaMemcpyHostToDevice, st));

  if (xcore) {
    // coated spheres: coated_mie_coeff, mie_coeffs.py:183-251
    if ((rc = W[15].ensure(sizeof(double) * n))) return rc;
    GM_CUDA_TRY(cudaMemcpyAsync(W[15].p, xcore, sizeof(double) * n, cudaMemcpyHostToDevice, st));
    // per-particle scratch: 3 D arrays + psi/chi at v, w (complex) and y (real)
    std::vector<long long> soff(n + 1, 0);
    for (int i = 0; i < n; ++i) soff[i + 1] = soff[i] + (long long)(nmax[i] + 1);
    if ((rc = W[3].ensure(sizeof(double) * 8 * soff[n]))) return rc;
    if ((rc = W[2].ensure(sizeof(long long) * (n + 1)))) return rc;
    GM_CUDA_TRY(cudaMemcpyAsync(W[2].p, soff.data(), sizeof(long long) * (n + 1), cudaMemcpyHostToDevice, st));
    k_coated_coeff<<<(n + 63) / 64, 64, 0, st>>>(n, W[15].as<double>(), W[0].as<double>(), W[5].as<double2>(), W[6].as<double2>(),
                                                  mat_stride, W[1].as<int>(), W[2].as<long long>(), W[3].as<double>(),
                                                  W[7].as<long long>(), W[8].as<double4>(), nullptr, 0, 0);
    GM_LAUNCH_CHECK(h);
    k_props_nat<<<(n + 127) / 128, 128, 0, st>>>(n, W[0].as<double>(), W[1].as<int>(), W[7].as<long long>(), W[8].as<double4>(),
                                                 W[9].as<double>());
    GM_LAUNCH_CHECK(h);
  } else {
    GM_CUDA_TRY(cudaMemsetAsync(W[3].p, 0, sizeof(double) * G.bessel_len, st));
    GM_CUDA_TRY(cudaMemsetAsync(W[4].p, 0, sizeof(double) * G.bessel_len, st));
    if (ajv && ayv) {
      // the reference's ajv/ayv hold orders k+1.5 (k = 0..nmax-1); psi_0, chi_0 come from sin/cos (mie_coeffs.py:108,112)
      GM_REQUIRE(bes_off != nullptr, "bes_off is NULL");
      std::vector<long long> off(n);
      std::vector<double> jh((size_t)nab + n), yh((size_t)nab + n);
      long long o = 0;
      for (int i = 0; i < n; ++i) {
        off[i] = o;
        jh[o] = 0.0;
        yh[o] = 0.0;
        for (int k = 0; k < nmax[i]; ++k) {
          jh[o + 1 + k] = ajv[bes_off[i] + k];
          yh[o + 1 + k] = ayv[bes_off[i] + k];
        }
        o += nmax[i] + 1;
      }
      if ((rc = W[12].ensure(sizeof(double) * jh.size())) || (rc = W[13].ensure(sizeof(double) * yh.size())) ||
          (rc = W[14].ensure(sizeof(long long) * n)))
        return rc;
      GM_CUDA_TRY(cudaMemcpyAsync(W[12].p, jh.data(), sizeof(double) * jh.size(), cudaMemcpyHostToDevice, st));
      GM_CUDA_TRY(cudaMemcpyAsync(W[13].p, yh.data(), sizeof(double) * yh.size(), cudaMemcpyHostToDevice, st));
      GM_CUDA_TRY(cudaMemcpyAsync(W[14].p, off.data(), sizeof(long long) * n, cudaMemcpyHostToDevice, st));
      k_bessel_from_jy<<<(n + 127) / 128, 128, 0, st>>>(n, W[0].as<double>(), W[1].as<int>(), W[2].as<long long>(),
                                                        W[14].as<long long>(), W[12].as<double>(), W[13].as<double>(),
                                                        W[3].as<double>(), W[4].as<double>());
      GM_LAUNCH_CHECK(h);
      GM_CUDA_TRY(cudaStreamSynchronize(st));  // jh/yh/off are stack-owned
    } else {
      k_bessel<<<(n + 127) / 128, 128, 0, st>>>(n, W[0].as<double>(), W[1].as<int>(), W[2].as<long long>(), W[3].as<double>(),
                                                W[4].as<double>());
      GM_LAUNCH_CHECK(h);
    }
    CoeffArgs A;
    memset(&A, 0, sizeof(A));
    A.nx = n;
    A.ngroup = G.ngroup;
    A.x = W[0].as<double>();
    A.nmax = W[1].as<int>();
    if ((rc = ensure_ntab(h, G.nmaxmax))) return rc;
    A.ntab = h->ntab.as<double2>();
    A.psi = W[3].as<double>();
    A.chi = W[4].as<double>();
    A.gboff = W[2].as<long long>();
    A.mz = W[5].as<double2>();
    A.mrel = W[6].as<double2>();
    A.mat_per_particle = mat_stride;
    A.aboff = W[7].as<long long>();
    A.ab = W[8].as<double4>();
    A.q = W[9].as<double>();
    GM_REQUIRE((G.ngroup + 3) / 4 <= 65535, "more than 8.3 million particles in one gm_mie_eval call: split the call");
    dim3 grid(1, (G.ngroup + 3) / 4);      // x = task, y = quad of groups (see k_coeff)
    A.ntask = 1;
    k_coeff<1><<<grid, 128, 0, st>>>(A);
    GM_LAUNCH_CHECK(h);
  }
  if (s12) {
    if ((rc = W[10].ensure(sizeof(double) * nang)) || (rc = W[11].ensure(sizeof(double) * 4 * (size_t)n * nang))) return rc;
    GM_CUDA_TRY(cudaMemcpyAsync(W[10].p, u, sizeof(double) * nang, cudaMemcpyHostToDevice, st));
    k_s12_direct<<<(n + 3) / 4, 128, 0, st>>>(n, W[1].as<int>(), W[7].as<long long>(), W[8].as<double4>(), nang, W[10].as<double>(),
                                              W[11].as<double>());
    GM_LAUNCH_CHECK(h);
    GM_CUDA_TRY(cudaMemcpyAsync(s12, W[11].p, sizeof(double) * 4 * (size_t)n * nang, cudaMemcpyDeviceToHost, st));
  }
  if (q) GM_CUDA_TRY(cudaMemcpyAsync(q, W[9].p, sizeof(double) * 6 * n, cudaMemcpyDeviceToHost, st));
  if (ab) GM_CUDA_TRY(cudaMemcpyAsync(ab, W[8].p, sizeof(double4) * nab, cudaMemcpyDeviceToHost, st));
  GM_CUDA_TRY(cudaStreamSynchronize(st));
  return GM_OK;
}

// ------------------------------------------------------------------------------------------------ tables
struct gm_table_s {
  gm_handle_t h = nullptr;
  int nx = 0, nang = 0, nrows = 0;
  Groups G;
  DevGroups D;
  std::vector<double> hx;
  std::vector<int32_t> hnmax;
  DevBuf T, cost, dr, psd_par, psd_frac;
  DevBuf g_list, g_skip, g_desc, g_items;   // Gram path (gm_gram.cuh)
  DevBuf norm_planes, norm_ang;              // gm_table_fetch_normalized
  DevBuf T2, W;                              // Gram evaluation through Chebyshev nodes: p/q table at the nodes, interpolation matrix
  int nnode = 0;
  DevBuf s_segs, s_big;                      // fused small-class path (gm_small.cuh): work-item segments, groups left to k_coeff
  DevBuf c_ab, c_scratch, c_soff, c_aboff, c_ratio;   // coated-sphere table path
  // fused GSF stage (gm_table_set_gsf): moments of every finished batch are expanded and downloaded behind the kernels
  std::vector<double> gsf_ang;
  int gsf_ng = 0, gsf_quant = 0;
  double* gsf_coef_host = nullptr;
  double* gsf_cnorm_host = nullptr;
  DevBuf gsf_coef, gsf_cnorm;
  // gm_table_set_mirror: second (peer-memory) destination of k_finalize's results
  double* mirror_scal = nullptr;
  double* mirror_phase = nullptr;
  long long c_nab = 0, c_nscr = 0;
  bool have_dr = false;
  bool psd_separate = false;
  // per-run buffers
  DevBuf chunk_start, mz, mrel, out_scal, out_phase, stats, q, s12;   // the large per-run scratch lives in the handle (shared by its tables)
  double last_stats[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  int timing = 0;
  cudaStream_t h2d_stream = nullptr, d2h_stream = nullptr;   // host-buffer calls: copies pipelined against the kernels
  std::vector<cudaEvent_t> io_events;
  std::vector<cudaEvent_t> evpool;   // pairs of events recorded around every launch of the last run (timing mode)
  std::vector<int> evkind;           // 0 coeff, 1 contract, 2 finalize, 3 gram, 4 gram_eval, 5 small
  size_t evused = 0;
  double ms_coeff = 0, ms_contract = 0, ms_finalize = 0, ms_gram = 0, ms_gram_eval = 0, ms_small = 0;
  int n_coeff = 0, n_contract = 0, n_finalize = 0, n_gram = 0, n_gram_eval = 0, n_small = 0;
};

extern "C" int gm_table_create(gm_handle_t h, int nx, const double* x, const int32_t* nmax, int nang, const double* cos_theta,
                               gm_table_t* out) {
  GM_REQUIRE(h != nullptr && out != nullptr, "handle / out is NULL");
  GM_REQUIRE(nang > 0 && nang <= GM_NANG_PAD, "nang must be in [1, 384]");
  GM_REQUIRE(cos_theta != nullptr, "cos_theta is NULL");
  int rc = check_particles(nx, x, nmax);
  if (rc) return rc;
  GM_REQUIRE(nx <= 65535 * 128, "more than 8.3 million grid points in one table");     // k_coeff: one grid.y entry per 4 groups of 32
  GM_CUDA_TRY(cudaSetDevice(h->device));
  cudaStream_t st = h->stream;
  gm_table_s* t = new gm_table_s();
  t->h = h;
  t->nx = nx;
  t->nang = nang;
  t->hx.assign(x, x + nx);
  t->hnmax.assign(nmax, nmax + nx);
  t->G.build(nx, nmax);
  t->nrows = ((t->G.nmaxmax + GM_KSTEP - 1) / GM_KSTEP) * GM_KSTEP;
  if ((rc = t->D.upload(t->G, x, nmax, st, h)) || (rc = t->cost.ensure(sizeof(double) * nang)) ||
      (rc = gm_pool_take(h, t->T, sizeof(double) * GM_NHALF * (size_t)t->nrows * GM_TROW))) {
    delete t;
    return rc;
  }
  GM_CUDA_TRY(cudaMemcpyAsync(t->cost.p, cos_theta, sizeof(double) * nang, cudaMemcpyHostToDevice, st));
  GM_CUDA_TRY(cudaMemsetAsync(t->T.p, 0, sizeof(double) * GM_NHALF * (size_t)t->nrows * GM_TROW, st));
  if (!t->G.glist.empty()) {
    if ((rc = t->g_list.ensure(sizeof(int) * t->G.glist.size())) || (rc = t->g_skip.ensure(t->G.ngroup))) {
      delete t;
      return rc;
    }
    GM_CUDA_TRY(cudaMemcpyAsync(t->g_list.p, t->G.glist.data(), sizeof(int) * t->G.glist.size(), cudaMemcpyHostToDevice, st));
    GM_CUDA_TRY(cudaMemcpyAsync(t->g_skip.p, t->G.gskip.data(), t->G.ngroup, cudaMemcpyHostToDevice, st));
  }
  k_bessel<<<(nx + 127) / 128, 128, 0, st>>>(nx, t->D.x.as<double>(), t->D.nmax.as<int>(), t->D.gboff.as<long long>(),
                                             t->D.psi_p, t->D.chi_p);
  GM_LAUNCH_CHECK(h);
  k_pt_table<<<(GM_NANG_PAD + 127) / 128, 128, 0, st>>>(nang, t->cost.as<double>(), t->nrows, t->T.as<double>());
  GM_LAUNCH_CHECK(h);
  k_xinv<<<(nx + 127) / 128, 128, 0, st>>>(nx, t->D.x.as<double>(), t->D.xinv.as<double>());
  GM_LAUNCH_CHECK(h);
  std::vector<double> hW;
  DevBuf nodes_d;
  if (!t->G.glist.empty() && getenv("GEOSMIE_EVAL_DIRECT") == nullptr) {
    // Chebyshev nodes u_j = cos(pi j / (M - 1)), M = 2 N + 1 with N = 8 x (largest Gram class): the quadratic forms of k_gram_eval are
    // polynomials of degree <= 2 N in u; W[j][a] = barycentric weight of node j for the table angle a (second-kind weights (-1)^j, halved
    // at the ends; a table angle that coincides with a node takes that node's value)
    const int N = 8 * t->G.gram_tgmax;
    const int M = 2 * N + 1;
    if (M <= GM_HALF_ANG && M < nang) {
      const double pi = 3.14159265358979323846;
      std::vector<double> un(M);
      for (int j = 0; j < M; ++j) un[j] = cos(pi * j / (M - 1));
      un[0] = 1.0;
      un[M - 1] = -1.0;
      hW.assign((size_t)M * GM_NANG_PAD, 0.0);
      for (int a = 0; a < nang; ++a) {
        const double ua = cos_theta[a];
        int hit = -1;
        for (int j = 0; j < M; ++j)
          if (ua == un[j]) hit = j;
        if (hit >= 0) {
          hW[(size_t)hit * GM_NANG_PAD + a] = 1.0;
          continue;
        }
        double sum = 0.0;
        for (int j = 0; j < M; ++j) {
          const double wj = ((j & 1) ? -1.0 : 1.0) * ((j == 0 || j == M - 1) ? 0.5 : 1.0) / (ua - un[j]);
          hW[(size_t)j * GM_NANG_PAD + a] = wj;
          sum += wj;
        }
        for (int j = 0; j < M; ++j) hW[(size_t)j * GM_NANG_PAD + a] /= sum;
      }
      if ((rc = t->T2.ensure(sizeof(double) * GM_NHALF * (size_t)t->nrows * GM_TROW)) || (rc = t->W.ensure(sizeof(double) * hW.size())) ||
          (rc = nodes_d.ensure(sizeof(double) * M))) {
        delete t;
        return rc;
      }
      GM_CUDA_TRY(cudaMemcpyAsync(nodes_d.p, un.data(), sizeof(double) * M, cudaMemcpyHostToDevice, st));
      GM_CUDA_TRY(cudaMemcpyAsync(t->W.p, hW.data(), sizeof(double) * hW.size(), cudaMemcpyHostToDevice, st));
      GM_CUDA_TRY(cudaMemsetAsync(t->T2.p, 0, sizeof(double) * GM_NHALF * (size_t)t->nrows * GM_TROW, st));
      k_pt_table<<<(GM_NANG_PAD + 127) / 128, 128, 0, st>>>(M, nodes_d.as<double>(), t->nrows, t->T2.as<double>());
      GM_LAUNCH_CHECK(h);
      t->nnode = M;
    }
  }
  GM_CUDA_TRY(cudaStreamSynchronize(st));
  nodes_d.release();
  *out = t;
  return GM_OK;
}

extern "C" int gm_table_destroy(gm_table_t t) {
  if (!t) return GM_OK;
  cudaSetDevice(t->h->device);
  for (DevBuf* b : {&t->out_phase, &t->norm_planes, &t->gsf_coef, &t->D.psichi, &t->T}) gm_pool_give(t->h, *b);
  t->D.release();
  for (DevBuf* b : {&t->g_list, &t->g_skip, &t->g_desc, &t->g_items, &t->s_segs, &t->s_big, &t->norm_planes, &t->norm_ang, &t->T2, &t->W, &t->gsf_coef, &t->gsf_cnorm, &t->c_ab, &t->c_scratch, &t->c_soff, &t->c_aboff, &t->c_ratio, &t->dr, &t->psd_par, &t->psd_frac, &t->T, &t->cost, &t->chunk_start, &t->mz, &t->mrel, &t->out_scal, &t->out_phase, &t->stats, &t->q, &t->s12})
    b->release();
  for (auto& e : t->evpool) cudaEventDestroy(e);
  for (auto& e : t->io_events) cudaEventDestroy(e);
  if (t->h2d_stream) cudaStreamDestroy(t->h2d_stream);
  if (t->d2h_stream) cudaStreamDestroy(t->d2h_stream);
  delete t;
  return GM_OK;
}

extern "C" int gm_table_nx(gm_table_t t) { return t ? t->nx : 0; }
extern "C" int gm_table_nang(gm_table_t t) { return t ? t->nang : 0; }

extern "C" int gm_table_set_bessel(gm_table_t t, const int64_t* off, const double* jv_half, const double* yv_half) {
  GM_REQUIRE(t && off && jv_half && yv_half, "NULL argument");
  gm_handle_t h = t->h;
  GM_CUDA_TRY(cudaSetDevice(h->device));
  cudaStream_t st = h->stream;
  size_t tot = 0;
  for (int i = 0; i < t->nx; ++i) tot = std::max(tot, (size_t)off[i] + t->hnmax[i] + 1);
  DevBuf dj, dy, doff;
  int rc;
  if ((rc = dj.ensure(sizeof(double) * tot)) || (rc = dy.ensure(sizeof(double) * tot)) || (rc = doff.ensure(sizeof(long long) * t->nx)))
    return rc;
  GM_CUDA_TRY(cudaMemcpyAsync(dj.p, jv_half, sizeof(double) * tot, cudaMemcpyHostToDevice, st));
  GM_CUDA_TRY(cudaMemcpyAsync(dy.p, yv_half, sizeof(double) * tot, cudaMemcpyHostToDevice, st));
  GM_CUDA_TRY(cudaMemcpyAsync(doff.p, off, sizeof(long long) * t->nx, cudaMemcpyHostToDevice, st));
  k_bessel_from_jy<<<(t->nx + 127) / 128, 128, 0, st>>>(t->nx, t->D.x.as<double>(), t->D.nmax.as<int>(), t->D.gboff.as<long long>(),
                                                        doff.as<long long>(), dj.as<double>(), dy.as<double>(), t->D.psi_p,
                                                        t->D.chi_p);
  GM_LAUNCH_CHECK(h);
  GM_CUDA_TRY(cudaStreamSynchronize(st));
  dj.release(); dy.release(); doff.release();
  return GM_OK;
}

extern "C" int gm_table_set_timing(gm_table_t t, int enable) {
  GM_REQUIRE(t != nullptr, "table is NULL");
  t->timing = enable;
  return GM_OK;
}

// record one event of a (begin, end) pair on the stream; events come from a grow-only pool, nothing synchronises here
static int ev_mark(gm_table_t t, int kind) {
  if (!t->timing) return GM_OK;
  if (t->evused == t->evpool.size()) {
    cudaEvent_t e;
    GM_CUDA_TRY(cudaEventCreate(&e));
    t->evpool.push_back(e);
    t->evkind.push_back(kind);
  }
  t->evkind[t->evused] = kind;
  GM_CUDA_TRY(cudaEventRecord(t->evpool[t->evused++], t->h->stream));
  return GM_OK;
}

// sums the CUDA-event durations of the launches of the last run (synchronises the stream)
extern "C" int gm_table_last_kernel_ms(gm_table_t t, double* a, double* b, double* c) {
  GM_REQUIRE(t != nullptr, "table is NULL");
  if (t->timing && t->evused) {
    GM_CUDA_TRY(cudaSetDevice(t->h->device));
    GM_CUDA_TRY(cudaStreamSynchronize(t->h->stream));
    t->ms_coeff = t->ms_contract = t->ms_finalize = t->ms_gram = t->ms_gram_eval = t->ms_small = 0;
    t->n_coeff = t->n_contract = t->n_finalize = t->n_gram = t->n_gram_eval = t->n_small = 0;
    for (size_t i = 0; i + 1 < t->evused; i += 2) {
      float ms = 0;
      GM_CUDA_TRY(cudaEventElapsedTime(&ms, t->evpool[i], t->evpool[i + 1]));
      if (t->evkind[i] == 0) { t->ms_coeff += ms; t->n_coeff++; }
      if (t->evkind[i] == 1) { t->ms_contract += ms; t->n_contract++; }
      if (t->evkind[i] == 2) { t->ms_finalize += ms; t->n_finalize++; }
      if (t->evkind[i] == 3) { t->ms_gram += ms; t->n_gram++; }
      if (t->evkind[i] == 4) { t->ms_gram_eval += ms; t->n_gram_eval++; }
      if (t->evkind[i] == 5) { t->ms_small += ms; t->n_small++; }
    }
  }
  if (a) *a = t->ms_coeff;
  if (b) *b = t->ms_contract + t->ms_gram + t->ms_gram_eval + t->ms_small;   // the whole angular stage (k_small: fused with its coefficients)
  if (c) *c = t->ms_finalize;
  return GM_OK;
}

// per-kernel split of the last run: ms[0..5] = k_coeff, k_contract, k_finalize, k_gram, k_gram_sum + k_gram_eval, k_small;
// n[0..4] = event-bracketed launch groups of each
extern "C" int gm_table_last_kernel_ms_ex(gm_table_t t, double ms[8], int32_t n[8]) {
  int rc = gm_table_last_kernel_ms(t, nullptr, nullptr, nullptr);
  if (rc) return rc;
  if (ms) {
    ms[0] = t->ms_coeff; ms[1] = t->ms_contract; ms[2] = t->ms_finalize; ms[3] = t->ms_gram; ms[4] = t->ms_gram_eval;
    ms[5] = t->ms_small;
    ms[6] = ms[7] = 0;
  }
  if (n) {
    n[0] = t->n_coeff; n[1] = t->n_contract; n[2] = t->n_finalize; n[3] = t->n_gram; n[4] = t->n_gram_eval;
    n[5] = t->n_small;
    n[6] = n[7] = 0;
  }
  return GM_OK;
}

static int fetch_stats(gm_table_t t);

extern "C" int gm_table_last_stats(gm_table_t t, double stats[8]) {
  GM_REQUIRE(t && stats, "NULL argument");
  if (t->stats.p) {
    GM_CUDA_TRY(cudaSetDevice(t->h->device));
    int rc = fetch_stats(t);
    if (rc) return rc;
  }
  for (int i = 0; i < 8; ++i) stats[i] = t->last_stats[i];
  return GM_OK;
}

// host buffers of a gm_table_run call whose transfers are pipelined batch by batch against the kernels
struct HostIO {
  const double* w_phase;
  const double* w_scal;
  double* out_scal;
  double* out_phase;
};

// Core of gm_table_run: all pointers are DEVICE pointers.  per_particle != 0 selects the S1/S2-per-particle epilogue.
// With `hio`, the weights of batch b are uploaded on a copy stream while batch b-1 computes, and the results of batch b
// are downloaded while batch b+1 computes.
static int table_run_core(gm_table_t t, int ntask, const double* d_mz, const double* d_mrel, int nmode, const double* d_wphase,
                          const double* d_wscal, int flags, double* d_out_scal, double* d_out_phase, double* d_q, double* d_s12,
                          bool per_particle, const HostIO* hio = nullptr, const double* d_core_ratio = nullptr) {
  gm_handle_t h = t->h;
  cudaStream_t st = h->stream;
  const Groups& G = t->G;
  int rc;
  const long long task_stride = G.task_rows * GM_SB;  // doubles
  const size_t per_task_bytes = (size_t)task_stride * 8;
  int tb = (int)std::max<size_t>(1, h->coef_budget_bytes / std::max<size_t>(per_task_bytes, 1));
  tb = std::min(tb, ntask);
  tb = std::min(tb, 32768);  // tasks per launch
  if (d_core_ratio) {
    // coated spheres: natural-layout a_n, b_n + recurrence scratch per (task, particle) are staged in HBM
    const size_t per_task = (size_t)t->c_nab * sizeof(double4) + (size_t)t->c_nscr * 8 * sizeof(double);
    tb = std::min(tb, (int)std::max<size_t>(1, ((size_t)1 << 30) / std::max<size_t>(per_task, 1)));
    if ((rc = t->c_ab.ensure((size_t)tb * t->c_nab * sizeof(double4))) || (rc = t->c_scratch.ensure((size_t)tb * t->c_nscr * 8 * sizeof(double))))
      return rc;
  }
  // Batch boundaries.  Device-pointer calls: as few batches as the scratch budget allows.  Host-buffer calls: at least
  // GM_HIO_MIN_BATCHES batches so that the copies overlap the kernels, of DECREASING size -- the download of the last batch is the
  // one transfer nothing can hide, so that batch is the smallest (GM_HIO_SPLIT, per cent of the tasks).
  std::vector<int> bstart(1, 0);
  if (hio && ntask >= 256) {
    static const int split[GM_HIO_MIN_BATCHES] = GM_HIO_SPLIT;
    const char* env = getenv("GEOSMIE_HIO_SPLIT");   // experiments: "45,35,20"
    int pct[GM_HIO_MIN_BATCHES], sum = 0;
    for (int b = 0; b < GM_HIO_MIN_BATCHES; ++b) pct[b] = split[b];
    if (env) {
      int k = 0;
      for (const char* p = env; *p && k < GM_HIO_MIN_BATCHES; ++k) {
        pct[k] = atoi(p);
        while (*p && *p != ',') ++p;
        if (*p == ',') ++p;
      }
    }
    for (int b = 0; b < GM_HIO_MIN_BATCHES; ++b) sum += std::max(pct[b], 1);
    int acc = 0;
    for (int b = 0; b + 1 < GM_HIO_MIN_BATCHES; ++b) {
      acc += std::max(pct[b], 1);
      const int e = (int)((long long)ntask * acc / sum);
      if (e > bstart.back() && e < ntask) bstart.push_back(e);
    }
    bstart.push_back(ntask);
    // a part larger than the scratch budget is cut further
    std::vector<int> cut(1, 0);
    for (size_t b = 0; b + 1 < bstart.size(); ++b)
      for (int t0 = bstart[b]; t0 < bstart[b + 1]; t0 += tb) cut.push_back(std::min(bstart[b + 1], t0 + tb));
    bstart.swap(cut);
    int largest = 0;
    for (size_t b = 0; b + 1 < bstart.size(); ++b) largest = std::max(largest, bstart[b + 1] - bstart[b]);
    tb = largest;
  } else {
    for (int t0 = tb; t0 < ntask; t0 += tb) bstart.push_back(t0);
    bstart.push_back(ntask);
  }
  const bool use_gram = !per_particle && !(flags & GM_F_NO_GRAM) && !G.glist.empty();
  // fused coefficient + Gram kernel for the groups with max nmax <= 8 (gm_small.cuh): homogeneous spheres, one PSD mode
  static const bool small_off = getenv("GEOSMIE_NO_SMALL") != nullptr;     // diagnostics: the round-1 path (k_coeff + k_gram) for every class
  const bool use_small = use_gram && nmode == 1 && !d_core_ratio && !d_q && !small_off && G.cls_begin[2] > 0;   // (k_small has no per-particle q output)
  std::vector<SmallSeg> small_segs;
  const int ndirect = use_gram ? G.ndirect : G.ngroup;
  // chunks of the per-angle contraction: enough CTAs to fill the machine ~8x over, cost-balanced by the k4 steps of the
  // groups it handles (groups taken by the Gram path cost nothing here); a chunk never holds more groups than the smem metadata
  std::vector<int> cstart(1, 0);
  int nchunk = 0;
  if (ndirect > 0) {
    const int want_items = GM_WANT_ITEMS_CFG * h->sm_count;   // equal-cost CTAs per SM: bounds the last-wave tail
    int want = (want_items + 2 * tb - 1) / (2 * tb);
    want = std::max(1, std::min(want, ndirect));
    long long tot = 0;
    for (int g = 0; g < G.ngroup; ++g) tot += (use_gram && G.gskip[g]) ? 0 : G.gk4[g];
    long long acc = 0;
    int made = 1;
    for (int g = 0; g + 1 < G.ngroup; ++g) {
      acc += (use_gram && G.gskip[g]) ? 0 : G.gk4[g];
      const bool cut = made < want && acc * want >= tot * made;
      if (cut || g + 1 - cstart.back() >= GM_MAX_CHUNK_GROUPS) {
        cstart.push_back(g + 1);
        if (cut) ++made;
      }
    }
    cstart.push_back(G.ngroup);
    nchunk = (int)cstart.size() - 1;
  }
  const int nchunk_total = nchunk + (use_gram ? 1 : 0);
  const int nbatch = (int)bstart.size() - 1;
  // Gram plan(s): descriptors (class, group range, partial slot) and CTA work items (descriptor, task range) per batch size
  struct GramPlan {
    int nt = 0, desc0 = 0, ndesc = 0, item0 = 0, nitem = 0;
    long long hstride = 0;
  };
  std::vector<GramPlan> plans;
  std::vector<GramDesc> all_desc;
  std::vector<GramItem> all_items;
  if (use_gram) {
    for (int b = 0; b < nbatch; ++b) {
      const int nt = bstart[b + 1] - bstart[b];
      bool have = false;
      for (auto& q : plans) have |= q.nt == nt;
      if (have) continue;
      GramPlan P;
      P.nt = nt;
      P.desc0 = (int)all_desc.size();
      P.item0 = (int)all_items.size();
      double ccost[GM_GRAM_MAX_TG + 1] = {0};
      double total = 0;
      auto dmma_per_group = [](int c) {   // DMMAs a team issues per group (16 k-steps)
        if (c == 0) return 32.0;
        if (c <= 4) return 64.0 * c * c;
        return 16.0 * 12.0 * c * ((c + 2) / 3);   // 12 warps x NI x CW tiles (ragged last third recomputes a tile)
      };
      // Classes 0 and 1 read 132 B of coefficient stream per DMMA and are bound by HBM, not by the tensor pipe (measured alone:
      // 0.42 / 0.56 of the DMMA peak at 4-5.4 TB/s, the others 0.76-0.84): they must not run all at the same time (see the
      // interleaving below).
      static const double slow[2] = {GM_GRAM_SLOW0, GM_GRAM_SLOW1};
      for (int c = 0; c <= GM_GRAM_MAX_TG; ++c) {
        const int nc = G.cls_begin[c + 1] - G.cls_begin[c];
        ccost[c] = (double)nc * (dmma_per_group(c) + 48.0) * (c <= 1 ? slow[c] : 1.0);   // DMMA issue slots per task (+ ring handling per group)
        if (use_small && c <= 1) continue;   // not k_gram's work
        total += ccost[c] * nt;
      }
      const double target = std::max(total / (h->sm_count * (double)GM_GRAM_ITEMS_PER_SM), 12000.0);
      std::vector<std::pair<double, GramItem>> items;
      for (int c = 0; c <= GM_GRAM_MAX_TG; ++c) {
        const int nc = G.cls_begin[c + 1] - G.cls_begin[c];
        if (nc == 0) continue;
        if (use_small && c <= 1) {
          // classes 0 / 1 are evaluated AND reduced by k_small: one descriptor whose "teams" are the segments (work items) of the
          // class; no k_gram work item.  The segment list does not depend on the batch size (classes 0, 1 come first in hstride).
          const int per = c == 0 ? GM_SMALL_SEG0 : GM_SMALL_SEG1;
          const int nseg = std::max(1, std::min(GM_SMALL_MAXSEG, (nc + per - 1) / per));
          GramDesc d;
          d.tg = c;
          d.nteam = nseg;
          d.gbegin = G.cls_begin[c];
          d.gend = G.cls_begin[c + 1];
          d.hoff = P.hstride;
          if (plans.empty())
            for (int k = 0; k < nseg; ++k) {
              SmallSeg sgm;
              sgm.cls = c;
              sgm.gbegin = G.cls_begin[c] + (int)((long long)nc * k / nseg);
              sgm.gend = G.cls_begin[c] + (int)((long long)nc * (k + 1) / nseg);
              sgm.seg = k;
              sgm.hoff = d.hoff;
              small_segs.push_back(sgm);
            }
          P.hstride += (long long)nseg * 4 * (c == 0 ? 16 : 64);
          all_desc.push_back(d);
          continue;
        }
        const int nteam = c <= 1 ? 12 : c == 2 ? 6 : c <= 4 ? 3 : 1;   // GramCfg<c>::NTEAM
        int nsplit = 1, tpc = 1;
        if (ccost[c] > 1.5 * target) nsplit = std::min(nc, (int)std::lround(ccost[c] / target));
        else tpc = std::max(1, std::min(nt, (int)(target / ccost[c])));
        tpc = (nt + ((nt + tpc - 1) / tpc) - 1) / ((nt + tpc - 1) / tpc);   // even task ranges
        tpc = std::min(nt, (tpc + nteam - 1) / nteam * nteam);              // whole tasks per team: a multiple of the team count
        for (int k = 0; k < nsplit; ++k) {
          GramDesc d;
          d.tg = c;
          d.nteam = 1;                                                      // a team owns whole tasks: one partial per (task, descriptor)
          d.gbegin = G.cls_begin[c] + (int)((long long)nc * k / nsplit);
          d.gend = G.cls_begin[c] + (int)((long long)nc * (k + 1) / nsplit);
          d.hoff = P.hstride;
          P.hstride += (long long)4 * (c == 0 ? 16 : 64 * c * c);           // GramCfg<c>::ND squared per block
          const int di = (int)all_desc.size() - P.desc0;
          all_desc.push_back(d);
          for (int t0 = 0; t0 < nt; t0 += tpc) {
            GramItem it = {di, t0, std::min(nt, t0 + tpc)};
            items.push_back({ccost[c] / nsplit * (it.t1 - it.t0), it});
          }
        }
      }
      // Launch order (the hardware hands out CTAs in index order as SMs become free): longest first inside each of two queues,
      // the HBM-bound items (classes 0, 1) and the pipe-bound ones, merged so that both queues advance at the same relative
      // pace -- at any moment only a fraction of the SMs streams small-class groups and the rest keeps the DMMA pipes busy.
      std::vector<std::pair<double, GramItem>> q[2];
      double qsum[2] = {0, 0};
      for (auto& e : items) {
        const int c = all_desc[P.desc0 + e.second.desc].tg;
        const int w = (GM_GRAM_INTERLEAVE && c <= 1) ? 1 : 0;
        q[w].push_back(e);
        qsum[w] += e.first;
      }
      for (int w = 0; w < 2; ++w)
        std::stable_sort(q[w].begin(), q[w].end(), [](const auto& a, const auto& b) { return a.first > b.first; });
      size_t qi[2] = {0, 0};
      double qdone[2] = {0, 0};
      while (qi[0] < q[0].size() || qi[1] < q[1].size()) {
        int w;
        if (qi[0] >= q[0].size()) w = 1;
        else if (qi[1] >= q[1].size()) w = 0;
        else w = (qdone[1] / qsum[1] < qdone[0] / qsum[0]) ? 1 : 0;
        qdone[w] += q[w][qi[w]].first;
        all_items.push_back(q[w][qi[w]++].second);
      }
      P.ndesc = (int)all_desc.size() - P.desc0;
      P.nitem = (int)all_items.size() - P.item0;
      plans.push_back(P);
    }
    long long hmax = 0;
    for (auto& P : plans) hmax = std::max(hmax, P.hstride * P.nt);
    if ((rc = t->g_desc.ensure(sizeof(GramDesc) * all_desc.size())) || (rc = t->g_items.ensure(sizeof(GramItem) * all_items.size())) ||
        (rc = t->h->scratch_g_hpart.ensure(sizeof(double) * (size_t)hmax)) ||
        (rc = t->h->scratch_g_hsum.ensure(sizeof(double) * (size_t)tb * 4 * 64 * G.gram_tgmax * G.gram_tgmax)))
      return rc;
    GM_CUDA_TRY(cudaMemcpyAsync(t->g_desc.p, all_desc.data(), sizeof(GramDesc) * all_desc.size(), cudaMemcpyHostToDevice, st));
    GM_CUDA_TRY(cudaMemcpyAsync(t->g_items.p, all_items.data(), sizeof(GramItem) * all_items.size(), cudaMemcpyHostToDevice, st));
  }
  std::vector<int> big_groups;
  if (use_small) {
    std::vector<unsigned char> is_small(G.ngroup, 0);
    for (int k = G.cls_begin[0]; k < G.cls_begin[2]; ++k) is_small[G.glist[k]] = 1;
    for (int g = 0; g < G.ngroup; ++g)
      if (!is_small[g]) big_groups.push_back(g);
    if ((rc = t->s_segs.ensure(sizeof(SmallSeg) * small_segs.size())) || (rc = t->s_big.ensure(sizeof(int) * std::max<size_t>(1, big_groups.size()))) ||
        (rc = t->h->scratch_taskc.ensure(sizeof(double2) * 2 * (size_t)tb)))
      return rc;
    GM_CUDA_TRY(cudaMemcpyAsync(t->s_segs.p, small_segs.data(), sizeof(SmallSeg) * small_segs.size(), cudaMemcpyHostToDevice, st));
    if (!big_groups.empty())
      GM_CUDA_TRY(cudaMemcpyAsync(t->s_big.p, big_groups.data(), sizeof(int) * big_groups.size(), cudaMemcpyHostToDevice, st));
  }
  // k_coeff instantiation: the 168-register build for launches of long groups (gm_mie_kernels.cuh)
  bool coeff_long = false;
  {
    long long rows = 0, cnt = 0;
    if (use_small) {
      for (int g : big_groups) rows += 4LL * G.gk4[g], ++cnt;
    } else {
      for (int g = 0; g < G.ngroup; ++g) rows += 4LL * G.gk4[g], ++cnt;
    }
    coeff_long = cnt > 0 && rows >= (long long)GM_COEFF_LONG_ROWS * cnt;
    static const char* force = getenv("GEOSMIE_COEFF_LONG");     // experiments: 0 / 1 forces the choice
    if (force) coeff_long = atoi(force) != 0;
  }
  if ((rc = t->h->scratch_coef.ensure(per_task_bytes * tb)) || (rc = t->h->scratch_gact.ensure((size_t)tb * G.ngroup)) ||
      (rc = t->h->scratch_scal_part.ensure(sizeof(double) * (size_t)tb * nmode * G.ngroup * GM_NSCAL)) ||
      (rc = t->h->scratch_part.ensure(sizeof(double) * (size_t)tb * nchunk_total * 4 * GM_NANG_PAD)) ||
      (rc = (use_gram && t->nnode > 0) ? t->h->scratch_nodepart.ensure(sizeof(double) * (size_t)tb * 4 * GM_NANG_PAD) : GM_OK) ||
      (rc = t->chunk_start.ensure(sizeof(int) * (nchunk + 1))) || (rc = t->stats.ensure(sizeof(unsigned long long) * 8)))
    return rc;
  if (t->gsf_ng > 0 && !per_particle) {
    if ((rc = gm_pool_take(t->h, t->gsf_coef, sizeof(double) * (size_t)ntask * 6 * t->gsf_ng)) || (rc = t->gsf_cnorm.ensure(sizeof(double) * ntask))) return rc;
  }
  GM_CUDA_TRY(cudaMemcpyAsync(t->chunk_start.p, cstart.data(), sizeof(int) * (nchunk + 1), cudaMemcpyHostToDevice, st));
  GM_CUDA_TRY(cudaMemsetAsync(t->stats.p, 0, sizeof(unsigned long long) * 8, st));
  const int smem = GM_CONTRACT_SMEM;
  for (int c = 0; c < nchunk; ++c)
    GM_REQUIRE(cstart[c + 1] - cstart[c] <= GM_MAX_CHUNK_GROUPS, "chunk has more particle groups than the smem metadata holds");
  if ((rc = ensure_ntab(h, G.nmaxmax))) return rc;
  const int64_t launches0 = h->launches;
  t->evused = 0;
  // Experiment (GEOSMIE_L2_PERSIST=<MB>, off by default): Riccati-Bessel tables pinned in L2 for the duration of the call -- k_coeff
  // reads one psi / chi row per order on its serial path while it streams gigabytes of coefficient rows through the same L2
  static const int l2_persist = getenv("GEOSMIE_L2_PERSIST") ? atoi(getenv("GEOSMIE_L2_PERSIST")) : 0;
  bool window_set = false;
  if (l2_persist && t->D.psichi_bytes > 0 && h->l2_persist_max > 0) {
    cudaStreamAttrValue av;
    memset(&av, 0, sizeof(av));
    av.accessPolicyWindow.base_ptr = t->D.psichi.p;
    av.accessPolicyWindow.num_bytes = std::min(t->D.psichi_bytes, (size_t)h->l2_window_max);
    av.accessPolicyWindow.hitRatio = (float)std::min(1.0, (double)h->l2_persist_max / (double)av.accessPolicyWindow.num_bytes);
    av.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
    av.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
    window_set = cudaStreamSetAttribute(st, cudaStreamAttributeAccessPolicyWindow, &av) == cudaSuccess;
    if (!window_set) cudaGetLastError();
  }
  if (hio) {
    if (!t->h2d_stream) {
      GM_CUDA_TRY(cudaStreamCreateWithFlags(&t->h2d_stream, cudaStreamNonBlocking));
      GM_CUDA_TRY(cudaStreamCreateWithFlags(&t->d2h_stream, cudaStreamNonBlocking));
    }
    while ((int)t->io_events.size() < 2 * nbatch + 1) {
      cudaEvent_t e;
      GM_CUDA_TRY(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
      t->io_events.push_back(e);
    }
    // the copy streams must not run ahead of work already queued on the compute stream (buffers are reused between calls)
    GM_CUDA_TRY(cudaEventRecord(t->io_events[2 * nbatch], st));
    GM_CUDA_TRY(cudaStreamWaitEvent(t->h2d_stream, t->io_events[2 * nbatch], 0));
    GM_CUDA_TRY(cudaStreamWaitEvent(t->d2h_stream, t->io_events[2 * nbatch], 0));
    for (int b = 0; b < nbatch; ++b) {
      const int t0 = bstart[b], nt = bstart[b + 1] - t0;
      if (hio->w_phase)
        GM_CUDA_TRY(cudaMemcpyAsync(const_cast<double*>(d_wphase) + (size_t)t0 * G.nx, hio->w_phase + (size_t)t0 * G.nx,
                                    sizeof(double) * (size_t)nt * G.nx, cudaMemcpyHostToDevice, t->h2d_stream));
      if (hio->w_scal)
        GM_CUDA_TRY(cudaMemcpyAsync(const_cast<double*>(d_wscal) + (size_t)t0 * nmode * G.nx, hio->w_scal + (size_t)t0 * nmode * G.nx,
                                    sizeof(double) * (size_t)nt * nmode * G.nx, cudaMemcpyHostToDevice, t->h2d_stream));
      GM_CUDA_TRY(cudaEventRecord(t->io_events[b], t->h2d_stream));
    }
  }

  for (int bi = 0; bi < nbatch; ++bi) {
    const int t0 = bstart[bi], nt = bstart[bi + 1] - t0;
    if (hio) GM_CUDA_TRY(cudaStreamWaitEvent(st, t->io_events[bi], 0));
    CoeffArgs A;
    memset(&A, 0, sizeof(A));
    A.nx = G.nx;
    A.ngroup = G.ngroup;
    A.x = t->D.x.as<double>();
    A.nmax = t->D.nmax.as<int>();
    A.ntab = h->ntab.as<double2>();
    A.psi = t->D.psi_p;
    A.chi = t->D.chi_p;
    A.gboff = t->D.gboff.as<long long>();
    A.mz = reinterpret_cast<const double2*>(d_mz) + t0;
    A.mrel = reinterpret_cast<const double2*>(d_mrel) + t0;
    A.mat_per_particle = 0;
    A.wphase = d_wphase ? d_wphase + (size_t)t0 * G.nx : nullptr;
    A.wscal = d_wscal ? d_wscal + (size_t)t0 * nmode * G.nx : nullptr;
    A.nmode = nmode;
    A.ntask = nt;
    A.dense = (flags & GM_F_ELIDE_ZERO_WEIGHT) ? 0 : 1;
    A.scale_sqrtw = per_particle ? 0 : 1;
    A.grow = t->D.grow.as<int>();
    A.gk4 = t->D.gk4.as<int>();
    A.coef = t->h->scratch_coef.as<double>();
    A.task_stride = task_stride;
    A.gact = t->h->scratch_gact.as<unsigned char>();
    A.scal_part = t->h->scratch_scal_part.as<double>();
    A.q = d_q ? d_q + (size_t)t0 * G.nx * 6 : nullptr;
#ifdef GM_NO_STATS
    A.stats = nullptr;
#else
    A.stats = t->stats.as<unsigned long long>();
#endif
    const GramPlan* P = nullptr;
    for (auto& q : plans)
      if (q.nt == nt) P = &q;
    if ((rc = ev_mark(t, 0))) return rc;
    if (d_core_ratio) {
      // mz carries the core index m1 = sqrt(eps1), mrel the shell index m2 = sqrt(eps2) (gm_mie_eval convention)
      k_coated_coeff<<<dim3((G.nx + 63) / 64, nt), 64, 0, st>>>(G.nx, nullptr, t->D.x.as<double>(), A.mz, A.mrel, 2, t->D.nmax.as<int>(),
                                                                 t->c_soff.as<long long>(), t->c_scratch.as<double>(),
                                                                 t->c_aboff.as<long long>(), t->c_ab.as<double4>(), d_core_ratio + t0,
                                                                 t->c_nscr * 8, t->c_nab);
      GM_LAUNCH_CHECK(h);
      A.aboff = t->c_aboff.as<long long>();
      A.ab = t->c_ab.as<double4>();
      A.ab_stride = t->c_nab;
      k_coeff<2><<<dim3((nt + GM_COEFF_TPC - 1) / GM_COEFF_TPC, (G.ngroup + 3) / 4), 128, 0, st>>>(A);
    } else if (use_small) {
      A.gsel = t->s_big.as<int>();
      A.nsel = (int)big_groups.size();
      if (A.nsel > 0) {
        if (coeff_long)
          k_coeff<0, GM_COEFF_MINB_LONG><<<dim3((nt + GM_COEFF_TPC - 1) / GM_COEFF_TPC, (A.nsel + 3) / 4), 128, 0, st>>>(A);
        else
          k_coeff<0><<<dim3((nt + GM_COEFF_TPC - 1) / GM_COEFF_TPC, (A.nsel + 3) / 4), 128, 0, st>>>(A);
      }
    } else {
      if (coeff_long)
        k_coeff<0, GM_COEFF_MINB_LONG><<<dim3((nt + GM_COEFF_TPC - 1) / GM_COEFF_TPC, (G.ngroup + 3) / 4), 128, 0, st>>>(A);
      else
        k_coeff<0><<<dim3((nt + GM_COEFF_TPC - 1) / GM_COEFF_TPC, (G.ngroup + 3) / 4), 128, 0, st>>>(A);
    }
    GM_LAUNCH_CHECK(h);
    if ((rc = ev_mark(t, 0))) return rc;
    if (use_small) {
      SmallArgs SA;
      memset(&SA, 0, sizeof(SA));
      SA.nx = G.nx;
      SA.ngroup = G.ngroup;
      SA.ntask = nt;
      SA.nseg = (int)small_segs.size();
      SA.segs = t->s_segs.as<SmallSeg>();
      SA.glist = t->g_list.as<int>();
      SA.x = A.x;
      SA.xinv = t->D.xinv.as<double>();
      SA.nmax = A.nmax;
      SA.psi = A.psi;
      SA.chi = A.chi;
      SA.gboff = A.gboff;
      SA.mz = A.mz;
      SA.mrel = A.mrel;
      SA.mzinv = t->h->scratch_taskc.as<double2>();
      SA.mrinv = SA.mzinv + tb;
      SA.wphase = A.wphase;
      SA.wscal = A.wscal;
      SA.dense = A.dense;
      SA.ntab = A.ntab;
      SA.hpart = t->h->scratch_g_hpart.as<double>();
      SA.hstride = P->hstride;
      SA.scal_part = A.scal_part;
      SA.stats = A.stats;
      if ((rc = ev_mark(t, 5))) return rc;
      k_task_prep<<<(nt + 127) / 128, 128, 0, st>>>(nt, SA.mz, SA.mrel, const_cast<double2*>(SA.mzinv), const_cast<double2*>(SA.mrinv));
      GM_LAUNCH_CHECK(h);
      int nseg0 = 0;                                 // segments are listed class 0 first
      for (const SmallSeg& sgm : small_segs) nseg0 += sgm.cls == 0 ? 1 : 0;
      const int gx = (nt + GM_SMALL_WARPS - 1) / GM_SMALL_WARPS;
      if (nseg0 > 0) {
        k_small<4, GM_SMALL_MINB4><<<dim3(gx, nseg0), GM_SMALL_WARPS * 32, 0, st>>>(SA);
        GM_LAUNCH_CHECK(h);
      }
      if (SA.nseg > nseg0) {
        SA.segs += nseg0;
        k_small<8, GM_SMALL_MINB8><<<dim3(gx, SA.nseg - nseg0), GM_SMALL_WARPS * 32, 0, st>>>(SA);
        GM_LAUNCH_CHECK(h);
      }
      if ((rc = ev_mark(t, 5))) return rc;
    }

    ContractArgs C;
    memset(&C, 0, sizeof(C));
    C.ntask = nt;
    C.ngroup = G.ngroup;
    C.nchunk = nchunk;
    C.nrows = t->nrows;
    C.T = t->T.as<double>();
    C.coef = t->h->scratch_coef.as<double>();
    C.task_stride = task_stride;
    C.grow = t->D.grow.as<int>();
    C.gk4 = t->D.gk4.as<int>();
    C.gact = t->h->scratch_gact.as<unsigned char>();
    C.gskip = use_gram ? t->g_skip.as<unsigned char>() : nullptr;
    C.chunk_start = t->chunk_start.as<int>();
    C.part = t->h->scratch_part.as<double>();
    C.nchunk_total = nchunk_total;
    C.nx = G.nx;
    C.nang = t->nang;
    C.s12 = d_s12 ? d_s12 + (size_t)t0 * G.nx * t->nang * 4 : nullptr;
    if (nchunk > 0) {
      if ((rc = ev_mark(t, 1))) return rc;
      if (per_particle)
        k_contract<true><<<nt * 2 * nchunk, GM_CONTRACT_THREADS, smem, st>>>(C);
      else
        k_contract<false><<<nt * 2 * nchunk, GM_CONTRACT_THREADS, smem, st>>>(C);
      GM_LAUNCH_CHECK(h);
      if ((rc = ev_mark(t, 1))) return rc;
    }
    if (use_gram) {
      GramArgs GA;
      memset(&GA, 0, sizeof(GA));
      GA.ngroup = G.ngroup;
      GA.items = t->g_items.as<GramItem>() + P->item0;
      GA.desc = t->g_desc.as<GramDesc>() + P->desc0;
      GA.glist = t->g_list.as<int>();
      GA.grow = t->D.grow.as<int>();
      GA.gk4 = t->D.gk4.as<int>();
      GA.gact = t->h->scratch_gact.as<unsigned char>();
      GA.coef = t->h->scratch_coef.as<double>();
      GA.task_stride = task_stride;
      GA.hpart = t->h->scratch_g_hpart.as<double>();
      GA.hstride = P->hstride;
      if (P->nitem > 0) {
        if ((rc = ev_mark(t, 3))) return rc;
        k_gram<<<P->nitem, GM_GRAM_THREADS, GM_GRAM_SMEM, st>>>(GA);
        GM_LAUNCH_CHECK(h);
        if ((rc = ev_mark(t, 3))) return rc;
      }
      const int N = 8 * G.gram_tgmax;
      GramSumArgs SA;
      memset(&SA, 0, sizeof(SA));
      SA.ndesc = P->ndesc;
      GM_REQUIRE(P->ndesc <= GM_GRAM_SUM_MAXDESC, "too many Gram descriptors");
      for (int k = 0; k < P->ndesc; ++k) SA.desc[k] = all_desc[P->desc0 + k];
      SA.hpart = GA.hpart;
      SA.hstride = P->hstride;
      SA.N = N;
      SA.hsum = t->h->scratch_g_hsum.as<double>();
      GramEvalArgs EA;
      memset(&EA, 0, sizeof(EA));
      EA.ntask = nt;
      EA.tasks_per_cta = std::max(1, (nt + (GM_EVAL_CTAS_PER_SM * h->sm_count / 4) - 1) / (GM_EVAL_CTAS_PER_SM * h->sm_count / 4));
      EA.hsum = SA.hsum;
      EA.N = N;
      EA.ntile = std::min((G.gram_nmax + 7) / 8, G.gram_tgmax);
      EA.nk4 = std::min((G.gram_nmax + 3) / 4, 2 * G.gram_tgmax);
      const int hbytes = 4 * N * (N + 4) * 8;
      EA.nbuf = 2 * hbytes <= GM_GRAM_EVAL_SMEM_MAX ? 2 : 1;
      EA.T = t->T.as<double>();
      EA.nrows = t->nrows;
      EA.part = t->h->scratch_part.as<double>();
      EA.nchunk = nchunk_total;
      EA.chunk = nchunk;
      int eblocks = 4;                                     // blocks of 96 angles
      if (t->nnode > 0) {
        // evaluate at the 2 N + 1 Chebyshev nodes only; k_gram_interp carries the four polynomials to the table's angles
        eblocks = (t->nnode + 95) / 96;
        EA.T = t->T2.as<double>();
        EA.part = t->h->scratch_nodepart.as<double>();
        EA.nchunk = 1;
        EA.chunk = 0;
        EA.tasks_per_cta = std::max(1, (nt * eblocks + GM_EVAL_CTAS_PER_SM * h->sm_count - 1) / (GM_EVAL_CTAS_PER_SM * h->sm_count));
      }
      if ((rc = ev_mark(t, 4))) return rc;
      k_gram_sum<<<dim3((2 * N * N + 255) / 256, nt), 256, 0, st>>>(SA);
      GM_LAUNCH_CHECK(h);
      const dim3 egrid(eblocks, (nt + EA.tasks_per_cta - 1) / EA.tasks_per_cta);
      switch (EA.ntile) {
        case 1: k_gram_eval<1><<<egrid, GM_GRAM_EVAL_THREADS, EA.nbuf * hbytes, st>>>(EA); break;
        case 2: k_gram_eval<2><<<egrid, GM_GRAM_EVAL_THREADS, EA.nbuf * hbytes, st>>>(EA); break;
        case 3: k_gram_eval<3><<<egrid, GM_GRAM_EVAL_THREADS, EA.nbuf * hbytes, st>>>(EA); break;
        case 4: k_gram_eval<4><<<egrid, GM_GRAM_EVAL_THREADS, EA.nbuf * hbytes, st>>>(EA); break;
        case 5: k_gram_eval<5><<<egrid, GM_GRAM_EVAL_THREADS, EA.nbuf * hbytes, st>>>(EA); break;
        case 6: k_gram_eval<6><<<egrid, GM_GRAM_EVAL_THREADS, EA.nbuf * hbytes, st>>>(EA); break;
        case 7: k_gram_eval<7><<<egrid, GM_GRAM_EVAL_THREADS, EA.nbuf * hbytes, st>>>(EA); break;
        default: k_gram_eval<8><<<egrid, GM_GRAM_EVAL_THREADS, EA.nbuf * hbytes, st>>>(EA); break;
      }
      GM_LAUNCH_CHECK(h);
      if (t->nnode > 0) {
        k_gram_interp<<<nt, GM_NANG_PAD, 0, st>>>(t->nang, t->nnode, t->W.as<double>(), t->h->scratch_nodepart.as<double>(),
                                                  t->h->scratch_part.as<double>(), nchunk_total, nchunk);
        GM_LAUNCH_CHECK(h);
      }
      if ((rc = ev_mark(t, 4))) return rc;
    }
    if (!per_particle) {
      if ((rc = ev_mark(t, 2))) return rc;
      k_finalize<<<nt, GM_NANG_PAD, 0, st>>>(nchunk_total, G.ngroup, nmode, t->nang, t->h->scratch_part.as<double>(), t->h->scratch_scal_part.as<double>(),
                                             d_out_phase + (size_t)t0 * 4 * t->nang, d_out_scal + (size_t)t0 * nmode * GM_NSCAL,
                                             t->mirror_phase ? t->mirror_phase + (size_t)t0 * 4 * t->nang : nullptr,
                                             t->mirror_scal ? t->mirror_scal + (size_t)t0 * nmode * GM_NSCAL : nullptr);
      GM_LAUNCH_CHECK(h);
      if ((rc = ev_mark(t, 2))) return rc;
      if (t->gsf_ng > 0) {
        // fused GSF stage on this batch's phase sums (same stream, so it runs right behind k_finalize)
        rc = gm_gsf_phase4_async(h, nt, t->nang, t->gsf_ang.data(), d_out_phase + (size_t)t0 * 4 * t->nang, t->gsf_ng,
                                 t->gsf_coef.as<double>() + (size_t)t0 * 6 * t->gsf_ng, t->gsf_cnorm.as<double>() + t0, t->gsf_quant);
        if (rc) return rc;
      }
      if (hio) {
        GM_CUDA_TRY(cudaEventRecord(t->io_events[nbatch + bi], st));
        GM_CUDA_TRY(cudaStreamWaitEvent(t->d2h_stream, t->io_events[nbatch + bi], 0));
        GM_CUDA_TRY(cudaMemcpyAsync(hio->out_scal + (size_t)t0 * nmode * GM_NSCAL, d_out_scal + (size_t)t0 * nmode * GM_NSCAL,
                                    sizeof(double) * (size_t)nt * nmode * GM_NSCAL, cudaMemcpyDeviceToHost, t->d2h_stream));
        if (hio->out_phase)
          GM_CUDA_TRY(cudaMemcpyAsync(hio->out_phase + (size_t)t0 * 4 * t->nang, d_out_phase + (size_t)t0 * 4 * t->nang,
                                      sizeof(double) * (size_t)nt * 4 * t->nang, cudaMemcpyDeviceToHost, t->d2h_stream));
        if (t->gsf_ng > 0 && t->gsf_coef_host) {
          GM_CUDA_TRY(cudaMemcpyAsync(t->gsf_coef_host + (size_t)t0 * 6 * t->gsf_ng, t->gsf_coef.as<double>() + (size_t)t0 * 6 * t->gsf_ng,
                                      sizeof(double) * (size_t)nt * 6 * t->gsf_ng, cudaMemcpyDeviceToHost, t->d2h_stream));
          if (t->gsf_cnorm_host)
            GM_CUDA_TRY(cudaMemcpyAsync(t->gsf_cnorm_host + t0, t->gsf_cnorm.as<double>() + t0, sizeof(double) * nt, cudaMemcpyDeviceToHost,
                                        t->d2h_stream));
        }
      }
    }
  }
  t->last_stats[4] = (double)(h->launches - launches0);
  if (window_set) {
    cudaStreamAttrValue av;
    memset(&av, 0, sizeof(av));
    av.accessPolicyWindow.num_bytes = 0;      // window off again for whatever else runs on the caller's stream
    cudaStreamSetAttribute(st, cudaStreamAttributeAccessPolicyWindow, &av);
  }
  if (hio) {
    GM_CUDA_TRY(cudaStreamSynchronize(t->d2h_stream));
    GM_CUDA_TRY(cudaStreamSynchronize(t->h2d_stream));
  }
  return GM_OK;
}

static int fetch_stats(gm_table_t t) {
  unsigned long long s[8];
  GM_CUDA_TRY(cudaMemcpyAsync(s, t->stats.p, sizeof(s), cudaMemcpyDeviceToHost, t->h->stream));
  GM_CUDA_TRY(cudaStreamSynchronize(t->h->stream));
  for (int i = 0; i < 4; ++i) t->last_stats[i] = (double)s[i];
  t->last_stats[5] = (double)s[5];
  if (s[5] != 0) {
    // sqrt(w_phase) is folded into the coefficients: a negative phase weight has no meaning there (the outputs hold NaN)
    gm_set_error("invalid argument: %llu negative w_phase value(s); split signed weights into w+ and w- (sums are linear in w)", s[5]);
    return GM_EINVAL;
  }
  return GM_OK;
}

extern "C" int gm_table_run_dev(gm_table_t t, int ntask, const double* mz, const double* mrel, int nmode, const double* w_phase,
                                const double* w_scal, int flags, double* out_scal, double* out_phase) {
  GM_REQUIRE(t != nullptr, "table is NULL");
  GM_REQUIRE(ntask > 0 && nmode >= 1 && nmode <= 32, "ntask / nmode out of range");
  GM_REQUIRE(mz && mrel && w_phase && out_scal && out_phase, "NULL argument");
  GM_REQUIRE(w_scal || nmode == 1, "w_scal is required when nmode > 1");
  GM_CUDA_TRY(cudaSetDevice(t->h->device));
  return table_run_core(t, ntask, mz, mrel, nmode, w_phase, w_scal, flags, out_scal, out_phase, nullptr, nullptr, false);
}

extern "C" int gm_table_run(gm_table_t t, int ntask, const double* mz, const double* mrel, int nmode, const double* w_phase,
                            const double* w_scal, int flags, double* out_scal, double* out_phase) {
  GM_REQUIRE(t != nullptr, "table is NULL");
  GM_REQUIRE(ntask > 0 && nmode >= 1 && nmode <= 32, "ntask / nmode out of range");
  GM_REQUIRE(mz && mrel && w_phase && out_scal && (out_phase || (flags & GM_F_PHASE_ON_DEVICE)), "NULL argument");
  GM_REQUIRE(w_scal || nmode == 1, "w_scal is required when nmode > 1");
  gm_handle_t h = t->h;
  GM_CUDA_TRY(cudaSetDevice(h->device));
  cudaStream_t st = h->stream;
  int rc;
  const size_t nw = (size_t)ntask * t->nx;
  if ((rc = t->mz.ensure(sizeof(double2) * ntask)) || (rc = t->mrel.ensure(sizeof(double2) * ntask)) ||
      (rc = t->h->scratch_wphase.ensure(sizeof(double) * nw)) || (rc = t->out_scal.ensure(sizeof(double) * (size_t)ntask * nmode * GM_NSCAL)) ||
      (rc = gm_pool_take(t->h, t->out_phase, sizeof(double) * (size_t)ntask * 4 * t->nang)))
    return rc;
  if (w_scal && (rc = t->h->scratch_wscal.ensure(sizeof(double) * nw * nmode))) return rc;
  GM_CUDA_TRY(cudaMemcpyAsync(t->mz.p, mz, sizeof(double2) * ntask, cudaMemcpyHostToDevice, st));
  GM_CUDA_TRY(cudaMemcpyAsync(t->mrel.p, mrel, sizeof(double2) * ntask, cudaMemcpyHostToDevice, st));
  HostIO hio = {w_phase, w_scal, out_scal, (flags & GM_F_PHASE_ON_DEVICE) ? nullptr : out_phase};
  rc = table_run_core(t, ntask, t->mz.as<double>(), t->mrel.as<double>(), nmode, t->h->scratch_wphase.as<double>(),
                      w_scal ? t->h->scratch_wscal.as<double>() : nullptr, flags, t->out_scal.as<double>(), t->out_phase.as<double>(), nullptr,
                      nullptr, false, &hio);
  if (rc) return rc;
  return fetch_stats(t);
}

extern "C" int gm_table_run_coated(gm_table_t t, int ntask, const double* m1, const double* m2, const double* core_ratio, int nmode,
                                   const double* w_phase, const double* w_scal, int flags, double* out_scal, double* out_phase) {
  GM_REQUIRE(t != nullptr, "table is NULL");
  GM_REQUIRE(ntask > 0 && nmode >= 1 && nmode <= 32, "ntask / nmode out of range");
  GM_REQUIRE(m1 && m2 && core_ratio && w_phase && out_scal && out_phase, "NULL argument");
  GM_REQUIRE(w_scal || nmode == 1, "w_scal is required when nmode > 1");
  for (int i = 0; i < ntask; ++i) GM_REQUIRE(core_ratio[i] > 0.0 && core_ratio[i] < 1.0, "core_ratio must be in (0, 1)");
  gm_handle_t h = t->h;
  GM_CUDA_TRY(cudaSetDevice(h->device));
  cudaStream_t st = h->stream;
  int rc;
  if (!t->c_soff.p) {
    std::vector<long long> soff(t->nx + 1, 0), aboff(t->nx + 1, 0);
    for (int i = 0; i < t->nx; ++i) {
      soff[i + 1] = soff[i] + t->hnmax[i] + 1;
      aboff[i + 1] = aboff[i] + t->hnmax[i];
    }
    t->c_nscr = soff[t->nx];
    t->c_nab = aboff[t->nx];
    if ((rc = t->c_soff.ensure(sizeof(long long) * (t->nx + 1))) || (rc = t->c_aboff.ensure(sizeof(long long) * (t->nx + 1)))) return rc;
    GM_CUDA_TRY(cudaMemcpyAsync(t->c_soff.p, soff.data(), sizeof(long long) * (t->nx + 1), cudaMemcpyHostToDevice, st));
    GM_CUDA_TRY(cudaMemcpyAsync(t->c_aboff.p, aboff.data(), sizeof(long long) * (t->nx + 1), cudaMemcpyHostToDevice, st));
    GM_CUDA_TRY(cudaStreamSynchronize(st));
  }
  const size_t nw = (size_t)ntask * t->nx;
  if ((rc = t->mz.ensure(sizeof(double2) * ntask)) || (rc = t->mrel.ensure(sizeof(double2) * ntask)) ||
      (rc = t->c_ratio.ensure(sizeof(double) * ntask)) || (rc = t->h->scratch_wphase.ensure(sizeof(double) * nw)) ||
      (rc = t->out_scal.ensure(sizeof(double) * (size_t)ntask * nmode * GM_NSCAL)) ||
      (rc = gm_pool_take(t->h, t->out_phase, sizeof(double) * (size_t)ntask * 4 * t->nang)))
    return rc;
  if (w_scal && (rc = t->h->scratch_wscal.ensure(sizeof(double) * nw * nmode))) return rc;
  GM_CUDA_TRY(cudaMemcpyAsync(t->mz.p, m1, sizeof(double2) * ntask, cudaMemcpyHostToDevice, st));
  GM_CUDA_TRY(cudaMemcpyAsync(t->mrel.p, m2, sizeof(double2) * ntask, cudaMemcpyHostToDevice, st));
  GM_CUDA_TRY(cudaMemcpyAsync(t->c_ratio.p, core_ratio, sizeof(double) * ntask, cudaMemcpyHostToDevice, st));
  HostIO hio = {w_phase, w_scal, out_scal, out_phase};
  rc = table_run_core(t, ntask, t->mz.as<double>(), t->mrel.as<double>(), nmode, t->h->scratch_wphase.as<double>(),
                      w_scal ? t->h->scratch_wscal.as<double>() : nullptr, flags, t->out_scal.as<double>(), t->out_phase.as<double>(), nullptr,
                      nullptr, false, &hio, t->c_ratio.as<double>());
  if (rc) return rc;
  return fetch_stats(t);
}

extern "C" int gm_table_set_gsf(gm_table_t t, const double* ang_deg, int ng, int quantize10, double* coef_host, double* cnorm_host) {
  GM_REQUIRE(t != nullptr, "table is NULL");
  if (ng <= 0 || !ang_deg) {   // switch the fused stage off
    t->gsf_ng = 0;
    t->gsf_coef_host = t->gsf_cnorm_host = nullptr;
    return GM_OK;
  }
  GM_REQUIRE(ng >= 3 && ng <= 2048, "ng out of range");
  t->gsf_ang.assign(ang_deg, ang_deg + t->nang);
  t->gsf_ng = ng;
  t->gsf_quant = quantize10;
  t->gsf_coef_host = coef_host;
  t->gsf_cnorm_host = cnorm_host;
  return GM_OK;
}

extern "C" int gm_table_set_mirror(gm_table_t t, double* scal_mirror, double* phase_mirror) {
  GM_REQUIRE(t != nullptr, "table is NULL");
  GM_REQUIRE((scal_mirror == nullptr) == (phase_mirror == nullptr), "give both mirror pointers or neither");
  t->mirror_scal = scal_mirror;
  t->mirror_phase = phase_mirror;
  return GM_OK;
}

extern "C" int gm_table_gsf_device(gm_table_t t, double** coef, double** cnorm) {
  GM_REQUIRE(t != nullptr, "table is NULL");
  if (coef) *coef = t->gsf_coef.as<double>();
  if (cnorm) *cnorm = t->gsf_cnorm.as<double>();
  return GM_OK;
}

extern "C" int gm_table_set_dr(gm_table_t t, const double* dr) {
  GM_REQUIRE(t && dr, "NULL argument");
  GM_CUDA_TRY(cudaSetDevice(t->h->device));
  int rc = t->dr.ensure(sizeof(double) * t->nx);
  if (rc) return rc;
  GM_CUDA_TRY(cudaMemcpyAsync(t->dr.p, dr, sizeof(double) * t->nx, cudaMemcpyHostToDevice, t->h->stream));
  GM_CUDA_TRY(cudaStreamSynchronize(t->h->stream));
  t->have_dr = true;
  return GM_OK;
}

extern "C" int gm_table_run_psd(gm_table_t t, int ntask, const double* mz, const double* mrel, int nmode, int psd_kind,
                                const double* psd_params, const double* frac, int flags, double* out_scal, double* out_phase) {
  GM_REQUIRE(t != nullptr, "table is NULL");
  GM_REQUIRE(ntask > 0 && nmode >= 1 && nmode <= 32, "ntask / nmode out of range");
  GM_REQUIRE(mz && mrel && psd_params && frac && out_scal && (out_phase || (flags & GM_F_PHASE_ON_DEVICE)), "NULL argument");
  GM_REQUIRE(psd_kind >= GM_PSD_LOGNORM && psd_kind <= GM_PSD_DU, "unknown psd_kind");
  GM_REQUIRE(psd_kind != GM_PSD_LOGNORM || t->have_dr, "GM_PSD_LOGNORM needs gm_table_set_dr first");
  GM_REQUIRE(t->nx >= 2, "need at least two grid points");
  gm_handle_t h = t->h;
  GM_CUDA_TRY(cudaSetDevice(h->device));
  cudaStream_t st = h->stream;
  int rc;
  const size_t nw = (size_t)ntask * t->nx;
  bool separate = nmode > 1;
  for (int i = 0; i < ntask * nmode && !separate; ++i) separate = frac[i] != 1.0;
  if ((rc = t->mz.ensure(sizeof(double2) * ntask)) || (rc = t->mrel.ensure(sizeof(double2) * ntask)) ||
      (rc = t->h->scratch_wphase.ensure(sizeof(double) * nw)) || (rc = t->out_scal.ensure(sizeof(double) * (size_t)ntask * nmode * GM_NSCAL)) ||
      (rc = gm_pool_take(t->h, t->out_phase, sizeof(double) * (size_t)ntask * 4 * t->nang)) ||
      (rc = t->psd_par.ensure(sizeof(double) * (size_t)ntask * nmode * GM_PSD_NPAR)) ||
      (rc = t->psd_frac.ensure(sizeof(double) * (size_t)ntask * nmode)))
    return rc;
  if (separate && (rc = t->h->scratch_wscal.ensure(sizeof(double) * nw * nmode))) return rc;
  GM_CUDA_TRY(cudaMemcpyAsync(t->mz.p, mz, sizeof(double2) * ntask, cudaMemcpyHostToDevice, st));
  GM_CUDA_TRY(cudaMemcpyAsync(t->mrel.p, mrel, sizeof(double2) * ntask, cudaMemcpyHostToDevice, st));
  GM_CUDA_TRY(cudaMemcpyAsync(t->psd_par.p, psd_params, sizeof(double) * (size_t)ntask * nmode * GM_PSD_NPAR, cudaMemcpyHostToDevice, st));
  GM_CUDA_TRY(cudaMemcpyAsync(t->psd_frac.p, frac, sizeof(double) * (size_t)ntask * nmode, cudaMemcpyHostToDevice, st));
  k_psd<<<ntask, 256, 0, st>>>(t->nx, nmode, psd_kind, t->D.x.as<double>(), t->dr.as<double>(), t->psd_par.as<double>(),
                               t->psd_frac.as<double>(), t->h->scratch_wscal.as<double>(), t->h->scratch_wphase.as<double>(), separate ? 1 : 0);
  GM_LAUNCH_CHECK(h);
  t->psd_separate = separate;
  HostIO hio = {nullptr, nullptr, out_scal, (flags & GM_F_PHASE_ON_DEVICE) ? nullptr : out_phase};   // results are downloaded batch by batch behind the kernels
  rc = table_run_core(t, ntask, t->mz.as<double>(), t->mrel.as<double>(), nmode, t->h->scratch_wphase.as<double>(),
                      separate ? t->h->scratch_wscal.as<double>() : nullptr, flags, t->out_scal.as<double>(), t->out_phase.as<double>(), nullptr,
                      nullptr, false, &hio);
  if (rc) return rc;
  return fetch_stats(t);
}

extern "C" int gm_table_get_weights(gm_table_t t, int ntask, int nmode, double* w) {
  GM_REQUIRE(t && w, "NULL argument");
  GM_CUDA_TRY(cudaSetDevice(t->h->device));
  const void* src = t->psd_separate ? t->h->scratch_wscal.p : t->h->scratch_wphase.p;
  GM_REQUIRE(src != nullptr, "no weights have been generated");
  GM_CUDA_TRY(cudaMemcpyAsync(w, src, sizeof(double) * (size_t)ntask * nmode * t->nx, cudaMemcpyDeviceToHost, t->h->stream));
  GM_CUDA_TRY(cudaStreamSynchronize(t->h->stream));
  return GM_OK;
}

// device copies of the outputs of the last gm_table_run (host-buffer variant), e.g. to chain gm_gsf_expand_phase4_dev
extern "C" int gm_table_device_outputs(gm_table_t t, double** out_scal, double** out_phase) {
  GM_REQUIRE(t != nullptr, "table is NULL");
  if (out_scal) *out_scal = t->out_scal.as<double>();
  if (out_phase) *out_phase = t->out_phase.as<double>();
  return GM_OK;
}

// k_phase_norm on the device-resident phase sums of the last run: block = planes [4][ntask][nang] (when want_planes) then pback [ntask][4]
static int normalize_on_device(gm_table_t t, int ntask, const double* theta_rad, const double* sin_theta, bool want_planes, double** planes,
                               double** pback) {
  GM_REQUIRE(ntask > 0 && t->out_phase.p && t->out_phase.cap >= sizeof(double) * (size_t)ntask * 4 * t->nang, "no phase sums of that many tasks on the device");
  gm_handle_t h = t->h;
  GM_CUDA_TRY(cudaSetDevice(h->device));
  cudaStream_t st = h->stream;
  int rc;
  const size_t plane = (size_t)ntask * t->nang;
  if ((rc = gm_pool_take(t->h, t->norm_planes, sizeof(double) * ((want_planes ? 4 * plane : 0) + 4 * (size_t)ntask))) || (rc = t->norm_ang.ensure(sizeof(double) * 2 * t->nang)))
    return rc;
  double* d_ang = t->norm_ang.as<double>();
  GM_CUDA_TRY(cudaMemcpyAsync(d_ang, theta_rad, sizeof(double) * t->nang, cudaMemcpyHostToDevice, st));
  GM_CUDA_TRY(cudaMemcpyAsync(d_ang + t->nang, sin_theta, sizeof(double) * t->nang, cudaMemcpyHostToDevice, st));
  *planes = want_planes ? t->norm_planes.as<double>() : nullptr;
  *pback = t->norm_planes.as<double>() + (want_planes ? 4 * plane : 0);
  k_phase_norm<<<ntask, GM_NANG_PAD, 0, st>>>(ntask, t->nang, t->out_phase.as<double>(), d_ang, d_ang + t->nang, *planes, *pback);
  GM_LAUNCH_CHECK(h);
  return GM_OK;
}

extern "C" int gm_table_fetch_normalized(gm_table_t t, int ntask, const double* theta_rad, const double* sin_theta, double* p11, double* p12,
                                         double* p33, double* p34, double* pback4) {
  GM_REQUIRE(t && theta_rad && sin_theta && pback4, "NULL argument");
  const bool want_planes = p11 != nullptr;
  GM_REQUIRE(want_planes ? (p12 && p33 && p34) : (!p12 && !p33 && !p34), "give all four planes or none of them");
  double *d_pl = nullptr, *d_pb = nullptr;
  int rc = normalize_on_device(t, ntask, theta_rad, sin_theta, want_planes, &d_pl, &d_pb);
  if (rc) return rc;
  cudaStream_t st = t->h->stream;
  const size_t plane = (size_t)ntask * t->nang;
  if (want_planes) {
    double* dst[4] = {p11, p12, p33, p34};
    for (int q = 0; q < 4; ++q)
      GM_CUDA_TRY(cudaMemcpyAsync(dst[q], d_pl + (size_t)q * plane, sizeof(double) * plane, cudaMemcpyDeviceToHost, st));
  }
  GM_CUDA_TRY(cudaMemcpyAsync(pback4, d_pb, sizeof(double) * 4 * (size_t)ntask, cudaMemcpyDeviceToHost, st));
  GM_CUDA_TRY(cudaStreamSynchronize(st));
  return GM_OK;
}

extern "C" int gm_table_normalize_device(gm_table_t t, int ntask, const double* theta_rad, const double* sin_theta, double** block) {
  GM_REQUIRE(t && theta_rad && sin_theta && block, "NULL argument");
  double *d_pl = nullptr, *d_pb = nullptr;
  int rc = normalize_on_device(t, ntask, theta_rad, sin_theta, true, &d_pl, &d_pb);
  if (rc) return rc;
  *block = d_pl;
  return GM_OK;
}

extern "C" int gm_table_particles(gm_table_t t, int ntask, const double* mz, const double* mrel, double* q, double* s12) {
  GM_REQUIRE(t != nullptr, "table is NULL");
  GM_REQUIRE(ntask > 0 && mz && mrel, "bad arguments");
  gm_handle_t h = t->h;
  GM_CUDA_TRY(cudaSetDevice(h->device));
  cudaStream_t st = h->stream;
  int rc;
  const size_t np = (size_t)ntask * t->nx;
  if ((rc = t->mz.ensure(sizeof(double2) * ntask)) || (rc = t->mrel.ensure(sizeof(double2) * ntask)) ||
      (rc = t->q.ensure(sizeof(double) * 6 * np)) || (rc = t->h->scratch_wphase.ensure(sizeof(double) * np)))
    return rc;
  if (s12 && (rc = t->s12.ensure(sizeof(double) * 4 * np * t->nang))) return rc;
  GM_CUDA_TRY(cudaMemcpyAsync(t->mz.p, mz, sizeof(double2) * ntask, cudaMemcpyHostToDevice, st));
  GM_CUDA_TRY(cudaMemcpyAsync(t->mrel.p, mrel, sizeof(double2) * ntask, cudaMemcpyHostToDevice, st));
  GM_CUDA_TRY(cudaMemsetAsync(t->h->scratch_wphase.p, 0, sizeof(double) * np, st));
  if ((rc = t->out_scal.ensure(sizeof(double) * (size_t)ntask * GM_NSCAL)) || (rc = gm_pool_take(t->h, t->out_phase, sizeof(double) * (size_t)ntask * 4 * t->nang)))
    return rc;
  rc = table_run_core(t, ntask, t->mz.as<double>(), t->mrel.as<double>(), 1, t->h->scratch_wphase.as<double>(), nullptr, 0,
                      t->out_scal.as<double>(), t->out_phase.as<double>(), t->q.as<double>(), s12 ? t->s12.as<double>() : nullptr,
                      s12 != nullptr);
  if (rc) return rc;
  if (q) GM_CUDA_TRY(cudaMemcpyAsync(q, t->q.p, sizeof(double) * 6 * np, cudaMemcpyDeviceToHost, st));
  if (s12) GM_CUDA_TRY(cudaMemcpyAsync(s12, t->s12.p, sizeof(double) * 4 * np * t->nang, cudaMemcpyDeviceToHost, st));
  return fetch_stats(t);
}
