// gm_small.cuh -- fused coefficient + Gram kernel for the particle groups with max nmax <= 8 (sm_100a, FP64 DMMA).
//
// What it replaces: for these groups, k_coeff (single_mie_coeff + mie_props_raw + the scalar part of integratePSD,
// mie_coeffs.py:83-130, mie_props.py:28-70, dointegration.py:1104-1190) AND the classes 0 / 1 of k_gram
// (mie_S12_backend_pt + calculateScatVals + the phase part of integratePSD, mie_props.py:133-150, dointegration.py:1044-1050,
// :1164-1166).  On optics_SU 79 % of the particle groups are in these two classes; their coefficient rows were 47 % of the
// 2.9 GB per step that k_coeff wrote to HBM and k_gram read back at 132 B per DMMA (HBM-bound, 0.42 / 0.56 of the DMMA peak).
// Here the rows never leave the SM: a warp evaluates a group of 32 particles (lane = particle) into its private
// shared-memory tile [ROWS][GM_SB] -- the same row layout as the coefficient stream -- and multiplies the tile into
// register-resident Gram accumulators on the FP64 tensor cores right away.
//
// Work item = one warp = (class, segment of that class's groups, task).  The warp walks through the groups of its segment:
//   * the order loop is fully unrolled and branch-free (ROWS = 4 or 8 orders, selects instead of branches), so that the
//     a_n / b_n arithmetic of order n overlaps the serial log-derivative recurrence of order n - 1 in one instruction stream;
//   * the eleven size-distribution sums are accumulated per lane over the whole segment and reduced once per work item
//     (k_coeff: one 16-value warp reduction per group);
//   * per-task constants (1/m_z, 1/m_rel) come from a tiny prep kernel, 1/x from the table.
// The partial Gram blocks go to the slots k_gram_sum already sums ([task][descriptor][team = segment][4][Nd][Nd]), the scalar
// sums to the scal_part slot of the segment's first group (zeros in the others), so k_gram_sum / k_gram_eval / k_finalize
// are unchanged.
#pragma once
#include "gm_gram.cuh"

#ifndef GM_SMALL_MINB4
#define GM_SMALL_MINB4 4         // CTAs of 128 threads per SM of the 4-order / 8-order instantiation (register budget 65536 / (128 * MINB))
#define GM_SMALL_MINB8 3
#endif
#ifndef GM_SMALL_SHORT_START
// 1: the D_n recurrence starts 4 + ceil(2.5 |z|) (at most the reference's 16) orders above max(nmax, |z|).  Starting from D = 0, the
// error of D_n falls by ~ |z|^2 / (4 j^2) per order j: for these groups (nmax <= 8, |z| <~ 5) D_n is converged to <= 1e-14 where the
// reference's fixed 16 orders (mie_coeffs.py:101) are; CPU error study in profiles/r01k_coeff_start_offset_study.txt, GPU parity
// of the complete optics_SU / optics_BC tables unchanged.  0 = exactly the reference's start order.
#define GM_SMALL_SHORT_START 1
#endif
constexpr int GM_SMALL_WARPS = 4;
constexpr int GM_SMALL_MAXSEG = 12;   // k_gram_sum adds up to 12 partial blocks per descriptor

struct SmallSeg {
  int cls;            // 0: stacked tile (max nmax <= 4), 1: max nmax <= 8
  int gbegin, gend;   // range in glist
  int seg;            // segment number inside its class = partial slot ("team") of the class's descriptor
  long long hoff;     // offset (doubles) of the class's descriptor inside a task's partial-H block
};

struct SmallArgs {
  int nx, ngroup, ntask, nseg;
  const SmallSeg* segs;
  const int* glist;
  const double* x;
  const double* xinv;
  const int* nmax;
  const double* psi;
  const double* chi;
  const long long* gboff;
  const double2* mz;      // [ntask] sqrt(eps mu)
  const double2* mrel;    // [ntask] sqrt(eps / mu)
  const double2* mzinv;   // [ntask] 1 / mz      (k_task_prep)
  const double2* mrinv;   // [ntask] 1 / mrel
  const double* wphase;   // [ntask][nx]
  const double* wscal;    // [ntask][nx] or null (nmode == 1)
  int dense;
  const double2* ntab;    // [n] = ((2n+1)/(n(n+1)), n(n+2)/(n+1))
  double* hpart;          // [ntask][hstride]
  long long hstride;
  double* scal_part;      // [ntask][1][ngroup][GM_NSCAL]
  unsigned long long* stats;
};

// 1 / m_z and 1 / m_rel per task (exact complex reciprocals; k_coeff forms them once per warp and task)
__global__ void __launch_bounds__(128) k_task_prep(int ntask, const double2* __restrict__ mz, const double2* __restrict__ mrel,
                                                   double2* __restrict__ mzinv, double2* __restrict__ mrinv) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= ntask) return;
  const double2 a = mz[t], b = mrel[t];
  const double da = a.x * a.x + a.y * a.y, db = b.x * b.x + b.y * b.y;
  mzinv[t] = make_double2(a.x / da, -a.y / da);
  mrinv[t] = make_double2(b.x / db, -b.y / db);
}

template <int ROWS>
__device__ __forceinline__ void small_item(const SmallArgs& A, const SmallSeg sg, const int task, double* tile) {
  constexpr bool STACK = ROWS == 4;
  const int lane = threadIdx.x & 31;
  const int lk = lane & 3, lr = lane >> 2;
  const double2 mzv = A.mz[task], mrv = A.mrel[task], mzi = A.mzinv[task], minv = A.mrinv[task];
  const double* wp_t = A.wphase + (size_t)task * A.nx;
  const double* ws_t = A.wscal ? A.wscal + (size_t)task * A.nx : nullptr;

  double S[GM_NSCAL];
#pragma unroll
  for (int k = 0; k < GM_NSCAL; ++k) S[k] = 0.0;
  double acc[STACK ? 2 : 4][2];
#pragma unroll
  for (int q = 0; q < (STACK ? 2 : 4); ++q) acc[q][0] = acc[q][1] = 0.0;
  unsigned st_ev = 0, st_nm = 0, st_nx = 0, st_k4 = 0, st_neg = 0;

  const double sgn = (lk & 1) ? -1.0 : 1.0;
  const int lane_off = STACK ? (lr & 3) * GM_SB + (lr >> 2) * 64 + lk : lr * GM_SB + lk;

  for (int gi = sg.gbegin; gi < sg.gend; ++gi) {
    const int g = A.glist[gi];
    const int i = g * 32 + lane;
    const bool valid = i < A.nx;
    const double xi = valid ? A.x[i] : 1.0;
    const int nm = valid ? A.nmax[i] : 0;
    const double wp = valid ? wp_t[i] : 0.0;
    const double ws = valid ? (ws_t ? ws_t[i] : wp) : 0.0;
    const bool act = valid && (A.dense || wp != 0.0 || ws != 0.0);
    st_neg += (valid && wp < 0.0) ? 1u : 0u;
    if (!__any_sync(0xffffffffu, act)) continue;   // the group adds exact zeros to every sum (x-moment sums included: w == 0)
    const double xinv = valid ? A.xinv[i] : 1.0;
    // Riccati-Bessel values of orders 0..ROWS, requested before the recurrence starts (their latency hides behind phase 1)
    const size_t bbase = (size_t)A.gboff[g] + lane;
    double psi[ROWS + 1], chi[ROWS + 1];
#pragma unroll
    for (int n = 0; n <= ROWS; ++n) {
      const bool have = act && n <= nm;
      psi[n] = have ? A.psi[bbase + (size_t)n * 32] : 0.0;
      chi[n] = have ? A.chi[bbase + (size_t)n * 32] : 0.0;
    }
    const double2 z = make_double2(mzv.x * xi, mzv.y * xi);                       // mie_coeffs.py:96
    const double2 zinv = make_double2(mzi.x * xinv, mzi.y * xinv);
    const double zabs = sqrt(fma(z.x, z.x, z.y * z.y));
#if GM_SMALL_SHORT_START
    const int nmx = act ? (int)rint(fmax((double)nm, zabs) + fmin(16.0, 4.0 + ceil(2.5 * zabs))) : 0;
#else
    const int nmx = act ? (int)rint(fmax((double)nm, zabs) + 16.0) : 0;          // mie_coeffs.py:101
#endif
    const int J = __reduce_max_sync(0xffffffffu, nmx);
    st_ev += act ? 1u : 0u;
    st_nm += act ? (unsigned)nm : 0u;
    st_nx += (unsigned)nmx;
    st_k4 += 1u;
    const double sw = sqrt(wp);

    // ---- phase 1: orders above ROWS, only the logarithmic-derivative recurrence (mie_coeffs.py:119-121) in the denominator
    // form of k_coeff: t_{n-1} = (2n+1)/z - 1/t_n, D_n = (n+1)/z - 1/t_n, started from D_{nmx} = 0 (t_{nmx-1} = nmx / z)
    double2 tt = make_double2((double)nmx * zinv.x, (double)nmx * zinv.y);
    int n = J - 1;
    double f2 = (double)(2 * n + 1);
    for (; n > ROWS; --n, f2 -= 2.0) {
      const double2 ti = crcp(tt);
      const bool on = n < nmx;
      tt.x = on ? fma(f2, zinv.x, -ti.x) : tt.x;
      tt.y = on ? fma(f2, zinv.y, -ti.y) : tt.y;
    }
    // ---- phase 2: orders ROWS..1, unrolled and branch-free.  A lane joins the recurrence at its own order nmx - 1 (`on`), which may
    // lie inside this range when its nmax is far below the group's (unsorted grids); lanes that are not active compute discarded values.
    double sext = 0.0, ssca = 0.0, qbr = 0.0, qbi = 0.0, sasy = 0.0;
    double2 a_next = make_double2(0.0, 0.0), b_next = make_double2(0.0, 0.0);
#pragma unroll
    for (int r = ROWS; r >= 1; --r) {
      const double fr2 = (double)(2 * r + 1);
      const double2 ti = crcp(tt);
      const bool on = r < nmx;
      const double2 D = make_double2(on ? fma((double)(r + 1), zinv.x, -ti.x) : 0.0, on ? fma((double)(r + 1), zinv.y, -ti.y) : 0.0);
      tt.x = on ? fma(fr2, zinv.x, -ti.x) : tt.x;
      tt.y = on ? fma(fr2, zinv.y, -ti.y) : tt.y;
      const bool emit = act && r <= nm;
      const bool wnz = sw != 0.0;   // a zero-weight particle of a dense run adds exact zeros even where a_n, b_n are not finite
      const double nox = (double)r * xinv;
      double2 da = cmul(D, minv);                                               // mie_coeffs.py:124
      da.x += nox;
      double2 db = cmul(D, mrv);                                                // mie_coeffs.py:125
      db.x += nox;
      const double psi_n = psi[r], psi_m = psi[r - 1], chi_n = chi[r], chi_m = chi[r - 1];
      // a_n = (da psi_n - psi_{n-1}) / (da xi_n - xi_{n-1}),  xi = psi - i chi          (mie_coeffs.py:113-114,127-128)
      double2 an = cdiv(make_double2(fma(da.x, psi_n, -psi_m), da.y * psi_n),
                        make_double2(fma(da.x, psi_n, fma(da.y, chi_n, -psi_m)), fma(da.y, psi_n, fma(-da.x, chi_n, chi_m))));
      double2 bn = cdiv(make_double2(fma(db.x, psi_n, -psi_m), db.y * psi_n),
                        make_double2(fma(db.x, psi_n, fma(db.y, chi_n, -psi_m)), fma(db.y, psi_n, fma(-db.x, chi_n, chi_m))));
      an.x = emit ? an.x : 0.0;
      an.y = emit ? an.y : 0.0;
      bn.x = emit ? bn.x : 0.0;
      bn.y = emit ? bn.y : 0.0;
      // efficiencies, mie_props.py:41-65 (orders above nmax add exact zeros)
      const double2 nt = __ldg(A.ntab + r);
      sext += fr2 * (an.x + bn.x);
      ssca += fr2 * (an.x * an.x + an.y * an.y + bn.x * bn.x + bn.y * bn.y);
      const double sgq = (r & 1) ? -fr2 : fr2;
      qbr += sgq * (an.x - bn.x);
      qbi += sgq * (an.y - bn.y);
      sasy += nt.y * (an.x * a_next.x + an.y * a_next.y + bn.x * b_next.x + bn.y * b_next.y) + nt.x * (an.x * bn.x + an.y * bn.y);
      a_next = an;
      b_next = bn;
      const double f = nt.x * sw;
      double* row = tile + (r - 1) * GM_SB + 2 * lane;
      *reinterpret_cast<double2*>(row) = make_double2(wnz ? (an.x + bn.x) * f : 0.0, wnz ? (an.y + bn.y) * f : 0.0);
      *reinterpret_cast<double2*>(row + 64) = make_double2(wnz ? (an.x - bn.x) * f : 0.0, wnz ? (an.y - bn.y) * f : 0.0);
    }
    // ---- per-particle efficiencies (mie_props.py:44-68) and this lane's share of the size-distribution sums
    {
      const double iy2 = xinv * xinv;
      const double qext = act ? 2.0 * sext * iy2 : 0.0, qsca = act ? 2.0 * ssca * iy2 : 0.0;
      const double qb = act ? (qbr * qbr + qbi * qbi) * iy2 : 0.0, gq = act ? 4.0 * iy2 * sasy : 0.0;
      const double x2 = xi * xi, x3 = x2 * xi, x4 = x2 * x2;
      const double w = valid ? ws : 0.0;
      const bool on = act && w != 0.0;
      const double x2w = x2 * w, x4w = x4 * w;
      S[GM_S_W] += w;
      S[GM_S_X2W] += valid ? x2w : 0.0;
      S[GM_S_X3W] += valid ? x3 * w : 0.0;
      S[GM_S_X4W] += valid ? x4w : 0.0;
      S[GM_S_QEXT] += on ? qext * x2w : 0.0;
      S[GM_S_QSCA] += on ? qsca * x2w : 0.0;
      S[GM_S_QABS] += on ? (qext - qsca) * x2w : 0.0;
      S[GM_S_QB] += on ? qb * x2w : 0.0;
      S[GM_S_G] += on ? gq * x2w : 0.0;
      S[GM_S_CSCA] += on ? qsca * qsca * x4w : 0.0;
      S[GM_S_CEXT] += on ? qext * qsca * x4w : 0.0;
    }
    __syncwarp();
    // ---- the tile times its own transpose on the FP64 tensor cores (same fragments as gram_cta<0> / <1>)
    const double* st = tile + lane_off;
    if constexpr (STACK) {
#pragma unroll 4
      for (int ks = 0; ks < 16; ++ks) {
        const double y = st[4 * ks];
        const double yx = __shfl_xor_sync(0xffffffffu, y, 1);
        const double yt = lr < 4 ? sgn * yx : 0.0;
        dmma884(acc[0][0], acc[0][1], y, y);    // Y Y^T: H1 | H3 / . | H2
        dmma884(acc[1][0], acc[1][1], yt, y);   // [X~+; 0] Y^T: . | H4
      }
    } else {
#pragma unroll 4
      for (int ks = 0; ks < 16; ++ks) {
        const double fp = st[4 * ks], fm = st[64 + 4 * ks];
        const double ft = sgn * __shfl_xor_sync(0xffffffffu, fp, 1);
        dmma884(acc[0][0], acc[0][1], fp, fp);   // H1
        dmma884(acc[1][0], acc[1][1], fm, fm);   // H2
        dmma884(acc[2][0], acc[2][1], fp, fm);   // H3
        dmma884(acc[3][0], acc[3][1], ft, fm);   // H4
      }
    }
    __syncwarp();   // the tile is rewritten by the next group
  }

  // ---- partial Gram blocks of this work item (slot = segment), layout of gram_cta's end-of-task store
  constexpr int N = STACK ? 4 : 8;
  double* hp = A.hpart + (size_t)task * A.hstride + sg.hoff + (size_t)sg.seg * 4 * N * N;
  if constexpr (STACK) {
    const int r4 = lr & 3, c4 = 2 * (lk & 1);
    const double2 gq = make_double2(acc[0][0], acc[0][1]), tq = make_double2(acc[1][0], acc[1][1]);
    if (lr < 4 && lk < 2) *reinterpret_cast<double2*>(hp + (0 * 4 + r4) * 4 + c4) = gq;      // H1
    if (lr >= 4 && lk >= 2) *reinterpret_cast<double2*>(hp + (1 * 4 + r4) * 4 + c4) = gq;    // H2
    if (lr < 4 && lk >= 2) {
      *reinterpret_cast<double2*>(hp + (2 * 4 + r4) * 4 + c4) = gq;                          // H3
      *reinterpret_cast<double2*>(hp + (3 * 4 + r4) * 4 + c4) = tq;                          // H4
    }
  } else {
#pragma unroll
    for (int q = 0; q < 4; ++q)
      *reinterpret_cast<double2*>(hp + ((size_t)q * N + lr) * N + 2 * lk) = make_double2(acc[q][0], acc[q][1]);
  }
  // ---- scalar sums: one 16-value warp reduction per work item, into the slot of the segment's first group; zeros elsewhere
  {
    double v[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) v[k] = k < GM_NSCAL ? S[k] : 0.0;
    const double r = warp_reduce16(v, lane);
    const int s = warp_reduce16_index(lane);
    double* base = A.scal_part + (size_t)task * A.ngroup * GM_NSCAL;
    if (!(lane & 1) && s < GM_NSCAL) base[(size_t)A.glist[sg.gbegin] * GM_NSCAL + s] = r;
    for (int gi = sg.gbegin + 1; gi < sg.gend; ++gi)
      if (lane < GM_NSCAL) base[(size_t)A.glist[gi] * GM_NSCAL + lane] = 0.0;
  }
  if (A.stats) {
    const unsigned ne = __reduce_add_sync(0xffffffffu, st_ev), snm = __reduce_add_sync(0xffffffffu, st_nm),
                   snx = __reduce_add_sync(0xffffffffu, st_nx), sneg = __reduce_add_sync(0xffffffffu, st_neg);
    if (lane == 0) {
      if (sneg) atomicAdd(&A.stats[5], (unsigned long long)sneg);
      atomicAdd(&A.stats[0], (unsigned long long)ne);
      atomicAdd(&A.stats[1], (unsigned long long)snm);
      atomicAdd(&A.stats[2], (unsigned long long)snx);
      atomicAdd(&A.stats[3], (unsigned long long)st_k4 * (STACK ? 1 : 2));
    }
  }
}

// grid = (ceil(ntask / 4), segments of the class), block = 128: the four warps of a CTA take four consecutive tasks of the same
// segment, so that the Bessel rows, x, 1/x and nmax of its groups are shared through L1.  One instantiation per class: the 4-order
// body fits 128 registers (16 warps per SM), the 8-order body needs 168 (12 warps per SM; at 128 it spills 300 B).
template <int ROWS, int MINB>
__global__ void __launch_bounds__(GM_SMALL_WARPS * 32, MINB) k_small(SmallArgs A) {
  __shared__ __align__(16) double tiles[GM_SMALL_WARPS][ROWS * GM_SB];
  const int warp = threadIdx.x >> 5;
  const int task = blockIdx.x * GM_SMALL_WARPS + warp;
  if (task >= A.ntask) return;
  small_item<ROWS>(A, A.segs[blockIdx.y], task, tiles[warp]);
}
