// gm_mie_kernels.cuh -- sm_100a device code of the Mie lookup-table hot path.
//
//   k_bessel     Riccati-Bessel psi_n(x), chi_n(x) per particle      (replaces scipy jv/yv pre-computation,
//                                                                     mie_coated.py:129-158)
//   k_pt_table   p_n = pi_n + tau_n, q_n = pi_n - tau_n per angle     (replaces preCalculatePT, mie_coated.py:160-179,
//                                                                     recurrences of mie_props.py:166-192)
//   k_coeff      D_n downward recurrence, a_n, b_n, efficiencies     (single_mie_coeff, mie_coeffs.py:83-130;
//                and the size-distribution scalar sums                mie_props_raw, mie_props.py:28-70;
//                                                                     integratePSD scalar part, dointegration.py:1104-1190)
//   k_contract   S+/S- angular contraction on the FP64 tensor cores  (mie_S12_backend_pt, mie_props.py:133-150;
//                (DMMA m8n8k4) + Mueller products + PSD reduction      calculateScatVals, dointegration.py:1044-1050;
//                                                                     phase part of integratePSD, :1164-1166)
//   k_finalize   deterministic reduction of the partial sums
//   k_s12_direct one warp per particle, lanes over angles             (mie_S12_backend, mie_props.py:119-131)
#pragma once
#include "gm_common.cuh"

// ================================================================================================ k_bessel
// One thread per particle.  chi_n by upward recurrence (stable).  psi_n upward while n + 1/2 < x (oscillatory
// region, stable), then psi_n = rho_n psi_{n-1} with the ratios rho_n = psi_n/psi_{n-1} obtained by the downward
// continued-fraction recurrence rho_n = 1/((2n+1)/x - rho_{n+1}) (stable for n + 1/2 >= x, where psi has no zeros).
// Table layout: value of particle (g*32 + lane) and order n at gboff[g] + n*32 + lane  (n = 0..nmax).
__global__ void __launch_bounds__(128) k_bessel(int nx, const double* __restrict__ x, const int* __restrict__ nmax,
                                                const long long* __restrict__ gboff, double* __restrict__ psi,
                                                double* __restrict__ chi) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nx) return;
  const double xi = x[i];
  const int nm = nmax[i];
  const size_t base = (size_t)gboff[i >> 5] + (i & 31);
#define BIDX(n) (base + (size_t)(n) * 32)
  double s, c;
  sincos(xi, &s, &c);
  const double xinv = 1.0 / xi;
  // chi upward: chi_0 = cos x, chi_1 = cos x / x + sin x, chi_{n+1} = (2n+1)/x chi_n - chi_{n-1}
  double c0 = c, c1 = c * xinv + s;
  chi[BIDX(0)] = c0;
  if (nm >= 1) chi[BIDX(1)] = c1;
  for (int n = 1; n < nm; ++n) {
    double c2 = (2 * n + 1) * xinv * c1 - c0;
    chi[BIDX(n + 1)] = c2;
    c0 = c1;
    c1 = c2;
  }
  // psi
  int nt = (int)ceil(xi - 0.5);
  if (nt < 0) nt = 0;
  if (nt > nm) nt = nm;
  double p0 = s, p1 = s * xinv - c;
  psi[BIDX(0)] = p0;
  if (nt >= 1) {
    psi[BIDX(1)] = p1;
    for (int n = 1; n < nt; ++n) {
      double p2 = (2 * n + 1) * xinv * p1 - p0;
      psi[BIDX(n + 1)] = p2;
      p0 = p1;
      p1 = p2;
    }
  }
  if (nt < nm) {
    int N = nm + (int)(4.3 * cbrt(xi)) + 20;
    double r = 0.0;
    for (int n = N; n > nt; --n) {
      r = 1.0 / ((2 * n + 1) * xinv - r);
      if (n <= nm) psi[BIDX(n)] = r;
    }
    double pp = psi[BIDX(nt)];
    for (int n = nt + 1; n <= nm; ++n) {
      pp = psi[BIDX(n)] * pp;
      psi[BIDX(n)] = pp;
    }
  }
#undef BIDX
}

// 1 / x per particle (the quotient k_coeff forms per task; k_small reads it)
__global__ void __launch_bounds__(128) k_xinv(int nx, const double* __restrict__ x, double* __restrict__ xinv) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < nx) xinv[i] = 1.0 / x[i];
}

// caller-supplied J_{k+1/2}(x), Y_{k+1/2}(x) -> psi, chi (mie_coeffs.py:106-112; psi_0 = sin x, chi_0 = cos x)
__global__ void __launch_bounds__(128) k_bessel_from_jy(int nx, const double* __restrict__ x, const int* __restrict__ nmax,
                                                        const long long* __restrict__ gboff, const long long* __restrict__ off,
                                                        const double* __restrict__ jv, const double* __restrict__ yv,
                                                        double* __restrict__ psi, double* __restrict__ chi) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nx) return;
  const double xi = x[i];
  const int nm = nmax[i];
  const size_t base = (size_t)gboff[i >> 5] + (i & 31);
  const double sx = sqrt(0.5 * 3.141592653589793238462643383279502884 * xi);
  double s, c;
  sincos(xi, &s, &c);
  psi[base] = s;
  chi[base] = c;
  for (int n = 1; n <= nm; ++n) {
    psi[base + (size_t)n * 32] = sx * jv[off[i] + n];
    chi[base + (size_t)n * 32] = -sx * yv[off[i] + n];
  }
}

// ================================================================================================ k_pt_table
// One thread per padded angle column.  Row n-1 of half h: T[h][n-1][0][col] = pi_n + tau_n, T[h][n-1][1][col] = pi_n - tau_n
// (un-normalised; the (2n+1)/(n(n+1)) factor of mie_ptnumba, mie_props.py:217-231, is folded into the coefficients).
__global__ void __launch_bounds__(128) k_pt_table(int nang, const double* __restrict__ cost, int nrows, double* __restrict__ T) {
  int a = blockIdx.x * blockDim.x + threadIdx.x;
  if (a >= GM_NANG_PAD) return;
  int half = a / GM_HALF_ANG, col = a % GM_HALF_ANG;
  double* base = T + (size_t)half * nrows * GM_TROW + col;
  if (a >= nang) {
    for (int n = 1; n <= nrows; ++n) {
      base[(size_t)(n - 1) * GM_TROW] = 0.0;
      base[(size_t)(n - 1) * GM_TROW + GM_LAH] = 0.0;
    }
    return;
  }
  const double u = cost[a];
  double pm2 = 1.0;        // pi_1
  double pm1 = 3.0 * u;    // pi_2
  double t1 = u, t2 = 6.0 * u * u - 3.0;
  base[0] = pm2 + t1;
  base[GM_LAH] = pm2 - t1;
  if (nrows >= 2) {
    base[GM_TROW] = pm1 + t2;
    base[GM_TROW + GM_LAH] = pm1 - t2;
  }
  for (int n = 3; n <= nrows; ++n) {
    double dn = (double)n;
    double pn = (2.0 * dn - 1.0) / (dn - 1.0) * pm1 * u - dn / (dn - 1.0) * pm2;   // mie_p, mie_props.py:166-177
    double tn = dn * u * pn - (dn + 1.0) * pm1;                                     // mie_t, mie_props.py:183-192
    base[(size_t)(n - 1) * GM_TROW] = pn + tn;
    base[(size_t)(n - 1) * GM_TROW + GM_LAH] = pn - tn;
    pm2 = pm1;
    pm1 = pn;
  }
}

// ================================================================================================ k_coeff
struct CoeffArgs {
  int nx, ngroup;
  const double* x;
  const int* nmax;
  const double* psi;
  const double* chi;
  const long long* gboff;
  const double2* mz;    // sqrt(eps*mu)
  const double2* mrel;  // sqrt(eps/mu)
  int mat_per_particle; // 0: indexed by task (blockIdx.y), 1: indexed by particle
  // ---- table mode
  const double* wphase;  // [ntask][nx]
  const double* wscal;   // [ntask][nmode][nx] or null
  int nmode;
  int dense;             // 1: evaluate zero-weight particles too
  int scale_sqrtw;       // 1: fold sqrt(w_phase) into the coefficients
  const int* grow;       // first coefficient row of group g
  const int* gk4;        // k4 steps of group g
  double* coef;          // [ntask][task_rows][GM_SB]
  long long task_stride; // doubles
  unsigned char* gact;   // [ntask][ngroup]
  double* scal_part;     // [ntask][nmode][ngroup][GM_NSCAL]
  // ---- natural mode / optional per-particle outputs
  const long long* aboff;  // prefix sum of nmax
  double4* ab;             // [sum nmax] a_n, b_n (MODE 1: output; MODE 2: input, + task * ab_stride)
  long long ab_stride;
  double* q;               // [ntask][nx][6] (nullable)
  unsigned long long* stats;  // [0] evals [1] sum nmax [2] sum nmx [3] padded k4 steps
  int ntask;                  // tasks of this launch (bound of the task loop when GM_COEFF_TPC > 1)
  const double2* ntab;        // [n] = ((2n+1)/(n(n+1)), n(n+2)/(n+1)): the order-dependent factors of mie_props_raw, read with a
                              // warp-uniform index instead of two reciprocals per order
  const int* gsel;            // nullable: the kernel works on the groups gsel[0..nsel) only (the others belong to k_small)
  int nsel;
};

// MODE 0: DMMA group layout (c+ = (a+b) f_n sqrt(w), c- = (a-b) f_n sqrt(w)); MODE 1: natural a_n, b_n;
// MODE 2: like MODE 0 but a_n, b_n are READ from the natural layout (coated spheres, produced by k_coated_coeff).
#ifndef GM_COEFF_MINB
#define GM_COEFF_MINB 4   // 16 warps per SM without spills; 5 / 6 (96 / 80 registers, 140 / 328 B of spills) run k_coeff in 99.8 / 126.6 ms instead of 67.8 ms on optics_SS (profiles/r02e_su_tuning_sweep.txt)
#endif
// GM_COEFF_TPC (experiment, default 1 = one task per CTA, unchanged code): tasks a CTA works through for its four particle
// groups.  With 1 the optics_SU step launches 76,860 CTAs of ~8 us each for this kernel; > 1 makes them proportionally fewer and
// longer and lets the Riccati-Bessel rows, x and nmax of the group stay in L1 / registers across the tasks.  Prepared for the
// rank-dependent slow mode of this kernel in multi-GPU jobs; not yet measured.
#ifndef GM_COEFF_TPC
#define GM_COEFF_TPC 1
#endif
#ifndef GM_COEFF_SHORT_START
#define GM_COEFF_SHORT_START 0
#endif
#ifndef GM_COEF_STCS
#define GM_COEF_STCS 1   // coefficient stream written with st.global.cs (evict-first): optics_SS k_coeff 120.6 -> 101.9 ms, optics_SU unchanged
#endif
// Grid: x = task (fastest: CTAs that are resident together work on the same four groups of consecutive tasks and share their Bessel rows
// through L1 / L2), y = quad of groups in DESCENDING order -- on a sorted size grid the cost of a group grows with its index (nmax ~ x), so
// the longest CTAs start first and the last wave holds the cheapest ones.  (Round 1 / early round 2: x = quad ascending, y = task; ncu of an
// optics_SS bin-5 launch showed the SMs active 74 % of the elapsed cycles: the tail was the largest groups of the last tasks.)
#if GM_COEFF_TPC == 1
#define GM_COEFF_TASK_LOOP const int task = blockIdx.x;
#define GM_COEFF_NEXT_TASK return
#else
#define GM_COEFF_TASK_LOOP \
  for (int task = blockIdx.x * GM_COEFF_TPC; task < min(A.ntask, (int)(blockIdx.x + 1) * GM_COEFF_TPC); ++task)
#define GM_COEFF_NEXT_TASK continue
#endif
// MINB = CTAs of 128 threads per SM: 4 (128 registers, 16 warps per SM) for short groups, GM_COEFF_MINB_LONG = 3 (168 registers, 12 warps,
// no spills) for launches whose groups are long: on optics_SS (mean nmax of the launched groups in the hundreds) the 168-register build
// runs 58.5 ms instead of 67.8 ms, on optics_SU (nmax 9..40) 0.445 instead of 0.428 ms; 2 (194 registers) is back at 67 ms
// (profiles/r02e_su_tuning_sweep.txt).  The host picks the instantiation per run from the mean number of rows per launched group.
#ifndef GM_COEFF_MINB_LONG
#define GM_COEFF_MINB_LONG 3
#endif
#ifndef GM_COEFF_PD_LONG
#define GM_COEFF_PD_LONG 4         // depth of the Riccati-Bessel prefetch ring (= unroll factor of the order loop) of the long build
#endif
#ifndef GM_COEFF_LONG_ROWS
#define GM_COEFF_LONG_ROWS 64     // mean coefficient rows (orders, padded to 4) per launched group from which the long build is used
#endif
template <int MODE, int MINB = GM_COEFF_MINB>
__global__ void __launch_bounds__(128, MINB) k_coeff(CoeffArgs A) {
  int g = (int)(gridDim.y - 1 - blockIdx.y) * 4 + (threadIdx.x >> 5);
  if (A.gsel) {
    if (g >= A.nsel) return;
    g = A.gsel[g];
  }
  if (g >= A.ngroup) return;
  const int lane = threadIdx.x & 31;
  const int i = g * 32 + lane;
  const bool valid = i < A.nx;
  const double xi = valid ? A.x[i] : 1.0;
  const int nm = valid ? A.nmax[i] : 0;
  GM_COEFF_TASK_LOOP {
  const int mi = A.mat_per_particle ? (valid ? i : 0) : task;
  const double2 mzv = A.mz[mi];
  const double2 mrv = A.mrel[mi];

  constexpr bool TABLE = (MODE != 1);
  double wp = 1.0;
  bool any = true;
  if (TABLE) {
    wp = valid ? A.wphase[(size_t)task * A.nx + i] : 0.0;
    any = wp != 0.0;
    if (A.wscal)
      for (int k = 0; k < A.nmode; ++k) any |= valid && A.wscal[((size_t)task * A.nmode + k) * A.nx + i] != 0.0;
  }
  const bool act = valid && (MODE == 1 || A.dense || any);
  if (TABLE && A.stats && A.scale_sqrtw) {
    const unsigned nneg = __popc(__ballot_sync(0xffffffffu, wp < 0.0));
    if (nneg && lane == 0) atomicAdd(&A.stats[5], (unsigned long long)nneg);     // reported as GM_EINVAL by the host-buffer calls
  }
  const double2 z = make_double2(mzv.x * xi, mzv.y * xi);                     // mie_coeffs.py:96
#if GM_COEFF_SHORT_START
  // experiment (default off): start the downward recurrence 4 + ceil(2.5 |z|) orders (at most the reference's 16) above
  // max(nmax, |z|).  The error of D_n from starting with D = 0 falls by ~(|z| / 2j)^2 per order, so for |z| << 1 three to four
  // orders already give D_n to 1 ulp (profiles/r01k_coeff_start_offset_study.txt); the reference always takes 16 (mie_coeffs.py:101).
  const double zabs = sqrt(fma(z.x, z.x, z.y * z.y));
  const int nmx = (act && MODE != 2) ? (int)rint(fmax((double)nm, zabs) + fmin(16.0, 4.0 + ceil(2.5 * zabs))) : 0;
#else
  const int nmx = (act && MODE != 2) ? (int)rint(fmax((double)nm, sqrt(fma(z.x, z.x, z.y * z.y))) + 16.0) : 0;  // mie_coeffs.py:101
#endif
  const int J = __reduce_max_sync(0xffffffffu, nmx);
  int rows = 0;
  if (TABLE) {
    rows = 4 * A.gk4[g];
    const bool gactive = __any_sync(0xffffffffu, act);
    if (lane == 0) A.gact[(size_t)task * A.ngroup + g] = gactive ? 1 : 0;
    if (!gactive) {
      for (int k = lane; k < A.nmode * GM_NSCAL; k += 32)
        A.scal_part[(((size_t)task * A.nmode + k / GM_NSCAL) * A.ngroup + g) * GM_NSCAL + k % GM_NSCAL] = 0.0;
      GM_COEFF_NEXT_TASK;
    }
  }
  const size_t bbase = (size_t)A.gboff[g] + lane;
  const double2 zinv = crcp(z);
  const double2 minv = crcp(mrv);
  const double xinv = 1.0 / xi;
  const double sw = (TABLE && A.scale_sqrtw) ? sqrt(wp) : 1.0;
  double* crow = nullptr;
  if (TABLE) crow = A.coef + (size_t)task * A.task_stride + (size_t)A.grow[g] * GM_SB + 2 * lane;
  const long long abo = (MODE != 0 && valid) ? A.aboff[i] : 0;
  const double4* ab_in = (MODE == 2) ? A.ab + (size_t)task * A.ab_stride : nullptr;

  double2 D = make_double2(0.0, 0.0);
  double psi_n = 0.0, chi_n = 0.0;
  double2 a_next = make_double2(0.0, 0.0), b_next = make_double2(0.0, 0.0);
  double sext = 0.0, ssca = 0.0, qbr = 0.0, qbi = 0.0, sasy = 0.0;
  // Riccati-Bessel values are fetched PD orders ahead of their use (the table read is the only memory access on the
  // serial recurrence's critical path); the first PD+1 orders are requested before the recurrence starts.
  constexpr int PD = (MODE == 0 && MINB == GM_COEFF_MINB_LONG && GM_COEFF_MINB_LONG != GM_COEFF_MINB) ? GM_COEFF_PD_LONG : 4;
  double qpsi[PD], qchi[PD];
#pragma unroll
  for (int k = 0; k < PD; ++k) qpsi[k] = qchi[k] = 0.0;
#ifndef GM_COEFF_BRANCHY
  if (act && MODE == 1) {
#else
  if (act && MODE != 2) {
#endif
    psi_n = A.psi[bbase + (size_t)nm * 32];
    chi_n = A.chi[bbase + (size_t)nm * 32];
#pragma unroll
    for (int k = 0; k < PD; ++k)
      if (nm - 1 - k >= 0) {
        qpsi[k] = A.psi[bbase + (size_t)(nm - 1 - k) * 32];
        qchi[k] = A.chi[bbase + (size_t)(nm - 1 - k) * 32];
      }
  }

  // ---- phase 1: orders above every particle of the group: only the logarithmic-derivative recurrence
  // mie_coeffs.py:119-121, D_n = r - 1/(D_{n+1} + r), r = (n+1)/z, started from D_{nmx} = 0.  It is carried in the
  // denominator t_n = D_{n+1} + (n+1)/z:  D_n = (n+1)/z - 1/t_n  and  t_{n-1} = D_n + n/z = (2n+1)/z - 1/t_n, i.e. one
  // complex reciprocal and two FMAs per order (D_n itself is only formed for the orders that emit coefficients).
  int nstart = J - 1;
  const int nemit = TABLE ? rows : __reduce_max_sync(0xffffffffu, act ? nm : 0);
  if (nemit > nstart) nstart = nemit;
  int n = nstart;
  double2 tt = make_double2((double)nmx * zinv.x, (double)nmx * zinv.y);   // t_{nmx-1}: D_{nmx} = 0
  double f2 = (double)(2 * nstart + 1);                                     // 2n + 1, warp-uniform, exact
  for (; n > nemit; --n, f2 -= 2.0) {
    if (n < nmx) {
      const double2 ti = crcp(tt);
      tt = make_double2(fma(f2, zinv.x, -ti.x), fma(f2, zinv.y, -ti.y));
    }
  }
#ifndef GM_COEFF_BRANCHY
  if constexpr (MODE == 0) {
    // ---- phase 2, table mode (round 2): ONE branch-free loop body.  The round-1 body below has two divergent regions per order (177
    // SASS instructions and two BSSY / BSYNC pairs per order); here every lane computes every order of the group's tile and selects
    // decide what counts: `on` = the lane's recurrence has started (n < nmx), `emit` = the order exists for the lane (n <= nmax).  The
    // Riccati-Bessel values come from a PD-deep ring that is loaded for every lane whatever its own nmax (rows up to the group's largest
    // nmax exist in the table; rows above a lane's nmax hold zeros), so the loads never depend on a predicate of the arithmetic.
    // No bounds checks on the ring loads: the tables have GM_BESSEL_SLACK_ROWS rows of slack at both ends, rows above the group's
    // largest nmax belong to the next group and rows below 0 to the previous one -- what is read there only ever feeds orders that no
    // lane emits (n > nmax) or that do not exist (n < 1).
    const double* pp = A.psi + bbase + (size_t)(n - 1) * 32;     // row of order n - 1
    const double* pc = A.chi + bbase + (size_t)(n - 1) * 32;
    double rp[PD], rc[PD];
#pragma unroll
    for (int k = 0; k < PD; ++k) {
      rp[k] = pp[-k * 32];
      rc[k] = pc[-k * 32];
    }
    double psi_c = pp[32], chi_c = pc[32];
    pp -= PD * 32;                                                // row of order n - 1 - PD: the next one the ring takes in
    pc -= PD * 32;
    // a zero-weight particle of a dense run adds exact zeros to every sum, even where its coefficients are not finite: not emitted
    // (unless the per-particle efficiencies are asked for, gm_table_particles: then every particle is evaluated and its rows carry f = 0)
    // `any`: some weight of the lane is non-zero -- the phase weight may be 0 where a (signed) scalar weight is not ('du' grid)
    const bool lane_on = act && (any || A.q != nullptr);
#pragma unroll(PD)
    for (; n >= 1; --n, f2 -= 2.0, pp -= 32, pc -= 32) {
      const double2 ti = crcp(tt);
      const bool on = n < nmx;
      const double f1 = 0.5 * f2 + 0.5;                                     // n + 1
      D.x = on ? fma(f1, zinv.x, -ti.x) : D.x;
      D.y = on ? fma(f1, zinv.y, -ti.y) : D.y;
      tt.x = on ? fma(f2, zinv.x, -ti.x) : tt.x;
      tt.y = on ? fma(f2, zinv.y, -ti.y) : tt.y;
      const bool emit = lane_on && n <= nm;
      const double psi_m = rp[0], chi_m = rc[0];                             // order n - 1
#pragma unroll
      for (int k = 0; k + 1 < PD; ++k) {
        rp[k] = rp[k + 1];
        rc[k] = rc[k + 1];
      }
      rp[PD - 1] = *pp;
      rc[PD - 1] = *pc;
      const double nox = (0.5 * f2 - 0.5) * xinv;                            // n / x
      double2 da = cmul(D, minv);                                            // mie_coeffs.py:124
      da.x += nox;
      double2 db = cmul(D, mrv);                                             // mie_coeffs.py:125
      db.x += nox;
      // a_n = (da psi_n - psi_{n-1}) / (da xi_n - xi_{n-1}),  xi = psi - i chi      (mie_coeffs.py:113-114,127-128)
      double2 an = cdiv(make_double2(fma(da.x, psi_c, -psi_m), da.y * psi_c),
                        make_double2(fma(da.x, psi_c, fma(da.y, chi_c, -psi_m)), fma(da.y, psi_c, fma(-da.x, chi_c, chi_m))));
      double2 bn = cdiv(make_double2(fma(db.x, psi_c, -psi_m), db.y * psi_c),
                        make_double2(fma(db.x, psi_c, fma(db.y, chi_c, -psi_m)), fma(db.y, psi_c, fma(-db.x, chi_c, chi_m))));
      psi_c = psi_m;
      chi_c = chi_m;
      an.x = emit ? an.x : 0.0;
      an.y = emit ? an.y : 0.0;
      bn.x = emit ? bn.x : 0.0;
      bn.y = emit ? bn.y : 0.0;
      // efficiencies, mie_props.py:41-65 (orders above a lane's nmax add exact zeros)
      const double2 nt = __ldg(A.ntab + n);
      sext += f2 * (an.x + bn.x);
      ssca += f2 * (an.x * an.x + an.y * an.y + bn.x * bn.x + bn.y * bn.y);
      const double sg = (n & 1) ? -f2 : f2;
      qbr += sg * (an.x - bn.x);
      qbi += sg * (an.y - bn.y);
      sasy += nt.y * (an.x * a_next.x + an.y * a_next.y + bn.x * b_next.x + bn.y * b_next.y) + nt.x * (an.x * bn.x + an.y * bn.y);
      a_next = an;
      b_next = bn;
      const double f = nt.x * sw;
      const double2 cp = make_double2((an.x + bn.x) * f, (an.y + bn.y) * f);
      const double2 cm = make_double2((an.x - bn.x) * f, (an.y - bn.y) * f);
#ifndef GM_COEFF_NOSTORE
      double* r = crow + (size_t)(n - 1) * GM_SB;
#if GM_COEF_STCS
      __stcs(reinterpret_cast<double2*>(r), cp);
      __stcs(reinterpret_cast<double2*>(r + 64), cm);
#else
      *reinterpret_cast<double2*>(r) = cp;
      *reinterpret_cast<double2*>(r + 64) = cm;
#endif
#endif
    }
  }
#endif
  // ---- phase 2: orders that emit coefficients (MODE 1 / 2; MODE 0 arrives here with n == 0)
  for (; n >= 1; --n, f2 -= 2.0) {
    if (n < nmx) {
      const double2 ti = crcp(tt);
      const double f1 = 0.5 * f2 + 0.5;                                     // n + 1
      D = make_double2(fma(f1, zinv.x, -ti.x), fma(f1, zinv.y, -ti.y));
      tt = make_double2(fma(f2, zinv.x, -ti.x), fma(f2, zinv.y, -ti.y));
    }
    double2 cp = make_double2(0.0, 0.0), cm = make_double2(0.0, 0.0);
    if (act && n <= nm) {
      double2 an, bn;
      const double dn = 0.5 * f2 - 0.5;   // n (f2 = 2n + 1 is carried as a double: no int -> double conversion per order)
      if (MODE == 2) {
        const double4 v = ab_in[abo + n - 1];
        an = make_double2(v.x, v.y);
        bn = make_double2(v.z, v.w);
      } else {
        const double psi_m = qpsi[0], chi_m = qchi[0];        // order n-1
#pragma unroll
        for (int k = 0; k + 1 < PD; ++k) {
          qpsi[k] = qpsi[k + 1];
          qchi[k] = qchi[k + 1];
        }
        if (n - 1 - PD >= 0) {
          qpsi[PD - 1] = A.psi[bbase + (size_t)(n - 1 - PD) * 32];
          qchi[PD - 1] = A.chi[bbase + (size_t)(n - 1 - PD) * 32];
        }
        const double nox = dn * xinv;
        double2 da = cmul(D, minv);                             // mie_coeffs.py:124
        da.x += nox;
        double2 db = cmul(D, mrv);                              // mie_coeffs.py:125
        db.x += nox;
        // a_n = (da psi_n - psi_{n-1}) / (da xi_n - xi_{n-1}),  xi = psi - i chi      (mie_coeffs.py:113-114,127-128)
        an = cdiv(make_double2(fma(da.x, psi_n, -psi_m), da.y * psi_n),
                  make_double2(fma(da.x, psi_n, fma(da.y, chi_n, -psi_m)), fma(da.y, psi_n, fma(-da.x, chi_n, chi_m))));
        bn = cdiv(make_double2(fma(db.x, psi_n, -psi_m), db.y * psi_n),
                  make_double2(fma(db.x, psi_n, fma(db.y, chi_n, -psi_m)), fma(db.y, psi_n, fma(-db.x, chi_n, chi_m))));
        psi_n = psi_m;
        chi_n = chi_m;
      }
      // efficiencies, mie_props.py:41-65
      const double cn = f2;                                   // 2n + 1
      const double2 nt = __ldg(A.ntab + n);
      const double c2n = nt.x;                                // (2n+1)/(n(n+1))
      sext += cn * (an.x + bn.x);
      ssca += cn * (an.x * an.x + an.y * an.y + bn.x * bn.x + bn.y * bn.y);
      const double sg = (n & 1) ? -cn : cn;
      qbr += sg * (an.x - bn.x);
      qbi += sg * (an.y - bn.y);
      sasy += nt.y * (an.x * a_next.x + an.y * a_next.y + bn.x * b_next.x + bn.y * b_next.y) +
              c2n * (an.x * bn.x + an.y * bn.y);
      a_next = an;
      b_next = bn;
      if (TABLE) {
        const double f = c2n * sw;
        if (sw != 0.0) {   // a zero-weight particle of a dense run adds exact zeros even where its coefficients are not finite
          cp = make_double2((an.x + bn.x) * f, (an.y + bn.y) * f);
          cm = make_double2((an.x - bn.x) * f, (an.y - bn.y) * f);
        }
      } else {
        A.ab[abo + n - 1] = make_double4(an.x, an.y, bn.x, bn.y);
      }
    }
#ifndef GM_COEFF_NOSTORE   // (diagnostic builds only: the kernel without its coefficient-stream stores)
    if (TABLE) {
      double* r = crow + (size_t)(n - 1) * GM_SB;
#if GM_COEF_STCS   // streaming (evict-first) stores: the stream is written once and read once, several L2 capacities later
      __stcs(reinterpret_cast<double2*>(r), cp);
      __stcs(reinterpret_cast<double2*>(r + 64), cm);
#else
      *reinterpret_cast<double2*>(r) = cp;
      *reinterpret_cast<double2*>(r + 64) = cm;
#endif
    }
#endif
  }
  // mie_props.py:44-68
  // (divisions by y^2 and qsca as multiplications by reciprocals: <= 2 ulp from the reference's quotients)
  double qv[6] = {0, 0, 0, 0, 0, 0};
  double gq = 0.0;   // asy * qsca
  if (act) {
    const double iy2 = xinv * xinv;
    qv[0] = 2.0 * sext * iy2;
    qv[1] = 2.0 * ssca * iy2;
    qv[2] = qv[0] - qv[1];
    qv[3] = (qbr * qbr + qbi * qbi) * iy2;
    gq = 4.0 * iy2 * sasy;
    if (A.q) {
      const double r1 = fast_rcp(qv[1]);
      qv[4] = gq * r1;
      qv[5] = qv[3] * r1;
    }
  }
  if (A.q && valid) {
    double* qo = A.q + ((size_t)task * A.nx + i) * 6;
#pragma unroll
    for (int k = 0; k < 6; ++k) qo[k] = qv[k];
  }
  if (A.stats) {
    unsigned ne = __popc(__ballot_sync(0xffffffffu, act));
    int snm = __reduce_add_sync(0xffffffffu, act ? nm : 0);
    int snx = __reduce_add_sync(0xffffffffu, act ? nmx : 0);
    if (lane == 0) {
      atomicAdd(&A.stats[0], (unsigned long long)ne);
      atomicAdd(&A.stats[1], (unsigned long long)snm);
      atomicAdd(&A.stats[2], (unsigned long long)snx);
      if (TABLE) atomicAdd(&A.stats[3], (unsigned long long)A.gk4[g]);
    }
  }
  if (TABLE) {
    // size-distribution scalar sums as warp-shuffle reductions (dointegration.py:1104-1107, :1133-1190)
    const double x2 = xi * xi, x3 = x2 * xi, x4 = x2 * x2;
    for (int k = 0; k < A.nmode; ++k) {
      double w = 0.0;
      if (valid) w = A.wscal ? A.wscal[((size_t)task * A.nmode + k) * A.nx + i] : wp;
      double v[16];
#pragma unroll
      for (int s = GM_NSCAL; s < 16; ++s) v[s] = 0.0;
      const bool on = act && w != 0.0;   // w == 0 adds exact zeros (also where a dense run met a non-finite efficiency)
      const double x2w = x2 * w, x4w = x4 * w;
      v[GM_S_W] = valid ? w : 0.0;
      v[GM_S_X2W] = valid ? x2w : 0.0;
      v[GM_S_X3W] = valid ? x3 * w : 0.0;
      v[GM_S_X4W] = valid ? x4w : 0.0;
      v[GM_S_QEXT] = on ? qv[0] * x2w : 0.0;
      v[GM_S_QSCA] = on ? qv[1] * x2w : 0.0;
      v[GM_S_QABS] = on ? qv[2] * x2w : 0.0;
      v[GM_S_QB] = on ? qv[3] * x2w : 0.0;
      v[GM_S_G] = on ? gq * x2w : 0.0;
      v[GM_S_CSCA] = on ? qv[1] * qv[1] * x4w : 0.0;
      v[GM_S_CEXT] = on ? qv[0] * qv[1] * x4w : 0.0;
      double* o = A.scal_part + (((size_t)task * A.nmode + k) * A.ngroup + g) * GM_NSCAL;
      const double r = warp_reduce16(v, lane);
      const int s = warp_reduce16_index(lane);
      if (!(lane & 1) && s < GM_NSCAL) o[s] = r;
    }
  }
  }   // GM_COEFF_TASK_LOOP
}

// ================================================================================================ k_contract
// mbarrier / bulk-copy (TMA 1-D) primitives
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  const uint32_t addr = smem_u32(bar);
  do {
    asm volatile(
        "{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
        : "=r"(ok)
        : "r"(addr), "r"(parity)
        : "memory");
  } while (!ok);
}
__device__ __forceinline__ bool mbar_test(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n.reg .pred p;\nmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

struct ContractArgs {
  int ntask, ngroup, nchunk, nrows;   // nrows = padded table rows
  const double* T;                    // [2][nrows][GM_TROW]
  const double* coef;                 // [ntask][task_rows][GM_SB]
  long long task_stride;
  const int* grow;
  const int* gk4;
  const unsigned char* gact;          // [ntask][ngroup]
  const unsigned char* gskip;         // [ngroup] or null: 1 = group is handled by the Gram path (gm_gram.cuh)
  const int* chunk_start;             // [nchunk + 1] group boundaries
  double* part;                       // [ntask][nchunk_total][4][GM_NANG_PAD]
  int nchunk_total;                   // nchunk + 1 when the Gram path adds its own slot of partial sums
  // per-particle variant
  double* s12;                        // [ntask][nx][nang][4]
  int nx, nang;
};

// One CTA = (task, angle half, chunk of particle groups).  12 warps; warp w owns angles [16w, 16w+16) of the half and
// all 32 particles of the current group:  D[16 x 64] += A[16 x 4] * B[4 x 64] per k4 step and per sign, as
// 2 (m-tiles) x 8 (n-tiles) x 2 (S+, S-) DMMA m8n8k4.  A = p/q table rows, B = coefficient rows, both streamed through a
// GM_STAGES-deep shared-memory ring by 1-D bulk copies (TMA) signalled through mbarriers.  The C fragment gives each
// lane (Re, Im) of S+ and S- of particle 4j + lane%4 at angle 8i + lane/4, so the Mueller products are lane-local.
template <bool PER_PARTICLE>
__global__ void __launch_bounds__(GM_CONTRACT_THREADS, 1) k_contract(ContractArgs A) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  double* stages = reinterpret_cast<double*>(smem_raw);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw + (size_t)GM_STAGES * GM_STAGE_DBL * 8);
  uint64_t* empty = full + GM_STAGES;
  int2* meta = reinterpret_cast<int2*>(empty + GM_STAGES);   // per group of the chunk: (k4 steps or 0 if inactive, first row)

  const int item = blockIdx.x;
  const int half = item & 1;
  const int chunk = (item >> 1) % A.nchunk;
  const int task = (item >> 1) / A.nchunk;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int lk = lane & 3, lr = lane >> 2;

  if (threadIdx.x == 0) {
    for (int s = 0; s < GM_STAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], GM_CONTRACT_WARPS);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  const int cs = A.chunk_start[chunk], ce = A.chunk_start[chunk + 1];
  const int ng = ce - cs;
  {
    const unsigned char* gact = A.gact + (size_t)task * A.ngroup;
    for (int g = threadIdx.x; g < ng; g += blockDim.x) {
      const bool on = !(A.gskip && A.gskip[cs + g]) && gact[cs + g];   // (gact is not written for the groups the Gram kernels own)
      meta[g] = make_int2(on ? A.gk4[cs + g] : 0, A.grow[cs + g]);
    }
  }
  __syncthreads();
  int nsteps = 0;
  for (int g = 0; g < ng; ++g) nsteps += meta[g].x;
  constexpr uint32_t TBYTES = GM_KSTEP * GM_TROW * 8, CBYTES = GM_KSTEP * GM_SB * 8;

#if GM_PRODUCER_WARP
  // ---- warp specialisation: the 4th warpgroup gives its registers back and one of its threads becomes the TMA producer
  if (warp >= GM_CONTRACT_WARPS) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 32;");
    if (warp == GM_CONTRACT_WARPS && lane == 0) {
      const double* Th = A.T + (size_t)half * A.nrows * GM_TROW;
      const double* coef_t = A.coef + (size_t)task * A.task_stride;
      int pg = 0, pk = 0;
      while (pg < ng && meta[pg].x == 0) ++pg;
      for (int pstep = 0; pstep < nsteps; ++pstep) {
        const int s = pstep % GM_STAGES;
        if (pstep >= GM_STAGES) mbar_wait(&empty[s], ((pstep / GM_STAGES) - 1) & 1);   // all 12 consumer warps released it
        double* dst = stages + (size_t)s * GM_STAGE_DBL;
        mbar_expect_tx(&full[s], TBYTES + CBYTES);
        bulk_g2s(dst, Th + (size_t)pk * GM_KSTEP * GM_TROW, TBYTES, &full[s]);
        bulk_g2s(dst + GM_KSTEP * GM_TROW, coef_t + ((size_t)meta[pg].y + (size_t)pk * GM_KSTEP) * GM_SB, CBYTES, &full[s]);
        if (++pk == meta[pg].x) {
          pk = 0;
          ++pg;
          while (pg < ng && meta[pg].x == 0) ++pg;
        }
      }
    }
    return;
  }
  asm volatile("setmaxnreg.inc.sync.aligned.u32 160;");
#else
  const double* Th = A.T + (size_t)half * A.nrows * GM_TROW;
  const double* coef_t = A.coef + (size_t)task * A.task_stride;
  // producer = thread 0 of warp 0: fetches as many steps as there are free stages, never blocking (a stage is free once
  // all 12 warps released its previous use: non-blocking mbarrier test)
  int pg = 0, pk = 0, pstep = 0;
  while (pg < ng && meta[pg].x == 0) ++pg;
  auto produce = [&](int consumed) {
    while (pstep < nsteps && pstep < consumed + GM_STAGES) {
      const int s = pstep % GM_STAGES;
      if (pstep >= GM_STAGES && !mbar_test(&empty[s], ((pstep / GM_STAGES) - 1) & 1)) break;
      double* dst = stages + (size_t)s * GM_STAGE_DBL;
      mbar_expect_tx(&full[s], TBYTES + CBYTES);
      bulk_g2s(dst, Th + (size_t)pk * GM_KSTEP * GM_TROW, TBYTES, &full[s]);
      bulk_g2s(dst + GM_KSTEP * GM_TROW, coef_t + ((size_t)meta[pg].y + (size_t)pk * GM_KSTEP) * GM_SB, CBYTES, &full[s]);
      ++pstep;
      if (++pk == meta[pg].x) {
        pk = 0;
        ++pg;
        while (pg < ng && meta[pg].x == 0) ++pg;
      }
    }
  };
  if (threadIdx.x == 0) produce(0);
#endif

  double accp[2][8][2], accm[2][8][2];
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) accp[i][j][0] = accp[i][j][1] = accm[i][j][0] = accm[i][j][1] = 0.0;
  double mu[2][4];   // per lane: sum w |S+|^2, sum w |S-|^2, sum w Re(S+ S-*), sum w Im(S+ S-*)
#pragma unroll
  for (int i = 0; i < 2; ++i) mu[i][0] = mu[i][1] = mu[i][2] = mu[i][3] = 0.0;

  const int a0 = warp * 16;
  int step = 0;
  for (int gi = 0; gi < ng; ++gi) {
    const int nk = meta[gi].x;
    if (nk == 0) continue;
    const int g = cs + gi;
    for (int k = 0; k < nk; ++k, ++step) {
      const int s = step % GM_STAGES;
#if !GM_PRODUCER_WARP
      if (threadIdx.x == 0) produce(step);
      __syncwarp();
#endif
      mbar_wait(&full[s], (step / GM_STAGES) & 1);
      const double* tb = stages + (size_t)s * GM_STAGE_DBL + lk * GM_TROW + a0 + lr;
      const double* cf = stages + (size_t)s * GM_STAGE_DBL + GM_KSTEP * GM_TROW + lk * GM_SB + lr;
      double ap[2], aq[2];
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        ap[i] = tb[8 * i];
        aq[i] = tb[GM_LAH + 8 * i];
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const double bp = cf[8 * j];
        const double bm = cf[64 + 8 * j];
#pragma unroll
        for (int i = 0; i < 2; ++i) {
          dmma884(accp[i][j][0], accp[i][j][1], ap[i], bp);
          dmma884(accm[i][j][0], accm[i][j][1], aq[i], bm);
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty[s]);
    }
    // group epilogue
#pragma unroll
    for (int i = 0; i < 2; ++i) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const double pr = accp[i][j][0], pi = accp[i][j][1], mr = accm[i][j][0], mi = accm[i][j][1];
        if (PER_PARTICLE) {
          const int pidx = g * 32 + 4 * j + lk;
          const int ang = half * GM_HALF_ANG + a0 + 8 * i + lr;
          if (pidx < A.nx && ang < A.nang) {
            double4 o = make_double4(0.5 * (pr + mr), 0.5 * (pi + mi), 0.5 * (pr - mr), 0.5 * (pi - mi));
            *reinterpret_cast<double4*>(A.s12 + (((size_t)task * A.nx + pidx) * A.nang + ang) * 4) = o;
          }
        } else {
          mu[i][0] = fma(pr, pr, fma(pi, pi, mu[i][0]));
          mu[i][1] = fma(mr, mr, fma(mi, mi, mu[i][1]));
          mu[i][2] = fma(pr, mr, fma(pi, mi, mu[i][2]));
          mu[i][3] = fma(pi, mr, fma(-pr, mi, mu[i][3]));
        }
        accp[i][j][0] = accp[i][j][1] = accm[i][j][0] = accm[i][j][1] = 0.0;
      }
    }
  }
  if (!PER_PARTICLE) {
    double* out = A.part + ((size_t)task * A.nchunk_total + chunk) * 4 * GM_NANG_PAD + half * GM_HALF_ANG + a0 + lr;
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        double v = mu[i][q];
        v += __shfl_xor_sync(0xffffffffu, v, 1);
        v += __shfl_xor_sync(0xffffffffu, v, 2);
        if (lk == 0) out[(size_t)q * GM_NANG_PAD + 8 * i] = v;
      }
  }
}

// ================================================================================================ k_finalize
// grid = ntask, block = GM_NANG_PAD.  Sums the chunk partials in chunk order and the group partials in group order.
//   p11 = (A + B)/4, p12 = -C_r/2, p33 = (A - B)/4, p34 = -C_i/2  with A = sum w|S+|^2, B = sum w|S-|^2, C = sum w S+ S-*
//   (S1 = (S+ + S-)/2, S2 = (S+ - S-)/2 in calculateScatVals, dointegration.py:1044-1050)
//   mirror_phase / mirror_scal (nullable): a second destination with the same layout -- the segment of this rank in rank 0's
//   gather buffer, mapped through CUDA IPC, so that the multi-GPU gather is part of this kernel (P2P stores over NVLink)
__global__ void __launch_bounds__(GM_NANG_PAD) k_finalize(int nchunk, int ngroup, int nmode, int nang, const double* __restrict__ part,
                                                         const double* __restrict__ scal_part, double* __restrict__ out_phase,
                                                         double* __restrict__ out_scal, double* __restrict__ mirror_phase,
                                                         double* __restrict__ mirror_scal) {
  const int task = blockIdx.x, a = threadIdx.x;
  if (a < nang) {
    double s[4] = {0, 0, 0, 0};
    for (int c = 0; c < nchunk; ++c) {
      const double* p = part + ((size_t)task * nchunk + c) * 4 * GM_NANG_PAD + a;
#pragma unroll
      for (int q = 0; q < 4; ++q) s[q] += p[(size_t)q * GM_NANG_PAD];
    }
    double* o = out_phase + (size_t)task * 4 * nang + a;
    const double p11 = 0.25 * (s[0] + s[1]), p12 = -0.5 * s[2], p33 = 0.25 * (s[0] - s[1]), p34 = -0.5 * s[3];
    o[0] = p11;
    o[(size_t)nang] = p12;
    o[(size_t)2 * nang] = p33;
    o[(size_t)3 * nang] = p34;
    if (mirror_phase) {
      double* m = mirror_phase + (size_t)task * 4 * nang + a;
      m[0] = p11;
      m[(size_t)nang] = p12;
      m[(size_t)2 * nang] = p33;
      m[(size_t)3 * nang] = p34;
    }
  }
  if (a < nmode * GM_NSCAL) {
    const int k = a / GM_NSCAL, q = a % GM_NSCAL;
    const double* p = scal_part + ((size_t)task * nmode + k) * ngroup * GM_NSCAL + q;
    double s = 0.0;
    for (int g = 0; g < ngroup; ++g) s += p[(size_t)g * GM_NSCAL];
    out_scal[((size_t)task * nmode + k) * GM_NSCAL + q] = s;
    if (mirror_scal) mirror_scal[((size_t)task * nmode + k) * GM_NSCAL + q] = s;
  }
}

// ================================================================================================ k_s12_direct
// One warp per particle, lanes over angles; a_n, b_n are broadcast from global memory (natural layout).  Uses the
// reference's pre-multiplied pi'_n, tau'_n and its summation (mie_S12_backend, mie_props.py:119-150).
__global__ void __launch_bounds__(128) k_s12_direct(int n, const int* __restrict__ nmax, const long long* __restrict__ aboff,
                                                    const double4* __restrict__ ab, int nang, const double* __restrict__ u,
                                                    double* __restrict__ s12) {
  const int p = blockIdx.x * 4 + (threadIdx.x >> 5);
  if (p >= n) return;
  const int lane = threadIdx.x & 31;
  const int nm = nmax[p];
  const double4* c = ab + aboff[p];
  for (int a = lane; a < nang; a += 32) {
    const double uu = u[a];
    double pm2 = 1.0, pm1 = 3.0 * uu;
    double s1r = 0, s1i = 0, s2r = 0, s2i = 0;
    for (int k = 1; k <= nm; ++k) {
      double pn, tn;
      const double dn = (double)k;
      if (k == 1) {
        pn = 1.0;
        tn = uu;
      } else if (k == 2) {
        pn = pm1;
        tn = 6.0 * uu * uu - 3.0;
      } else {
        pn = (2.0 * dn - 1.0) / (dn - 1.0) * pm1 * uu - dn / (dn - 1.0) * pm2;
        tn = dn * uu * pn - (dn + 1.0) * pm1;
        pm2 = pm1;
        pm1 = pn;
      }
      const double f = (2.0 * dn + 1.0) / (dn * (dn + 1.0));
      const double pf = pn * f, tf = tn * f;
      const double4 v = c[k - 1];
      s1r += v.x * pf + v.z * tf;
      s1i += v.y * pf + v.w * tf;
      s2r += v.x * tf + v.z * pf;
      s2i += v.y * tf + v.w * pf;
    }
    double* o = s12 + ((size_t)p * nang + a) * 4;
    o[0] = s1r;
    o[1] = s1i;
    o[2] = s2r;
    o[3] = s2i;
  }
}

// efficiencies from natural-layout coefficients (used by the coated path), mie_props_raw, mie_props.py:28-70
__global__ void __launch_bounds__(128) k_props_nat(int n, const double* __restrict__ y, const int* __restrict__ nmax,
                                                   const long long* __restrict__ aboff, const double4* __restrict__ ab,
                                                   double* __restrict__ q) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n) return;
  const int nm = nmax[p];
  const double4* c = ab + aboff[p];
  double sext = 0, ssca = 0, qbr = 0, qbi = 0, sasy = 0;
  for (int k = 1; k <= nm; ++k) {
    const double4 v = c[k - 1];
    double4 w = make_double4(0, 0, 0, 0);
    if (k < nm) w = c[k];
    const double dn = (double)k, cn = 2.0 * dn + 1.0;
    sext += cn * (v.x + v.z);
    ssca += cn * (v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w);
    const double sg = (k & 1) ? -cn : cn;
    qbr += sg * (v.x - v.z);
    qbi += sg * (v.y - v.w);
    sasy += dn * (dn + 2.0) / (dn + 1.0) * (v.x * w.x + v.y * w.y + v.z * w.z + v.w * w.w) +
            cn / (dn * (dn + 1.0)) * (v.x * v.z + v.y * v.w);
  }
  const double y2 = y[p] * y[p];
  double* o = q + (size_t)p * 6;
  o[0] = 2.0 * sext / y2;
  o[1] = 2.0 * ssca / y2;
  o[2] = o[0] - o[1];
  o[3] = (qbr * qbr + qbi * qbi) / y2;
  o[4] = 4.0 / y2 * sasy / o[1];
  o[5] = o[3] / o[1];
}

// ================================================================================================ k_phase_norm
// grid = ntask, block = GM_NANG_PAD.  dointegration.py:977-988 on the raw phase sums [task][4][nang] -> planes [4][ntask][nang] and
// pback4 [task][4].  The trapezoid terms d_i (y_i + y_{i+1}) / 2 are summed per warp by a fixed shuffle tree, then in warp order.
__global__ void __launch_bounds__(GM_NANG_PAD) k_phase_norm(int ntask, int nang, const double* __restrict__ phase, const double* __restrict__ theta,
                                                           const double* __restrict__ sint, double* __restrict__ planes, double* __restrict__ pback4) {
  __shared__ double y[GM_NANG_PAD];
  __shared__ double wsum[GM_NANG_PAD / 32];
  __shared__ double total;
  const int task = blockIdx.x, a = threadIdx.x;
  const double* P = phase + (size_t)task * 4 * nang;
  double p[4] = {0, 0, 0, 0};
  if (a < nang) {
#pragma unroll
    for (int q = 0; q < 4; ++q) p[q] = P[(size_t)q * nang + a];
    y[a] = p[0] * sint[a];
  }
  __syncthreads();
  double term = 0.0;
  if (a + 1 < nang) term = (theta[a + 1] - theta[a]) * (y[a + 1] + y[a]) / 2.0;
  term = warp_sum(term);
  if ((a & 31) == 0) wsum[a >> 5] = term;
  __syncthreads();
  if (a == 0) {
    double s = 0.0;
    for (int w = 0; w < GM_NANG_PAD / 32; ++w) s += wsum[w];
    total = s;
  }
  __syncthreads();
  if (a < nang) {
    const double p11n = 2.0 * p[0] / total;
    double o[4];
    o[0] = p11n;
#pragma unroll
    for (int q = 1; q < 4; ++q) o[q] = p[q] * p11n / p[0];
    if (planes) {
#pragma unroll
      for (int q = 0; q < 4; ++q) planes[((size_t)q * ntask + task) * nang + a] = o[q];
    }
    if (a == nang - 1) {
#pragma unroll
      for (int q = 0; q < 4; ++q) pback4[(size_t)task * 4 + q] = o[q];
    }
  }
}
