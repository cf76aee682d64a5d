// gm_coated.cuh -- coated (core + shell) sphere Mie coefficients, one thread per particle.
// Restates coated_mie_coeff (src/pymiecoated/pymiecoated/mie_coeffs.py:183-251) without complex-argument Bessel library
// calls: the three logarithmic derivatives D_n(u), D_n(v), D_n(w) come from the reference's own downward recurrence
// (:214-218, shared nmx :204-205); psi_n at the complex arguments v, w follows from psi_n = psi_{n-1} / (D_n + n/z)
// (the definition of D_n), chi_n from the upward recurrence, psi_n(y) for the real shell size parameter as in k_bessel.
#pragma once
#include "gm_common.cuh"

__device__ __forceinline__ double2 csin_d(double2 z) {
  double s, c;
  sincos(z.x, &s, &c);
  return make_double2(s * cosh(z.y), c * sinh(z.y));
}
__device__ __forceinline__ double2 ccos_d(double2 z) {
  double s, c;
  sincos(z.x, &s, &c);
  return make_double2(c * cosh(z.y), -s * sinh(z.y));
}
__device__ __forceinline__ double2 cadd(double2 a, double2 b) { return make_double2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ double2 csub(double2 a, double2 b) { return make_double2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ double2 cscale(double2 a, double s) { return make_double2(a.x * s, a.y * s); }

// scratch per particle: (nmax+1) * 8 doubles, particle i at soff[i]*8 (+ task * scratch_stride)
// mat_stride: 0 one material pair for all, 1 per particle, 2 per task (blockIdx.y).
// core_ratio != NULL (table mode): the core size parameter is core_ratio[task] * y (RH-dependent shell growth).
__global__ void __launch_bounds__(64) k_coated_coeff(int n, const double* __restrict__ xcore, const double* __restrict__ yshell,
                                                     const double2* __restrict__ m1a, const double2* __restrict__ m2a, int mat_stride,
                                                     const int* __restrict__ nmax, const long long* __restrict__ soff,
                                                     double* __restrict__ scratch, const long long* __restrict__ aboff,
                                                     double4* __restrict__ ab, const double* __restrict__ core_ratio,
                                                     long long scratch_stride, long long ab_stride) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n) return;
  const int task = blockIdx.y;
  const double y = yshell[p];
  const double x = core_ratio ? core_ratio[task] * y : xcore[p];
  const int mi = mat_stride == 1 ? p : (mat_stride == 2 ? task : 0);
  const double2 m1 = m1a[mi], m2 = m2a[mi];
  const int nm = nmax[p];
  scratch += (size_t)task * scratch_stride;
  ab += (size_t)task * ab_stride;
  const double2 m = cdiv(m2, m1);                 // :197
  const double2 u = cscale(m1, x), v = cscale(m2, x), w = cscale(m2, y);   // :198-200
  const double mx = fmax(hypot(m1.x * y, m1.y * y), hypot(w.x, w.y));       // :203
  const int nmx = (int)rint(fmax((double)nm, mx) + 16.0);                   // :204
  double* sc = scratch + (size_t)soff[p] * 8;
  double2* Du = reinterpret_cast<double2*>(sc);
  double2* Dv = Du + (nm + 1);
  double2* Dw = Dv + (nm + 1);
  double* ry = reinterpret_cast<double*>(Dw + (nm + 1));
  // downward recurrences :214-218 (dnx[j-1] = r - 1/(dnx[j] + r), r = (j+1)/z; D_n = dnx[n-1])
  const double2 zi[3] = {crcp(u), crcp(v), crcp(w)};
  double2* Dz[3] = {Du, Dv, Dw};
  for (int k = 0; k < 3; ++k) {
    double2 D = make_double2(0.0, 0.0);
    for (int j = nmx - 1; j >= 1; --j) {
      const double2 r = cscale(zi[k], (double)(j + 1));
      D = csub(r, crcp(cadd(D, r)));
      if (j <= nm) Dz[k][j] = D;
    }
  }
  // ratios psi_n(y)/psi_{n-1}(y) for the real argument (see k_bessel)
  const double yinv = 1.0 / y;
  int nt = (int)ceil(y - 0.5);
  if (nt < 0) nt = 0;
  if (nt > nm) nt = nm;
  if (nt < nm) {
    int N = nm + (int)(4.3 * cbrt(y)) + 20;
    double r = 0.0;
    for (int k = N; k > nt; --k) {
      r = 1.0 / ((2 * k + 1) * yinv - r);
      if (k <= nm) ry[k] = r;
    }
  }
  // upward sweep over the order k
  double sy, cy;
  sincos(y, &sy, &cy);
  double pm1 = sy, pm2 = 0.0;                   // psi_{k-1}(y), psi_{k-2}(y)
  double cm1 = cy, cm2 = 0.0;                   // chi_{k-1}(y), chi_{k-2}(y)
  double2 pv = csin_d(v), pw = csin_d(w);       // psi_{k-1} at v, w
  double2 cvm1 = ccos_d(v), cvm2 = make_double2(0.0, 0.0);
  double2 cwm1 = ccos_d(w), cwm2 = make_double2(0.0, 0.0);
  const double2 m2inv = crcp(m2);
  const double2 minv = crcp(m);
  const long long abo = aboff[p];
  for (int k = 1; k <= nm; ++k) {
    const double dk = (double)k;
    const double tk = (double)(2 * k - 1);
    // real argument: psi_k(y), chi_k(y)
    double psi_y, chi_y;
    if (k == 1) {
      psi_y = (nt >= 1) ? sy * yinv - cy : ry[1] * pm1;
      chi_y = cy * yinv + sy;
    } else {
      psi_y = (k <= nt) ? tk * yinv * pm1 - pm2 : ry[k] * pm1;
      chi_y = tk * yinv * cm1 - cm2;
    }
    const double p1y = pm1, ch1y = cm1;          // :226-227
    pm2 = pm1; pm1 = psi_y;
    cm2 = cm1; cm1 = chi_y;
    // complex arguments: psi_k = psi_{k-1} / (D_k + k/z); chi_k upward
    pv = cdiv(pv, cadd(Dv[k], cscale(zi[1], dk)));
    pw = cdiv(pw, cadd(Dw[k], cscale(zi[2], dk)));
    double2 chvk, chwk;
    if (k == 1) {
      chvk = cadd(cmul(cvm1, zi[1]), csin_d(v));
      chwk = cadd(cmul(cwm1, zi[2]), csin_d(w));
    } else {
      chvk = csub(cscale(cmul(cvm1, zi[1]), tk), cvm2);
      chwk = csub(cscale(cmul(cwm1, zi[2]), tk), cwm2);
    }
    cvm2 = cvm1; cvm1 = chvk;
    cwm2 = cwm1; cwm1 = chwk;
    // Matzler / Bohren-Huffman combination, mie_coeffs.py:230-249
    const double2 dnu = Du[k], dnv = Dv[k], dnw = Dw[k];
    const double2 uu = csub(cmul(m, dnu), dnv);            // :230
    const double2 vv = csub(cmul(dnu, minv), dnv);         // :231
    const double2 fv = cdiv(pv, chvk);                     // :232
    const double2 ku1 = cdiv(cmul(uu, fv), pw);            // :234
    const double2 kv1 = cdiv(cmul(vv, fv), pw);            // :235
    const double2 pt = csub(pw, cmul(chwk, fv));           // :236
    const double2 prat = cdiv(cdiv(pw, pv), chvk);         // :237
    const double2 ku2 = cadd(cmul(uu, pt), prat);          // :238
    const double2 kv2 = cadd(cmul(vv, pt), prat);          // :239
    const double2 dns = cadd(cdiv(ku1, ku2), dnw);         // :240,:243
    const double2 gns = cadd(cdiv(kv1, kv2), dnw);         // :241,:244
    const double nrat = dk * yinv;                         // :245
    double2 a1 = cmul(dns, m2inv);                         // :246
    a1.x += nrat;
    double2 b1 = cmul(m2, gns);                            // :247
    b1.x += nrat;
    // an = (py a1 - p1y) / (gsy a1 - gs1y), gsy = py - i chy      :248-249
    const double2 an = cdiv(make_double2(psi_y * a1.x - p1y, psi_y * a1.y),
                            make_double2(psi_y * a1.x + chi_y * a1.y - p1y, psi_y * a1.y - chi_y * a1.x + ch1y));
    const double2 bn = cdiv(make_double2(psi_y * b1.x - p1y, psi_y * b1.y),
                            make_double2(psi_y * b1.x + chi_y * b1.y - p1y, psi_y * b1.y - chi_y * b1.x + ch1y));
    ab[abo + k - 1] = make_double4(an.x, an.y, bn.x, bn.y);
  }
}
