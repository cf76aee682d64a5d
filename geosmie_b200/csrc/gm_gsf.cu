// gm_gsf.cu -- generalized-spherical-function expansion of the six phase-matrix elements (north-star item 4).
// Replaces one run of the Fortran program src/gsf/spher_expan.f per (bin, wavelength, RH) cell.
//   host (once per call): Gauss-Legendre nodes/weights (GAUSS :520-579, N = ng, weights x2 as for IND1 = 0), the
//     interpolation bracket of every node in the input angle grid (LINTERPOL :593-623) and the table of generalized
//     spherical functions P^l_00, P^l_22, P^l_2-2, P^l_02 at the nodes (GENER :363-407, coefficients :293-303);
//   device (one CTA per cell): interpolate the six elements to the nodes, weight, contract with the table
//     (SPHER_EXPAN :304-347) in the Fortran accumulation order, normalise by AL1(0) (main :95-103).
#include <math.h>
#include <stdlib.h>

#include "gm_common.cuh"

#define GM_LAUNCH_CHECK(h)               \
  do {                                   \
    (h)->launches++;                     \
    GM_CUDA_TRY(cudaGetLastError());     \
  } while (0)

namespace {

// GAUSS(N, IND1=0, IND2=0, Z, W), spher_expan.f:520-579
void gauss_nodes(int N, std::vector<double>& Z, std::vector<double>& W) {
  Z.assign(N, 0.0);
  W.assign(N, 0.0);
  const double A = 1.0, B = 2.0, C = 3.0;
  const int IND = N % 2;
  const int K = N / 2 + IND;
  const double F = (double)N;
  for (int I = 1; I <= K; ++I) {
    const int M = N + 1 - I;
    double X = 0.0;
    if (I == 1) X = A - B / ((F + A) * F);
    if (I == 2) X = (Z[N - 1] - A) * 4.0 + Z[N - 1];
    if (I == 3) X = (Z[N - 2] - Z[N - 1]) * 1.6 + Z[N - 2];
    if (I > 3) X = (Z[M] - Z[M + 1]) * C + Z[M + 2];
    if (I == K && IND == 1) X = 0.0;
    int NITER = 0;
    double CHECK = 1e-16;
    double PA, PB, PC;
    for (;;) {
      PB = 1.0;
      NITER++;
      if (NITER > 100) CHECK = CHECK * 10.0;
      PC = X;
      double DJ = A;
      for (int J = 2; J <= N; ++J) {
        DJ = DJ + A;
        PA = PB;
        PB = PC;
        PC = X * PB + (X * PB - PA) * (DJ - A) / DJ;
      }
      PA = A / ((PB - X * PC) * F);
      PB = PA * PC * (A - X * X);
      X = X - PB;
      if (!(fabs(PB) > CHECK * fabs(X))) break;
    }
    Z[M - 1] = X;
    W[M - 1] = PA * PA * (A - X * X);
    W[M - 1] = B * W[M - 1];  // IND1 == 0
    if (I == K && IND == 1) continue;
    Z[I - 1] = -Z[M - 1];
    W[I - 1] = W[M - 1];
  }
}

// GENER(U, L1MAX) for all nodes: table G[4][node i][l] (l fastest), spher_expan.f:293-303 and :363-407
void gsf_table(int ng, const std::vector<double>& X, std::vector<double>& G) {
  const int L1MAX = ng;
  std::vector<double> c1(L1MAX + 2), c2(L1MAX + 2), c3(L1MAX + 2), c4(L1MAX + 2), c5(L1MAX + 2), c6(L1MAX + 2), c7(L1MAX + 2),
      c8(L1MAX + 2);
  for (int L1 = 3; L1 <= L1MAX; ++L1) {
    const int L = L1 - 1;
    c1[L1] = 1.0 / (double)(L + 1);
    c2[L1] = (double)(2 * L + 1);
    c3[L1] = 1.0 / sqrt((double)((L + 1) * (L + 1) - 4));
    c4[L1] = sqrt((double)(L * L - 4));
    c5[L1] = 1.0 / ((double)L * (double)((L + 1) * (L + 1) - 4));
    c6[L1] = (double)(2 * L + 1) * (double)(L * (L + 1));
    c7[L1] = (double)((2 * L + 1) * 4);
    c8[L1] = (double)(L + 1) * (double)(L * L - 4);
  }
  const double D6 = 0.25 * sqrt(6.0);
  G.assign((size_t)4 * ng * ng, 0.0);
  std::vector<double> P1(L1MAX + 2), P2(L1MAX + 2), P3(L1MAX + 2), P4(L1MAX + 2);
  for (int i = 0; i < ng; ++i) {
    const double U = X[i];
    const double DUP = 1.0 + U, DUM = 1.0 - U, DU = U * U;
    P1[1] = 1.0; P1[2] = U; P1[3] = 0.5 * (3.0 * DU - 1.0);
    P2[1] = 0.0; P2[2] = 0.0; P2[3] = 0.25 * DUP * DUP;
    P3[1] = 0.0; P3[2] = 0.0; P3[3] = 0.25 * DUM * DUM;
    P4[1] = 0.0; P4[2] = 0.0; P4[3] = D6 * (DU - 1.0);
    const int LMAX = L1MAX - 1;
    for (int L1 = 3; L1 <= LMAX; ++L1) {
      const double CU1 = c2[L1] * U, CU2 = c6[L1] * U;
      const int L2 = L1 + 1, L3 = L1 - 1;
      const double DL = (double)L3;
      P1[L2] = c1[L1] * (CU1 * P1[L1] - DL * P1[L3]);
      P2[L2] = c5[L1] * ((CU2 - c7[L1]) * P2[L1] - c8[L1] * P2[L3]);
      P3[L2] = c5[L1] * ((CU2 + c7[L1]) * P3[L1] - c8[L1] * P3[L3]);
      P4[L2] = c3[L1] * (CU1 * P4[L1] - c4[L1] * P4[L3]);
    }
    for (int l = 0; l < ng; ++l) {
      G[((size_t)0 * ng + i) * ng + l] = (l + 1 <= L1MAX && (l < 3 || true)) ? P1[l + 1] : 0.0;
      G[((size_t)1 * ng + i) * ng + l] = P2[l + 1];
      G[((size_t)2 * ng + i) * ng + l] = P3[l + 1];
      G[((size_t)3 * ng + i) * ng + l] = P4[l + 1];
    }
  }
}

struct GsfNode {
  double dxinv_num;  // X - XX(I-1)   (radians)
  double dx;         // XX(I) - XX(I-1)
  double w;          // Gauss weight
  int i0, i1;        // indices of YY(I-1), YY(I) (0-based)
};

// nrow = 6: F rows are F11,F22,F33,F44,F12,F34.  nrow = 4: rows are P11,P12,P33,P34 of gm_table_run (P22 = P11, P44 = P33).
__global__ void __launch_bounds__(256) k_gsf(int nrow, int nang, int ng, const double* __restrict__ F, const GsfNode* __restrict__ nodes,
                                             const double* __restrict__ G, double* __restrict__ coef, double* __restrict__ cnorm,
                                             int quantize10, double* __restrict__ raw) {
  extern __shared__ double sm[];
  double* f = sm;                 // [6][nang]
  double* ff = sm + 6 * nang;     // [6][ng]: FF11, FP, FM, FF44, FF12, FF34
  double* res = ff + 6 * ng;      // [6][ng]
  const int cell = blockIdx.x;
  const double* Fc = F + (size_t)cell * nrow * nang;
  for (int k = threadIdx.x; k < 6 * nang; k += blockDim.x) {
    int row = k / nang;
    if (nrow == 4) row = (row == 0 || row == 1) ? 0 : (row == 2 || row == 3) ? 2 : (row == 4 ? 1 : 3);
    f[k] = Fc[row * nang + k % nang];
  }
  __syncthreads();
  for (int i = threadIdx.x; i < ng; i += blockDim.x) {
    const GsfNode nd = nodes[i];
    double v[6];
#pragma unroll
    for (int k = 0; k < 6; ++k) {
      const double y0 = f[k * nang + nd.i0], y1 = f[k * nang + nd.i1];
      // LINTERPOL: Y = (YY(I)-YY(I-1))/(XX(I)-XX(I-1))*(X-XX(I-1))+YY(I-1)   (no FMA contraction)
      v[k] = __dadd_rn(__dmul_rn(__ddiv_rn(__dsub_rn(y1, y0), nd.dx), nd.dxinv_num), y0);
      v[k] = __dmul_rn(v[k], nd.w);                        // FFxx = Fxx(I)*WI, :316-321
    }
    // input order F11,F22,F33,F44,F12,F34
    ff[0 * ng + i] = v[0];
    ff[1 * ng + i] = __dadd_rn(v[1], v[2]);               // FP = FF22+FF33
    ff[2 * ng + i] = __dsub_rn(v[1], v[2]);               // FM = FF22-FF33
    ff[3 * ng + i] = v[3];
    ff[4 * ng + i] = v[4];
    ff[5 * ng + i] = v[5];
  }
  __syncthreads();
  // accumulation over the nodes in the Fortran order (DO 300 I / DO 260 L1), one thread per (series, l)
  for (int o = threadIdx.x; o < 6 * ng; o += blockDim.x) {
    const int s = o / ng, l = o % ng;
    // series: 0 AL1 (FF11,P1) 1 AL2acc (FP,P2) 2 AL3acc (FM,P3) 3 AL4 (FF44,P1) 4 BET1 (FF12,P4) 5 BET2 (FF34,P4)
    const int gsel = (s == 0 || s == 3) ? 0 : (s == 1 ? 1 : (s == 2 ? 2 : 3));
    const double* g = G + (size_t)gsel * ng * ng + l;
    const double* w = ff + s * ng;
    double acc = 0.0;
    for (int i = 0; i < ng; ++i) acc = __dadd_rn(acc, __dmul_rn(w[i], g[(size_t)i * ng]));
    res[o] = acc;
  }
  __syncthreads();
  // DO 350: scaling by (l + 1/2) and the AL2/AL3 recombination, then CNORM = 1/AL1(1)
  const double cn = 1.0 / (res[0] * 0.5);
  for (int l = threadIdx.x; l < ng; l += blockDim.x) {
    const double CL = (double)l + 0.5;
    const double al1 = res[0 * ng + l] * CL;
    const double a2 = res[1 * ng + l] * CL * 0.5;
    const double a3 = res[2 * ng + l] * CL * 0.5;
    double o[6] = {al1, a2 + a3, a2 - a3, res[3 * ng + l] * CL, res[4 * ng + l] * CL, res[5 * ng + l] * CL};
#pragma unroll
    for (int k = 0; k < 6; ++k) {
      double v = o[k] * cn;
      if (quantize10) v = rint(v * 1e10) / 1e10;
      coef[((size_t)cell * 6 + k) * ng + l] = v;
      if (raw) raw[((size_t)cell * 6 + k) * ng + l] = o[k];   // as SPHER_EXPAN leaves them: the input of MATR (one_calc :168-173)
    }
  }
  if (threadIdx.x == 0 && cnorm) cnorm[cell] = cn;
}

// k_gsf_multi: the same arithmetic in the same order as k_gsf (bit-identical results), with GSF_CB cells per CTA.  k_gsf reads
// the whole table of generalized spherical functions (4 x ng x ng doubles = 532 KB for ng = 129) through L1 once per cell and is
// bound by L2 -> SM bandwidth (optics_SU: 1.2 GB per step).  Here every table value a thread loads is used for GSF_CB cells
// (GSF_CB independent accumulation chains per thread, each summed over the nodes in the Fortran order with separate multiply
// and add roundings).  (Rejected: staging the table through shared memory in node chunks with one thread per moment index --
// 0.223 ms against 0.126 ms for k_gsf on 2196 cells: two barriers per chunk and 16 shared loads per 24 FP64 instructions.)
#ifndef GSF_CB
#define GSF_CB 2   // 2196 optics_SU cells: k_gsf 0.127 ms, CB = 2: 0.101, 3: 0.105, 4: 0.196 (fewer, fatter CTAs: wave quantisation)
#endif

__global__ void __launch_bounds__(256) k_gsf_multi(int nrow, int nang, int ng, int ncell, const double* __restrict__ F,
                                                   const GsfNode* __restrict__ nodes, const double* __restrict__ G,
                                                   double* __restrict__ coef, double* __restrict__ cnorm, int quantize10,
                                                   double* __restrict__ raw) {
  extern __shared__ double sm[];
  double* f = sm;                              // [6][nang]            one cell at a time
  double* ff = f + 6 * nang;                   // [CB][6][ng]          FF11, FP, FM, FF44, FF12, FF34 at the nodes
  double* res = ff + GSF_CB * 6 * ng;          // [CB][6][ng]
  const int cell0 = blockIdx.x * GSF_CB;
  const int ncb = min(GSF_CB, ncell - cell0);
  // ---- interpolation to the Gauss nodes, cell by cell (LINTERPOL, one_calc :159-166; weights :316-321)
  for (int c = 0; c < GSF_CB; ++c) {
    if (c >= ncb) {
      for (int k = threadIdx.x; k < 6 * ng; k += blockDim.x) ff[(size_t)c * 6 * ng + k] = 0.0;
      continue;
    }
    const double* Fc = F + (size_t)(cell0 + c) * nrow * nang;
    __syncthreads();
    for (int k = threadIdx.x; k < 6 * nang; k += blockDim.x) {
      int row = k / nang;
      if (nrow == 4) row = (row == 0 || row == 1) ? 0 : (row == 2 || row == 3) ? 2 : (row == 4 ? 1 : 3);
      f[k] = Fc[row * nang + k % nang];
    }
    __syncthreads();
    for (int i = threadIdx.x; i < ng; i += blockDim.x) {
      const GsfNode nd = nodes[i];
      double v[6];
#pragma unroll
      for (int k = 0; k < 6; ++k) {
        const double y0 = f[k * nang + nd.i0], y1 = f[k * nang + nd.i1];
        v[k] = __dadd_rn(__dmul_rn(__ddiv_rn(__dsub_rn(y1, y0), nd.dx), nd.dxinv_num), y0);
        v[k] = __dmul_rn(v[k], nd.w);
      }
      double* o = ff + (size_t)c * 6 * ng + i;
      o[0 * ng] = v[0];
      o[1 * ng] = __dadd_rn(v[1], v[2]);
      o[2 * ng] = __dsub_rn(v[1], v[2]);
      o[3 * ng] = v[3];
      o[4 * ng] = v[4];
      o[5 * ng] = v[5];
    }
  }
  __syncthreads();
  // ---- accumulation over the nodes in the Fortran order (DO 300 I / DO 260 L1), one thread per (series, l), GSF_CB cells
  for (int o = threadIdx.x; o < 6 * ng; o += blockDim.x) {
    const int s6 = o / ng, l = o % ng;
    // series: 0 AL1 (FF11,P1) 1 AL2acc (FP,P2) 2 AL3acc (FM,P3) 3 AL4 (FF44,P1) 4 BET1 (FF12,P4) 5 BET2 (FF34,P4)
    const int gsel = (s6 == 0 || s6 == 3) ? 0 : (s6 == 1 ? 1 : (s6 == 2 ? 2 : 3));
    const double* g = G + (size_t)gsel * ng * ng + l;
    const double* w = ff + s6 * ng;
    double acc[GSF_CB];
#pragma unroll
    for (int c = 0; c < GSF_CB; ++c) acc[c] = 0.0;
#pragma unroll 8
    for (int i = 0; i < ng; ++i) {
      const double gv = g[(size_t)i * ng];
#pragma unroll
      for (int c = 0; c < GSF_CB; ++c) acc[c] = __dadd_rn(acc[c], __dmul_rn(w[(size_t)c * 6 * ng + i], gv));
    }
#pragma unroll
    for (int c = 0; c < GSF_CB; ++c) res[(size_t)c * 6 * ng + o] = acc[c];
  }
  __syncthreads();
  // ---- DO 350: scaling by (l + 1/2), AL2/AL3 recombination, CNORM = 1/AL1(1)
  for (int e = threadIdx.x; e < ncb * ng; e += blockDim.x) {
    const int c = e / ng, lo = e % ng;
    const double* r = res + (size_t)c * 6 * ng;
    const double cn = 1.0 / (r[0] * 0.5);
    const double CL = (double)lo + 0.5;
    const double al1 = r[0 * ng + lo] * CL;
    const double a2 = r[1 * ng + lo] * CL * 0.5;
    const double a3 = r[2 * ng + lo] * CL * 0.5;
    double o[6] = {al1, a2 + a3, a2 - a3, r[3 * ng + lo] * CL, r[4 * ng + lo] * CL, r[5 * ng + lo] * CL};
    const size_t cell = (size_t)(cell0 + c);
#pragma unroll
    for (int k = 0; k < 6; ++k) {
      double v = o[k] * cn;
      if (quantize10) v = rint(v * 1e10) / 1e10;
      coef[(cell * 6 + k) * ng + lo] = v;
      if (raw) raw[(cell * 6 + k) * ng + lo] = o[k];
    }
    if (lo == 0 && cnorm) cnorm[cell] = cn;
  }
}

// ------------------------------------------------------------------------------------------------ diagnostics
// The diagnostic half of spher_expan.f: MATR (:419-517) re-synthesises the matrix from the un-normalised coefficients at the
// input angles (what main writes to <file>.expan_matr, :84-90) and at READMATRIX's alternative grid (:237-258, USE_ALT_ANG = 1);
// fiterr = max |F11 - F11OUT| over both (ERREVAL, ERRTYP = MAXABS, range [0, 180] deg; one_calc :168-177).
struct GsfAlt {
  double u;          // cos(angl(i))       (host libm, like the oracle)
  double ualt;       // cos(alt_angl(i))
  double num, dx;    // LINTERPOL bracket of alt_angl(i) in the input grid, in DEGREES (READMATRIX interpolates before the D2R scaling)
  int i0, i1;
  int inrange;       // ang_min <= angl(i) <= ang_max
  int pad;
};

// one MATR angle; no FMA contraction (the oracle is compiled without FMA)
__device__ __forceinline__ void gsf_matr_angle(int ng, const double* __restrict__ c, double U, double (&F)[6]) {
  const double* A1 = c, *A2 = c + ng, *A3 = c + 2 * ng, *A4 = c + 3 * ng, *B1 = c + 4 * ng, *B2 = c + 5 * ng;
  const int LMAX = ng - 1;
  const double D6 = __dmul_rn(sqrt(6.0), 0.25);
  double F11 = 0, F2 = 0, F3 = 0, F44 = 0, F12 = 0, F34 = 0, P1 = 0, P2 = 0, P3 = 0, P4 = 0;
  double PP1 = 1.0;
  const double up = __dadd_rn(1.0, U), um = __dsub_rn(1.0, U);
  double PP2 = __dmul_rn(__dmul_rn(0.25, up), up);
  double PP3 = __dmul_rn(__dmul_rn(0.25, um), um);
  double PP4 = __dmul_rn(D6, __dsub_rn(__dmul_rn(U, U), 1.0));
  for (int L1 = 1; L1 <= ng; ++L1) {
    const int L = L1 - 1;
    const double DL = (double)L, DL1 = (double)L1, PL1 = (double)(2 * L + 1);
    F11 = __dadd_rn(F11, __dmul_rn(A1[L], PP1));
    F44 = __dadd_rn(F44, __dmul_rn(A4[L], PP1));
    if (L != LMAX) {
      const double P = __ddiv_rn(__dsub_rn(__dmul_rn(__dmul_rn(PL1, U), PP1), __dmul_rn(DL, P1)), DL1);
      P1 = PP1;
      PP1 = P;
    }
    if (L < 2) continue;
    F2 = __dadd_rn(F2, __dmul_rn(__dadd_rn(A2[L], A3[L]), PP2));
    F3 = __dadd_rn(F3, __dmul_rn(__dsub_rn(A2[L], A3[L]), PP3));
    F12 = __dadd_rn(F12, __dmul_rn(B1[L], PP4));
    F34 = __dadd_rn(F34, __dmul_rn(B2[L], PP4));
    if (L == LMAX) continue;
    const double PL2 = __dmul_rn(__dmul_rn(DL, DL1), U);
    const double PL3 = __dmul_rn(DL1, __dsub_rn(__dmul_rn(DL, DL), 4.0));
    const double PL4 = __ddiv_rn(1.0, __dmul_rn(DL, __dsub_rn(__dmul_rn(DL1, DL1), 4.0)));
    double P = __dmul_rn(__dsub_rn(__dmul_rn(__dmul_rn(PL1, __dsub_rn(PL2, 4.0)), PP2), __dmul_rn(PL3, P2)), PL4);
    P2 = PP2;
    PP2 = P;
    P = __dmul_rn(__dsub_rn(__dmul_rn(__dmul_rn(PL1, __dadd_rn(PL2, 4.0)), PP3), __dmul_rn(PL3, P3)), PL4);
    P3 = PP3;
    PP3 = P;
    P = __ddiv_rn(__dsub_rn(__dmul_rn(__dmul_rn(PL1, U), PP4), __dmul_rn(__dsqrt_rn(__dsub_rn(__dmul_rn(DL, DL), 4.0)), P4)),
                  __dsqrt_rn(__dsub_rn(__dmul_rn(DL1, DL1), 4.0)));
    P4 = PP4;
    PP4 = P;
  }
  F[0] = F11;
  F[1] = __dmul_rn(__dadd_rn(F2, F3), 0.5);
  F[2] = __dmul_rn(__dsub_rn(F2, F3), 0.5);
  F[3] = F44;
  F[4] = F12;
  F[5] = F34;
}

// CTA per cell, thread per (grid, angle).  F11 of the cell is row 0 for both input layouts (nrow = 6 or 4).
__global__ void __launch_bounds__(256) k_gsf_matr(int nrow, int nang, int ng, const double* __restrict__ F, const double* __restrict__ raw,
                                                  const GsfAlt* __restrict__ alt, double* __restrict__ fout, double* __restrict__ fiterr) {
  extern __shared__ double sm[];
  double* c = sm;              // [6][ng]
  double* f11 = sm + 6 * ng;   // [nang]
  __shared__ double red[256];
  const int cell = blockIdx.x;
  for (int k = threadIdx.x; k < 6 * ng; k += blockDim.x) c[k] = raw[(size_t)cell * 6 * ng + k];
  for (int k = threadIdx.x; k < nang; k += blockDim.x) f11[k] = F[(size_t)cell * nrow * nang + k];
  __syncthreads();
  double err = 0.0;
  for (int e = threadIdx.x; e < 2 * nang; e += blockDim.x) {
    const int which = e / nang, i = e % nang;
    const GsfAlt a = alt[i];
    double out[6];
    gsf_matr_angle(ng, c, which ? a.ualt : a.u, out);
    double ref;
    if (which == 0) {
      ref = f11[i];
      if (fout)
#pragma unroll
        for (int k = 0; k < 6; ++k) fout[((size_t)cell * 6 + k) * nang + i] = out[k];
    } else {
      const double y0 = f11[a.i0], y1 = f11[a.i1];
      ref = __dadd_rn(__dmul_rn(__ddiv_rn(__dsub_rn(y1, y0), a.dx), a.num), y0);   // F11ALT(i), LINTERPOL in degrees
    }
    if (a.inrange) err = fmax(err, fabs(__dsub_rn(ref, out[0])));
  }
  red[threadIdx.x] = err;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) red[threadIdx.x] = fmax(red[threadIdx.x], red[threadIdx.x + o]);
    __syncthreads();
  }
  if (threadIdx.x == 0 && fiterr) fiterr[cell] = red[0];
}

int gsf_upload_alt(gm_handle_t h, int nang, const double* ang) {
  const double PI = acos(-1.0), D2R = PI / 180.0;
  const double ang_min = 0.0 * D2R, ang_max = 180.0 * D2R;   // params.h:9-10
  std::vector<double> altd(nang);
  altd[0] = ang[0];
  altd[nang - 1] = ang[nang - 1];
  for (int i = 2; i <= nang / 2; ++i) altd[i - 1] = 0.5 * (ang[i - 2] + ang[i - 1]);
  for (int i = nang / 2 + 1; i <= nang - 1; ++i) altd[i - 1] = 0.5 * (ang[i] + ang[i - 1]);
  std::vector<GsfAlt> A(nang);
  for (int i = 0; i < nang; ++i) {
    const double x = altd[i];
    GsfAlt& a = A[i];
    if (x < ang[0]) {
      a.i0 = 0; a.i1 = 1; a.dx = ang[1] - ang[0]; a.num = x - ang[0];
    } else if (x > ang[nang - 1]) {
      a.i0 = nang - 2; a.i1 = nang - 1; a.dx = ang[nang - 1] - ang[nang - 2]; a.num = x - ang[nang - 2];
    } else {
      int I;
      for (I = 2; I <= nang; ++I)
        if (ang[I - 1] > x) break;
      if (I > nang) I = nang;
      a.i0 = I - 2; a.i1 = I - 1; a.dx = ang[I - 1] - ang[I - 2]; a.num = x - ang[I - 2];
    }
    const double r = ang[i] * D2R;
    a.u = cos(r);
    a.ualt = cos(altd[i] * D2R);
    a.inrange = !(r < ang_min || r > ang_max);
    a.pad = 0;
  }
  int rc = h->gsf_alt.ensure(sizeof(GsfAlt) * nang);
  if (rc) return rc;
  GM_CUDA_TRY(cudaMemcpyAsync(h->gsf_alt.p, A.data(), sizeof(GsfAlt) * nang, cudaMemcpyHostToDevice, h->stream));
  GM_CUDA_TRY(cudaStreamSynchronize(h->stream));
  return GM_OK;
}

int gsf_upload_constants(gm_handle_t h, int nang, const double* h_ang_deg, int ng) {
  cudaStream_t st = h->stream;
  std::vector<double> X, W, G;
  gauss_nodes(ng, X, W);
  gsf_table(ng, X, G);
  // angles in radians: angl(i) = angl(i)*D2R, D2R = PI/180 with PI = DACOS(-1) (params.h:3-4, spher_expan.f:260-263)
  const double PI = acos(-1.0), D2R = PI / 180.0;
  std::vector<double> XX(nang);
  for (int i = 0; i < nang; ++i) XX[i] = h_ang_deg[i] * D2R;
  std::vector<GsfNode> nodes(ng);
  for (int i = 0; i < ng; ++i) {
    const double x = acos(X[i]);
    int I;  // 1-based index of the upper bracket as in LINTERPOL
    if (x < XX[0]) {
      // Y = (YY(1)-YY(2))/(XX(1)-XX(2))*(X-XX(1))+YY(1)
      nodes[i].i0 = 0; nodes[i].i1 = 1;
      // rewrite in the generic form with (y1 - y0)/(dx) * num + y0: (YY(1)-YY(2))/(XX(1)-XX(2)) = (y1 - y0)/(XX(2)-XX(1)) up to rounding
      nodes[i].dx = XX[1] - XX[0];
      nodes[i].dxinv_num = x - XX[0];
    } else if (x > XX[nang - 1]) {
      nodes[i].i0 = nang - 2; nodes[i].i1 = nang - 1;
      nodes[i].dx = XX[nang - 1] - XX[nang - 2];
      // Y = slope*(X-XX(NN)) + YY(NN) == slope*(X - XX(NN-1)) + YY(NN-1) up to rounding
      nodes[i].dxinv_num = x - XX[nang - 2];
    } else {
      for (I = 2; I <= nang; ++I)
        if (XX[I - 1] > x) break;
      if (I > nang) I = nang;  // X == XX(NN): the Fortran loop leaves I = NN+1; guard against reading past the end
      nodes[i].i0 = I - 2; nodes[i].i1 = I - 1;
      nodes[i].dx = XX[I - 1] - XX[I - 2];
      nodes[i].dxinv_num = x - XX[I - 2];
    }
    nodes[i].w = W[i];
  }
  int rc;
  if ((rc = h->gsf_nodes.ensure(sizeof(GsfNode) * ng)) || (rc = h->gsf_table.ensure(sizeof(double) * G.size()))) return rc;
  GM_CUDA_TRY(cudaMemcpyAsync(h->gsf_nodes.p, nodes.data(), sizeof(GsfNode) * ng, cudaMemcpyHostToDevice, st));
  GM_CUDA_TRY(cudaMemcpyAsync(h->gsf_table.p, G.data(), sizeof(double) * G.size(), cudaMemcpyHostToDevice, st));
  GM_CUDA_TRY(cudaStreamSynchronize(st));  // nodes/G are stack-owned
  return GM_OK;
}

int gsf_core(gm_handle_t h, int ncell, int nang, const double* h_ang_deg, const double* d_F, int ng, double* d_coef,
             double* d_cnorm, int quantize10, int nrow = 6, double* d_fout = nullptr, double* d_fiterr = nullptr) {
  GM_REQUIRE(ng >= 3 && ng <= 2048, "ng out of range");
  GM_REQUIRE(nang >= 2 && nang <= 1000, "nang out of range (NANG_MAX = 1000, params.h:1)");
  cudaStream_t st = h->stream;
  const bool diag = d_fout || d_fiterr;
  // the constants depend only on (ng, angle grid): build and upload them once per grid
  std::vector<double> key(h_ang_deg, h_ang_deg + nang);
  key.push_back((double)ng);
  if (key != h->gsf_key) {
    int rc0 = gsf_upload_constants(h, nang, h_ang_deg, ng);
    if (rc0) return rc0;
    h->gsf_key = key;
    h->gsf_alt_valid = false;
  }
  double* d_raw = nullptr;
  if (diag) {
    int rc0;
    if (!h->gsf_alt_valid) {
      if ((rc0 = gsf_upload_alt(h, nang, h_ang_deg))) return rc0;
      h->gsf_alt_valid = true;
    }
    if ((rc0 = h->gsf_raw.ensure(sizeof(double) * (size_t)ncell * 6 * ng))) return rc0;
    d_raw = h->gsf_raw.as<double>();
  }
  const bool force_simple = getenv("GEOSMIE_GSF_SIMPLE") != nullptr;   // cross-check of the two kernels (tests)
  const size_t smem_multi = sizeof(double) * (6 * (size_t)nang + (size_t)(2 * GSF_CB * 6) * ng);
  if (ncell >= 2 * GSF_CB && smem_multi <= 200 * 1024 && !force_simple) {
    GM_CUDA_TRY(cudaFuncSetAttribute(k_gsf_multi, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_multi));
    k_gsf_multi<<<(ncell + GSF_CB - 1) / GSF_CB, 256, smem_multi, st>>>(nrow, nang, ng, ncell, d_F, h->gsf_nodes.as<GsfNode>(),
                                                                      h->gsf_table.as<double>(), d_coef, d_cnorm, quantize10, d_raw);
  } else {
    const size_t smem = sizeof(double) * (6 * (size_t)nang + 12 * (size_t)ng);
    GM_CUDA_TRY(cudaFuncSetAttribute(k_gsf, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_gsf<<<ncell, 256, smem, st>>>(nrow, nang, ng, d_F, h->gsf_nodes.as<GsfNode>(), h->gsf_table.as<double>(), d_coef, d_cnorm, quantize10,
                                    d_raw);
  }
  GM_LAUNCH_CHECK(h);
  if (diag) {
    const size_t smem2 = sizeof(double) * (6 * (size_t)ng + (size_t)nang);
    GM_CUDA_TRY(cudaFuncSetAttribute(k_gsf_matr, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2));
    k_gsf_matr<<<ncell, 256, smem2, st>>>(nrow, nang, ng, d_F, d_raw, h->gsf_alt.as<GsfAlt>(), d_fout, d_fiterr);
    GM_LAUNCH_CHECK(h);
  }
  return GM_OK;
}

}  // namespace

extern "C" int gm_gsf_expand_dev(gm_handle_t h, int ncell, int nang, const double* ang_deg, const double* F, int ng, double* coef,
                                 double* cnorm, int quantize10) {
  GM_REQUIRE(h && ang_deg && F && coef, "NULL argument");
  GM_REQUIRE(ncell > 0, "ncell must be > 0");
  GM_CUDA_TRY(cudaSetDevice(h->device));
  return gsf_core(h, ncell, nang, ang_deg, F, ng, coef, cnorm, quantize10);
}

int gm_gsf_phase4_async(gm_handle_s* h, int ncell, int nang, const double* h_ang_deg, const double* d_P4, int ng, double* d_coef,
                        double* d_cnorm, int quantize10) {
  return gsf_core(h, ncell, nang, h_ang_deg, d_P4, ng, d_coef, d_cnorm, quantize10, 4);
}

extern "C" int gm_gsf_expand_phase4_dev(gm_handle_t h, int ncell, int nang, const double* ang_deg, const double* P4, int ng,
                                        double* coef, double* cnorm, int quantize10) {
  GM_REQUIRE(h && ang_deg && P4 && coef, "NULL argument");
  GM_REQUIRE(ncell > 0, "ncell must be > 0");
  GM_CUDA_TRY(cudaSetDevice(h->device));
  return gsf_core(h, ncell, nang, ang_deg, P4, ng, coef, cnorm, quantize10, 4);
}

extern "C" int gm_gsf_expand(gm_handle_t h, int ncell, int nang, const double* ang_deg, const double* F, int ng, double* coef,
                             double* cnorm, int quantize10) {
  GM_REQUIRE(h && ang_deg && F && coef, "NULL argument");
  GM_REQUIRE(ncell > 0, "ncell must be > 0");
  GM_CUDA_TRY(cudaSetDevice(h->device));
  cudaStream_t st = h->stream;
  int rc;
  const size_t nf = (size_t)ncell * 6 * nang, nc = (size_t)ncell * 6 * ng;
  if ((rc = h->ws[2].ensure(sizeof(double) * nf)) || (rc = h->ws[3].ensure(sizeof(double) * nc)) ||
      (rc = h->ws[4].ensure(sizeof(double) * ncell)))
    return rc;
  GM_CUDA_TRY(cudaMemcpyAsync(h->ws[2].p, F, sizeof(double) * nf, cudaMemcpyHostToDevice, st));
  rc = gsf_core(h, ncell, nang, ang_deg, h->ws[2].as<double>(), ng, h->ws[3].as<double>(), h->ws[4].as<double>(), quantize10);
  if (rc) return rc;
  GM_CUDA_TRY(cudaMemcpyAsync(coef, h->ws[3].p, sizeof(double) * nc, cudaMemcpyDeviceToHost, st));
  if (cnorm) GM_CUDA_TRY(cudaMemcpyAsync(cnorm, h->ws[4].p, sizeof(double) * ncell, cudaMemcpyDeviceToHost, st));
  GM_CUDA_TRY(cudaStreamSynchronize(st));
  return GM_OK;
}

extern "C" int gm_gsf_diagnose_dev(gm_handle_t h, int ncell, int nang, const double* ang_deg, const double* F, int nrow, int ng,
                                   double* coef, double* cnorm, int quantize10, double* fout, double* fiterr) {
  GM_REQUIRE(h && ang_deg && F && coef, "NULL argument");
  GM_REQUIRE(ncell > 0, "ncell must be > 0");
  GM_REQUIRE(nrow == 6 || nrow == 4, "nrow must be 6 (F11,F22,F33,F44,F12,F34) or 4 (P11,P12,P33,P34)");
  GM_REQUIRE(fout || fiterr, "give fout and/or fiterr (use gm_gsf_expand_dev for the moments alone)");
  GM_CUDA_TRY(cudaSetDevice(h->device));
  return gsf_core(h, ncell, nang, ang_deg, F, ng, coef, cnorm, quantize10, nrow, fout, fiterr);
}

extern "C" int gm_gsf_diagnose(gm_handle_t h, int ncell, int nang, const double* ang_deg, const double* F, int ng, double* coef,
                               double* cnorm, int quantize10, double* fout, double* fiterr) {
  GM_REQUIRE(h && ang_deg && F && coef, "NULL argument");
  GM_REQUIRE(ncell > 0, "ncell must be > 0");
  GM_REQUIRE(fout || fiterr, "give fout and/or fiterr (use gm_gsf_expand for the moments alone)");
  GM_CUDA_TRY(cudaSetDevice(h->device));
  cudaStream_t st = h->stream;
  int rc;
  const size_t nf = (size_t)ncell * 6 * nang, nc = (size_t)ncell * 6 * ng;
  if ((rc = h->ws[2].ensure(sizeof(double) * nf)) || (rc = h->ws[3].ensure(sizeof(double) * nc)) ||
      (rc = h->ws[4].ensure(sizeof(double) * ncell)) || (rc = h->ws[5].ensure(sizeof(double) * nf)) ||
      (rc = h->ws[6].ensure(sizeof(double) * ncell)))
    return rc;
  GM_CUDA_TRY(cudaMemcpyAsync(h->ws[2].p, F, sizeof(double) * nf, cudaMemcpyHostToDevice, st));
  rc = gsf_core(h, ncell, nang, ang_deg, h->ws[2].as<double>(), ng, h->ws[3].as<double>(), h->ws[4].as<double>(), quantize10, 6,
                h->ws[5].as<double>(), h->ws[6].as<double>());
  if (rc) return rc;
  GM_CUDA_TRY(cudaMemcpyAsync(coef, h->ws[3].p, sizeof(double) * nc, cudaMemcpyDeviceToHost, st));
  if (cnorm) GM_CUDA_TRY(cudaMemcpyAsync(cnorm, h->ws[4].p, sizeof(double) * ncell, cudaMemcpyDeviceToHost, st));
  if (fout) GM_CUDA_TRY(cudaMemcpyAsync(fout, h->ws[5].p, sizeof(double) * nf, cudaMemcpyDeviceToHost, st));
  if (fiterr) GM_CUDA_TRY(cudaMemcpyAsync(fiterr, h->ws[6].p, sizeof(double) * ncell, cudaMemcpyDeviceToHost, st));
  GM_CUDA_TRY(cudaStreamSynchronize(st));
  return GM_OK;
}
