/*
 * oracle/gsf_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Plain-C restatement of the Fortran-77 program src/gsf/spher_expan.f (Mishchenko's generalized-spherical-function
 * expansion, as driven by src/gsf/convertncdf.py:173-189) for one scattering matrix.
 *
 * Parity status: "PARITY UNPINNED" by the reference -- the reference ships no test, golden file or built executable
 * for this program and no Fortran compiler exists in this image (gfortran/flang/f2c absent), so the program cannot
 * be run here.  This restatement is pinned instead by (a) the analytic Rayleigh expansion coefficients and
 * (b) re-synthesis of the input matrix from the coefficients with the MATR recurrences (spher_expan.f:419-517);
 * see tests/test_oracle.py.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

/* GAUSS(N,IND1,IND2,Z,W), spher_expan.f:520-579 (IND2 printing omitted) */
void orc_gauss(int N, int IND1, double* Z, double* W) {
  double A = 1.0, B = 2.0, C = 3.0;
  int IND = N % 2;
  int K = N / 2 + IND;
  double F = (double)N;
  for (int I = 1; I <= K; ++I) {
    int M = N + 1 - I;
    double X = 0, PA = 0, PB, PC, DJ, CHECK;
    int NITER;
    if (I == 1) X = A - B / ((F + A) * F);                       /* :531 */
    if (I == 2) X = (Z[N - 1] - A) * 4.0 + Z[N - 1];              /* :532 */
    if (I == 3) X = (Z[N - 2] - Z[N - 1]) * 1.6 + Z[N - 2];        /* :533 */
    if (I > 3) X = (Z[M] - Z[M + 1]) * C + Z[M + 2];              /* :534  Z(M+1), Z(M+2), Z(M+3) 1-based */
    if (I == K && IND == 1) X = 0.0;                             /* :535 */
    NITER = 0;
    CHECK = 1e-16;
    do {
      PB = 1.0;
      NITER = NITER + 1;
      if (NITER > 100) CHECK = CHECK * 10.0;
      PC = X;
      DJ = A;
      for (int J = 2; J <= N; ++J) {
        DJ = DJ + A;
        PA = PB;
        PB = PC;
        PC = X * PB + (X * PB - PA) * (DJ - A) / DJ;
      }
      PA = A / ((PB - X * PC) * F);
      PB = PA * PC * (A - X * X);
      X = X - PB;
    } while (fabs(PB) > CHECK * fabs(X));
    Z[M - 1] = X;
    W[M - 1] = PA * PA * (A - X * X);
    if (IND1 == 0) W[M - 1] = B * W[M - 1];
    if (I == K && IND == 1) continue;
    Z[I - 1] = -Z[M - 1];
    W[I - 1] = W[M - 1];
  }
  if (IND1 != 0)
    for (int I = 0; I < N; ++I) Z[I] = (A + Z[I]) / B;
}

/* LINTERPOL(NN,XX,YY,X,IEX=1,IERR), spher_expan.f:593-623 */
double orc_linterpol(int NN, const double* XX, const double* YY, double X) {
  if (X < XX[0]) return (YY[0] - YY[1]) / (XX[0] - XX[1]) * (X - XX[0]) + YY[0];
  if (X > XX[NN - 1]) return (YY[NN - 1] - YY[NN - 2]) / (XX[NN - 1] - XX[NN - 2]) * (X - XX[NN - 1]) + YY[NN - 1];
  int I;
  for (I = 2; I <= NN; ++I)
    if (XX[I - 1] > X) break;
  if (I > NN) I = NN; /* X == XX(NN): Fortran would index one past the end; clamp */
  return (YY[I - 1] - YY[I - 2]) / (XX[I - 1] - XX[I - 2]) * (X - XX[I - 2]) + YY[I - 2];
}

/* GENER(U,L1MAX), spher_expan.f:363-407 with COEF1..8 of SPHER_EXPAN :293-303; arrays 1-based of size L1MAX+2 */
static void gener(double U, int L1MAX, const double* const* CO, double D6, double* P1, double* P2, double* P3, double* P4) {
  double DUP = 1.0 + U, DUM = 1.0 - U, DU = U * U;
  P1[1] = 1.0; P1[2] = U; P1[3] = 0.5 * (3.0 * DU - 1.0);
  P2[1] = 0.0; P2[2] = 0.0; P2[3] = 0.25 * DUP * DUP;
  P3[1] = 0.0; P3[2] = 0.0; P3[3] = 0.25 * DUM * DUM;
  P4[1] = 0.0; P4[2] = 0.0; P4[3] = D6 * (DU - 1.0);
  int LMAX = L1MAX - 1;
  for (int L1 = 3; L1 <= LMAX; ++L1) {
    double C1 = CO[0][L1], C2 = CO[1][L1], C3 = CO[2][L1], C4 = CO[3][L1], C5 = CO[4][L1], C6 = CO[5][L1], C7 = CO[6][L1], C8 = CO[7][L1];
    double CU1 = C2 * U, CU2 = C6 * U;
    int L2 = L1 + 1, L3 = L1 - 1;
    double DL = (double)L3;
    P1[L2] = C1 * (CU1 * P1[L1] - DL * P1[L3]);
    P2[L2] = C5 * ((CU2 - C7) * P2[L1] - C8 * P2[L3]);
    P3[L2] = C5 * ((CU2 + C7) * P3[L1] - C8 * P3[L3]);
    P4[L2] = C3 * (CU1 * P4[L1] - C4 * P4[L3]);
  }
}

/*
 * The expansion itself: one_calc's interpolation to the Gauss nodes (spher_expan.f:158-167) + SPHER_EXPAN (:269-358).
 *   angl[nang] in radians, F[6][nang] in the order F11,F22,F33,F44,F12,F34 (convertncdf.py:177);
 *   raw[6][ng] = AL1,AL2,AL3,AL4,BET1,BET2 as SPHER_EXPAN leaves them (NOT yet multiplied by CNORM).
 */
static void gsf_raw(int nang, const double* angl, const double* F, int ng, double* raw) {
  int NG = ng, L1MAX = ng;
  double* X = (double*)malloc(sizeof(double) * NG * 2);
  double* W = X + NG;
  orc_gauss(NG, 0, X, W);                              /* one_calc :158 */
  double* FN = (double*)malloc(sizeof(double) * 6 * NG);
  for (int i = 0; i < NG; ++i) {
    double ang = acos(X[i]);                           /* :160 */
    for (int k = 0; k < 6; ++k) FN[k * NG + i] = orc_linterpol(nang, angl, F + (size_t)k * nang, ang);  /* :161-166 */
  }
  size_t LN = (size_t)L1MAX + 2;
  double* buf = (double*)calloc(LN * (8 + 4 + 6), sizeof(double));
  double* CO[8];
  for (int k = 0; k < 8; ++k) CO[k] = buf + LN * k;
  double* P1 = buf + LN * 8, *P2 = P1 + LN, *P3 = P2 + LN, *P4 = P3 + LN;
  double* AL1 = P4 + LN, *AL2 = AL1 + LN, *AL3 = AL2 + LN, *AL4 = AL3 + LN, *BET1 = AL4 + LN, *BET2 = BET1 + LN;
  for (int L1 = 3; L1 <= L1MAX; ++L1) {                 /* DO 150, :293-303 */
    int L = L1 - 1;
    CO[0][L1] = 1.0 / (double)(L + 1);
    CO[1][L1] = (double)(2 * L + 1);
    CO[2][L1] = 1.0 / sqrt((double)((L + 1) * (L + 1) - 4));
    CO[3][L1] = sqrt((double)(L * L - 4));
    CO[4][L1] = 1.0 / ((double)L * (double)((L + 1) * (L + 1) - 4));
    CO[5][L1] = (double)(2 * L + 1) * (double)(L * (L + 1));
    CO[6][L1] = (double)((2 * L + 1) * 4);
    CO[7][L1] = (double)(L + 1) * (double)(L * L - 4);
  }
  double D6 = 0.25 * sqrt(6.0);                        /* :312 */
  for (int I = 0; I < NG; ++I) {                        /* DO 300 */
    gener(X[I], L1MAX, (const double* const*)CO, D6, P1, P2, P3, P4);
    double WI = W[I];
    double FF11 = FN[0 * NG + I] * WI, FF22 = FN[1 * NG + I] * WI, FF33 = FN[2 * NG + I] * WI;
    double FF44 = FN[3 * NG + I] * WI, FF12 = FN[4 * NG + I] * WI, FF34 = FN[5 * NG + I] * WI;
    double FP = FF22 + FF33, FM = FF22 - FF33;
    for (int L1 = 1; L1 <= L1MAX; ++L1) {               /* DO 260 */
      AL1[L1] += FF11 * P1[L1];
      AL4[L1] += FF44 * P1[L1];
      AL2[L1] += FP * P2[L1];
      AL3[L1] += FM * P3[L1];
      BET1[L1] += FF12 * P4[L1];
      BET2[L1] += FF34 * P4[L1];
    }
  }
  for (int L1 = 1; L1 <= L1MAX; ++L1) {                 /* DO 350 */
    double CL = (double)(L1 - 1) + 0.5;
    AL1[L1] *= CL;
    double A2 = AL2[L1] * CL * 0.5, A3 = AL3[L1] * CL * 0.5;
    AL2[L1] = A2 + A3;
    AL3[L1] = A2 - A3;
    AL4[L1] *= CL;
    BET1[L1] *= CL;
    BET2[L1] *= CL;
  }
  double* S[6] = {AL1, AL2, AL3, AL4, BET1, BET2};
  for (int k = 0; k < 6; ++k)
    for (int L1 = 1; L1 <= L1MAX; ++L1) raw[(size_t)k * ng + (L1 - 1)] = S[k][L1];
  free(buf); free(FN); free(X);
}

/*
 * one_calc (spher_expan.f:120-180) + SPHER_EXPAN (:269-358) + the normalisation/output block of main (:93-107).
 *   ang_deg[nang], F[6][nang] in the order F11,F22,F33,F44,F12,F34 (convertncdf.py:177);
 *   coef[6][ng] = AL1,AL2,AL3,AL4,BET1,BET2 (times CNORM); returns CNORM = 1/AL1(1).
 *   quantize10 != 0 rounds to 10 decimals like the '(X,I5,6F17.10)' record (:96,:104).
 */
double orc_gsf_expand(int nang, const double* ang_deg, const double* F, int ng, double* coef, int quantize10) {
  const double PI = acos(-1.0), D2R = PI / 180.0;      /* params.h:3-4 */
  double* angl = (double*)malloc(sizeof(double) * nang);
  for (int i = 0; i < nang; ++i) angl[i] = ang_deg[i] * D2R;   /* READMATRIX :260-263 */
  double* raw = (double*)malloc(sizeof(double) * 6 * ng);
  gsf_raw(nang, angl, F, ng, raw);
  double CNORM = 1.0 / raw[0];                          /* main :95 */
  for (int k = 0; k < 6; ++k)
    for (int l = 0; l < ng; ++l) {
      double v = raw[(size_t)k * ng + l] * CNORM;
      if (quantize10) v = rint(v * 1e10) / 1e10;
      coef[(size_t)k * ng + l] = v;
    }
  free(raw); free(angl);
  return CNORM;
}

void orc_gsf_matr(int ng, const double* coef, int nang, const double* angl_rad, double* out);

/*
 * The diagnostic half of the program: READMATRIX's alternative angle grid (:235-258, USE_ALT_ANG = 1, params.h:13),
 * one_calc's two MATR + ERREVAL calls (:168-177; ERRTYP = MAXABS over [ang_min, ang_max] = [0, 180] deg, params.h:8-10)
 * and what main writes to <file>.expan_matr (:84-90): the matrix re-synthesised from the UN-normalised coefficients at the
 * input angles.  fout[6][nang] = F11OUT,F22OUT,F33OUT,F44OUT,F12OUT,F34OUT; returns fiterr.
 */
double orc_gsf_one_calc(int nang, const double* ang_deg, const double* F, int ng, double* fout) {
  const double PI = acos(-1.0), D2R = PI / 180.0;
  const double ang_min = 0.0 * D2R, ang_max = 180.0 * D2R;
  double* angl = (double*)malloc(sizeof(double) * nang * 4);
  double* alt = angl + nang, *f11alt = alt + nang, *tmp = f11alt + nang;
  /* READMATRIX: alternative grid and F11ALT, both still in degrees (:237-251; 1-based i) */
  alt[0] = ang_deg[0];
  alt[nang - 1] = ang_deg[nang - 1];
  for (int i = 2; i <= nang / 2; ++i) alt[i - 1] = 0.5 * (ang_deg[i - 2] + ang_deg[i - 1]);
  for (int i = nang / 2 + 1; i <= nang - 1; ++i) alt[i - 1] = 0.5 * (ang_deg[i] + ang_deg[i - 1]);
  for (int i = 0; i < nang; ++i) f11alt[i] = orc_linterpol(nang, ang_deg, F, alt[i]);
  for (int i = 0; i < nang; ++i) {                      /* :260-263 */
    angl[i] = ang_deg[i] * D2R;
    alt[i] = alt[i] * D2R;
  }
  double* raw = (double*)malloc(sizeof(double) * 6 * ng);
  gsf_raw(nang, angl, F, ng, raw);
  double* out = (double*)malloc(sizeof(double) * 6 * nang);
  double err_alt = 0.0, err = 0.0;
  orc_gsf_matr(ng, raw, nang, alt, out);                /* :169 */
  for (int i = 0; i < nang; ++i) {                      /* ERREVAL(NANG,angl,F11ALT,F11OUT,...), :170: range test on angl */
    if (angl[i] < ang_min || angl[i] > ang_max) continue;
    double t = fabs(f11alt[i] - out[i]);
    if (t > err_alt) err_alt = t;
  }
  orc_gsf_matr(ng, raw, nang, angl, fout);              /* :173 */
  for (int i = 0; i < nang; ++i) {
    if (angl[i] < ang_min || angl[i] > ang_max) continue;
    double t = fabs(F[i] - fout[i]);
    if (t > err) err = t;
  }
  (void)tmp;
  free(out); free(raw); free(angl);
  return err > err_alt ? err : err_alt;                 /* :176 */
}

/*
 * MATR (spher_expan.f:419-517): re-synthesise F11,F22,F33,F44,F12,F34 at angles angl_rad[nang] from the coefficients.
 * Used by the tests as a round-trip check of the expansion.  out[6][nang].
 */
void orc_gsf_matr(int ng, const double* coef, int nang, const double* angl_rad, double* out) {
  const double* A1 = coef, *A2 = coef + ng, *A3 = coef + 2 * ng, *A4 = coef + 3 * ng, *B1 = coef + 4 * ng, *B2 = coef + 5 * ng;
  int LMAX = ng - 1, L1MAX = ng;
  double D6 = sqrt(6.0) * 0.25;
  for (int I1 = 0; I1 < nang; ++I1) {
    double U = cos(angl_rad[I1]);
    double F11 = 0, F2 = 0, F3 = 0, F44 = 0, F12 = 0, F34 = 0, P1 = 0, P2 = 0, P3 = 0, P4 = 0;
    double PP1 = 1.0, PP2 = 0.25 * (1.0 + U) * (1.0 + U), PP3 = 0.25 * (1.0 - U) * (1.0 - U), PP4 = D6 * (U * U - 1.0);
    for (int L1 = 1; L1 <= L1MAX; ++L1) {
      int L = L1 - 1;
      double DL = (double)L, DL1 = (double)L1, PL1 = (double)(2 * L + 1), P;
      F11 += A1[L1 - 1] * PP1;
      F44 += A4[L1 - 1] * PP1;
      if (L != LMAX) {
        P = (PL1 * U * PP1 - DL * P1) / DL1;
        P1 = PP1;
        PP1 = P;
      }
      if (L < 2) continue;
      F2 += (A2[L1 - 1] + A3[L1 - 1]) * PP2;
      F3 += (A2[L1 - 1] - A3[L1 - 1]) * PP3;
      F12 += B1[L1 - 1] * PP4;
      F34 += B2[L1 - 1] * PP4;
      if (L == LMAX) continue;
      double PL2 = DL * DL1 * U, PL3 = DL1 * (DL * DL - 4.0), PL4 = 1.0 / (DL * (DL1 * DL1 - 4.0));
      P = (PL1 * (PL2 - 4.0) * PP2 - PL3 * P2) * PL4; P2 = PP2; PP2 = P;
      P = (PL1 * (PL2 + 4.0) * PP3 - PL3 * P3) * PL4; P3 = PP3; PP3 = P;
      P = (PL1 * U * PP4 - sqrt(DL * DL - 4.0) * P4) / sqrt(DL1 * DL1 - 4.0); P4 = PP4; PP4 = P;
    }
    out[0 * nang + I1] = F11;
    out[1 * nang + I1] = (F2 + F3) * 0.5;
    out[2 * nang + I1] = (F2 - F3) * 0.5;
    out[3 * nang + I1] = F44;
    out[4 * nang + I1] = F12;
    out[5 * nang + I1] = F34;
  }
}
