// gm_gram.cuh -- Gram-matrix form of the angular contraction + size-distribution reduction (sm_100a, FP64 DMMA).
//
// What it replaces: the same reference code as k_contract (mie_S12_backend_pt, mie_props.py:133-150; calculateScatVals,
// dointegration.py:1044-1050; the phase part of integratePSD, :1164-1166), for particle groups with nmax <= 64.
//
// Idea.  With S+(u) = sum_n c+_n p_n(u), S-(u) = sum_n c-_n q_n(u) (see k_contract) the four weighted sums a table cell needs,
//     A(u) = sum_p |S+|^2,  B(u) = sum_p |S-|^2,  Cr(u) + i Ci(u) = sum_p S+ conj(S-),
// are quadratic forms in the angle functions:
//     A = p^T H1 p,  B = q^T H2 q,  Cr = p^T H3 q,  Ci = p^T H4 q,
//     H1 = X+ X+^T, H2 = X- X-^T, H3 = X+ X-^T, H4 = X~+ X-^T            (N x N, N = max nmax)
// where row n of X+ (X-) holds (Re, Im) of c+_n (c-_n) of every particle and X~+ is X+ with (re, im) -> (im, -re).
// The reduction over particles becomes the K dimension of a small GEMM (cost ~ 8 N^2 FMA per particle instead of
// (4 N + 8) N_ang), the angle grid enters only once per cell (k_gram_eval).  For optics_SU (mean nmax 7.4, 371 angles) this
// is ~9x less FP64 work than the per-angle contraction; numerically the two forms agree to ~1e-14 of max P11.
//
//   k_gram       CTA = (class of particle groups, range of tasks).  Groups are classed by tg = ceil(max nmax / 8) (number of
//                8-row DMMA tiles); a class fixes how the 4 tg^2 accumulator tiles are split over `S` warps (a team) and how
//                many teams (12 / S) work on different groups at the same time.  One TMA producer warp streams the
//                coefficient rows of each group (contiguous in the stream written by k_coeff) into a shared-memory ring
//                (one lane of the producer warp per team); end-of-task markers travel through the same ring, so a CTA runs
//                through several tasks without any CTA-wide barrier.
//   k_gram_eval  CTA = task: sums the partial H (fixed order: deterministic), D = L^T H on DMMA, out = rowsum(D .* R).
#pragma once
#include "gm_mie_kernels.cuh"

#ifndef GM_GRAM_SYMMETRIC
#define GM_GRAM_SYMMETRIC 0   // bit mask (1: 2-warp teams, 2: 4-warp teams, 4: 12-warp teams; 8: skip the ragged duplicate tiles): H1 = X+ X+^T, H2 = X- X-^T from the tiles on and above the diagonal only (mirrored at the flush; -20 % DMMAs on
                              // optics_SU).  Measured SLOWER (k_gram 0.871 -> 0.989 ms: the skipped tiles leave the H1/H2 warps of a team idle while the
                              // H3/H4 warps set its pace, and the predicated tile loops schedule worse), so it stays off.
#endif
constexpr int GM_GRAM_MAX_TG = 8;                                    // groups with max nmax <= 64 take the Gram path
constexpr int GM_GRAM_THREADS = (GM_CONTRACT_WARPS + 1) * 32;        // 12 consumer warps + 1 producer warp
constexpr int GM_GRAM_MAX_SLOTS = 48;
constexpr int GM_GRAM_RING_DBL = 24 * 8 * GM_SB;                     // 202 752 B
constexpr int GM_GRAM_SMEM = GM_GRAM_RING_DBL * 8 + 2 * GM_GRAM_MAX_SLOTS * 8 + GM_GRAM_MAX_SLOTS * 4 + 64 * 4;

struct GramDesc {
  int tg;          // template class: 1..8 tiles, 0 = stacked c+/c- tile of groups with nmax <= 4
  int nteam;       // teams writing separate partials
  int gbegin, gend;  // range in glist
  long long hoff;  // offset (doubles) of this descriptor's partials inside a task's partial-H block
};
struct GramItem {
  int desc, t0, t1;
};

struct GramArgs {
  int ngroup;
  const GramItem* items;
  const GramDesc* desc;
  const int* glist;
  const int* grow;
  const int* gk4;
  const unsigned char* gact;   // [ntask][ngroup]
  const double* coef;
  long long task_stride;
  double* hpart;               // [ntask][hstride]
  long long hstride;
};

// TG = number of 8-row DMMA tiles of the class (1..8); TG = 0 is the STACKED class of groups with max nmax <= 4: rows 0-3 of
// the single 8-row tile hold c+ of orders 1..4 and rows 4-7 hold c- (Y = [X+; X-]), so that ONE product Y Y^T delivers H1
// (top-left 4 x 4), H3 (top-right) and H2 (bottom-right) and a second one, [X~+; 0] Y^T, delivers H4: 2 DMMAs per k-step
// instead of the 4 of class 1 (over half of the optics_SU groups are in this class).
template <int TG>
struct GramCfg {
  static constexpr int TGE = TG == 0 ? 1 : TG;                                  // tiles per side
  static constexpr int S = TG <= 1 ? 1 : TG == 2 ? 2 : TG <= 4 ? 4 : 12;        // warps per team
  static constexpr int NTEAM = GM_CONTRACT_WARPS / S;
  static constexpr int DEPTH = TG == 0 ? 4 : TG <= 4 ? 2 : TG <= 6 ? 4 : 3;     // ring slots per team
  static constexpr int R = NTEAM * DEPTH;
  static constexpr int SLOT_DBL = (TG == 0 ? 4 : 8 * TG) * GM_SB;
  static constexpr int CW = (TG + 2) / 3;                                       // column tiles per warp when S == 12
  static constexpr int NJOB = TG == 0 ? 2 : S == 1 ? 4 : S == 2 ? 2 : 1;
  static constexpr int NI = TGE;
  static constexpr int NJ = S == 12 ? CW : TGE;
  static constexpr int ND = TG == 0 ? 4 : 8 * TG;                               // side of the partial Gram blocks written by a team
  static_assert(R <= GM_GRAM_MAX_SLOTS && R * SLOT_DBL <= GM_GRAM_RING_DBL, "ring does not fit");
};

// Ring protocol.  Team t owns the DEPTH slots [t DEPTH, (t + 1) DEPTH) and whole TASKS of the work item (t0 + t, t0 + t + NTEAM, ...);
// lane t of the producer warp feeds them (the lanes run independently, so a slow team never blocks the others).  tags[slot] = number of coefficient rows copied into the slot
// (rows beyond are treated as zero by the consumers, nothing is copied for them), 0 = end-of-task marker.
template <int TG>
__device__ __forceinline__ void gram_cta(const GramArgs& A, const GramItem it, const GramDesc d, double* ring, uint64_t* full,
                                         uint64_t* empty, volatile int* tags, volatile int* cand) {
  using C = GramCfg<TG>;
  constexpr int S = C::S, NTEAM = C::NTEAM, DEPTH = C::DEPTH, R = C::R, SLOT_DBL = C::SLOT_DBL, NI = C::NI, NJ = C::NJ, NJOB = C::NJOB;
  constexpr int N = C::ND;
  constexpr bool GM_GRAM_SYM_S = (GM_GRAM_SYMMETRIC & (S == 2 ? 1 : S == 4 ? 2 : S == 12 ? 4 : 0)) != 0;   // mirrored tiles in this class
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int lk = lane & 3, lr = lane >> 2;

  if (threadIdx.x == 0) {
    for (int s = 0; s < R; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], S);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  if (warp == GM_CONTRACT_WARPS) {
    // ------------------------------------------------------------------------------------------ producer warp
    // Team t works through the tasks it.t0 + t, it.t0 + t + NTEAM, ... of the work item, each with ALL groups of the descriptor:
    // one partial H per (task, descriptor) instead of one per team, and NTEAM times fewer end-of-task flushes (with 2..12 groups per
    // task and class, as on optics_SU, the flushes of the round-robin group deal cost as much as the DMMAs).  Lane t of this warp
    // feeds team t; the lanes run independently.
    if (lane < NTEAM) {
      int m = 0;      // slots produced so far by this lane
      for (int task = it.t0 + lane; task < it.t1; task += NTEAM) {
        const unsigned char* ga = A.gact + (size_t)task * A.ngroup;
        const double* coef_t = A.coef + (size_t)task * A.task_stride;
        for (int idx = d.gbegin; idx < d.gend; ++idx) {
          const int g = A.glist[idx];
          if (ga[g] == 0) continue;
          const int row = A.grow[g], nr = GM_KSTEP * A.gk4[g];
          const int s = lane * DEPTH + m % DEPTH;
          if (m >= DEPTH) mbar_wait(&empty[s], ((m / DEPTH) - 1) & 1);
          ++m;
          tags[s] = nr;
          mbar_expect_tx(&full[s], (uint32_t)(nr * GM_SB * 8));
          bulk_g2s(ring + (size_t)s * SLOT_DBL, coef_t + (size_t)row * GM_SB, (uint32_t)(nr * GM_SB * 8), &full[s]);
        }
        const int s = lane * DEPTH + m % DEPTH;       // end-of-task marker
        if (m >= DEPTH) mbar_wait(&empty[s], ((m / DEPTH) - 1) & 1);
        ++m;
        tags[s] = 0;
        mbar_arrive(&full[s]);
      }
    }
    return;
  }

  // ---------------------------------------------------------------------------------------------- consumer warps
  const int team = warp / S, r = warp % S;
  // role of this warp: which block(s) of H, which operand halves of a coefficient row (c+ at 0, c- at 64)
  int blk = 0, aoff = 0, boff = 0, j0 = 0;
  bool tilde = false;
  if (S >= 4) {
    blk = S == 12 ? r / 3 : r;
    j0 = S == 12 ? (r % 3) * C::CW : 0;
    aoff = blk == 1 ? 64 : 0;
    boff = blk == 0 ? 0 : 64;
    tilde = blk == 3;
  }
  const double sgn = (lk & 1) ? -1.0 : 1.0;
  // stacked class: tile row lr = order (lr & 3) + 1 of c+ (lr < 4) or c- (lr >= 4, 64 doubles further in the coefficient row)
  const int lane_off = TG == 0 ? (lr & 3) * GM_SB + (lr >> 2) * 64 + lk : lr * GM_SB + lk;

  double acc[NJOB][NI][NJ][2];
#pragma unroll
  for (int q = 0; q < NJOB; ++q)
#pragma unroll
    for (int i = 0; i < NI; ++i)
#pragma unroll
      for (int j = 0; j < NJ; ++j) acc[q][i][j][0] = acc[q][i][j][1] = 0.0;

  int task = it.t0 + team;
  if (task >= it.t1) return;                      // fewer tasks than teams in this work item
  for (int m = 0;; ++m) {
    const int s = team * DEPTH + m % DEPTH;
    mbar_wait(&full[s], (m / DEPTH) & 1);
    const int nr = tags[s];
    if (nr > 0) {
      const double* st = ring + (size_t)s * SLOT_DBL + lane_off;
      const int nrl = nr - lr;   // row 8 i + lr of the slot was copied iff 8 i < nrl (otherwise it is a zero row)
      if constexpr (TG == 0) {
        const bool have = (lr & 3) < nr;
#pragma unroll 4
        for (int ks = 0; ks < 16; ++ks) {
          const double y = have ? st[4 * ks] : 0.0;
          const double yx = __shfl_xor_sync(0xffffffffu, y, 1);
          const double yt = lr < 4 ? sgn * yx : 0.0;
          dmma884(acc[0][0][0][0], acc[0][0][0][1], y, y);    // Y Y^T: H1 | H3 / . | H2
          dmma884(acc[1][0][0][0], acc[1][0][0][1], yt, y);   // [X~+; 0] Y^T: . | H4
        }
      } else if constexpr (S == 1) {
#pragma unroll 4
        for (int ks = 0; ks < 16; ++ks) {
          const double fp = 0 < nrl ? st[4 * ks] : 0.0, fm = 0 < nrl ? st[64 + 4 * ks] : 0.0;
          const double ft = sgn * __shfl_xor_sync(0xffffffffu, fp, 1);
          dmma884(acc[0][0][0][0], acc[0][0][0][1], fp, fp);   // H1
          dmma884(acc[1][0][0][0], acc[1][0][0][1], fm, fm);   // H2
          dmma884(acc[2][0][0][0], acc[2][0][0][1], fp, fm);   // H3
          dmma884(acc[3][0][0][0], acc[3][0][0][1], ft, fm);   // H4
        }
      } else if constexpr (S == 2) {
        // r = 0: H1 = (X+, X+), H2 = (X-, X-);  r = 1: H3 = (X+, X-), H4 = (X~+, X-)
#pragma unroll 2
        for (int ks = 0; ks < 16; ++ks) {
          double fp[NI], fm[NI], a1[NI], b0[NI];
#pragma unroll
          for (int i = 0; i < NI; ++i) {
            fp[i] = 8 * i < nrl ? st[i * 8 * GM_SB + 4 * ks] : 0.0;
            fm[i] = 8 * i < nrl ? st[i * 8 * GM_SB + 64 + 4 * ks] : 0.0;
            const double ft = sgn * __shfl_xor_sync(0xffffffffu, fp[i], 1);
            a1[i] = r ? ft : fm[i];
            b0[i] = r ? fm[i] : fp[i];
          }
#pragma unroll
          for (int i = 0; i < NI; ++i)
#pragma unroll
            for (int j = 0; j < NI; ++j) {
              if ((GM_GRAM_SYMMETRIC & 1) && r == 0 && j < i) continue;      // H1, H2 are symmetric: tile (i, j) = tile (j, i)^T
              dmma884(acc[0][i][j][0], acc[0][i][j][1], fp[i], b0[j]);
              dmma884(acc[1][i][j][0], acc[1][i][j][1], a1[i], fm[j]);
            }
        }
      } else {
        const double* sa = st + aoff;
        const double* sb = st + boff;
        const int jlast = (S == 12 && j0 + NJ > TG) ? TG - 1 - j0 : NJ - 1;   // ragged last third: its column tiles beyond TG do not exist
        const bool sym = (GM_GRAM_SYMMETRIC & (S == 12 ? 4 : 2)) && blk < 2;                       // H1, H2: only the tiles on and above the diagonal
#pragma unroll 2
        for (int ks = 0; ks < 16; ++ks) {
          double b[NJ];
#pragma unroll
          for (int j = 0; j < NJ; ++j) {
            const int jt = j0 + (j < jlast ? j : jlast);
            b[j] = 8 * jt < nrl ? sb[jt * 8 * GM_SB + 4 * ks] : 0.0;
          }
#pragma unroll
          for (int i = 0; i < NI; ++i) {
            double a = 8 * i < nrl ? sa[i * 8 * GM_SB + 4 * ks] : 0.0;
            const double at = sgn * __shfl_xor_sync(0xffffffffu, a, 1);
            a = tilde ? at : a;
#pragma unroll
            for (int j = 0; j < NJ; ++j) {
              if (((GM_GRAM_SYMMETRIC & 8) && j > jlast) || (sym && i > j0 + j)) continue;
              dmma884(acc[0][i][j][0], acc[0][i][j][1], a, b[j]);
            }
          }
        }
      }
    } else {
      // end of task: this team's partial H -> global (fixed slot: summed in fixed order by k_gram_sum), then reset
      double* hp = A.hpart + (size_t)task * A.hstride + d.hoff;     // the only partial of this (task, descriptor)
      if constexpr (TG == 0) {
        // 4 x 4 blocks: lane (lr, lk) holds rows lr, columns 2 lk, 2 lk + 1 of the two 8 x 8 products
        const int r4 = lr & 3, c4 = 2 * (lk & 1);
        const double2 g = make_double2(acc[0][0][0][0], acc[0][0][0][1]), tq = make_double2(acc[1][0][0][0], acc[1][0][0][1]);
        if (lr < 4 && lk < 2) *reinterpret_cast<double2*>(hp + (0 * 4 + r4) * 4 + c4) = g;      // H1
        if (lr >= 4 && lk >= 2) *reinterpret_cast<double2*>(hp + (1 * 4 + r4) * 4 + c4) = g;    // H2
        if (lr < 4 && lk >= 2) {
          *reinterpret_cast<double2*>(hp + (2 * 4 + r4) * 4 + c4) = g;                          // H3
          *reinterpret_cast<double2*>(hp + (3 * 4 + r4) * 4 + c4) = tq;                         // H4
        }
        acc[0][0][0][0] = acc[0][0][0][1] = acc[1][0][0][0] = acc[1][0][0][1] = 0.0;
      }
#pragma unroll
      for (int q = 0; q < (TG == 0 ? 0 : NJOB); ++q) {
        int b = blk;
        if (S == 1) b = q;
        if (S == 2) b = q == 0 ? (r ? 2 : 0) : (r ? 3 : 1);
#pragma unroll
        for (int i = 0; i < NI; ++i)
#pragma unroll
          for (int j = 0; j < NJ; ++j) {
            const int jc = j0 + j;
            if (jc < TG) {
              if (!(GM_GRAM_SYM_S && b < 2 && i > jc))
                *reinterpret_cast<double2*>(hp + ((size_t)b * N + 8 * i + lr) * N + 8 * jc + 2 * lk) =
                    make_double2(acc[q][i][j][0], acc[q][i][j][1]);
              if (GM_GRAM_SYM_S && b < 2 && i < jc) {       // the mirror tile (same products, same order: bit-identical)
                hp[((size_t)b * N + 8 * jc + 2 * lk) * N + 8 * i + lr] = acc[q][i][j][0];
                hp[((size_t)b * N + 8 * jc + 2 * lk + 1) * N + 8 * i + lr] = acc[q][i][j][1];
              }
            }
            acc[q][i][j][0] = acc[q][i][j][1] = 0.0;
          }
      }
      task += NTEAM;
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(&empty[s]);
    if (nr == 0 && task >= it.t1) break;
  }
}

// 128 registers (220 B of spills in the widest instantiations) is the ceiling: the 13 warps of the CTA land 4 / 3 / 3 / 3 on the four SM
// sub-partitions, and four warps of 144 registers (the next step, which would not spill) exceed one sub-partition's 16,384-register file --
// a __maxnreg__(144) build fails at launch with "too many resources requested".
__global__ void __launch_bounds__(GM_GRAM_THREADS, 1) k_gram(GramArgs A) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  double* ring = reinterpret_cast<double*>(smem_raw);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw + (size_t)GM_GRAM_RING_DBL * 8);
  uint64_t* empty = full + GM_GRAM_MAX_SLOTS;
  volatile int* tags = reinterpret_cast<volatile int*>(empty + GM_GRAM_MAX_SLOTS);
  volatile int* cand = tags + GM_GRAM_MAX_SLOTS;   // producer scratch: (first row, rows) of up to 32 candidate groups
  const GramItem it = A.items[blockIdx.x];
  const GramDesc d = A.desc[it.desc];
  switch (d.tg) {
    case 0: gram_cta<0>(A, it, d, ring, full, empty, tags, cand); break;
    case 1: gram_cta<1>(A, it, d, ring, full, empty, tags, cand); break;
    case 2: gram_cta<2>(A, it, d, ring, full, empty, tags, cand); break;
    case 3: gram_cta<3>(A, it, d, ring, full, empty, tags, cand); break;
    case 4: gram_cta<4>(A, it, d, ring, full, empty, tags, cand); break;
    case 5: gram_cta<5>(A, it, d, ring, full, empty, tags, cand); break;
    case 7: gram_cta<7>(A, it, d, ring, full, empty, tags, cand); break;
    case 6: gram_cta<6>(A, it, d, ring, full, empty, tags, cand); break;
    default: gram_cta<8>(A, it, d, ring, full, empty, tags, cand); break;
  }
}

// ================================================================================================ k_gram_sum
// H[task][4][N][N] = sum of the partial Gram blocks of every (descriptor, team) in fixed order (deterministic).  One thread
// per pair of adjacent columns; the partials of the teams of a descriptor are independent loads (latency overlapped).
constexpr int GM_GRAM_SUM_MAXDESC = 48;
struct GramSumArgs {
  int ndesc;
  GramDesc desc[GM_GRAM_SUM_MAXDESC];   // by value: read from the constant bank, no dependent global load per descriptor
  const double* hpart;
  long long hstride;
  int N;          // 8 * largest class
  double* hsum;   // [ntask][4][N][N]
};

__global__ void __launch_bounds__(256) k_gram_sum(GramSumArgs A) {
  const int task = blockIdx.y;
  const int N = A.N, half = N / 2;
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= 4 * N * half) return;
  const int b = e / (N * half), n = (e / half) % N, c = 2 * (e % half);
  const double* hp = A.hpart + (size_t)task * A.hstride;
  double2 s = make_double2(0.0, 0.0);
  for (int k = 0; k < A.ndesc; ++k) {
    const GramDesc d = A.desc[k];
    const int Nd = d.tg ? 8 * d.tg : 4;   // class 0 (stacked, nmax <= 4) writes 4 x 4 blocks
    if (n < Nd && c < Nd) {
      const double2* p = reinterpret_cast<const double2*>(hp + d.hoff + ((size_t)b * Nd + n) * Nd + c);
      const size_t stride = (size_t)2 * Nd * Nd;   // 4 Nd^2 doubles per slot
      // slots in order (deterministic); k_gram's descriptors have ONE slot since a team owns whole tasks, k_small's up to 12: the loads of
      // four slots are issued together
      int t = 0;
      for (; t + 4 <= d.nteam; t += 4) {
        const double2 v0 = p[t * stride], v1 = p[(t + 1) * stride], v2 = p[(t + 2) * stride], v3 = p[(t + 3) * stride];
        s.x = (((s.x + v0.x) + v1.x) + v2.x) + v3.x;
        s.y = (((s.y + v0.y) + v1.y) + v2.y) + v3.y;
      }
      for (; t < d.nteam; ++t) {
        const double2 v = p[t * stride];
        s.x += v.x;
        s.y += v.y;
      }
    }
  }
  if (b < 2) {
    // H1, H2 are symmetric: hand k_gram_eval the upper triangle with doubled off-diagonal (p^T U p = p^T H p), so that it can
    // skip the tiles below the diagonal
    s.x = c > n ? 2.0 * s.x : (c == n ? s.x : 0.0);
    s.y = c + 1 > n ? 2.0 * s.y : (c + 1 == n ? s.y : 0.0);
  }
  *reinterpret_cast<double2*>(A.hsum + (((size_t)task * 4 + b) * N + n) * N + c) = s;
}

// ================================================================================================ k_gram_eval
// Angle-stationary evaluation of the four quadratic forms.  CTA = (block of 96 angles, range of tasks); each warp owns 8
// angles and keeps its p_n / q_n DMMA A-fragments in registers for the whole task range; H of the next task is fetched into
// shared memory (cp.async, double-buffered when it fits) while the current one is multiplied:  D = L^T H (DMMA), then
// out = rowsum(D .* R) with the right factors taken from the same fragments by warp shuffles.
struct GramEvalArgs {
  int ntask, tasks_per_cta;
  const double* hsum;   // [ntask][4][N][N]
  int N;                // 8 * largest class
  int ntile;            // column tiles that can be non-zero (template parameter of the launch)
  int nk4;              // k4 steps whose p/q table rows exist (<= 2 ntile, 4 nk4 <= nrows); fragments beyond are zero
  int nbuf;             // shared-memory buffers for H (1 or 2)
  const double* T;      // p/q angle table of the bin, [2][nrows][GM_TROW]
  int nrows;
  double* part;         // [ntask][nchunk][4][GM_NANG_PAD]
  int nchunk, chunk;
};

constexpr int GM_GRAM_EVAL_THREADS = 384;
constexpr int GM_GRAM_EVAL_SMEM_MAX = 220 * 1024;

__device__ __forceinline__ void cp_async16(void* dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
}

template <int NTILE>
__global__ void __launch_bounds__(GM_GRAM_EVAL_THREADS, 1) k_gram_eval(GramEvalArgs A) {
  constexpr int NK4 = 2 * NTILE;   // compile-time trip counts: guarding DMMAs with run-time bounds only predicates them off
  extern __shared__ __align__(16) double Hs[];   // [nbuf][4][N][NP]
  const int N = A.N, NP = N + 4;
  const size_t buf_dbl = (size_t)4 * N * NP;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int lk = lane & 3, lr = lane >> 2;
  const int mt = blockIdx.x * 12 + warp;          // m-tile of 8 angles in the padded 384-angle space
  const int half = mt / 24, acol = (mt % 24) * 8;
  const int t0 = blockIdx.y * A.tasks_per_cta, t1 = min(A.ntask, t0 + A.tasks_per_cta);

  auto prefetch = [&](int task, int buf) {
    const double* src = A.hsum + (size_t)task * 4 * N * N;
    double* dst = Hs + (size_t)buf * buf_dbl;
    const int halfN = N / 2;
    for (int e = threadIdx.x; e < 4 * N * halfN; e += blockDim.x) {
      const int row = e / halfN, c = 2 * (e % halfN);
      cp_async16(dst + (size_t)row * NP + c, src + (size_t)row * N + c);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  if (t0 < t1) prefetch(t0, 0);

  // A fragments: lane (lr, lk) holds L[angle acol + lr][n = 4 ks + lk] for p (ap) and q (aq)
  double ap[NK4], aq[NK4];
  {
    const double* Th = A.T + (size_t)half * A.nrows * GM_TROW + acol + lr;
#pragma unroll
    for (int ks = 0; ks < NK4; ++ks) {
      ap[ks] = aq[ks] = 0.0;
      if (ks < A.nk4) {
        ap[ks] = Th[(size_t)(4 * ks + lk) * GM_TROW];
        aq[ks] = Th[(size_t)(4 * ks + lk) * GM_TROW + GM_LAH];
      }
    }
  }
  const int src0 = lr * 4 + 2 * (lk & 1);
  const bool hi = (lk >> 1) != 0;

  for (int task = t0; task < t1; ++task) {
    const int buf = A.nbuf == 2 ? ((task - t0) & 1) : 0;
    if (A.nbuf == 2 && task + 1 < t1) {
      prefetch(task + 1, buf ^ 1);
      asm volatile("cp.async.wait_group 1;" ::: "memory");
    } else {
      asm volatile("cp.async.wait_group 0;" ::: "memory");
    }
    __syncthreads();
    const double* Hb = Hs + (size_t)buf * buf_dbl + (size_t)lk * NP + lr;
    double out[4];
#pragma unroll
    for (int pair = 0; pair < 2; ++pair) {
      // pair 0: H1 (left p, right p) and H3 (left p, right q);  pair 1: H2 (left q, right q) and H4 (left p, right q)
      const int b0 = pair, b1 = pair + 2;
      double acc0[NTILE][2], acc1[NTILE][2];
#pragma unroll
      for (int j = 0; j < NTILE; ++j) acc0[j][0] = acc0[j][1] = acc1[j][0] = acc1[j][1] = 0.0;
      const double* H0 = Hb + (size_t)b0 * N * NP;
      const double* H1 = Hb + (size_t)b1 * N * NP;
#pragma unroll
      for (int ks = 0; ks < NK4; ++ks) {
        const double a0 = pair == 0 ? ap[ks] : aq[ks];
        const double a1 = ap[ks];
#pragma unroll
        for (int j = 0; j < NTILE; ++j) {
          if (8 * j + 7 >= 4 * ks)   // block b0 (H1 / H2) is upper triangular (k_gram_sum): tiles below the diagonal are zero
            dmma884(acc0[j][0], acc0[j][1], a0, H0[(size_t)4 * ks * NP + 8 * j]);
          dmma884(acc1[j][0], acc1[j][1], a1, H1[(size_t)4 * ks * NP + 8 * j]);
        }
      }
      // right factors R[angle][c], c = 8 j + 2 lk (+1): held by lane (lr, c % 4) in fragment ks = c / 4
      double o0 = 0.0, o1 = 0.0;
#pragma unroll
      for (int j = 0; j < NTILE; ++j) {
          const double pe0 = __shfl_sync(0xffffffffu, ap[2 * j], src0), pe1 = __shfl_sync(0xffffffffu, ap[2 * j + 1], src0);
          const double po0 = __shfl_sync(0xffffffffu, ap[2 * j], src0 + 1), po1 = __shfl_sync(0xffffffffu, ap[2 * j + 1], src0 + 1);
          const double qe0 = __shfl_sync(0xffffffffu, aq[2 * j], src0), qe1 = __shfl_sync(0xffffffffu, aq[2 * j + 1], src0);
          const double qo0 = __shfl_sync(0xffffffffu, aq[2 * j], src0 + 1), qo1 = __shfl_sync(0xffffffffu, aq[2 * j + 1], src0 + 1);
          const double pc = hi ? pe1 : pe0, pc1 = hi ? po1 : po0;   // p at columns c, c + 1
          const double qc = hi ? qe1 : qe0, qc1 = hi ? qo1 : qo0;   // q at columns c, c + 1
          const double r0a = pair == 0 ? pc : qc, r0b = pair == 0 ? pc1 : qc1;   // right factor of block b0
          o0 = fma(acc0[j][0], r0a, fma(acc0[j][1], r0b, o0));
          o1 = fma(acc1[j][0], qc, fma(acc1[j][1], qc1, o1));                     // blocks H3, H4: right factor q
      }
      o0 += __shfl_xor_sync(0xffffffffu, o0, 1);
      o0 += __shfl_xor_sync(0xffffffffu, o0, 2);
      o1 += __shfl_xor_sync(0xffffffffu, o1, 1);
      o1 += __shfl_xor_sync(0xffffffffu, o1, 2);
      out[b0] = o0;
      out[b1] = o1;
    }
    if (lk == 0) {
      double* o = A.part + (((size_t)task * A.nchunk + A.chunk) * 4) * GM_NANG_PAD + half * GM_HALF_ANG + acol + lr;
#pragma unroll
      for (int b = 0; b < 4; ++b) o[(size_t)b * GM_NANG_PAD] = out[b];
    }
    __syncthreads();   // everyone is done with `buf` before the next prefetch overwrites it
    if (A.nbuf == 1 && task + 1 < t1) prefetch(task + 1, 0);
  }
}


// ================================================================================================ k_gram_interp
// The four quadratic forms are POLYNOMIALS of degree <= 2 N in u = cos(theta) (p_n = pi_n + tau_n and q_n = pi_n - tau_n are polynomials
// of degree n), so k_gram_eval evaluates them at M = 2 N + 1 Chebyshev nodes only (81 instead of 371 angles for optics_SU) and this kernel
// carries them to the table's angles with the barycentric interpolation matrix W [M][GM_NANG_PAD] built on the host (exact for these
// polynomials; Lebesgue constant 3.7: numpy prototype 9e-15 of max |S+|^2).  grid = ntask, block = GM_NANG_PAD; fixed summation order.
__global__ void __launch_bounds__(GM_NANG_PAD) k_gram_interp(int nang, int M, const double* __restrict__ W, const double* __restrict__ node,
                                                            double* __restrict__ part, int nchunk, int chunk) {
  __shared__ __align__(16) double nv[GM_HALF_ANG][4];      // [node][form]: one thread reads its four node values with two 16-byte loads
  const int task = blockIdx.x, a = threadIdx.x;
  const double* src = node + (size_t)task * 4 * GM_NANG_PAD;
  for (int e = a; e < 4 * M; e += GM_NANG_PAD) nv[e % M][e / M] = src[(size_t)(e / M) * GM_NANG_PAD + e % M];
  __syncthreads();
  if (a >= nang) return;
  double s[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll 4
  for (int j = 0; j < M; ++j) {
    const double w = W[(size_t)j * GM_NANG_PAD + a];
    const double2 v01 = *reinterpret_cast<const double2*>(&nv[j][0]), v23 = *reinterpret_cast<const double2*>(&nv[j][2]);
    s[0] = fma(w, v01.x, s[0]);
    s[1] = fma(w, v01.y, s[1]);
    s[2] = fma(w, v23.x, s[2]);
    s[3] = fma(w, v23.y, s[3]);
  }
  double* o = part + (((size_t)task * nchunk + chunk) * 4) * GM_NANG_PAD + a;
#pragma unroll
  for (int b = 0; b < 4; ++b) o[(size_t)b * GM_NANG_PAD] = s[b];
}
