/*
 * geosmie_b200.h -- C ABI of libgeosmie_b200.so (hand-written sm_100a CUDA; no torch types, no exceptions).
 *
 * This is the drop-in boundary for the GEOSmie Mie lookup-table hot path.  Every entry point names the
 * reference interface it replaces (paths relative to the GEOS-ESM/GEOSmie tree).  Conventions:
 *   - all floating point is IEEE double; complex numbers are (re, im) pairs of doubles; m = n + i k, k > 0 absorbing;
 *   - all arrays are caller-owned, C-contiguous; the library never frees or retains caller memory;
 *   - functions return 0 on success or a negative GM_E* code; gm_last_error() gives a thread-local message;
 *   - `*_dev` variants take DEVICE pointers and enqueue asynchronously on the handle's stream (gm_set_stream),
 *     the plain variants take HOST pointers, copy in/out and synchronise before returning;
 *   - one handle per GPU, a handle must not be used from two threads at once.
 */
#ifndef GEOSMIE_B200_H
#define GEOSMIE_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GM_OK 0
#define GM_EINVAL (-1)
#define GM_ECUDA (-2)
#define GM_ENOMEM (-3)

/* number of raw size-distribution sums returned per (task, mode) by gm_table_run, see GM_S_* below */
#define GM_NSCAL 11
#define GM_S_W 0      /* sum w                      (integratePSD `num`, dointegration.py:1104) */
#define GM_S_X2W 1    /* sum x^2 w                  (rarr2, :1105; r = x*lam/2pi applied by the caller) */
#define GM_S_X3W 2    /* sum x^3 w                  (rarr3, :1106) */
#define GM_S_X4W 3    /* sum x^4 w                  (rarr4, :1107) */
#define GM_S_QEXT 4   /* sum qext x^2 w             (:1188-1190 'dumix', area weighting :1111-1112) */
#define GM_S_QSCA 5   /* sum qsca x^2 w */
#define GM_S_QABS 6   /* sum qabs x^2 w */
#define GM_S_QB 7     /* sum qb   x^2 w             (:1133-1145) */
#define GM_S_G 8      /* sum g qsca x^2 w           (thisweight *= qsca aliasing quirk, :1157) */
#define GM_S_CSCA 9   /* sum qsca^2 x^4 w           (csca = qsca pi r^2 with the (area*qsca) weight) */
#define GM_S_CEXT 10  /* sum qext qsca x^4 w */

/* flags for gm_table_run */
#define GM_F_ELIDE_ZERO_WEIGHT 1 /* skip particles whose weights are all exactly 0 (result-neutral) */
#define GM_F_NO_GRAM 2           /* evaluate every particle group with the per-angle contraction (k_contract); default: groups
                                    with nmax <= 64 go through the Gram-matrix form (k_gram + k_gram_eval), same sums */
#define GM_F_PHASE_ON_DEVICE 4   /* host-buffer calls: do not download the raw phase sums (out_phase may be NULL); they stay in the
                                    table's device buffer for gm_table_fetch_normalized / gm_gsf_expand_phase4_dev */

typedef struct gm_handle_s* gm_handle_t;
typedef struct gm_table_s* gm_table_t;

/* ---- lifetime ---------------------------------------------------------------------------------------------------- */
int gm_version(void);
const char* gm_last_error(void);
int gm_init(int device, gm_handle_t* out);
int gm_destroy(gm_handle_t h);
/* `cuda_stream` is a cudaStream_t (e.g. torch.cuda.current_stream().cuda_stream); NULL = default stream */
int gm_set_stream(gm_handle_t h, void* cuda_stream);
int gm_sync(gm_handle_t h);
/* kernels launched by this handle since gm_init (for bench.py's gpu_launches) */
int64_t gm_launch_count(gm_handle_t h);

/* ---- B1/B2: per-particle Mie ----------------------------------------------------------------------------------------
 * Replaces pymiecoated mie_coeffs(params) + mie_props + mie_S12 (mie_coeffs.py:36-73, :132-180, :183-251;
 * mie_props.py:28-70, :119-150) and the batch loop MultipleMie.calculateS12SizeRange (mie_coated.py:61-89).
 *   n         particles
 *   x[n]      size parameter used in the homogeneous formulas (the reference passes y when it falls back to
 *             single_mie_coeff, mie_coeffs.py:66-69 -- the Python wrapper performs that dispatch)
 *   xcore[n]  NULL for homogeneous spheres; else core size parameter (then x[] is the shell size parameter y)
 *   mz[n][2]  sqrt(eps*mu), mrel[n][2] sqrt(eps/mu)  (mie_coeffs.py:96-97); coated: m1 = sqrt(eps1) in mz, m2 = sqrt(eps2) in mrel
 *   mat_stride 0: one material broadcast to all particles, 1: per-particle material
 *   nmax[n]   number of terms, computed by the caller with the reference expression (mie_coeffs.py:99)
 *   bes_off/ajv/ayv  optional caller-supplied J_{k+1.5}(x), Y_{k+1.5}(x), k = 0..nmax-1, particle i at bes_off[i]
 *             (the reference's ajv/ayv inputs, mie_coated.py:260-263); NULL = computed on the device
 *   nang,u    cosines of the scattering angles (may be 0 / NULL)
 * outputs
 *   q[n][6]   qext,qsca,qabs,qb,asy,qratio  (mie_props_raw return order, mie_props.py:70)
 *   s12[n][nang][4]  Re S1, Im S1, Re S2, Im S2 (nullable)
 *   ab[sum nmax][4]  Re a_n, Im a_n, Re b_n, Im b_n concatenated per particle (nullable)
 */
int gm_mie_eval(gm_handle_t h, int n, const double* x, const double* xcore, const double* mz, const double* mrel,
                int mat_stride, const int32_t* nmax, const int64_t* bes_off, const double* ajv, const double* ayv,
                int nang, const double* u, double* q, double* s12, double* ab);

/* ---- B3: fused table cells ------------------------------------------------------------------------------------------
 * A table object holds the per-bin constants of dointegration.fun's bin loop (dointegration.py:787-798):
 * the size-parameter grid, nmax, the Riccati-Bessel tables psi_n(x), chi_n(x) (replaces
 * MultipleMie.preCalculateBessel, mie_coated.py:153-158) and the pi_n/tau_n angle table (replaces
 * preCalculatePT, mie_coated.py:160-179).
 */
int gm_table_create(gm_handle_t h, int nx, const double* x, const int32_t* nmax, int nang, const double* cos_theta,
                    gm_table_t* out);
int gm_table_destroy(gm_table_t t);
int gm_table_nx(gm_table_t t);
int gm_table_nang(gm_table_t t);
/* optional validation mode: replace the device-computed Riccati-Bessel tables by caller-supplied
 * J_{k+0.5}(x_i), Y_{k+0.5}(x_i), k = 0..nmax_i (note: nmax_i + 1 values per particle, particle i at off[i]) */
int gm_table_set_bessel(gm_table_t t, const int64_t* off, const double* jv_half, const double* yv_half);

/*
 * gm_table_run: for each task (one complex refractive index + one weight vector over the bin's x grid) evaluate
 * Mie at every x, form the Mueller elements at every angle and reduce over the size distribution.
 * Replaces rawMie + calculateScatVals + the reductions of integratePSD (dointegration.py:1211-1254, :1044-1050,
 * :1104-1107, :1164-1166, :1188-1190); the caller combines the raw sums exactly like integratePSD does.
 *   mz/mrel [ntask][2]          as in gm_mie_eval (mu = 1: both sqrt(eps))
 *   w_phase [ntask][nx]         number weights for the phase-matrix sums; must be >= 0 (sqrt(w) is folded into the coefficients).
 *                               Signed weights (the reference's non-monotonic 'du' grid): call once with max(w, 0) and once
 *                               with max(-w, 0) and subtract the phase sums -- the sums are linear in w
 *   w_scal  [ntask][nmode][nx]  number weights for the scalar sums (NULL: nmode must be 1 and w_phase is used)
 *   out_scal [ntask][nmode][GM_NSCAL]
 *   out_phase [ntask][4][nang]  sum_x w P(x,theta) for P11(=P22), P12, P33(=P44), P34
 * Deterministic: fixed reduction order, bitwise identical results run to run.
 * The host-buffer variant pipelines its transfers batch by batch against the kernels (pass pinned memory to benefit).
 */
int gm_table_run(gm_table_t t, int ntask, const double* mz, const double* mrel, int nmode, const double* w_phase,
                 const double* w_scal, int flags, double* out_scal, double* out_phase);
int gm_table_run_dev(gm_table_t t, int ntask, const double* mz, const double* mrel, int nmode, const double* w_phase,
                     const double* w_scal, int flags, double* out_scal, double* out_phase);
/*
 * gm_table_run_coated: like gm_table_run for coated (core + shell) spheres -- the table's x grid is the SHELL size
 * parameter y, task i has core index m1[i], shell index m2[i] (both sqrt(eps), as in gm_mie_eval) and core size
 * parameter core_ratio[i] * y (RH-dependent shell growth: core_ratio = 1 / growth factor).  Coefficients follow
 * coated_mie_coeff (mie_coeffs.py:183-251); the contraction, Mueller and size-distribution stages are shared with the
 * homogeneous path.  Extension: the reference has no coated table driver (mie_coated.py:80 is a dead branch).
 */
int gm_table_run_coated(gm_table_t t, int ntask, const double* m1, const double* m2, const double* core_ratio, int nmode,
                        const double* w_phase, const double* w_scal, int flags, double* out_scal, double* out_phase);

/*
 * gm_table_run_psd: like gm_table_run, but the number weights are generated on the device from per-(task, mode)
 * parameters (replaces the O(nx) numpy work of dointegration.calculatePSD :539-664 / particleparams.getLogNormPSD
 * :113-127; the scalars below are still derived on the host with the reference's formulas):
 *   GM_PSD_LOGNORM  params = xmode, xmin, xmax, ln(sigma)   in size-parameter space (r * 2 pi / lambda); needs gm_table_set_dr
 *   GM_PSD_SS       params = xconv (2 pi / lambda), rMinUse, rMaxUse, rrat     (Gong sea-salt, :585-624)
 *   GM_PSD_DU       params = xconv, rMinMaj, rMaxMaj, 0                        (r^-4 sub-bins, :626-662)
 *   psd_params [ntask][nmode][GM_PSD_NPAR], frac [ntask][nmode] (phase weight = sum_mode frac * w_mode)
 */
#define GM_PSD_LOGNORM 1
#define GM_PSD_SS 2
#define GM_PSD_DU 3
#define GM_PSD_NPAR 4
int gm_table_set_dr(gm_table_t t, const double* dr /*[nx] drarr of initializeXarr, dointegration.py:425-463*/);
int gm_table_run_psd(gm_table_t t, int ntask, const double* mz, const double* mrel, int nmode, int psd_kind,
                     const double* psd_params, const double* frac, int flags, double* out_scal, double* out_phase);
/* the generated weights of the last gm_table_run_psd (host copy, for validation): w [ntask][nmode][nx] */
int gm_table_get_weights(gm_table_t t, int ntask, int nmode, double* w);

/*
 * Fused GSF stage: after gm_table_set_gsf every gm_table_run* call also expands the phase sums of each finished batch
 * of tasks into generalized-spherical-function moments (the work of rungsf.py / gm_gsf_expand) right behind the
 * reduction kernel, and -- for the host-buffer calls -- downloads them batch by batch while later batches compute.
 *   ang_deg [nang of the table] scattering angles in degrees; ng moments (129 = NSPHER, params.h:14);
 *   coef_host [ntask][6][ng] / cnorm_host [ntask] must stay valid for the following run calls (may be NULL: device only,
 *   see gm_table_gsf_device).  ng = 0 switches the stage off.
 */
int gm_table_set_gsf(gm_table_t t, const double* ang_deg, int ng, int quantize10, double* coef_host, double* cnorm_host);
int gm_table_gsf_device(gm_table_t t, double** coef, double** cnorm);

/* device copies of the outputs of the last host-buffer gm_table_run (valid until the next call on this table), so that
 * gm_gsf_expand_phase4_dev can be chained without a host round trip */
int gm_table_device_outputs(gm_table_t t, double** out_scal, double** out_phase);
/* The a-posteriori normalisation of the table driver on the device-resident phase sums of the last host-buffer gm_table_run* call
 * of this table (replaces the numpy statements of dointegration.fun, dointegration.py:977-988, over [cells x 371] arrays):
 *   I = trapz(p11 sin(theta), theta);  p11n = 2 p11 / I;  pXX = pXX p11n / p11 for XX = 12, 33, 34;  pback = the values at the
 *   last angle.  theta_rad, sin_theta [nang] (host; the caller's own sin values, so the products match the reference's bits).
 * Host outputs, each one CONTIGUOUS plane: p11, p12, p33, p34 [ntask][nang] (the caller's final arrays: p22 = p11 and p44 = p33
 * for spheres), pback4 [ntask][4] in the order 11, 12, 33, 34.  The four plane pointers may all be NULL: only pback4 is delivered
 * (tables whose phase matrices are not kept, e.g. fine spectral grids for band averaging).  The trapezoid sum is taken in a fixed
 * order (deterministic). */
int gm_table_fetch_normalized(gm_table_t t, int ntask, const double* theta_rad, const double* sin_theta, double* p11, double* p12,
                              double* p33, double* p34, double* pback4);
/* The same normalisation without the download: *block = device pointer of [4][ntask][nang] (p11, p12, p33, p34) followed by
 * pback4 [ntask][4], valid until the next call on this table -- the source of a multi-GPU gm_peer_put. */
int gm_table_normalize_device(gm_table_t t, int ntask, const double* theta_rad, const double* sin_theta, double** block);
/* per-particle outputs through the table (DMMA) path: q [ntask][nx][6], s12 [ntask][nx][nang][4] (host pointers) */
int gm_table_particles(gm_table_t t, int ntask, const double* mz, const double* mrel, double* q, double* s12);
/* statistics of the last gm_table_run*: [0] particle evaluations, [1] sum of nmax over evaluated particles,
 * [2] sum of nmx, [3] executed DMMA k4-steps x 32 particles (padded work), [4] kernels launched */
int gm_table_last_stats(gm_table_t t, double stats[8]);
/* CUDA-event time (ms) of the contraction kernel launches of the last run (0 if timing disabled) */
int gm_table_set_timing(gm_table_t t, int enable);
int gm_table_last_kernel_ms(gm_table_t t, double* coeff_ms, double* contract_ms, double* finalize_ms);
/* per-kernel split: ms[0..4] / n[0..4] = time and event-bracketed launch groups of k_coeff, k_contract, k_finalize, k_gram,
 * k_gram_sum + k_gram_eval (contract_ms above = k_contract + k_gram + k_gram_sum + k_gram_eval); entries 5..7 reserved */
int gm_table_last_kernel_ms_ex(gm_table_t t, double ms[8], int32_t n[8]);

/* ---- B4: generalized-spherical-function expansion ---------------------------------------------------------------------
 * Replaces one run of ./spher_expan.x per cell (src/gsf/spher_expan.f main :1-110, one_calc :120-180,
 * GAUSS :520-579, LINTERPOL :593-623, SPHER_EXPAN :269-358, GENER :363-407) as spawned by convertncdf.convertData
 * (src/gsf/convertncdf.py:173-189).
 *   F [ncell][6][nang]   order F11,F22,F33,F44,F12,F34 (convertncdf.py:177) at angles ang_deg[nang] (ascending)
 *   coef [ncell][6][ng]  AL1,AL2,AL3,AL4,BET1,BET2 divided by AL1(0) (spher_expan.f:95-103); cnorm[ncell] = 1/AL1(0)
 *   quantize10           round to 10 decimals like the Fortran '(X,I5,6F17.10)' output (:96,:104)
 */
int gm_gsf_expand(gm_handle_t h, int ncell, int nang, const double* ang_deg, const double* F, int ng, double* coef,
                  double* cnorm, int quantize10);
int gm_gsf_expand_dev(gm_handle_t h, int ncell, int nang, const double* ang_deg, const double* F, int ng, double* coef,
                      double* cnorm, int quantize10);

/* same expansion fed directly with gm_table_run's out_phase layout: P4 [ncell][4][nang] = P11,P12,P33,P34 (device
 * pointers; F22 = P11, F44 = P33 for spheres, calculateScatVals dointegration.py:1044-1050).  The expansion is
 * normalised by AL1(0), so the a-posteriori P11 normalisation of fun (:978-985, a per-cell constant) cancels. */
int gm_gsf_expand_phase4_dev(gm_handle_t h, int ncell, int nang, const double* ang_deg, const double* P4, int ng, double* coef,
                             double* cnorm, int quantize10);

/* The diagnostic half of the same program: the matrix re-synthesised from the UN-normalised coefficients at the input angles
 * (MATR, spher_expan.f:419-517 -- what main writes to <file>.expan_matr, :84-90) and the fit error one_calc returns and main
 * prints (:168-177): fiterr = max |F11 - F11OUT| (ERREVAL :636-688 with ERRTYP = MAXABS over [0, 180] deg, params.h:8-10) over
 * the input grid and READMATRIX's alternative (mid-point) grid (:237-258, USE_ALT_ANG = 1).  The moments are returned as by
 * gm_gsf_expand.  fout [ncell][6][nang] (order F11,F22,F33,F44,F12,F34; nullable), fiterr [ncell] (nullable; not both NULL).
 * The _dev variant also accepts gm_table_run's 4-row phase layout (nrow = 4). */
int gm_gsf_diagnose(gm_handle_t h, int ncell, int nang, const double* ang_deg, const double* F, int ng, double* coef,
                    double* cnorm, int quantize10, double* fout, double* fiterr);
int gm_gsf_diagnose_dev(gm_handle_t h, int ncell, int nang, const double* ang_deg, const double* F, int nrow, int ng,
                        double* coef, double* cnorm, int quantize10, double* fout, double* fiterr);

/* ---- B5: band averaging ----------------------------------------------------------------------------------------------
 * Replaces bandaverage.doAverage (src/geosmie/bandaverage.py:18-50) over all (variable, bin, rh) columns.
 *   v [ncol][nlam] at wavelengths lam[nlam] (metres, ascending); bands lo/hi [nband] in cm^-1 (use_wavenum=1) or metres
 *   out [ncol][nband]
 */
int gm_band_average(gm_handle_t h, int ncol, int nlam, const double* lam, const double* v, int nband, const double* lo,
                    const double* hi, int use_wavenum, double* out);

/* ---- multi-GPU exchange over NVLink peer memory -------------------------------------------------------------
 * The reference has no multi-process table build (its only parallelism is one process per species,
 * proc.v2.1.0.csh:30-33; dointegration.py:804-810 notes that the (wavelength, RH) cells are independent).  Here cells are
 * sharded over one process per GPU and the finished rows of every rank land in ONE device buffer on rank 0, which every
 * rank maps through CUDA IPC:
 *   rank 0:  gm_peer_alloc  -> device pointer + a 64-byte IPC handle (sent to the other ranks by the caller, e.g. through
 *            torch.distributed broadcast);      other ranks: gm_peer_open(handle) -> the same memory in their address space.
 *   gm_peer_put   copy-engine transfer (no SMs on either side) src -> dst, asynchronous on the handle's exchange stream and
 *                 ordered after everything enqueued so far on the handle's compute stream; src/dst: device, peer or pinned
 *                 host memory (cudaMemcpyDefault), so it also serves rank 0's read-out.
 *   gm_peer_join  the compute stream waits for all puts issued so far (e.g. before an event that closes a timed region);
 *   gm_peer_sync  the host waits for them.  After gm_peer_sync on every rank + a barrier the data is complete on rank 0.
 *   gm_table_set_mirror  fused variant: k_finalize of every following gm_table_run* call ALSO stores its results through
 *                 the given (peer) pointers, [ntask][nmode][GM_NSCAL] and [ntask][4][nang], so the transfer is part of the
 *                 producing kernel (P2P stores over NVLink) and needs no extra pass; gm_gsf_expand_phase4_dev accepts a
 *                 peer pointer for `coef` likewise.  NULL, NULL switches it off.
 *   gm_peer_mark / gm_peer_wait   mark(i) remembers "all puts issued so far"; wait(i) makes the compute stream wait for that
 *                 point (i = 0..3): the fence of a double-buffered producer before it overwrites the source of an older put.
 */
#define GM_IPC_HANDLE_BYTES 64
int gm_peer_alloc(gm_handle_t h, size_t bytes, void** dptr, unsigned char ipc_handle[GM_IPC_HANDLE_BYTES]);
int gm_peer_free(gm_handle_t h, void* dptr);
int gm_peer_open(gm_handle_t h, const unsigned char ipc_handle[GM_IPC_HANDLE_BYTES], void** dptr);
int gm_peer_close(gm_handle_t h, void* dptr);
int gm_peer_put(gm_handle_t h, void* dst, const void* src, size_t bytes);
/* strided read-out on the exchange stream: `height` rows of `width` bytes, device / peer memory -> host memory with row pitch
 * `dpitch` (rank 0 scatters the segment of rank r into rows r, r + W, ... of the final table arrays in ONE pass) */
int gm_peer_get2d(gm_handle_t h, void* dst_host, size_t dpitch, const void* src_dev, size_t spitch, size_t width, size_t height);
int gm_peer_join(gm_handle_t h);
int gm_peer_sync(gm_handle_t h);
int gm_peer_mark(gm_handle_t h, int idx);
int gm_peer_wait(gm_handle_t h, int idx);
int gm_table_set_mirror(gm_table_t t, double* scal_mirror, double* phase_mirror);

#ifdef __cplusplus
}
#endif
#endif /* GEOSMIE_B200_H */
