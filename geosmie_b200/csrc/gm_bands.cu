// gm_bands.cu -- band averaging of table columns (replaces bandaverage.doAverage, src/geosmie/bandaverage.py:18-50).
//   host: wavenumber grid (:31), 100-point np.linspace sub-grid per band clipped to the table range (:33-43) and the
//         interp1d bracket of every sub-point (scipy _call_linear: searchsorted, clip to [1, n-1]);
//   device: one warp per (column, band): linear interpolation at the 100 sub-points and their arithmetic mean (:46-49).
#include <math.h>

#include <algorithm>

#include "gm_common.cuh"

namespace {
constexpr int NSUB = 100;  // num_subbin, bandaverage.py:32

struct SubPoint {
  double xn;     // clipped sub-grid wavenumber
  double xlo, xhi;
  int ilo, ihi;  // indices into the ORIGINAL wavelength axis
};

__global__ void __launch_bounds__(128) k_band(int ncol, int nlam, int nband, const double* __restrict__ v,
                                              const SubPoint* __restrict__ sp, double* __restrict__ out) {
  const int item = blockIdx.x * 4 + (threadIdx.x >> 5);
  if (item >= ncol * nband) return;
  const int lane = threadIdx.x & 31;
  const int col = item / nband, b = item % nband;
  const double* y = v + (size_t)col * nlam;
  double s = 0.0;
  for (int k = lane; k < NSUB; k += 32) {
    const SubPoint p = sp[b * NSUB + k];
    const double ylo = y[p.ilo], yhi = y[p.ihi];
    const double slope = (yhi - ylo) / (p.xhi - p.xlo);
    s += slope * (p.xn - p.xlo) + ylo;
  }
  s = warp_sum(s);
  if (lane == 0) out[(size_t)col * nband + b] = s / (double)NSUB;
}
}  // namespace

extern "C" int gm_band_average(gm_handle_t h, int ncol, int nlam, const double* lam, const double* v, int nband, const double* lo,
                               const double* hi, int use_wavenum, double* out) {
  GM_REQUIRE(h && lam && v && lo && hi && out, "NULL argument");
  GM_REQUIRE(ncol > 0 && nlam >= 2 && nband > 0, "bad sizes");
  GM_CUDA_TRY(cudaSetDevice(h->device));
  cudaStream_t st = h->stream;
  // wnum = (100*lam)**-1 and its ascending sort permutation (interp1d sorts x when assume_sorted=False)
  std::vector<double> wn(nlam);
  std::vector<int> perm(nlam);
  for (int i = 0; i < nlam; ++i) {
    wn[i] = 1.0 / (100.0 * lam[i]);
    perm[i] = i;
  }
  std::stable_sort(perm.begin(), perm.end(), [&](int a, int b) { return wn[a] < wn[b]; });
  std::vector<double> xs(nlam);
  for (int i = 0; i < nlam; ++i) xs[i] = wn[perm[i]];
  const double wmin = xs.front(), wmax = xs.back();
  std::vector<SubPoint> sp((size_t)nband * NSUB);
  for (int b = 0; b < nband; ++b) {
    double bl = lo[b], br = hi[b];
    if (!use_wavenum) {  // bandaverage.py:26-29
      bl = 1.0 / (100.0 * bl);
      br = 1.0 / (100.0 * br);
    }
    const double beg = bl < br ? bl : br, end = bl < br ? br : bl;
    const double step = (end - beg) / (double)(NSUB - 1);
    for (int k = 0; k < NSUB; ++k) {
      double x = (k == NSUB - 1) ? end : beg + (double)k * step;  // np.linspace
      x = std::min(std::max(x, wmin), wmax);                      // np.clip
      int idx = (int)(std::lower_bound(xs.begin(), xs.end(), x) - xs.begin());  // searchsorted(side='left')
      idx = std::min(std::max(idx, 1), nlam - 1);
      SubPoint& p = sp[(size_t)b * NSUB + k];
      p.xn = x;
      p.xlo = xs[idx - 1];
      p.xhi = xs[idx];
      p.ilo = perm[idx - 1];
      p.ihi = perm[idx];
    }
  }
  int rc;
  const size_t nv = (size_t)ncol * nlam, no = (size_t)ncol * nband;
  if ((rc = h->ws[0].ensure(sizeof(SubPoint) * sp.size())) || (rc = h->ws[1].ensure(sizeof(double) * nv)) ||
      (rc = h->ws[2].ensure(sizeof(double) * no)))
    return rc;
  GM_CUDA_TRY(cudaMemcpyAsync(h->ws[0].p, sp.data(), sizeof(SubPoint) * sp.size(), cudaMemcpyHostToDevice, st));
  GM_CUDA_TRY(cudaMemcpyAsync(h->ws[1].p, v, sizeof(double) * nv, cudaMemcpyHostToDevice, st));
  const int items = ncol * nband;
  k_band<<<(items + 3) / 4, 128, 0, st>>>(ncol, nlam, nband, h->ws[1].as<double>(), h->ws[0].as<SubPoint>(), h->ws[2].as<double>());
  h->launches++;
  GM_CUDA_TRY(cudaGetLastError());
  GM_CUDA_TRY(cudaMemcpyAsync(out, h->ws[2].p, sizeof(double) * no, cudaMemcpyDeviceToHost, st));
  GM_CUDA_TRY(cudaStreamSynchronize(st));
  return GM_OK;
}
