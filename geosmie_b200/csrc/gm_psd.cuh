// gm_psd.cuh -- size-distribution number weights on the device (replaces the per-cell numpy work of
// dointegration.calculatePSD, src/geosmie/dointegration.py:539-664, and particleparams.getLogNormPSD, :113-127).
// The host still derives the handful of per-cell scalars (humidified mode radius, cut-offs, growth ratio) with the
// reference's formulas; the O(nx) weight vectors never cross PCIe.
#pragma once
#include "gm_common.cuh"

// deterministic block sum (fixed shuffle tree, then warp partials in warp order)
__device__ __forceinline__ double block_sum_256(double v, double* sh) {
  v = warp_sum(v);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  __syncthreads();
  if (lane == 0) sh[warp] = v;
  __syncthreads();
  double t = 0.0;
  for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += sh[w];
  return t;
}

// grid = ntask, block = 256.  params [ntask][nmode][4], frac [ntask][nmode].
// wscal [ntask][nmode][nx] (may alias wphase when nmode == 1 and frac == 1), wphase [ntask][nx].
__global__ void __launch_bounds__(256) k_psd(int nx, int nmode, int kind, const double* __restrict__ x, const double* __restrict__ dr,
                                             const double* __restrict__ params, const double* __restrict__ frac,
                                             double* __restrict__ wscal, double* __restrict__ wphase, int separate) {
  __shared__ double sh[8];
  const int task = blockIdx.x;
  for (int k = 0; k < nmode; ++k) {
    const double* P = params + ((size_t)task * nmode + k) * GM_PSD_NPAR;
    double* w = separate ? wscal + ((size_t)task * nmode + k) * nx : wphase + (size_t)task * nx;
    double local = 0.0;
    for (int i = threadIdx.x; i < nx; i += blockDim.x) {
      const double xi = x[i];
      double v = 0.0;
      if (kind == GM_PSD_LOGNORM) {
        // dNdx = 1/(x sqrt(2pi) ln s) exp(-ln(x/xmode)^2 / (2 ln(s)^2)); 0 outside (xmin, xmax); dN = dNdx * dr
        const double xmode = P[0], xmin = P[1], xmax = P[2], lns = P[3];
        if (!(xi >= xmax) && !(xi <= xmin)) {
          const double l = log(xi / xmode);
          v = 1.0 / (xi * 2.5066282746310002 * lns) * exp(-(l * l) / (2.0 * lns * lns)) * dr[i];
        }
      } else {
        const double xconv = P[0], rlo = P[1], rhi = P[2];
        const double ri = xi / xconv;
        // getDR (dointegration.py:93-101) on the radius grid (ss) or the size-parameter grid (du)
        const double s = (kind == GM_PSD_SS) ? xconv : 1.0;
        double d;
        if (i == 0) d = x[1] / s - x[0] / s;
        else if (i == nx - 1) d = x[nx - 1] / s - x[nx - 2] / s;
        else d = ((xi / s - x[i - 1] / s) + (x[i + 1] / s - xi / s)) / 2.0;
        if (!(ri < rlo) && !(ri > rhi)) {
          if (kind == GM_PSD_SS) {
            // Gong (2003) sea-salt dN/dr80 (dointegration.py:603-608)
            const double r80rat = 1.65 * P[3];
            const double r80 = ri * r80rat * 1e6;
            const double aFac = 4.7 * pow(1.0 + 30.0 * r80, -0.017 * pow(r80, -1.44));
            const double bFac = (0.433 - log10(r80)) / 0.433;
            v = 1.373 * pow(r80, -aFac) * (1.0 + 0.057 * pow(r80, 3.45)) * pow(10.0, 1.607 * exp(-bFac * bFac)) * r80rat * d;
          } else {
            v = 1.0 / (ri * ri * ri * ri) * d;   // dndr = r^-4 (dointegration.py:640)
          }
        }
      }
      w[i] = v;
      local += v;
    }
    const double tot = block_sum_256(local, sh);
    const double f = frac[(size_t)task * nmode + k];
    for (int i = threadIdx.x; i < nx; i += blockDim.x) {
      const double wn = w[i] / tot;            // psd /= np.sum(psd)
      w[i] = wn;
      if (separate) {
        double* wp = wphase + (size_t)task * nx + i;
        *wp = (k == 0) ? f * wn : *wp + f * wn;
      }
    }
    __syncthreads();
  }
}
