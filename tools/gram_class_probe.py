#!/usr/bin/env python3
"""Per-class efficiency of k_gram: synthetic size grids whose particle groups all fall into ONE Gram class (max nmax of a
group in (8 (c-1), 8 c], class 0: <= 4), evaluated dense for `ntask` tasks; prints k_gram's time, the DMMAs it issues and
the fraction of the measured DMMA peak (37.1 TFLOP/s = 0.2455 DMMA/clk/SM at 1.965 GHz).

    python tools/gram_class_probe.py [ntask [class ...]]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np

from geosmie_b200 import _lib, dointegration as DI
from geosmie_b200.pymiecoated.mie_coeffs import nmax_of

PEAK = 37.1e12 / 512.0     # DMMA m8n8k4 per second (tools/fp64_peak.cu)


def x_range_for_nmax(lo, hi):
    """Largest x interval whose nmax = round(2 + x + 4 x^(1/3)) lies in [lo, hi]."""
    x = np.geomspace(1e-4, 100.0, 200001)
    nm = nmax_of(x)
    ok = (nm >= lo) & (nm <= hi)
    return x[ok][0], x[ok][-1]


def main():
    ntask = int(sys.argv[1]) if len(sys.argv) > 1 else 549
    only = [int(a) for a in sys.argv[2:]] or list(range(9))
    h = _lib.Handle.get(0)
    cost = np.cos(np.radians(DI.table_angles()))
    rng = np.random.default_rng(0)
    m = (1.3 + 0.3 * rng.random(ntask)) + 1j * 10 ** rng.uniform(-6, -1, ntask)
    nx = 4480
    for c in only:
        lo, hi = (2, 4) if c == 0 else (max(5, 8 * c - 7), 8 * c)
        x0, x1 = x_range_for_nmax(lo, hi)
        x = np.geomspace(x0, x1, nx)
        nm = nmax_of(x)
        w = np.full((ntask, nx), 1.0 / nx)
        t = _lib.Table(x, nm, cost, h)
        t.set_timing(True)
        for _ in range(3):
            t.run(m, m, w, None, elide=False)
        k = t.last_kernel_ms()
        gm = nm.reshape(-1, 32).max(axis=1)
        tg = np.ceil(gm / 8.0)
        dm = np.where(gm <= 4, 32.0, np.where(tg <= 4, 64.0 * tg * tg, 16.0 * 12.0 * tg * np.floor((tg + 2) / 3)))
        need = 16.0 * float((nm.astype(float) ** 2).sum()) / 512.0
        ndm = float(dm.sum()) * ntask
        ms = k["k_gram"]
        print("class %d  nmax %2d..%2d  groups %d  k_gram %.3f ms  DMMA %.3g  executed %.2f of peak  necessary %.2f   (k_coeff %.3f ms, sum+eval %.3f ms)"
              % (c, nm.min(), nm.max(), len(gm), ms, ndm, ndm / (ms * 1e-3) / PEAK, need * ntask / (ms * 1e-3) / PEAK, k["k_coeff"],
                 k["k_gram_sum_eval"]), flush=True)
        t.close()


if __name__ == "__main__":
    main()
