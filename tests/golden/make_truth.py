#!/usr/bin/env python3
"""High-precision (mpmath, 60 digits) coated-sphere efficiencies and S1/S2 for the coated cases of mie_single.npz.

The reference evaluates coated_mie_coeff (mie_coeffs.py:183-251) with scipy's complex-argument jv/yv, whose error
reaches ~4e-7 in a_n for |z| ~ 100 (two of the 30 golden cases are off by 3e-9 in Qext).  This fixture lets the tests
show that a deviation from the reference at that level is the reference's own Bessel error.
    python tests/golden/make_truth.py      (about ten minutes)
"""
import os
import sys

import mpmath as mp
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import mie_oracle as mo  # noqa: E402

mp.mp.dps = 60


def coated_mp(eps1, eps2, x, y):
    m1, m2 = mp.sqrt(eps1), mp.sqrt(eps2)
    m = m2 / m1
    u, v, w = m1 * x, m2 * x, m2 * y
    nmax = int(round(2 + float(y) + 4 * float(y) ** (1 / 3.)))
    psi = lambda n, z: mp.sqrt(mp.pi * z / 2) * mp.besselj(n + 0.5, z)
    chi = lambda n, z: -mp.sqrt(mp.pi * z / 2) * mp.bessely(n + 0.5, z)
    D = lambda n, z: psi(n - 1, z) / psi(n, z) - n / z
    an, bn = [], []
    for n in range(1, nmax + 1):
        dnu, dnv, dnw = D(n, u), D(n, v), D(n, w)
        pv, pw, py = psi(n, v), psi(n, w), psi(n, y)
        chv, chw, chy = chi(n, v), chi(n, w), chi(n, y)
        p1y, ch1y = psi(n - 1, y), chi(n - 1, y)
        gsy, gs1y = py - 1j * chy, p1y - 1j * ch1y
        uu, vv, fv = m * dnu - dnv, dnu / m - dnv, pv / chv
        pt, prat = pw - chw * fv, pw / pv / chv
        dns = (uu * fv / pw) / (uu * pt + prat) + dnw
        gns = (vv * fv / pw) / (vv * pt + prat) + dnw
        a1, b1 = dns / m2 + n / y, m2 * gns + n / y
        an.append(complex((py * a1 - p1y) / (gsy * a1 - gs1y)))
        bn.append(complex((py * b1 - p1y) / (gsy * b1 - gs1y)))
    return np.array(an), np.array(bn), nmax


if __name__ == "__main__":
    d = np.load(os.path.join(HERE, "mie_single.npz"))
    par = d["par"]
    idx, q, s12 = [], [], []
    us = d["us"]
    for i in range(par.shape[0]):
        x, y = par[i, 0], par[i, 1]
        if np.isnan(y):
            continue
        an, bn, nmax = coated_mp(mp.mpc(complex(par[i, 2], par[i, 3])), mp.mpc(complex(par[i, 6], par[i, 7])), mp.mpf(x), mp.mpf(y))
        idx.append(i)
        q.append(mo.mie_props(an, bn, nmax, y))
        # S1, S2 at the fixture's angles from the 60-digit coefficients (mie_S12_backend, mie_props.py:119-131)
        s12.append([[v.real, v.imag, w.real, w.imag] for v, w in (mo.mie_s12(an, bn, nmax, float(u)) for u in us)])
        print(i, x, y, q[-1][:2], flush=True)
    np.savez_compressed(os.path.join(HERE, "coated_truth.npz"), idx=np.array(idx), q=np.array(q), s12=np.array(s12))
