#!/usr/bin/env python3
"""Generalized-spherical-function basis from the REFERENCE's own re-synthesis code (src/utils/eval_gsfun.py), as a fixture.

The reference ships no test or golden for src/gsf/spher_expan.f, but it does ship the INVERSE of that program: eval_gsfun.py
rebuilds P11 ... P44 from the `pmom` moments with its Wigner-d recursion `d_mn` (:43-72) and scipy's Legendre polynomials
(OPTICS.getPmatrix / calcP11 / calcP12 / calcP22_P33 / calcP34, :126-296).  This script imports that module unmodified
(xarray and matplotlib, which only its file reader and plots need, are stubbed) and records, on the 371 table angles,

    leg[s]    = eval_legendre(s, cos theta)                                        (calcP11: P11 and P44)
    d02[s]    = d_mn(0, 2, s, theta)  / (1j ** 2).real                             (calcP12, calcP34)
    d22[s]    = d_mn(2, 2, s, theta)  / (1j ** 0).real                             (calcP22_P33: a2 + a3)
    d2m2[s]   = d_mn(2, -2, s, theta) / (1j ** -4).real                            (calcP22_P33: a2 - a3)

for s = 0..128, walking the recursion exactly as calcP12 does (and the first 13 orders on a uniform 0.2-degree grid).  tests/test_gsf_pin.py re-synthesises the phase matrix of real
table cells from OUR moments with this basis and compares it with the phase matrix the moments were computed from.

    python tests/golden/make_gsf_basis.py        # writes tests/golden/gsf_basis.npz
"""
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("GEOSMIE_REFERENCE", "/root/reference")
NMOM = 129


def reference_eval_gsfun():
    for name in ("xarray", "matplotlib", "matplotlib.pyplot"):
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    import importlib.util
    spec = importlib.util.spec_from_file_location("ref_eval_gsfun", os.path.join(REF, "src", "utils", "eval_gsfun.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def basis(ang_deg, nmom=NMOM):
    from scipy.special import eval_legendre
    G = reference_eval_gsfun()
    theta = np.radians(np.asarray(ang_deg, dtype=float))
    out = {"leg": np.stack([eval_legendre(s, np.cos(theta)) for s in range(nmom)])}
    for key, (m, n) in (("d02", (0, 2)), ("d22", (2, 2)), ("d2m2", (2, -2))):
        norm = (1j ** (n - m)).real
        b = np.zeros((nmom, theta.size))
        for i, t in enumerate(theta):          # the walk of OPTICS.calcP12 (eval_gsfun.py:178-196)
            d0, dneg1 = G.d_mn(m, n, 0, t)
            b[0, i] = d0 / norm
            d1, d0 = G.d_mn(m, n, 1, t, dm1=d0, dm2=dneg1)
            b[1, i] = d1 / norm
            dm2, dm1 = d0, d1
            for s in range(2, nmom):
                dfunc, dm2 = G.d_mn(m, n, s, t, dm1=dm1, dm2=dm2)
                b[s, i] = dfunc / norm
                dm1 = dfunc
        out[key] = b
    return out


if __name__ == "__main__":
    ang = np.concatenate([np.linspace(0., 1., 100, endpoint=False), np.linspace(1., 10., 100, endpoint=False),
                          np.linspace(10., 180., 171, endpoint=True)])
    # a second, uniform 0.2-degree grid with the first 13 orders only: known-answer expansions whose linear-interpolation error
    # (spher_expan.f LINTERPOL) is ~1e-5 instead of the ~1e-4 of the 1-degree table grid
    fine = np.linspace(0., 180., 901)
    fb = basis(fine, 13)
    np.savez_compressed(os.path.join(HERE, "gsf_basis.npz"), ang=ang, fine_ang=fine, **basis(ang), **{"fine_" + k: v for k, v in fb.items()})
    print("wrote gsf_basis.npz")
