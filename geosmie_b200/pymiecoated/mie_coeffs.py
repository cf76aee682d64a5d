"""Mie coefficients a_n, b_n on the GPU behind the reference's interface (src/pymiecoated/pymiecoated/mie_coeffs.py).

Validation and dispatch (which formula for which arguments, which ValueError for which mistake) follow
mie_coeffs(params) :36-73 on the host; the arithmetic runs in libgeosmie_b200 (k_coeff / k_coated_coeff).
"""
import numpy as np

from .. import _lib


def nmax_of(x):
    """nmax = int(round(2+x+4*x**(1/3))) (mie_coeffs.py:99, :148, :203; mie_coated.py:131) -- half-to-even rounding."""
    x = np.asarray(x, dtype=float)
    return np.round(2 + x + 4 * x ** (1.0 / 3.0)).astype(np.int32)


class MieCoeffs(object):
    """Wrapper for the Mie coefficients (mie_coeffs.py:30-34)."""

    def __init__(self, par):
        (self.an, self.bn, self.nmax) = mie_coeffs(par)


def _resolve(params):
    """Validation + dispatch of mie_coeffs.py:40-71.  Returns (kind, size, xcore, mz, mrel, ajv, ayv)."""
    eps = complex(params["eps"]) if params["eps"] is not None else None
    x = float(params["x"]) if params["x"] is not None else None
    ajv = np.array(params.get("ajv", ()))
    ayv = np.array(params.get("ayv", ()))
    if (x is None) or (eps is None):
        raise ValueError("Must specify x and either eps or m.")
    mu = complex(params["mu"]) if params["mu"] is not None else complex(1.0)
    y = float(params["y"]) if params["y"] is not None else None
    eps2 = complex(params["eps2"]) if params["eps2"] is not None else None
    coated = (y is not None)
    if coated == (eps2 is None):
        raise ValueError("Must specify both y and m2 for coated particles.")
    if coated and mu != complex(1.0):
        raise ValueError("Multilayer calculations for magnetic particles are not currently supported.")
    if not coated:
        y = x
        eps2 = eps
    # Do not use the coated version if it is not necessary (:65-71)
    if x == y or eps == eps2:
        return ("single", y, None, np.sqrt(eps * mu), np.sqrt(eps / mu), ajv, ayv)
    elif x == 0:
        return ("single", y, None, np.sqrt(eps2 * mu), np.sqrt(eps2 / mu), ajv, ayv)
    return ("coated", y, x, np.sqrt(eps), np.sqrt(eps2), (), ())


def mie_coeffs(params):
    """Input validation and function selection for the Mie coefficients (mie_coeffs.py:36-73).  Returns (an, bn, nmax)."""
    kind, size, xcore, mz, mrel, ajv, ayv = _resolve(params)
    nmax = int(nmax_of(size))
    h = _lib.Handle.get()
    kw = {}
    if len(ajv) and len(ayv):
        kw = dict(ajv=ajv[:nmax], ayv=ayv[:nmax])
    _, _, ab = h.mie_eval([size], [mz], [mrel], [nmax], xcore=None if xcore is None else [xcore], want_s12=False, want_ab=True, **kw)
    an = ab[:, 0] + 1j * ab[:, 1]
    bn = ab[:, 2] + 1j * ab[:, 3]
    return (an, bn, nmax)


def single_mie_coeff(eps, mu, x, ajv=(), ayv=()):
    """Mie coefficients for the single-layered sphere (mie_coeffs.py:132-180)."""
    return mie_coeffs({"eps": eps, "mu": mu, "x": x, "y": None, "eps2": None, "ajv": ajv, "ayv": ayv})


single_mie_coeff_numba = single_mie_coeff


def coated_mie_coeff(eps1, eps2, x, y):
    """Mie coefficients for the dual-layered (coated) sphere (mie_coeffs.py:183-251)."""
    nmax = int(nmax_of(y))
    _, _, ab = _lib.Handle.get().mie_eval([y], [np.sqrt(complex(eps1))], [np.sqrt(complex(eps2))], [nmax], xcore=[x],
                                          want_s12=False, want_ab=True)
    return (ab[:, 0] + 1j * ab[:, 1], ab[:, 2] + 1j * ab[:, 3], nmax)
