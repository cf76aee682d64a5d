"""Scattering properties from Mie coefficients (src/pymiecoated/pymiecoated/mie_props.py) -- evaluated on the GPU.

The functions that take a coefficient object re-evaluate particle -> (a_n, b_n) -> property inside one C-ABI call
(gm_mie_eval); `coeffs` therefore carries the particle parameters next to an/bn/nmax.
"""
import numpy as np

def mie_props(coeffs, y):
    """The scattering properties (mie_props.py:72-75)."""
    q = coeffs._eval()[0][0]
    return {"qext": float(q[0]), "qsca": float(q[1]), "qabs": float(q[2]), "qb": float(q[3]), "asy": float(q[4]),
            "qratio": float(q[5])}

def mie_S12(coeffs, u):
    """The amplitude scattering matrix (mie_props.py:108-110, :119-131)."""
    s = coeffs._eval(u=[u])[1][0, 0]
    return (complex(s[0], s[1]), complex(s[2], s[3]))

def mie_S12_pt(coeffs, pin, tin):
    """S1,S2 from caller-supplied pre-multiplied pi_n/tau_n arrays (mie_props.py:112-113, mie_S12_backend_pt :133-150):
    S1 = sum a_n pin_n + sum b_n tin_n, S2 = sum a_n tin_n + sum b_n pin_n over the first nmax entries.  The coefficients
    come from the GPU (gm_mie_eval); the four nmax-long dot products with the CALLER's arrays are formed here in the
    reference's order, so arbitrary (truncated, rescaled) pin / tin behave exactly as in the reference."""
    pin = np.asarray(pin, dtype=float)
    tin = np.asarray(tin, dtype=float)
    nmax = coeffs.nmax
    if pin.shape[0] < nmax or tin.shape[0] < nmax:
        raise ValueError("pin / tin must hold at least nmax = %d entries" % nmax)
    an, bn = coeffs.an[:nmax], coeffs.bn[:nmax]
    s1 = np.dot(an, pin[:nmax]) + np.dot(bn, tin[:nmax])
    s2 = np.dot(an, tin[:nmax]) + np.dot(bn, pin[:nmax])
    return (complex(s1), complex(s2))

def mie_pt(u, nmax):
    """pi_n, tau_n pre-multiplied by (2n+1)/(n(n+1)) (mie_props.py:194-195, :217-231).  Host-side helper: these
    arrays are inputs of the S12_pt API, not part of the accelerated path (the GPU keeps its own table)."""
    u = float(u)
    p = np.zeros(max(nmax, 2))
    t = np.zeros(max(nmax, 2))
    p[0], p[1] = 1.0, 3 * u
    t[0], t[1] = u, 6 * u ** 2 - 3
    for ni in range(2, nmax):
        n = float(ni)
        p[ni] = (2 * n + 1) / n * p[ni - 1] * u - (n + 1) / n * p[ni - 2]
        t[ni] = (n + 1) * u * p[ni] - (n + 2) * p[ni - 1]
    k = np.arange(1, max(nmax, 2) + 1, dtype=float)
    n2 = (2 * k + 1) / (k * (k + 1))
    return (p * n2)[:nmax], (t * n2)[:nmax]

mie_ptnumba = mie_pt
