"""Python face of the GSF oracle (TEST INFRASTRUCTURE; C restatement of src/gsf/spher_expan.f in gsf_oracle.c).

"PARITY UNPINNED" by the reference: no Fortran compiler in this image, no reference test/golden for this program.
Pinned by the analytic Rayleigh coefficients and the MATR round trip (tests/test_oracle.py)."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "_build", "liboracle_gsf.so")
        if not os.path.exists(path):
            subprocess.check_call(["make", "-s", "-C", _HERE])
        L = C.CDLL(path)
        L.orc_gsf_expand.restype = C.c_double
        L.orc_gsf_one_calc.restype = C.c_double
        _LIB = L
    return _LIB


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def gauss(n):
    z = np.zeros(n)
    w = np.zeros(n)
    lib().orc_gauss(C.c_int(n), C.c_int(0), _p(z), _p(w))
    return z, w


def expand(ang_deg, F, ng=129, quantize10=False):
    """F [6][nang] in the order F11,F22,F33,F44,F12,F34 -> (coef [6][ng] = AL1..BET2, CNORM)."""
    ang = np.ascontiguousarray(ang_deg, dtype=float)
    F = np.ascontiguousarray(F, dtype=float)
    assert F.shape == (6, ang.size)
    coef = np.zeros((6, ng))
    cn = lib().orc_gsf_expand(C.c_int(ang.size), _p(ang), _p(F), C.c_int(ng), _p(coef), C.c_int(int(quantize10)))
    return coef, cn


def matr(coef, ang_deg):
    coef = np.ascontiguousarray(coef, dtype=float)
    ang = np.ascontiguousarray(np.radians(ang_deg), dtype=float)
    out = np.zeros((6, ang.size))
    lib().orc_gsf_matr(C.c_int(coef.shape[1]), _p(coef), C.c_int(ang.size), _p(ang), _p(out))
    return out


def one_calc(ang_deg, F, ng=129):
    """The diagnostic half of spher_expan.f (alternative angle grid, MATR, ERREVAL): F [6][nang] ->
    (fout [6][nang] = the .expan_matr columns, fiterr)."""
    ang = np.ascontiguousarray(ang_deg, dtype=float)
    F = np.ascontiguousarray(F, dtype=float)
    assert F.shape == (6, ang.size)
    out = np.zeros((6, ang.size))
    err = lib().orc_gsf_one_calc(C.c_int(ang.size), _p(ang), _p(F), C.c_int(ng), _p(out))
    return out, float(err)
