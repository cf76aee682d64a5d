"""Object API of pymiecoated on sm_100a CUDA: Mie, MultipleMie, MieScatterProps
(src/pymiecoated/pymiecoated/mie_coated.py).  Same keyword arguments, attributes, return types and ValueErrors as the
reference; every number is produced by libgeosmie_b200 (no CPU fallback)."""
import numpy as np
from numpy import sqrt

from .. import _lib
from .mie_aux import Cache
from .mie_coeffs import _resolve, nmax_of


class _Coeffs(object):
    """Device-evaluated stand-in of MieCoeffs: keeps the resolved particle and evaluates properties on demand."""

    def __init__(self, par):
        (self.kind, self.size, self.xcore, self.mz, self.mrel, self.ajv, self.ayv) = _resolve(par)
        self.nmax = int(nmax_of(self.size))
        self._ab = None

    def _eval(self, u=None, want_ab=False):
        h = _lib.Handle.get()
        kw = {}
        if len(self.ajv) and len(self.ayv):
            kw = dict(ajv=np.asarray(self.ajv)[:self.nmax], ayv=np.asarray(self.ayv)[:self.nmax])
        return h.mie_eval([self.size], [self.mz], [self.mrel], [self.nmax], u=u,
                          xcore=None if self.xcore is None else [self.xcore], want_s12=u is not None, want_ab=want_ab, **kw)

    def _coeffs(self):
        if self._ab is None:
            self._ab = self._eval(want_ab=True)[2]
        return self._ab

    @property
    def an(self):
        ab = self._coeffs()
        return ab[:, 0] + 1j * ab[:, 1]

    @property
    def bn(self):
        ab = self._coeffs()
        return ab[:, 2] + 1j * ab[:, 3]


class MultipleMie(object):
    """Size-range batch (mie_coated.py:41-179).  preCalculate() builds the per-bin device table (Riccati-Bessel and
    pi/tau tables, replacing the scipy/numba pre-computation), calculateS12SizeRange(mr, mi) returns the reference's
    dict of lists."""

    def __init__(self, xArr, yArr, costarr):
        self.xArr = xArr
        self.yArr = yArr
        self.costarr = costarr
        self.parr = {}
        self.tarr = {}
        self.jvdic = {}
        self.yvdic = {}
        self._table = None

    def preCalculate(self):
        self.preCalculateBessel()
        self.preCalculatePT()

    def preCalculateBessel(self):
        self._ensure_table()

    def preCalculatePT(self):
        self._ensure_table()

    MAX_TABLE_ANGLES = 384      # angle capacity of the DMMA table path (2 halves x 192)

    def _direct(self, mr, mi):
        """More angles than the table path holds: one warp per particle, lanes over angles (k_s12_direct)."""
        eps = complex(mr, mi) ** 2
        x = np.asarray(self.xArr, dtype=float)
        q, s12, _ = _lib.Handle.get().mie_eval(x, [np.sqrt(eps * 1.0)], [np.sqrt(eps / 1.0)], nmax_of(x),
                                               u=np.asarray(self.costarr, dtype=float))
        return q, s12

    def _ensure_table(self):
        if self.yArr is not None or len(self.costarr) > self.MAX_TABLE_ANGLES:
            return None
        if self._table is None:
            x = np.asarray(self.xArr, dtype=float)
            self._table = _lib.Table(x, nmax_of(x), np.asarray(self.costarr, dtype=float))
        return self._table

    def device_table(self):
        return self._ensure_table()

    def calculateS12SizeRange(self, mr, mi):
        eps = complex(mr, mi) ** 2          # mie_coated.py:62
        mu = 1.0
        prokeys = ['qext', 'qsca', 'qabs', 'asy', 'qb', 'qratio']
        x = np.asarray(self.xArr, dtype=float)
        cost = np.asarray(self.costarr, dtype=float)
        if self.yArr is None:
            q, s12 = self.calculateS12SizeRangeArrays(mr, mi)
        else:
            # the reference calls an undefined coated_mie_coeff_numba here (mie_coated.py:80); this build evaluates the
            # coated sphere (core xArr, shell yArr; core index (mr, mi) is not enough for two materials, so the batch
            # coated path takes the shell index from self.m2)
            y = np.asarray(self.yArr, dtype=float)
            m2 = complex(getattr(self, "m2", complex(mr, mi)))
            q, s12, _ = _lib.Handle.get().mie_eval(y, [np.sqrt(eps)], [m2], nmax_of(y), u=cost, xcore=x)
        ret = {'s12': [[(complex(a[0], a[1]), complex(a[2], a[3])) for a in row] for row in s12]}
        cols = {'qext': 0, 'qsca': 1, 'qabs': 2, 'qb': 3, 'asy': 4, 'qratio': 5}
        for k in prokeys:
            ret[k] = [float(v) for v in q[:, cols[k]]]
        self._last_arrays = (q, s12)
        return ret

    def calculateS12SizeRangeArrays(self, mr, mi):
        """Extension: same evaluation, numpy arrays (q [nx][6], s12 [nx][nang][4]) instead of lists of tuples."""
        eps = complex(mr, mi) ** 2
        t = self._ensure_table()
        if t is None:
            return self._direct(mr, mi)
        q, s12 = t.particles([np.sqrt(eps * 1.0)], [np.sqrt(eps / 1.0)], want_s12=True)     # DMMA table path
        return q[0], s12[0]


class MieScatterProps(object):
    """Stores the mie coefficients and the corresponding parameters (mie_coated.py:181-211)."""

    def __init__(self, params, ajv, ayv):
        par = dict(zip(("eps", "mu", "x", "y", "eps2"), params[:5]))
        par["ajv"] = ajv
        par["ayv"] = ayv
        self._coeffs = _Coeffs(par)
        self._props = None
        self._S12 = None
        self.size = par["x"] if par["y"] is None else par["y"]
        self.ajv = ajv
        self.ayv = ayv

    def prop(self, prop_name):
        if self._props is None:
            q = self._coeffs._eval()[0][0]
            self._props = {"qext": float(q[0]), "qsca": float(q[1]), "qabs": float(q[2]), "qb": float(q[3]),
                           "asy": float(q[4]), "qratio": float(q[5])}
        return self._props[prop_name]

    def S12(self, u):
        s = self._coeffs._eval(u=[float(u)])[1][0, 0]
        self._S12 = (complex(s[0], s[1]), complex(s[2], s[3]))
        return self._S12

    def S12_array(self, u):
        """Extension: S1, S2 at many cosines in one GPU call -> complex arrays."""
        s = self._coeffs._eval(u=np.asarray(u, dtype=float))[1][0]
        return s[:, 0] + 1j * s[:, 1], s[:, 2] + 1j * s[:, 3]

    def S12_pt(self, pin, tin):
        from .mie_props import mie_S12_pt
        self._S12_pt = mie_S12_pt(self._coeffs, pin, tin)
        return self._S12_pt


class Mie(object):
    """Mie scattering of a homogeneous or coated sphere (mie_coated.py:214-396): the reference's object API on top of the GPU library.

    Keyword arguments / attributes as in the reference: x, y (core and shell size parameters), eps, mu, eps2 (permittivity, permeability,
    shell permittivity) or m, m2 (refractive indices; `mc` is an alias of `m2`), and the validation inputs ajv, ayv.  Results are cached
    per parameter set in a mie_aux.Cache, like the reference's."""

    _PLAIN = ("eps", "mu", "eps2")              # stored as given
    _CHECKED = ("m", "m2", "mc", "x", "y")      # go through the property setters, in this order (y is validated against x)

    def __init__(self, **kwargs):
        self._cache = Cache()
        self.eps, self.mu, self.eps2 = None, 1.0, None
        self._x = self._y = None
        self.ajv = self.ayv = ()
        for k in self._PLAIN:
            if k in kwargs:
                self.__dict__[k] = kwargs[k]
        for k in self._CHECKED:
            if k in kwargs:
                setattr(self, k, kwargs[k])
        for k in ("ajv", "ayv"):
            if k in kwargs:
                setattr(self, k, tuple(kwargs[k]))

    # ---- parameters -------------------------------------------------------------------------------------------------
    def _set_m(self, m):
        self.mu, self.eps = 1.0, m ** 2

    def _set_m2(self, m2):
        self.eps2 = m2 ** 2

    def _set_x(self, x):
        if not x >= 0.0:
            raise ValueError("The size x cannot be smaller than 0.")
        self._x = x

    def _set_y(self, y):
        if not y >= self.x:
            raise ValueError("The size y cannot be smaller than x.")
        self._y = y

    m = property(lambda self: sqrt(self.eps / self.mu), _set_m, doc="refractive index sqrt(eps / mu); setting it sets mu = 1")
    m2 = property(lambda self: sqrt(self.eps2), _set_m2, doc="refractive index of the shell")
    mc = m2
    x = property(lambda self: self._x, _set_x, doc="size parameter of the sphere (of the core when y is given)")
    y = property(lambda self: self._y, _set_y, doc="size parameter of the shell")

    # ---- results ----------------------------------------------------------------------------------------------------
    def _params_signature(self):
        return (self.eps, self.mu, self.x, self.y, self.eps2)

    def _entry(self):
        sig = self._params_signature()
        if sig not in self._cache:
            self._cache[sig] = MieScatterProps(sig, self.ajv, self.ayv)
        return self._cache[sig]

    def _get_scatt_prop(self, prop):
        return self._entry().prop(prop)

    def S12(self, u):
        """Amplitude scattering matrix elements S1, S2 (Bohren and Huffman conventions) at the cosine u of the scattering angle."""
        if abs(u) > 1:
            raise ValueError("The cosine u must be between -1 and 1.")
        return self._entry().S12(u)

    def S12_array(self, u):
        """Extension: S1, S2 at an array of cosines in one GPU call."""
        u = np.asarray(u, dtype=float)
        if np.any(np.abs(u) > 1):
            raise ValueError("The cosine u must be between -1 and 1.")
        return self._entry().S12_array(u)

    def S12_pt(self, pin, tin):
        """S1, S2 from caller-supplied angle functions pi_n, tau_n (mie_props.py:133-150)."""
        return self._entry().S12_pt(pin, tin)


def _efficiency(name, what):
    def f(self):
        return self._get_scatt_prop(name)
    f.__name__, f.__doc__ = name, what
    return f


for _name, _what in (("qext", "extinction efficiency"), ("qsca", "scattering efficiency"), ("qabs", "absorption efficiency"),
                     ("qb", "backscattering efficiency"), ("asy", "asymmetry parameter <cos(theta)>"),
                     ("qratio", "backscattering ratio qb / qsca")):
    setattr(Mie, _name, _efficiency(_name, _what))
del _name, _what
