"""Species configuration loader and the small physical helper laws (mirror of src/geosmie/particleparams.py).

Everything here is cheap scalar/vector host bookkeeping that feeds the GPU path: JSON species files, refractive-index
table readers (GADS/OPAC, csv, wsv, HITRAN water), humidity growth and the lognormal number distribution.  Paths are
CWD-relative like the reference (`data/...`, `geosparticles/...`).
"""
import json
import sys

import numpy as np


def getPPJSON(partid):
    """Species JSON + refractive-index tables -> dict (particleparams.py:18-56).  Adds data['mList'] (one
    [wavelength[m], n, k-as-stored] array per entry of ri.path) and normalises the 'du' sub-bin shorthand."""
    with open("%s" % partid) as fp:
        data = json.load(fp)
    ri = data['ri']
    readers = {'gads': getM, 'csv': lambda p: getMSep(p, ','), 'wsv': lambda p: getMSep(p, None)}
    if ri['format'] not in readers:
        print('refractive index format %s not yet supported' % ri['format'])
        sys.exit()
    data['mList'] = [readers[ri['format']](path) for path in ri['path']]
    psd = data['psd']
    if psd['type'] == 'du':
        lo, hi = [], []
        for a, b in zip(psd['params']['rMinMaj'], psd['params']['rMaxMaj']):
            lo.append(a if isinstance(a, list) else [a])
            hi.append(b if isinstance(a, list) else [b])
        psd['params']['rMinMaj'], psd['params']['rMaxMaj'] = lo, hi
    return data


def getParticleParams(partID, datatype):
    """Top level call (particleparams.py:59-64); only JSON is supported."""
    if datatype == 'json':
        return getPPJSON(partID)
    print("bad datatype: %s" % datatype)
    sys.exit()


def getM(fn):
    """OPAC/GADS table: 17 header lines, 61 rows, columns 1, 8, 9 = wavelength[um], n, k (particleparams.py:67-70)."""
    data = np.loadtxt(fn, skiprows=17, max_rows=61, comments=None, unpack=True, usecols=[1, 8, 9])
    data[0] *= 1e-6
    return data


def getMSep(fn, sep):
    """'wavelength[um] n k' separated table (particleparams.py:73-76)."""
    data = np.loadtxt(fn, unpack=True, delimiter=sep)
    data[0] *= 1e-6
    return data


def getWaterM():
    """HITRAN water table data/refrac.water.txt (particleparams.py:79-82)."""
    data = np.loadtxt('data/refrac.water.txt', skiprows=13, unpack=True, usecols=[0, 1, 2])
    data[0] *= 1e-6
    return data


_growth_memo = None      # dict while a growth_memo() scope is open, else None (plain evaluation)


class growth_memo(object):
    """Scope in which humidityGrowth remembers its results.  The humidified size of a dry size does not depend on the
    wavelength, while the table build asks for it in every (wavelength, RH) cell of a bin (3 x 10,980 calls for optics_SS):
    BinPlan opens this scope around its cell loop.  The memo lives only inside the scope, so a caller that edits the
    parameter dict between two table builds never sees stale values."""

    def __enter__(self):
        global _growth_memo
        self._outer = _growth_memo
        _growth_memo = {} if _growth_memo is None else _growth_memo
        return self

    def __exit__(self, *exc):
        global _growth_memo
        _growth_memo = self._outer
        return False


def humidityGrowth(params, siz0, rh, allrh):
    """Humidified size of a dry size siz0 at relative humidity rh (particleparams.py:85-108); same arithmetic as the
    reference, evaluated once per (parameter dict, dry size, RH) inside a growth_memo() scope."""
    memo = _growth_memo
    if memo is None:
        return _humidityGrowth(params, siz0, rh, allrh)
    key = (id(params), siz0, rh)
    try:
        hit = memo.get(key)
    except TypeError:          # unhashable size (array): no memo
        return _humidityGrowth(params, siz0, rh, allrh)
    if hit is not None and hit[0] is params and hit[1] is allrh:
        return hit[2]
    val = _humidityGrowth(params, siz0, rh, allrh)
    memo[key] = (params, allrh, val)
    return val


def _humidityGrowth(params, siz0, rh, allrh):
    rhi = list(allrh).index(rh)
    rhtype = params['type']
    rhp = params['params']
    if rhtype in ('simple', 'trivial'):
        return siz0 * rhp['gf'][rhi]
    if rhtype == 'ss':      # Gerber [1985]
        if rh == 0.0:
            return siz0
        siz = siz0 * 100    # cm
        return (rhp['c1'] * siz ** rhp['c2'] / (rhp['c3'] * siz ** rhp['c4'] - np.log10(rh)) + siz ** 3.) ** (1. / 3.) / 100.
    if rhtype == 'su':
        if rh == 0.0:
            return siz0
        from . import carma_growth
        return siz0 * float(carma_growth.grow_v75(rh, siz0, temp=rhp['temp']))
    raise ValueError("unknown rhDep type %r" % rhtype)


def getLogNormPSD(rmode, sigma, xxArr, lambd, rmax, rmin):
    """dN/dx of a lognormal number distribution in size-parameter space, truncated to (xmin, xmax)
    (particleparams.py:113-127)."""
    xconv = 2 * np.pi / lambd
    xmode, xmax, xmin = rmode * xconv, rmax * xconv, rmin * xconv
    dNdx = 1. / (xxArr * (2 * np.pi) ** 0.5 * np.log(sigma)) * np.exp(-(np.log(xxArr / xmode) ** 2) / (2. * np.log(sigma) ** 2))
    dNdx[np.where(xxArr >= xmax)] = 0.
    dNdx[np.where(xxArr <= xmin)] = 0.
    return dNdx
