"""GPU table driver: mirror of src/geosmie/dointegration.py (Mie branch) on top of libgeosmie_b200.

The reference walks bin -> wavelength -> RH and, per cell, calls rawMie (per-particle Mie at every x, S1/S2 at 371
angles, Mueller elements) and integratePSD (size-distribution reduction).  Here the per-cell inputs (refractive index,
number weights) are still produced on the host with the reference's own formulas, but ALL cells of a bin go to the GPU
in one call (gm_table_run): the per-particle S1/S2 are never materialised, only the reduced sums come back, and
`combine_modes` / `postprocess` then apply integratePSD's closing arithmetic and the a-posteriori steps of `fun`
vectorised over cells.  Kernel-mode (GRASP/Saito) dust is out of scope (no Mie; external kernel files).
"""
import os
import sys

import numpy as np
from scipy.interpolate import interp1d

from . import _lib, ncio
from . import particleparams as pp
from .pymiecoated.mie_coated import MultipleMie  # noqa: F401  (re-exported like the reference module does)
from .pymiecoated.mie_coeffs import nmax_of

# key groups, same names and order as the reference (dointegration.py:16-24); the order of scalarkeys matters (:1156-1157)
scatkeys = ['p11', 'p12', 'p22', 'p33', 'p34', 'p44']
scalarkeys = ['qext', 'qsca', 'qabs', 'qb', 'g', 'csca', 'cext']
extrakeys = ['ssa', 'bsca', 'bext', 'bbck', 'refreal', 'refimag', 'lidar_ratio']
elekeys = ['pback']
nlscalarkeys = ['mass', 'volume', 'area', 'rEff', 'rMass', 'rhop', 'growth_factor', 'rLow', 'rUp']
allkeys = scatkeys + scalarkeys + extrakeys + elekeys + nlscalarkeys

_S = _lib  # raw-sum indices S_W ... S_CEXT
_trapz = getattr(np, "trapezoid", None) or np.trapz


# ----------------------------------------------------------------------------------------------- grids
def table_angles():
    """The 371 output scattering angles in degrees (dointegration.py:739-742)."""
    return np.concatenate([np.linspace(0., 1., 100, endpoint=False), np.linspace(1., 10., 100, endpoint=False),
                           np.linspace(10., 180., 171, endpoint=True)])


def getDR(arr):
    """Centred bin widths of a grid (dointegration.py:93-101)."""
    arr = np.asarray(arr, dtype=float)
    out = np.empty_like(arr)
    out[0] = arr[1] - arr[0]
    out[-1] = arr[-1] - arr[-2]
    out[1:-1] = ((arr[1:-1] - arr[:-2]) + (arr[2:] - arr[1:-1])) / 2.
    return out


def getXArrCarma(minx, maxx, nbinperdecade):
    """CARMA-style grid, geometric in volume (dointegration.py:103-153).  The loop form keeps the reference's
    floating-point sequence (rvolmin * rmRat ** i evaluated per point)."""
    rat = maxx / minx
    nbin = np.log10(maxx / minx) * nbinperdecade
    rmRat = (rat ** 3) ** (1. / nbin)
    rMin = minx * ((1. + rmRat) / 2.) ** (1. / 3.)
    cpi = 4. / 3. * np.pi
    rvolmin = cpi * rMin ** 3.
    vrfact = ((3. / 2. / np.pi / (rmRat + 1)) ** (1. / 3.)) * (rmRat ** (1. / 3.) - 1.)
    # per point: rvol = rvolmin * rmRat ** i, r = (rvol / cpi) ** (1/3), dr = vrfact * rvol ** (1/3) -- as Python floats (the same
    # libm pow and IEEE products as the reference's numpy scalars, without their per-operation overhead)
    rvolmin, rmRat, vrfact, third = float(rvolmin), float(rmRat), float(vrfact), 1. / 3.
    rvol = [rvolmin * rmRat ** i for i in range(int(nbin))]
    return np.array([(v / cpi) ** third for v in rvol]), np.array([vrfact * v ** third for v in rvol])


def initializeXarr(params, radind, minlam, maxlam):
    """Per-bin size-parameter grid shared by all wavelengths and RH (dointegration.py:425-463)."""
    pparam = params['psd']['params']
    psdtype = params['psd']['type']
    if psdtype == 'lognorm':
        lo = pparam['rmin0'][radind][0] * 2 * np.pi / maxlam
        hi = pparam['rmax0'][radind][-1] * 2 * np.pi / minlam * 3.0
        if 'numperdec' not in pparam:
            print("numperdec missing from dict")
            sys.exit()
        return getXArrCarma(lo, hi, pparam['numperdec'][radind])
    if psdtype == 'ss':
        lo = pparam['rMinMaj'][radind] * 2 * np.pi / maxlam
        hi = pparam['rMaxMaj'][radind] * 2 * np.pi / minlam * 10
        return getXArrCarma(lo, hi, pparam['numperdec'][radind])
    if psdtype == 'du':
        lo = pparam['rMinMaj'][radind][0] * 2 * np.pi / maxlam
        hi = pparam['rMaxMaj'][radind][-1] * 2 * np.pi / minlam
        return (np.linspace(np.log10(lo), np.log10(hi), 1000)) ** 10., None
    raise ValueError("unknown psd type %r" % psdtype)


# ----------------------------------------------------------------------------------------------- per-cell host inputs
def getHumidRefractiveIndex(params, radind, rhi, rh, nref0, nrefwater):
    """Volume mixing with water, n = n_w + (n_0 - n_w) rrat^3 (dointegration.py:493-536).  Returns mr, mi, gf, rrat."""
    pparam = params['psd']['params']
    rparams = params['rhDep']
    typ = rparams['type']
    if typ == 'simple' or (typ == 'trivial' and rhi == 0):
        gf = rparams['params']['gf'][rhi]
        rrat = (1. / gf)
    elif typ in ('ss', 'su'):
        if typ == 'ss':
            try:
                rMinMaj = pparam['rMinMaj'][radind]
            except Exception:
                rMinMaj = pparam['rmin0'][radind][0]
        else:
            rMinMaj = pparam['rmin0'][radind][0]
        rMinUse = pp.humidityGrowth(rparams, rMinMaj, rh[rhi], rh)
        rrat = (rMinMaj / rMinUse)
        gf = 1. / rrat
    else:
        print("PROBLEM! NO RHDEP DEFINED!")
        sys.exit()
    nrefUse = [nrefwater + (n0 - nrefwater) * (rrat) ** 3. for n0 in nref0]
    return [n.real for n in nrefUse], [n.imag for n in nrefUse], gf, rrat


def calculatePSD(params, radind, onerh, rh, xxarr, drarr, rrat, lam):
    """Number weights per grid point for every mode of the bin (dointegration.py:539-664).
    Returns psd (list of arrays summing to 1), ref (mass-effective radii), rLow, rUp."""
    pparam = params['psd']['params']
    psdtype = params['psd']['type']
    rparams = params['rhDep']
    rarr = xxarr / (2 * np.pi / lam)
    rr = xxarr * lam / 2. / np.pi
    psd, ref = [], []
    if psdtype == 'lognorm':
        rmodes = [pp.humidityGrowth(rparams, r0, onerh, rh) for r0 in pparam['r0'][radind]]
        rmaxs0 = list(pparam['rmax0'][radind])
        rmaxs = [pp.humidityGrowth(rparams, r, onerh, rh) for r in rmaxs0]
        rmins = list(pparam['rmin0'][radind])      # rmin does not grow with humidity
        rLow, rUp = rmins[0], rmaxs0[0]
        sigmas = pparam['sigma'][radind]
        for k in range(len(rmodes)):
            w = pp.getLogNormPSD(rmodes[k], sigmas[k], xxarr, lam, rmaxs[k], rmins[k])   # dN/dr
            w *= drarr                                                                   # -> dN
            w /= np.sum(w)
            psd.append(w)
            ref.append(np.sum(rr ** 4. * w) / np.sum(rr ** 3. * w))
    elif psdtype == 'ss':
        rMinMaj, rMaxMaj = pparam['rMinMaj'][radind], pparam['rMaxMaj'][radind]
        rLow, rUp = rMinMaj, rMaxMaj
        rMinUse = pp.humidityGrowth(rparams, rMinMaj, onerh, rh)
        rMaxUse = pp.humidityGrowth(rparams, rMaxMaj, onerh, rh)
        rrat = (rMinMaj / rMinUse)
        dr = getDR(rarr)
        r80rat = 1.65 * rrat
        r80 = rarr * r80rat * 1e6
        aFac = 4.7 * (1. + 30. * r80) ** (-0.017 * r80 ** (-1.44))
        bFac = (0.433 - np.log10(r80)) / 0.433
        dndr = 1.373 * r80 ** (-aFac) * (1. + 0.057 * r80 ** 3.45) * 10. ** (1.607 * np.exp(-bFac ** 2.)) * r80rat
        dndr[rarr < rMinUse] = 0.
        dndr[rarr > rMaxUse] = 0.
        dndr = dndr / np.sum(dndr)
        w = dndr * dr
        w /= np.sum(w)
        psd = [w]
        ref = [np.sum(rr ** 4 * w) / np.sum(rr ** 3 * w)]
    elif psdtype == 'du':
        for rMinMaj, rMaxMaj in zip(pparam['rMinMaj'][radind], pparam['rMaxMaj'][radind]):
            rLow, rUp = rMinMaj, rMaxMaj
            dndr = rarr ** (-4.)
            dndr[rarr < rMinMaj] = 0.
            dndr[rarr > rMaxMaj] = 0.
            if np.max(dndr) == 0.:
                print('prc:', np.min(xxarr), np.max(xxarr), np.min(rarr), np.max(rarr), lam, rMinMaj, rMaxMaj)
                sys.exit()
            w = dndr * getDR(xxarr)
            w /= np.sum(w)
            psd.append(w)
            ref.append(np.sum(rr ** 4. * w) / np.sum(rr ** 3. * w))
    else:
        raise ValueError("unknown psd type %r" % psdtype)
    return psd, ref, rLow, rUp


# ----------------------------------------------------------------------------------------------- reductions
def combine_modes(scal, phase, fracs, lam, reff0, rhop0, rhop):
    """integratePSD's closing arithmetic (dointegration.py:1104-1200) from the GPU raw sums.

    scal  [..., nmode, GM_NSCAL] raw sums in size-parameter space (r = x lam / 2 pi is applied here)
    phase [..., 4, nang]         sum_x w P(x, theta), already weighted by the mode fractions
    fracs, reff0 [nmode]; lam, rhop0, rhop broadcastable to the leading dims.  Returns the `ret` dict of integratePSD.
    """
    scal = np.asarray(scal, dtype=float)
    lead = scal.shape[:-2]
    nmode = scal.shape[-2]
    c = np.asarray(lam, dtype=float) / (2. * np.pi)
    c = np.broadcast_to(c, lead)[..., None]
    rhop = np.broadcast_to(np.asarray(rhop, dtype=float), lead)[..., None]
    rhop0 = np.broadcast_to(np.asarray(rhop0, dtype=float), lead)[..., None]
    fr = np.asarray(fracs, dtype=float).reshape((1,) * len(lead) + (nmode,))
    reff0 = np.asarray(reff0, dtype=float)
    if reff0.ndim == 1:
        reff0 = reff0.reshape((1,) * len(lead) + (nmode,))
    S = lambda k: scal[..., k]
    num = S(_S.S_W)
    r2, r3, r4 = c ** 2 * S(_S.S_X2W), c ** 3 * S(_S.S_X3W), c ** 4 * S(_S.S_X4W)
    t = {}
    t['num'] = num
    t['area'] = np.pi * r2 / num
    t['volume'] = 4. / 3. * np.pi * r3 / num
    t['mass'] = t['volume'] * rhop
    t['rEff'] = r3 / r2
    reff_mass = r4 / r3
    t['rMass'] = 4. / 3. * np.pi * rhop * reff_mass ** 3.
    rMass0 = 4. / 3. * np.pi * rhop0 * reff0 ** 3
    # area-weighted efficiencies ('dumix', :1188-1190); the (lam/2pi)^2 factors cancel in the ratios
    t['qext'] = S(_S.S_QEXT) / S(_S.S_X2W)
    t['qsca'] = S(_S.S_QSCA) / S(_S.S_X2W)
    t['qabs'] = S(_S.S_QABS) / S(_S.S_X2W)
    # qb through the backed-out P11(180) (:1133-1145, :1168-1169)
    p11back = 4. * np.pi * S(_S.S_QB) / S(_S.S_QSCA)
    t['qb'] = p11back * t['qsca'] / (4. * np.pi)
    # keys after 'g' carry the (area * qsca) weight because of the in-place `thisweight *= qsca` (:1111-1112, :1157)
    t['g'] = S(_S.S_G) / S(_S.S_QSCA)
    t['csca'] = np.pi * c ** 2 * S(_S.S_CSCA) / S(_S.S_QSCA)
    t['cext'] = np.pi * c ** 2 * S(_S.S_CEXT) / S(_S.S_QSCA)
    conv = 1. / rhop / t['rEff'] * t['rMass'] / rMass0            # :1193
    t['bsca'] = 3. / 4. * t['qsca'] * conv
    t['bext'] = 3. / 4. * t['qext'] * conv
    t['bbck'] = 3. / 4. * t['qb'] * conv / (4 * np.pi)
    ret = {k: np.sum(fr * v, axis=-1) for k, v in t.items()}       # ret[key] += frac * thisret[key], :1199-1200
    ret['lidar_ratio'] = np.zeros(lead)
    if phase is None:          # the phase matrix stays on the device (gm_table_fetch_normalized delivers it normalised)
        return ret
    ph = np.asarray(phase, dtype=float)
    ret['p11'] = ph[..., 0, :]
    ret['p12'] = ph[..., 1, :]
    ret['p22'] = ph[..., 0, :].copy()
    ret['p33'] = ph[..., 2, :]
    ret['p34'] = ph[..., 3, :]
    ret['p44'] = ph[..., 2, :].copy()
    return ret


def postprocess(ret, ang):
    """Extra variables and the a-posteriori phase-matrix normalisation of `fun` (dointegration.py:950-988)."""
    ret['lidar_ratio'] = ret['qext'] / ret['qb'] * 4 * np.pi
    ret['ssa'] = ret['qsca'] / ret['qext']
    theta = np.radians(ang)
    p11 = ret['p11']
    p11n = 2. * p11 / _trapz(p11 * np.sin(theta), theta)[..., None]
    same44 = ret['p44'] is ret['p33'] or np.array_equal(ret['p44'], ret['p33'])
    for k in ('p12', 'p22', 'p33', 'p34'):
        ret[k] = ret[k] * p11n / p11
    # spheres: p44 == p33 on input (calculateScatVals), so the same arithmetic gives the same bits -- evaluated once
    ret['p44'] = ret['p33'].copy() if same44 else ret['p44'] * p11n / p11
    ret['p11'] = p11n
    ret['pback'] = np.stack([ret[k][..., -1] for k in ('p11', 'p12', 'p33', 'p34', 'p22', 'p44')], axis=-1)
    return ret


# ----------------------------------------------------------------------------------------------- reference-shaped API
class RawMie(dict):
    """Return value of rawMie: the reference's dict of per-particle arrays, plus what integratePSD needs to run the
    fused GPU reduction instead of re-reducing the arrays on the host."""
    mm = None
    m = None


def rawMie(mm, scatkeys_, scalarkeys_, lam, mr, mi, psd, costarr):
    """Per-particle Mie over the grid of `mm` at one refractive index (dointegration.py:1211-1254): efficiencies,
    csca/cext and the six Mueller elements [nx, nang], computed on the GPU (DMMA per-particle path)."""
    x = np.asarray(mm.xArr, dtype=float)
    rr = x * lam / (2. * np.pi)
    q, s12 = mm.calculateS12SizeRangeArrays(mr, mi)
    s1 = s12[..., 0] + 1j * s12[..., 1]
    s2 = s12[..., 2] + 1j * s12[..., 3]
    ret = RawMie()
    ret.mm, ret.m = mm, (mr, mi)
    a1, a2 = np.abs(s1) ** 2, np.abs(s2) ** 2
    vals = {'p11': 0.5 * (a1 + a2), 'p12': 0.5 * (a2 - a1), 'p22': 0.5 * (a1 + a2), 'p33': (s1 * np.conj(s2)).real,
            'p34': -(np.conj(s1) * s2).imag, 'p44': (s1 * np.conj(s2)).real,
            'qext': q[:, 0], 'qsca': q[:, 1], 'qabs': q[:, 2], 'qb': q[:, 3], 'g': q[:, 4],
            'csca': q[:, 1] * np.pi * rr ** 2, 'cext': q[:, 0] * np.pi * rr ** 2}
    for k in list(scatkeys_) + list(scalarkeys_):
        ret[k] = vals[k]
    return ret


def integratePSD(xxarr, rawret, psd, fracs, lam, reff0, rhop0, rhop):
    """Size-distribution integration of rawMie results (dointegration.py:1064-1209).  `rawret` must come from this
    module's rawMie: the reduction is re-run fused on the GPU (gm_table_run) from the (table, m) each entry carries."""
    tables = [r.mm.device_table() for r in rawret]
    nmode = len(fracs)
    ws = np.asarray(psd, dtype=float)
    same = all(r.m == rawret[0].m for r in rawret)
    if same:
        m = complex(*rawret[0].m)
        mz = [np.sqrt(m ** 2 * 1.0)]
        wp = np.tensordot(np.asarray(fracs, dtype=float), ws, axes=(0, 0))[None]
        scal, phase = tables[0].run(mz, mz, wp, ws[None])
        return combine_modes(scal[0], phase[0], fracs, lam, reff0, rhop0, rhop)
    mz = [np.sqrt(complex(*r.m) ** 2 * 1.0) for r in rawret]
    wp = np.asarray(fracs, dtype=float)[:, None] * ws
    scal, phase = tables[0].run(mz, mz, wp, ws[:, None, :])
    return combine_modes(scal[:, 0, :], phase.sum(axis=0), fracs, lam, reff0, rhop0, rhop)


# ----------------------------------------------------------------------------------------------- file layout
_VARMETA = {
    'p11': ('dimensionless', 'P11 element of the normalized scattering matrix'),
    'p12': ('dimensionless', 'P12 element of the normalized scattering matrix'),
    'p22': ('dimensionless', 'P22 element of the normalized scattering matrix'),
    'p33': ('dimensionless', 'P33 element of the normalized scattering matrix'),
    'p34': ('dimensionless', 'P34 element of the normalized scattering matrix'),
    'p44': ('dimensionless', 'P44 element of the normalized scattering matrix'),
    'qsca': ('dimensionless', 'scattering efficiency'),
    'qabs': ('dimensionless', 'absorption efficiency'),
    'qext': ('dimensionless', 'extinction efficiency'),
    'g': ('dimensionless', 'asymmetry factor'),
    'ssa': ('dimensionless', 'single-scattering albedo'),
    'qb': ('dimensionless', 'backscattering efficiency'),
    'bsca': ('m2 (kg dry mass)-1', 'mass scattering efficiency'),
    'bext': ('m2 (kg dry mass)-1', 'mass extinction efficiency'),
    'csca': ('m2', 'mass scattering cross-section'),
    'cext': ('m2', 'mass extinction cross-section'),
    'bbck': ('m2 (kg dry mass)-1 sr-1', 'mass backscatter efficiency'),
    'lidar_ratio': ('', 'lidar ratio'),
    'mass': ('kg', 'particle mass'),
    'volume': ('m3 kg-1', 'particle volume per kg dry mass'),
    'area': ('m2 kg-1', 'particle cross sectional area per kg dry mass'),
    'rEff': ('m', 'effective radius of bin'),
    'rMass': ('kg', 'effective mass of wet particle'),
    'rUp': ('m', 'upper edge of radius bin'),
    'rLow': ('m', 'lower edge of radius bin'),
    'pback': ('dimensionless', 'phase function in backscatter direction, ordered as P11, P12, P33, P34, P22, P44'),
    'rhop': ('kg m-3', 'wet particle density'),
    'growth_factor': ('fraction', 'growth factor = ratio of wet to dry particle radius'),
    'refreal': ('dimensionless', 'real refractive index of wet particle'),
    'refimag': ('dimensionless', 'imaginary refractive index of wet particle'),
}
# creation order of the data variables in the reference file (dointegration.py:366-397)
_VARORDER = ['p11', 'p12', 'p22', 'p33', 'p34', 'p44', 'qsca', 'qabs', 'qext', 'g', 'ssa', 'qb', 'bsca', 'bext', 'csca',
             'cext', 'bbck', 'lidar_ratio', 'mass', 'volume', 'area', 'rEff', 'rMass', 'rUp', 'rLow', 'pback', 'rhop',
             'growth_factor', 'refreal', 'refimag']


def _dimtypes(oppclassic):
    radius, lamb, npol = ('radius', 'lambda', 'nPol') if oppclassic else ('bin', 'wavelength', 'p')
    if oppclassic:
        kinds = {"scal": (radius, "rh", lamb), "nl": (radius, "rh"), "scat": (radius, "rh", lamb, "ang"),
                 "ele": (npol, radius, "rh", lamb)}
    else:
        kinds = {"scal": (radius, lamb, "rh"), "nl": (radius, "rh"), "scat": (radius, lamb, "rh", "ang"),
                 "ele": (radius, lamb, "rh", "p")}
    return radius, lamb, npol, kinds


def _kind_of(key):
    if key in scatkeys:
        return "scat"
    if key in elekeys:
        return "ele"
    if key in nlscalarkeys:
        return "nl"
    return "scal"


def createNCDF(ncdfID, oppfx, rarr, rharr, lambarr, ang, oppclassic):
    """Create the output file with the reference's dimensions, variables, attributes and order
    (dointegration.py:205-422): optics_<id>.nomom[.legacy].nc4."""
    radius, lamb, npol, kinds = _dimtypes(oppclassic)
    fn = 'optics_%s.nomom.legacy.nc4' % ncdfID if oppclassic else 'optics_%s.nomom.nc4' % ncdfID
    nc = ncio.Dataset(os.path.join(oppfx, fn), 'w')
    nc.createDimension('rh', len(rharr))
    nc.createDimension(lamb, len(lambarr))
    nc.createDimension(radius, len(rarr))
    nc.createDimension(npol, 6)
    nc.createDimension('ang', len(ang))
    coord = [('rh', 'f8', 'fraction', 'relative humidity', rharr),
             (lamb, 'f8', 'm', 'wavelength', lambarr),
             (radius, 'i8', 'dimensionless', 'radius bin index (1-indexed)', range(1, len(rarr) + 1)),
             (npol, 'i8', 'dimensionless', 'Scattering matrix element index, ordered as P11, P12, P33, P34, P22, P44',
              [11, 12, 33, 34, 22, 44]),
             ('ang', 'f8', 'degrees', 'scattering angle', ang)]
    for name, dt, units, long_name, values in coord:
        v = nc.createVariable(name, dt, (name), compression='zlib')
        v.long_name = long_name
        v.units = units
        v[:] = np.asarray(list(values))
    for key in _VARORDER:
        v = nc.createVariable(key, 'f8', kinds[_kind_of(key)], compression='zlib')
        v.long_name = _VARMETA[key][1]
        v.units = _VARMETA[key][0]
    return nc


# ----------------------------------------------------------------------------------------------- the table build
def bins_of(params):
    """Bin indices and nominal radii (dointegration.py:713-725)."""
    p = params['psd']['params']
    t = params['psd']['type']
    if t == 'lognorm':
        return list(range(len(p['r0']))), list(p['r0'])
    if t == 'ss':
        return list(range(len(p['rMinMaj']))), list(p['rMinMaj'])
    if t == 'du':
        return list(range(len(p['rMinMaj']))), [xx[0] for xx in p['rMinMaj']]
    raise ValueError("unknown psd type %r" % t)


def psd_parameters_rh(params, radind, onerh, rh):
    """The humidity-dependent scalars of calculatePSD (dointegration.py:539-664): everything psd_parameters needs except the
    wavelength.  Returns (kind, modes, rLow, rUp); modes[k] = the lengths [m] of mode k that get multiplied by 2 pi / lambda
    (lognorm: r_mode, r_min, r_max, then ln sigma as is; ss: r_min_use, r_max_use, r_min / r_min_use as they are)."""
    pparam = params['psd']['params']
    psdtype = params['psd']['type']
    rparams = params['rhDep']
    if psdtype == 'lognorm':
        rmodes = [pp.humidityGrowth(rparams, r0, onerh, rh) for r0 in pparam['r0'][radind]]
        rmaxs0 = list(pparam['rmax0'][radind])
        rmaxs = [pp.humidityGrowth(rparams, r, onerh, rh) for r in rmaxs0]
        rmins = list(pparam['rmin0'][radind])
        sig = pparam['sigma'][radind]
        return _lib.PSD_LOGNORM, [[rmodes[k], rmins[k], rmaxs[k], np.log(sig[k])] for k in range(len(rmodes))], rmins[0], rmaxs0[0]
    if psdtype == 'ss':
        rMinMaj, rMaxMaj = pparam['rMinMaj'][radind], pparam['rMaxMaj'][radind]
        rMinUse = pp.humidityGrowth(rparams, rMinMaj, onerh, rh)
        rMaxUse = pp.humidityGrowth(rparams, rMaxMaj, onerh, rh)
        return _lib.PSD_SS, [[1.0, rMinUse, rMaxUse, rMinMaj / rMinUse]], rMinMaj, rMaxMaj
    if psdtype == 'du':
        lo, hi = pparam['rMinMaj'][radind], pparam['rMaxMaj'][radind]
        return _lib.PSD_DU, [[1.0, a, b, 0.0] for a, b in zip(lo, hi)], lo[-1], hi[-1]
    raise ValueError("unknown psd type %r" % psdtype)


def psd_parameters(params, radind, onerh, rh, lam):
    """The per-cell scalars of calculatePSD (dointegration.py:539-664) without its O(nx) array work: what the device
    kernel k_psd needs to generate the number weights.  Returns (kind, par [nmode][4], rLow, rUp)."""
    kind, modes, rLow, rUp = psd_parameters_rh(params, radind, onerh, rh)
    xconv = 2 * np.pi / lam
    if kind == _lib.PSD_LOGNORM:
        return kind, [[m[0] * xconv, m[1] * xconv, m[2] * xconv, m[3]] for m in modes], rLow, rUp
    return kind, [[xconv, m[1], m[2], m[3]] for m in modes], rLow, rUp


class WavelengthSubset(object):
    """`cells` argument of BinPlan: every RH cell of the given wavelength indices (what a rank of a multi-GPU build owns)."""

    def __init__(self, li):
        self.li = np.asarray(li, dtype=np.int64)

    def as_set(self, nr, trivial):
        return {(int(l), r) for l in self.li for r in range(nr) if not (trivial and r > 0)}


class BinPlan(object):
    """Host-side description of one size bin: the grid and, for every (wavelength, RH) cell that must be computed, the
    refractive indices, the size-distribution inputs and the scalars that the reference derives inside its lambda/RH
    loops (dointegration.py:811-889).  `tasks` flattens the cells into gm_table_run tasks.

    device_psd=True (default in `fun`): only the per-cell PSD parameters are derived here and the weight vectors are
    generated on the GPU (gm_table_run_psd); device_psd=False builds the weights with numpy exactly like the reference
    (calculatePSD) and uploads them -- used for validation."""

    def __init__(self, params, radind, lambarr, rh_used, part_m, water_m, cells=None, device_psd=False):
        self.radind = radind
        if device_psd and float(np.asarray(rh_used)[0]) != 0.0:
            # reff_mass0 is defined at the humidity VALUE 0.0 (calculatePSD(..., onerh=0., ...), dointegration.py:837-838); the
            # device-PSD path takes it from the RH-index-0 cell, which is the same thing only when the list starts at 0.0
            device_psd = False
        if params['psd']['type'] == 'du':
            # the reference's 'du' grid, (linspace(log10 x))**10 (dointegration.py:456-460), is not monotonic, so getDR and with
            # it the number weights change sign along the grid (:640-648): build them with numpy like the reference and let
            # evaluate() treat the two signs separately (the device folds sqrt(w) into the coefficients)
            device_psd = False
        self.xx, self.dr = initializeXarr(params, radind, lambarr[0], lambarr[-1])
        self.nmax = nmax_of(self.xx)
        pparam = params['psd']['params']
        trivial = params['rhDep']['type'] == 'trivial'
        rhop00 = params['rhop0']
        self.rhop0 = rhop00[radind] if isinstance(rhop00, list) else rhop00
        self.fracs = list(pparam['fracs'][radind])
        self.nri = len(part_m)
        self.cells = []          # (li, rhi)
        self.trivial = trivial
        self.device_psd = device_psd
        self.psd_kind = None
        mr_l, w_l, par_l, meta = [], [], [], []
        if isinstance(cells, WavelengthSubset):
            want = cells if device_psd else cells.as_set(len(rh_used), trivial)      # whole wavelengths: the RH-index-0 cells are included
        else:
            want = None if cells is None else set(cells)
            if device_psd and want:
                want |= {(li, 0) for (li, _) in want}      # reff_mass0 comes from the RH-index-0 cell of the same wavelength
        if device_psd:
            # the wavelength and the humidity enter the per-cell inputs separately: 61 + 36 evaluations instead of 61 x 36
            self._scan_cells_separable(params, radind, lambarr, rh_used, part_m, water_m, want, trivial)
            return
        with pp.growth_memo():
            self._scan_cells(params, radind, lambarr, rh_used, part_m, water_m, want, trivial, device_psd, mr_l, w_l, par_l, meta)
        ncell, nmode = len(self.cells), len(self.fracs)
        self.m = np.array(mr_l, dtype=np.complex128).reshape(ncell, self.nri)
        self.w = None if device_psd else np.array(w_l, dtype=float).reshape(ncell, nmode, self.xx.size)
        self.psd_par = np.array(par_l, dtype=float).reshape(ncell, nmode, 4) if device_psd else None
        self.lam = np.array([a[0] for a in meta])
        self.rhop = np.array([a[1] for a in meta])
        self.gf = np.array([a[2] for a in meta])
        self.rLow = np.array([a[3] for a in meta])
        self.rUp = np.array([a[4] for a in meta])
        self.reff0 = None if device_psd else np.array([a[5] for a in meta]).reshape(ncell, nmode)

    def _scan_cells_separable(self, params, radind, lambarr, rh_used, part_m, water_m, want, trivial):
        """The lambda / RH loops of dointegration.fun (:811-889) for the device-PSD path, without a Python loop over cells.
        Every per-cell input is a product of a wavelength-only and a humidity-only factor: n_0(lambda), n_water(lambda) and
        2 pi / lambda on one side, the growth factor, rrat^3, the wet density and the grown mode radii on the other.  The
        humidity-only scalars are evaluated with the reference's own scalar statements (one call per RH level), the outer
        combination is the same IEEE multiply / add per element, so the result is bit-identical to the cell loop
        (tests/test_host_logic.py, tests/test_live_reference.py)."""
        lam = np.asarray(lambarr, dtype=float)
        nl, nr = lam.size, len(rh_used)
        rlist = [0] if trivial else list(range(nr))
        # humidity-only part (getHumidRefractiveIndex :493-536 with a dummy index, psd_parameters_rh)
        gf, r3, rhop, modes = [], [], [], []
        kind = rLow = rUp = None
        with pp.growth_memo():
            for rhi in rlist:
                _, _, g, rrat = getHumidRefractiveIndex(params, radind, rhi, rh_used, [0j], 0j)
                kind, md, rLow, rUp = psd_parameters_rh(params, radind, rh_used[rhi], rh_used)
                gf.append(g)
                r3.append(rrat ** 3.)
                rhop.append(rrat ** 3. * self.rhop0 + (1. - rrat ** 3.) * 1000.)
                modes.append(md)
        self.psd_kind = kind
        gf, r3, rhop, modes = np.array(gf, dtype=float), np.array(r3, dtype=float), np.array(rhop, dtype=float), np.array(modes, dtype=float)
        # wavelength-only part (:813-826): the stored -k is flipped, water is used as is
        n0_re = np.stack([np.asarray(f[0](lam), dtype=float) for f in part_m], axis=1)            # [nl][nri]
        n0_im = np.stack([-np.asarray(f[1](lam), dtype=float) for f in part_m], axis=1)
        if trivial:
            nw_re, nw_im = np.ones(nl), np.zeros(nl)
        else:
            nw_re, nw_im = np.asarray(water_m[0](lam), dtype=float), np.asarray(water_m[1](lam), dtype=float)
        # n = n_w + (n_0 - n_w) rrat^3 component by component (a Python complex times a float is the same two products)
        mre = nw_re[:, None, None] + (n0_re - nw_re[:, None])[:, None, :] * r3[None, :, None]       # [nl][nrh][nri]
        mim = nw_im[:, None, None] + (n0_im - nw_im[:, None])[:, None, :] * r3[None, :, None]
        xconv = 2 * np.pi / lam
        nmode = modes.shape[1]
        par = np.empty((nl, len(rlist), nmode, 4))
        if kind == _lib.PSD_LOGNORM:
            par[..., :3] = modes[None, :, :, :3] * xconv[:, None, None, None]
            par[..., 3] = modes[None, :, :, 3]
        else:
            par[..., 0] = xconv[:, None, None]
            par[..., 1:] = modes[None, :, :, 1:]
        sel = np.ones((nl, len(rlist)), dtype=bool)
        if isinstance(want, WavelengthSubset):
            sel[:] = False
            sel[want.li, :] = True
        elif want is not None:
            sel[:] = False
            for (li, rhi) in want:
                if rhi in rlist:
                    sel[li, rlist.index(rhi)] = True
        li_idx, rj = np.nonzero(sel)                      # wavelength-major, RH-minor: the order of the reference's loops
        ri_idx = np.asarray(rlist, dtype=np.int64)[rj]
        self.cells = list(zip(li_idx.tolist(), ri_idx.tolist()))
        self.m = (mre[li_idx, rj] + 1j * mim[li_idx, rj]).reshape(len(self.cells), self.nri)
        self.w = None
        self.psd_par = np.ascontiguousarray(par[li_idx, rj])
        self.lam = lam[li_idx]
        self.rhop = rhop[rj]
        self.gf = gf[rj]
        self.rLow = np.full(len(self.cells), rLow, dtype=float)
        self.rUp = np.full(len(self.cells), rUp, dtype=float)
        self.reff0 = None

    def _scan_cells(self, params, radind, lambarr, rh_used, part_m, water_m, want, trivial, device_psd, mr_l, w_l, par_l, meta):
        """The lambda / RH loops of dointegration.fun (:811-889): per-cell refractive indices and size-distribution inputs."""
        for li, lam in enumerate(lambarr):
            if want is not None and not any(c[0] == li for c in want):
                continue
            nref0 = [complex(f[0](lam), -f[1](lam)) for f in part_m]      # sign flip of the stored -k (:813-816)
            nrefwater = complex(1, 0) if trivial else complex(water_m[0](lam), water_m[1](lam))
            reff_mass0 = None
            if not device_psd:
                _, _, _, rrat0 = getHumidRefractiveIndex(params, radind, 0, rh_used, nref0, nrefwater)
                _, reff_mass0, _, _ = calculatePSD(params, radind, 0., rh_used, self.xx, self.dr, rrat0, lam)
            for rhi, onerh in enumerate(rh_used):
                if trivial and rhi > 0:
                    continue
                if want is not None and (li, rhi) not in want:
                    continue
                mr, mi, gf, rrat = getHumidRefractiveIndex(params, radind, rhi, rh_used, nref0, nrefwater)
                if device_psd:
                    self.psd_kind, par, rLow, rUp = psd_parameters(params, radind, onerh, rh_used, lam)
                    par_l.append(par)
                else:
                    psd, _, rLow, rUp = calculatePSD(params, radind, onerh, rh_used, self.xx, self.dr, rrat, lam)
                    w_l.append(np.asarray(psd))
                rhop = rrat ** 3. * self.rhop0 + (1. - rrat ** 3.) * 1000.
                self.cells.append((li, rhi))
                mr_l.append([complex(a, b) for a, b in zip(mr, mi)])
                meta.append((lam, rhop, gf, rLow, rUp, None if reff_mass0 is None else list(reff_mass0)))

    # ---- flatten to GPU tasks
    def tasks(self):
        """Uploaded-weights path.  Returns (mz [ntask], w_phase [ntask][nx], w_scal [ntask][nmode_t][nx], tasks_per_cell)."""
        fr = np.asarray(self.fracs, dtype=float)
        ncell, nmode, nx = self.w.shape
        if self.nri == 1:
            # one Mie evaluation per cell, shared by all modes (allret replicated, dointegration.py:882-884)
            mz = np.sqrt(self.m[:, 0] ** 2 * 1.0)
            wp = np.tensordot(self.w, fr, axes=(1, 0))
            return mz, wp, self.w, 1
        assert self.nri == nmode, "one refractive index per PSD mode expected"
        mz = np.sqrt(self.m.reshape(-1) ** 2 * 1.0)
        wp = (self.w * fr[None, :, None]).reshape(ncell * nmode, nx)
        return mz, wp, self.w.reshape(ncell * nmode, 1, nx), nmode

    def tasks_psd(self):
        """Device-PSD path.  Returns (mz [ntask], par [ntask][nmode_t][4], frac [ntask][nmode_t], tasks_per_cell)."""
        fr = np.asarray(self.fracs, dtype=float)
        ncell, nmode, _ = self.psd_par.shape
        if self.nri == 1:
            return np.sqrt(self.m[:, 0] ** 2 * 1.0), self.psd_par, np.broadcast_to(fr, (ncell, nmode)).copy(), 1
        assert self.nri == nmode, "one refractive index per PSD mode expected"
        return (np.sqrt(self.m.reshape(-1) ** 2 * 1.0), self.psd_par.reshape(ncell * nmode, 1, 4),
                np.tile(fr, ncell).reshape(ncell * nmode, 1), nmode)

    def evaluate(self, table, elide=True, phase_on_device=False):
        """Run every task of the bin on `table`; returns (scal, phase, tasks_per_cell).  phase_on_device=True (device-PSD path with
        one task per cell only): the raw phase sums stay on the GPU and `phase` is None -- see Table.fetch_normalized."""
        if self.device_psd:
            mz, par, fr, tpc = self.tasks_psd()
            if self.psd_kind == _lib.PSD_LOGNORM:
                table.set_dr(self.dr)
            scal, phase = table.run_psd(mz, mz, self.psd_kind, par, fr, elide=elide, phase_on_device=phase_on_device and tpc == 1)
        else:
            mz, wp, ws, tpc = self.tasks()
            scal, phase = table.run(mz, mz, wp, ws, elide=elide)      # signed weights ('du' grid) are split inside Table.run
        return scal, phase, tpc

    def reduce(self, scal, phase, tasks_per_cell):
        """Raw sums of the tasks -> the integratePSD `ret` dict with a leading cell axis."""
        ncell = len(self.cells)
        if tasks_per_cell == 1:
            sc, ph = scal, phase
        else:
            sc = scal.reshape(ncell, tasks_per_cell, scal.shape[-1])
            ph = None if phase is None else phase.reshape(ncell, tasks_per_cell, 4, phase.shape[-1]).sum(axis=1)
        reff0 = self.reff0
        if reff0 is None:
            # reff_mass0 = sum r^4 w / sum r^3 w of the RH-index-0 weights at the same wavelength (:837-838, :1118-1121)
            ref = (self.lam / (2. * np.pi))[:, None] * sc[..., _S.S_X4W] / sc[..., _S.S_X3W]
            dry = {li: i for i, (li, rhi) in enumerate(self.cells) if rhi == 0}
            reff0 = ref[[dry[li] for li, _ in self.cells]]
        return combine_modes(sc, ph, self.fracs, self.lam, reff0, self.rhop0, self.rhop)


def run_bin(plan, costarr, handle=None, elide=True, table=None):
    """GPU evaluation of every cell of a bin.  Returns (ret dict with leading cell axis, table)."""
    own = table is None
    if own:
        table = _lib.Table(plan.xx, plan.nmax, costarr, handle)
    scal, phase, tpc = plan.evaluate(table, elide=elide)
    ret = plan.reduce(scal, phase, tpc)
    return ret, table


def fun(partID0, datatype, oppfx, oppclassic, elide=True, write=True, comm=None, device_psd=True, keep_phase=True):
    """Main table build called from runoptics.py (dointegration.py:672-1036).  Same arguments as the reference plus
    `elide` (skip exactly-zero-weight particles; result-neutral), `write` (False: return the arrays only), `comm`
    (a geosmie_b200.dist.Comm: cells are sharded across ranks and gathered to rank 0) and `device_psd` (True: number
    weights generated on the GPU from per-cell parameters; False: numpy weights exactly like calculatePSD, uploaded);
    `keep_phase=False` (only with write=False) drops the six [bin, wavelength, rh, ang] arrays from the result -- for
    fine spectral grids whose only consumer is the band averaging (pback is still filled).  Returns the dict of arrays
    written to optics_<id>.nomom[.legacy].nc4 on rank 0 (None elsewhere)."""
    partID = partID0.split('/')[-1].replace(".json", "")
    print("\n ####################\n Starting case %s\n ####################\n" % partID)
    partID2 = partID0.replace('-orig', '') if '-orig' in partID else partID0
    params = pp.getParticleParams(partID2, datatype)
    mode = params.get('mode', 'mie')
    if mode != 'mie':
        raise NotImplementedError("mode %r (GRASP/Saito kernels) is outside the Mie hot path of this build" % mode)
    mList = params['mList']
    lambarr = mList[0][0]
    part_m = [(interp1d(m[0], m[1]), interp1d(m[0], m[2])) for m in mList]
    water_m = None
    if params['rhDep']['type'] != 'trivial':
        wl = pp.getWaterM()
        water_m = (interp1d(wl[0], wl[1]), interp1d(wl[0], wl[2]))
    radindarr, radiusarr = bins_of(params)
    rh = params['rh']
    ang = table_angles()
    costarr = np.cos(np.radians(ang))
    # the file's rh coordinate is the un-capped list; the physics uses the capped one (createNCDF is called before
    # the maxrh cap, dointegration.py:753 vs :828-832)
    rh_used = rh
    if 'maxrh' in params:
        rh_used = np.array(rh)
        rh_used[np.where(rh_used > params['maxrh'])[0]] = params['maxrh']
    nb, nl, nr, na = len(radiusarr), len(lambarr), len(rh), len(ang)
    vals = {}
    if not keep_phase and write:
        raise ValueError("keep_phase=False needs write=False (the file layout contains the phase matrices)")
    for key in allkeys:
        if not keep_phase and key in scatkeys:
            continue
        shape = {"scat": (nb, nl, nr, na), "ele": (nb, nl, nr, 6), "nl": (nb, nr), "scal": (nb, nl, nr)}[_kind_of(key)]
        vals[key] = np.zeros(shape)

    rank, world = (0, 1) if comm is None else (comm.rank, comm.world)
    keys = list(vals.keys())
    width = {k: {"scat": na, "ele": 6}.get(_kind_of(k), 1) for k in keys}
    for radind in radindarr:
        print("=== === === USING RADIND %d" % radind)
        trivial = params['rhDep']['type'] == 'trivial'
        # Multi-GPU: rank r owns the wavelengths r, r+W, ... with all their RH cells (the RH-index-0 cell that defines
        # mass0 / reff_mass0 stays local), reduces and post-processes them, and only finished rows travel to rank 0.
        mine = None if world == 1 else WavelengthSubset(np.arange(rank, nl, world))
        plan = BinPlan(params, radind, lambarr, rh_used, part_m, water_m, cells=mine, device_psd=device_psd)
        ncell = len(plan.cells)
        li = np.array([c[0] for c in plan.cells], dtype=np.int64)
        ri = np.array([c[1] for c in plan.cells], dtype=np.int64)
        ret, in_place, phase_block, table = None, False, None, None
        # Multi-GPU with the phase matrices normalised on the device: the block of every rank goes to rank 0's peer-mapped buffer
        # straight from device memory and rank 0 reads every segment with ONE strided copy per plane into rows r, r + W, ... of the
        # final arrays -- no host staging, no row matrix, no scatter (decided identically on every rank)
        peer_phase = False
        if world > 1 and keep_phase and plan.device_psd and plan.nri == 1:
            nrr = 1 if trivial else nr
            seg_bytes = (4 * na + 4) * 8 * ((nl + world - 1) // world) * nrr
            pg = comm.peer_gather(seg_bytes)
            peer_phase = pg is not None
        if ncell:
            table = _lib.Table(plan.xx, plan.nmax, costarr)
            # One task per cell and weights made on the device (every shipped Mie species except the 'du' grid and multi-index
            # bins): the a-posteriori normalisation runs on the GPU (k_phase_norm) and the four distinct phase-matrix planes arrive
            # normalised -- straight in their final arrays when this rank owns the whole bin in file order (one GPU).
            on_device = plan.device_psd and plan.nri == 1
            scal, phase, tpc = plan.evaluate(table, elide=elide, phase_on_device=on_device)
            ret = plan.reduce(scal, phase, tpc)
            if phase is None:
                ret['lidar_ratio'] = ret['qext'] / ret['qb'] * 4 * np.pi
                ret['ssa'] = ret['qsca'] / ret['qext']
                in_place = world == 1 and ncell == nl * nr                 # this rank owns the whole bin, cells in file order
                if keep_phase and peer_phase:
                    phase_block = table.normalize_device(ang)         # stays on the GPU: it travels to rank 0 over NVLink below
                    pb = None
                elif keep_phase:
                    dst = [vals[k][radind].reshape(ncell, na) for k in ('p11', 'p12', 'p33', 'p34')] if in_place else None
                    planes, pb = table.fetch_normalized(ang, out=dst)
                    ret['p11'], ret['p12'], ret['p33'], ret['p34'] = planes
                    ret['p22'], ret['p44'] = planes[0], planes[2]      # spheres (calculateScatVals, dointegration.py:1044-1050)
                else:
                    pb = table.fetch_pback(ang)                        # the phase matrices never leave the GPU
                if pb is not None:
                    ret['pback'] = pb[:, [0, 1, 2, 3, 0, 2]]
            else:
                ret = postprocess(ret, ang)
            # mass0 = volume(RH index 0) * rhop0 of the same (bin, lambda) (dointegration.py:992-997)
            vol0 = np.zeros(nl)
            sel0 = ri == 0
            vol0[li[sel0]] = ret['volume'][sel0]
            mass0 = vol0[li] * plan.rhop0
            ret['area'] = ret['area'] / mass0
            ret['volume'] = ret['volume'] / mass0
            ret['rhop'] = plan.rhop
            ret['growth_factor'] = plan.gf
            ret['rLow'] = plan.rLow
            ret['rUp'] = plan.rUp
            ret['refreal'] = plan.m[:, 0].real
            ret['refimag'] = -np.abs(plan.m[:, 0].imag)
        if peer_phase:
            if phase_block is not None:
                pg.put(0, phase_block[0], phase_block[1])
            pg.complete()                                          # every rank's block has landed in rank 0's buffer
            if rank == 0:
                h0 = pg.h
                row = nrr * na * 8                                 # one wavelength of one plane: nrr RH x na angles
                pbs = []
                for r in range(world):
                    nlr = len(range(r, nl, world))
                    ncr = nlr * nrr
                    if ncr == 0:
                        pbs.append(np.zeros((0, 4)))
                        continue
                    seg = pg.seg_ptr(0, 0, rank=r)
                    for q, key in enumerate(('p11', 'p12', 'p33', 'p34')):
                        dst = vals[key][radind, r]                 # rows r, r + W, ... of (wavelength, rh, ang)
                        h0.peer_get2d(dst.ctypes.data, world * nr * na * 8, seg + q * ncr * na * 8, row, row, nlr)
                    pbr = np.empty((ncr, 4))
                    h0.peer_put(pbr.ctypes.data, seg + 4 * ncr * na * 8, pbr.nbytes)
                    pbs.append(pbr)
                h0.peer_sync()
                for r in range(world):
                    if pbs[r].shape[0]:
                        vals['pback'][radind, r::world, :nrr] = pbs[r][:, [0, 1, 2, 3, 0, 2]].reshape(-1, nrr, 6)
                vals['p22'][radind] = vals['p11'][radind]
                vals['p44'][radind] = vals['p33'][radind]
            comm.barrier()                                         # the buffer may be reused (next bin) only now
        if table is not None:
            table.close()
        if world > 1:
            # only the distinct columns travel: p22 / p44 are rebuilt from p11 / p33 on rank 0 when the device normalised them;
            # with the peer path above the phase matrices and pback have already arrived
            dup = {'p22': 'p11', 'p44': 'p33'} if (plan.device_psd and plan.nri == 1 and keep_phase) else {}     # same on every rank
            if peer_phase:
                dup = {k: None for k in scatkeys + ['pback']}
            send = [k for k in keys if k not in dup]
            rows = np.zeros((ncell, 2 + sum(width[k] for k in send)))
            if ncell:
                rows[:, 0], rows[:, 1] = li, ri
                o = 2
                for k in send:
                    rows[:, o:o + width[k]] = np.asarray(ret[k]).reshape(ncell, width[k])
                    o += width[k]
            rows = comm.gather_rows(rows)          # over NVLink (peer puts or NCCL gather); gloo on CPU
            if rank != 0:
                continue
            li, ri = rows[:, 0].astype(np.int64), rows[:, 1].astype(np.int64)
            ret, o = {}, 2
            for k in send:
                ret[k] = rows[:, o:o + width[k]] if width[k] > 1 else rows[:, o]
                o += width[k]
            for k, src in dup.items():
                if src is not None:
                    ret[k] = ret[src]
            # the rows arrive rank by rank, each rank's in (wavelength, rh) order of its wavelengths r, r + W, ...: one strided block
            # assignment per rank and variable instead of an element-wise scatter
            nrr_ = 1 if trivial else nr
            counts = [len(range(r, nl, world)) * nrr_ for r in range(world)]
            if sum(counts) == rows.shape[0]:
                start = 0
                for r in range(world):
                    blk = slice(start, start + counts[r])
                    start += counts[r]
                    if counts[r] == 0:
                        continue
                    for key in keys:
                        if key not in ret or _kind_of(key) == "nl":
                            continue
                        col = np.asarray(ret[key])[blk]
                        vals[key][radind, r::world, :nrr_] = col.reshape((-1, nrr_) + col.shape[1:])
                block_done = True
            else:
                block_done = False
        if ret is not None:
            for key in keys:
                if key not in ret:
                    continue                                   # arrived through the peer path
                col = np.asarray(ret[key])
                if world > 1 and block_done and _kind_of(key) != "nl":
                    continue                                   # written rank block by rank block above
                if _kind_of(key) == "nl":
                    # (bin, rh) variables are overwritten at every wavelength: the last one wins (dointegration.py:1026-1027)
                    last = li == li.max()
                    vals[key][radind, ri[last]] = col[last]
                elif in_place and key in ('p11', 'p12', 'p33', 'p34'):
                    continue                                   # written by the device-to-host copy itself
                elif in_place:
                    vals[key][radind] = col.reshape(vals[key][radind].shape)      # cells are in (wavelength, rh) order
                else:
                    vals[key][radind, li, ri] = col
        if trivial:
            # copyDryValues (dointegration.py:465-491)
            for key in vals:
                if _kind_of(key) == "nl":
                    vals[key][radind, 1:] = vals[key][radind, 0]
                else:
                    vals[key][radind, :, 1:] = vals[key][radind, :, :1]
    if rank != 0:
        return None
    out = {'rh': np.asarray(rh, dtype=float), 'wavelength': np.asarray(lambarr), 'ang': ang, 'vals': vals,
           'radius': radiusarr}
    if write:
        write_table(partID, oppfx, out, oppclassic)
    return out


def write_table(ncdfID, oppfx, out, oppclassic, hydrophobic_bin=False):
    """Write the arrays of `fun` with createNCDF's layout (new: (bin, wavelength, rh[, ang|p]); legacy:
    (radius, rh, lambda[, ang]) and pback (nPol, radius, rh, lambda), dointegration.py:356-363, :1006-1031).
    hydrophobic_bin=True applies hydrophobic.doConversion's rule in memory (a new first bin holding the RH-index-0
    values at every RH, hydrophobic.py:38-113) instead of the reference's write / rename / re-read / re-write cycle."""
    from . import hydrophobic
    radius = list(out['radius'])
    if hydrophobic_bin:
        radius = [radius[0], radius[0]]
    nc = createNCDF(ncdfID, oppfx, radius, out['rh'], out['wavelength'], out['ang'], oppclassic)
    rname = 'radius' if oppclassic else 'bin'
    for key, a in out['vals'].items():
        kind = _kind_of(key)
        if oppclassic:
            if kind == "scat":
                a = a.transpose(0, 2, 1, 3)
            elif kind == "ele":
                a = a.transpose(3, 0, 2, 1)
            elif kind == "scal":
                a = a.transpose(0, 2, 1)
        if hydrophobic_bin:
            a = hydrophobic._convert_var(key, a, tuple(nc.variables[key].dimensions), rname)
        nc.variables[key][:] = a
    nc.close()
