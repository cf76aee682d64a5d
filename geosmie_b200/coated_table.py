"""Coated (core + shell) table build -- an EXTENSION beyond the reference, which has coated_mie_coeff for single particles
(mie_coeffs.py:183-251) but no coated table driver (MultipleMie's yArr branch calls an undefined function, mie_coated.py:80).

BASELINE config 4 / SURVEY 8d: black-carbon cores on the dry bc.json grid (core size parameter x, index m_soot(lambda))
that take up a water shell with relative humidity: shell size parameter y = gf(RH) * x with the species' growth-factor
list, shell index m_water(lambda).  Number weights are those of the dry lognormal distribution (every dry particle keeps
its number and grows); areas/volumes/efficiencies refer to the coated particle.  RH levels with gf == 1 are homogeneous
spheres of the core material (mie_coeffs.py:66 dispatch: x == y).
"""
import numpy as np

from . import _lib
from . import dointegration as DI
from .pymiecoated.mie_coeffs import nmax_of


def coated_cells(params, lambarr, part_m, water_m, radind=0, cells=None):
    """Per-cell inputs: (li, rhi), core index, shell index, growth factor, dry weights, dry mass-effective radius."""
    xx, dr = DI.initializeXarr(params, radind, lambarr[0], lambarr[-1])
    rh = params['rh']
    gfs = params['rhDep']['params']['gf']
    out = []
    want = None if cells is None else set(cells)
    for li, lam in enumerate(lambarr):
        if want is not None and not any(c[0] == li for c in want):
            continue
        m1 = complex(part_m[0][0](lam), -part_m[0][1](lam))
        m2 = complex(water_m[0](lam), water_m[1](lam))
        psd0, ref0, rLow, rUp = DI.calculatePSD(params, radind, rh[0], rh, xx, dr, 1.0, lam)
        for rhi in range(len(rh)):
            if want is not None and (li, rhi) not in want:
                continue
            out.append(dict(cell=(li, rhi), lam=lam, m1=m1, m2=m2, gf=float(gfs[rhi]), w=psd0[0], reff0=ref0[0]))
    return xx, out


def build(params, lambarr, part_m, water_m, radind=0, cells=None, elide=True, handle=None):
    """Evaluate the coated table cells on the GPU.  Returns (cells, ret) with ret = integratePSD-style dict (leading cell
    axis) after the a-posteriori normalisation of dointegration.fun."""
    xx, cl = coated_cells(params, lambarr, part_m, water_m, radind, cells)
    ang = DI.table_angles()
    cost = np.cos(np.radians(ang))
    rhop0 = params['rhop0'][radind] if isinstance(params['rhop0'], list) else params['rhop0']
    ncell = len(cl)
    scal = np.zeros((ncell, 1, _lib.GM_NSCAL))
    phase = np.zeros((ncell, 4, ang.size))
    for gf in sorted(set(c['gf'] for c in cl)):
        idx = [i for i, c in enumerate(cl) if c['gf'] == gf]
        y = gf * xx                                   # shell size parameter grid of this RH level
        t = _lib.Table(y, nmax_of(y), cost, handle)
        m1 = np.array([cl[i]['m1'] for i in idx])
        w = np.array([cl[i]['w'] for i in idx])
        if gf == 1.0:
            mz = np.sqrt(m1 ** 2 * 1.0)
            s, p = t.run(mz, mz, w, elide=elide)
        else:
            m2 = np.array([cl[i]['m2'] for i in idx])
            s, p = t.run_coated(m1, m2, 1.0 / gf, w, elide=elide)
        scal[idx], phase[idx] = s, p
        t.close()
    lam = np.array([c['lam'] for c in cl])
    gfa = np.array([c['gf'] for c in cl])
    rrat = 1.0 / gfa
    rhop = rrat ** 3. * rhop0 + (1. - rrat ** 3.) * 1000.
    reff0 = np.array([[c['reff0']] for c in cl])
    ret = DI.combine_modes(scal, phase, [1.0], lam, reff0, rhop0, rhop)
    ret = DI.postprocess(ret, ang)
    return [c['cell'] for c in cl], ret, (scal, phase)
