"""optics_X.nomom.nc4 -> optics_X.nc4 with the `pmom` GSF moments (mirror of src/gsf/convertncdf.py, mode 'pygeos').

The reference dumps every (wavelength, RH, bin) cell to a text file, spawns ./spher_expan.x and parses its output
(convertncdf.py:173-189, :369-394).  Here all cells go to the GPU in one gm_gsf_expand call; the column -> `p` index
mapping [0,4,2,5,1,3] and the file layout are the reference's.  Modes 'legendre', 'ice' and 'physical' (foreign input
formats) are outside the hot path.
"""
import os
import shutil

import numpy as np

from .. import _lib, ncio

NUM_EXPAND = 129                      # numExpand, convertncdf.py:323 == NSPHER, params.h:14
MISH_KEYS = ['p11', 'p22', 'p33', 'p44', 'p12', 'p34']   # order Mishchenko's code expects (:177)
NPOL_OF_COLUMN = [0, 4, 2, 5, 1, 3]   # expansion column ii -> index along p / nPol (:381)
PMOM_LONG_NAME = ('Moments of the generalized spherical functions order over nPol as the 11, 12, 33, 34, 22, 44 '
                  'elements of the scattering law matrix for a Fourier decomposition of the phase matrix such as those '
                  'used in discrete ordinate solutions to the RTE')


def createVariablesPyGeosMie(ncdf, numExpand, oppclassic):
    """`pmom` variable and its moment dimension (convertncdf.py:70-79)."""
    if oppclassic:
        ncdf.createDimension('nMom', numExpand)
        ncdf.createVariable('pmom', 'f8', ('nPol', 'nMom', 'radius', 'rh', 'lambda'), compression='zlib')
    else:
        ncdf.createDimension('m', numExpand)
        ncdf.createVariable('pmom', 'f8', ('bin', 'wavelength', 'rh', 'p', 'm'), compression='zlib')
    ncdf.variables['pmom'].long_name = PMOM_LONG_NAME


def expand_table(ang, elements, oppclassic, quantize=True, handle=None):
    """elements: {key: array} for the six MISH_KEYS in file layout -> pmom array in file layout.
    quantize=True rounds to 10 decimals like the Fortran F17.10 text round trip (spher_expan.f:96,104)."""
    first = elements[MISH_KEYS[0]]
    if oppclassic:                       # (radius, rh, lambda, ang) -> (radius, lambda, rh, ang)
        nb, nr, nl, na = first.shape
    else:
        nb, nl, nr, na = first.shape
    # the six elements of a cell side by side: every file variable is decoded straight into its column of F (one pass each)
    F = np.empty((nb, nl, nr, 6, na))
    for k, key in enumerate(MISH_KEYS):
        dst = F[:, :, :, k, :]
        if oppclassic:
            dst = dst.transpose(0, 2, 1, 3)
        v = elements[key]
        if hasattr(v, "read_into"):
            v.read_into(dst)
        else:
            np.copyto(dst, np.asarray(v[:], dtype=float))
    h = handle or _lib.Handle.get()
    coef, _ = h.gsf_expand(ang, F.reshape(nb * nl * nr, 6, na), NUM_EXPAND, quantize10=quantize)
    inv = np.argsort(NPOL_OF_COLUMN)     # file order (nPol) <- Mishchenko's column order
    pm = np.take(coef.reshape(nb, nl, nr, 6, NUM_EXPAND), inv, axis=3)
    if oppclassic:                       # ('nPol','nMom','radius','rh','lambda')
        pm = pm.transpose(3, 4, 0, 2, 1)
    return pm


NUM_GAUSS_LEGENDRE = 960              # num_gauss, convertncdf.py:323 (spher_expan.x accepts fewer than 1000 input angles)


def hbleg(x, pmoms):
    """sum_l pmoms[..., l] P_l(x) with the reference's recurrence (convertncdf.py:31-43), vectorised: pmoms [..., nmom], x [nx]
    -> [..., nx].  (The stored Legendre moments carry their own normalisation: no (2l+1) factor is applied.)"""
    x = np.asarray(x, dtype=float)
    nmom = pmoms.shape[-1]
    leg = np.zeros((nmom, x.size))
    leg[0] = 1.
    if nmom > 1:
        leg[1] = x
    for imom in range(2, nmom):
        leg[imom] = 1. / imom * ((2. * imom - 1.) * x * leg[imom - 1] - (imom - 1.) * leg[imom - 2])
    return np.asarray(pmoms, dtype=float) @ leg


def process_legendre(nc, quantize=True, handle=None):
    """rungsf mode 'legendre' (convertncdf.py:190-219, :46-67, :377-394): a legacy table whose `pmom (nPol, nMom, radius, rh, lambda)`
    holds LEGENDRE moments of the six elements.  Every cell's series is evaluated at the 960 Gauss angles (the input the expansion
    program gets) and at 181 linear angles, all cells are expanded in one GPU call, and the file receives
        pmom2         the original Legendre moments,
        pmom          the generalized-spherical-function moments, zero-padded to nMom,
        phase_matrix  (nPol, scattering_angle, radius, rh, lambda) and scattering_angle = 0..180 degrees.
    The reference indexes pmom2 / phase_matrix with the new-style order although it creates them with the legacy dimensions
    (:379, :392) and stops with an index error on its own files; the variables are written here as they are dimensioned."""
    from numpy.polynomial.legendre import leggauss
    pm = np.array(nc.variables['pmom'][:])                    # (nPol, nMom, radius, rh, lambda)
    npol, nmom, nrad, nrh, nlam = pm.shape
    linangs = np.linspace(0, 180, 181)
    gpoints, _ = leggauss(NUM_GAUSS_LEGENDRE)
    gpoints = gpoints[::-1]
    gdeg = np.degrees(np.arccos(gpoints))
    moms = pm.transpose(2, 3, 4, 0, 1)                        # (radius, rh, lambda, nPol, nMom)
    order = NPOL_OF_COLUMN                                    # Mishchenko's column ii <- nPol index (convertncdf.py:201)
    F = hbleg(gpoints, moms[..., order, :]).reshape(nrad * nrh * nlam, 6, NUM_GAUSS_LEGENDRE)
    lin = hbleg(np.cos(np.radians(linangs)), moms)            # (radius, rh, lambda, nPol, 181), nPol order as stored
    h = handle or _lib.Handle.get()
    coef, _ = h.gsf_expand(gdeg, F, NUM_EXPAND, quantize10=quantize)
    coef = coef.reshape(nrad, nrh, nlam, 6, NUM_EXPAND)
    new = np.zeros((npol, nmom, nrad, nrh, nlam))
    nkeep = min(nmom, NUM_EXPAND)
    for ii, p_ in enumerate(order):
        new[p_, :nkeep] = coef[..., ii, :nkeep].transpose(3, 0, 1, 2)
    nc.createVariable('pmom2', 'f8', ('nPol', 'nMom', 'radius', 'rh', 'lambda'))
    nc.createDimension('nMom-GSF', NUM_EXPAND)
    nc.variables['pmom2'].long_name = getattr(nc.variables['pmom'], 'long_name', '')
    nc.variables['pmom2'][:] = pm
    nc.variables['pmom'].long_name = PMOM_LONG_NAME
    nc.variables['pmom'][:] = new
    nc.createDimension('scattering_angle', linangs.size)
    nc.createVariable('scattering_angle', 'f8', ('scattering_angle'))
    nc.variables['scattering_angle'][:] = linangs
    nc.variables['scattering_angle'].long_name = 'Scattering angle in degrees'
    nc.createVariable('phase_matrix', 'f8', ('nPol', 'scattering_angle', 'radius', 'rh', 'lambda'))
    nc.variables['phase_matrix'].long_name = 'Phase matrix elements, ordered over nPol as P11, P12, P33, P34, P22, P44'
    nc.variables['phase_matrix'][:] = lin.transpose(3, 4, 0, 1, 2)
    return new


def processFileRaw(infile, outdir, whichproc, rhop0, mode, ice, quantize=True):
    """Copy <name>.nomom.nc4 to <name>.nc4 and add pmom (convertncdf.py:319-401)."""
    if mode not in ('pygeos', 'legendre') or ice:
        raise NotImplementedError("rungsf mode %r reads a foreign table format (ice crystals / GRASP kernels) that is outside the Mie "
                                  "hot path; 'pygeos' and 'legendre' are available" % ('ice' if ice else mode))
    if mode == 'legendre':
        fn = os.path.basename(infile)
        outfile = os.path.join(outdir, fn.replace('nomom.', ''))
        nc = ncio.open_copy(infile, outfile)
        print('mode %s' % mode)
        process_legendre(nc, quantize=quantize)
        nc.close()
        print("%s done" % fn)
        return outfile
    fn = os.path.basename(infile)
    outfile = os.path.join(outdir, fn.replace('nomom.', ''))
    nc = ncio.open_copy(infile, outfile)      # copy of the .nomom file that gains `pmom` (convertncdf.py:331-332)
    oppclassic = 'wavelength' not in nc.variables
    if oppclassic:
        print("Operating on a legacy file")
    print('mode %s' % mode)
    createVariablesPyGeosMie(nc, NUM_EXPAND, oppclassic)
    ang = np.array(nc.variables['ang'][:])
    elements = {k: nc.variables[k] for k in MISH_KEYS}       # decoded lazily, straight into expand_table's input array
    nc.variables['pmom'][:] = expand_table(ang, elements, oppclassic, quantize=quantize)
    nc.close()
    print("%s done" % fn)
    return outfile


def convertFile(filepath, opdir, mode, rhop0):
    """Entry point used by rungsf.py (convertncdf.py:492-506)."""
    if mode in ['pygeos', 'legendre']:
        rhop0 = None
    ice = False
    if mode == 'ice':
        mode, ice = 'legendre', True
    return processFileRaw(filepath, opdir, 0, rhop0, mode, ice)
