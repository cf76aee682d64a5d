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
    a = [np.asarray(elements[k], dtype=float) for k in MISH_KEYS]
    if oppclassic:                       # (radius, rh, lambda, ang) -> (radius, lambda, rh, ang)
        a = [v.transpose(0, 2, 1, 3) for v in a]
    nb, nl, nr, na = a[0].shape
    F = np.stack(a, axis=3).reshape(nb * nl * nr, 6, na)
    h = handle or _lib.Handle.get()
    coef, _ = h.gsf_expand(ang, F, NUM_EXPAND, quantize10=quantize)
    pm = np.zeros((nb, nl, nr, 6, NUM_EXPAND))
    coef = coef.reshape(nb, nl, nr, 6, NUM_EXPAND)
    for ii, npol in enumerate(NPOL_OF_COLUMN):
        pm[:, :, :, npol, :] = coef[:, :, :, ii, :]
    if oppclassic:                       # ('nPol','nMom','radius','rh','lambda')
        pm = pm.transpose(3, 4, 0, 2, 1)
    return pm


def processFileRaw(infile, outdir, whichproc, rhop0, mode, ice, quantize=True):
    """Copy <name>.nomom.nc4 to <name>.nc4 and add pmom (convertncdf.py:319-401)."""
    if mode != 'pygeos':
        raise NotImplementedError("rungsf mode %r reads a foreign table format; only 'pygeos' is on the Mie hot path" % mode)
    fn = os.path.basename(infile)
    outfile = os.path.join(outdir, fn.replace('nomom.', ''))
    shutil.copyfile(infile, outfile)
    nc = ncio.Dataset(outfile, 'r+')
    oppclassic = 'wavelength' not in nc.variables
    if oppclassic:
        print("Operating on a legacy file")
    print('mode %s' % mode)
    createVariablesPyGeosMie(nc, NUM_EXPAND, oppclassic)
    ang = np.array(nc.variables['ang'][:])
    elements = {k: np.array(nc.variables[k][:]) for k in MISH_KEYS}
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
