"""Band averaging of an optics table (mirror of src/geosmie/bandaverage.py) with the per-column averaging on the GPU.

getBands keeps the reference's band tables (GEOS5 / RRTMG / RRTMGP / PURDUE); doAverage's work for ALL
(variable, bin, RH) columns is one gm_band_average call instead of a Python quadruple loop.
"""
import os

import numpy as np

from . import _lib, ncio

varsToAverage = ['qsca', 'qext', 'bsca', 'bext', 'g', 'bbck', 'refreal', 'refimag']   # bandaverage.py:155


def getBands(mode):
    """Band edges (bandaverage.py:73-124): wavenumbers [cm-1] for RRTMG/RRTMGP, wavelengths [m] for GEOS5/PURDUE."""
    useWavenum = True
    numBandsMod = 0
    if mode == 'GEOS5':
        useWavenum = False
        lo = np.array([.175, .225, .285, .300, .325, .400, .690, 1.220, 2.270, 29.412, 18.519, 12.5, 10.204, 9.091, 8.230,
                       7.246, 5.263, 3.333, 16.129]) * 1.e-6
        up = np.array([.225, .285, .300, .325, .400, .690, 1.220, 2.270, 3.850, 40., 29.412, 18.519, 12.5, 10.204, 9.091,
                       8.230, 7.246, 5.263, 18.519]) * 1.e-6
        numBandsMod = -1
    elif mode == 'RRTMG':
        sw_l = [2600., 3250., 4000., 4650., 5150., 6150., 7700., 8050., 12850., 16000., 22650., 29000., 38000., 820.]
        sw_r = [3250., 4000., 4650., 5150., 6150., 7700., 8050., 12850., 16000., 22650., 29000., 38000., 50000., 2600.]
        lw_l = [10., 350., 500., 630., 700., 820., 980., 1080., 1180., 1390., 1480., 1800., 2080., 2250., 2380., 2600.]
        lw_r = [350., 500., 630., 700., 820., 980., 1080., 1180., 1390., 1480., 1800., 2080., 2250., 2380., 2600., 3250.]
        lo, up = sw_l + lw_l, sw_r + lw_r
    elif mode == 'RRTMGP':
        sw_l = [820., 2680., 3250., 4000., 4650., 5150., 6150., 7700., 8050., 12850., 16000., 22650., 29000., 38000.]
        sw_r = [2680., 3250., 4000., 4650., 5150., 6150., 7700., 8050., 12850., 16000., 22650., 29000., 38000., 50000.]
        lw_l = [10., 250., 500., 630., 700., 820., 980., 1080., 1180., 1390., 1480., 1800., 2080., 2250., 2390., 2680.]
        lw_r = [250., 500., 630., 700., 820., 980., 1080., 1180., 1390., 1480., 1800., 2080., 2250., 2390., 2680., 3250.]
        lo, up = sw_l + lw_l, sw_r + lw_r
    elif mode == 'PURDUE':
        useWavenum = False
        lo = np.array([.175, .225, .245, .280, .295, .310, .325, .400, .700, 1.220, 2.270, 3.33, 5.26, 7.25, 8.23, 9.09,
                       10.2, 12.5, 16.13, 18.52, 29.41]) * 1.e-6
        up = np.array([.225, .280, .260, .295, .310, .320, .400, .700, 1.220, 2.270, 10.00, 5.26, 7.25, 8.23, 9.09, 10.2,
                       12.5, 18.52, 18.52, 29.41, 40.0]) * 1.e-6
    else:
        raise ValueError("unknown band mode %r" % mode)
    bandMeanM = [(x + y) / 2. for x, y in zip(lo, up)]
    if useWavenum:
        bandMeanM = [(x * 100.) ** (-1) for x in bandMeanM]
    return lo, up, bandMeanM, useWavenum, len(lo) + numBandsMod


def doAverage(lam, varIn, bandl, bandr, useWavenum, solardata):
    """One band of one column (bandaverage.py:18-50) -- evaluated on the GPU; `solardata` is unused as in the reference."""
    out = _lib.Handle.get().band_average(lam, np.asarray(varIn, dtype=float)[None, :], [bandl], [bandr], useWavenum)
    return float(out[0, 0])


def average_columns(lam, cols, mode):
    """cols [ncol][nlam] -> [ncol][nbands_orig] for band mode `mode` in one GPU call."""
    lo, up, _, useWavenum, _ = getBands(mode)
    return _lib.Handle.get().band_average(lam, cols, np.asarray(lo, dtype=float), np.asarray(up, dtype=float), useWavenum)


def fun(data, part, opfn, mode, useSolar, noIR):
    """Band-average the 8 scalar fields of an open table and write opticsBands_*.nc4 (bandaverage.py:126-294)."""
    try:
        lam = np.array(data.variables['wavelength'][:])
        oppclassic, radiusNm, lamNm = False, 'bin', 'wavelength'
    except Exception:
        lamNm = 'lambda'
        lam = np.array(data.variables[lamNm][:])
        oppclassic, radiusNm = True, 'radius'
        print("Operating on a legacy file")
    rh = np.array(data.variables['rh'][:])
    reff = np.array(data.variables['rEff'][:])
    radius = np.array(data.variables[radiusNm][:])
    lBandLow, lBandUp, bandMeanM, useWavenum, nbands = getBands(mode)
    nbo, nrh, nbin = len(lBandLow), len(rh), len(reff[:, 0])
    output = {}
    for varName in varsToAverage:
        a = np.array(data.variables[varName][:])
        a = a if oppclassic else a.transpose(0, 2, 1)        # -> (bin, rh, lambda)
        avg = average_columns(lam, a.reshape(nbin * nrh, len(lam)), mode).reshape(nbin, nrh, nbo)
        output[varName] = avg if oppclassic else avg.transpose(0, 2, 1)
    if noIR and mode == 'GEOS5':
        ind = np.where(np.asarray(lBandLow) > 3e-6)[0]
        for k, v in (('qsca', 0), ('bsca', 0), ('qext', 1e-32), ('bext', 1e-32)):
            if oppclassic:
                output[k][:, :, ind] = v
            else:
                output[k][:, ind, :] = v
    if mode == 'GEOS5':
        # bands 0 and 2 are merged (wavelength-width weighted) and band 0 is dropped (bandaverage.py:203-220)
        bw1, bw2 = lBandUp[0] - lBandLow[0], lBandUp[2] - lBandLow[2]
        for key, val in output.items():
            if oppclassic:
                val[:, :, 2] = (bw1 * val[:, :, 0] + bw2 * val[:, :, 2]) / (bw1 + bw2)
                output[key] = val[:, :, 1:]
            else:
                val[:, 2, :] = (bw1 * val[:, 0, :] + bw2 * val[:, 2, :]) / (bw1 + bw2)
                output[key] = val[:, 1:, :]
        # The reference drops the first band centre INSIDE the loop over the eight variables (bandaverage.py:220), leaves 11 of
        # the 19 values for an 18-long wavelength variable and stops with a broadcast error when it writes the file: its GEOS5 mode
        # cannot finish.  The evident intent -- the centre of the dropped band 0 goes once -- is what is done here.
        bandMeanM = bandMeanM[1:]
    nc = ncio.Dataset(opfn, 'w')
    nc.createDimension('rh', nrh)
    nc.createDimension(lamNm, nbands)
    nc.createDimension(radiusNm, nbin)
    nc.createDimension('nchar', 80)
    for varName in output.keys():
        dims = (radiusNm, 'rh', lamNm) if oppclassic else (radiusNm, lamNm, 'rh')
        nc.createVariable(varName, 'f8', dims, compression='zlib')
        nc.variables[varName][:] = output[varName]
    nc.createVariable(lamNm, 'f8', (lamNm))
    nc.variables[lamNm][:] = bandMeanM
    nc.createVariable('rh', 'f8', ('rh'))
    nc.variables['rh'][:] = rh
    nc.createVariable('qname', 'c', (radiusNm, 'nchar'))
    qname = ['%s%03d' % (part, i) for i in range(1, len(radius) + 1)]
    if mode == 'RRTMG':
        lo = np.array(lBandLow) ** (-1) * 0.01
        up = np.array(lBandUp) ** (-1) * 0.01
        lBandLow, lBandUp = up[::-1], lo[::-1]
    lBandLow = np.asarray(lBandLow) * 1e6
    lBandUp = np.asarray(lBandUp) * 1e6
    nc.createVariable('bandLow', 'f8', (lamNm))
    nc.variables['bandLow'][:] = lBandLow[:nbands] if len(lBandLow) != nbands else lBandLow
    nc.variables['bandLow'].long_name = 'Lower edges of the bands'
    nc.variables['bandLow'].units = 'micrometers'
    nc.createVariable('bandUp', 'f8', (lamNm))
    # the reference writes lBandLow into bandUp as well (bandaverage.py:268) -- reproduced so files compare equal
    nc.variables['bandUp'][:] = lBandLow[:nbands] if len(lBandLow) != nbands else lBandLow
    nc.variables['bandUp'].long_name = 'Upper edges of the bands'
    nc.variables['bandUp'].units = 'micrometers'
    for qi, qq in enumerate(qname):
        for ci, char in enumerate(qq):
            nc.variables['qname'][qi, ci] = char
    nc.createVariable(radiusNm, 'f8', (radiusNm))
    nc.variables[radiusNm][:] = radius
    for var in ['rLow', 'rUp']:
        nc.createVariable(var, 'f8', (radiusNm))
        nc.variables[var][:] = np.array(data.variables[var][:])[:, 0]
    for var in ['rEff', 'rMass']:
        nc.createVariable(var, 'f8', (radiusNm, 'rh'))
        nc.variables[var][:] = np.array(data.variables[var][:])
    nc.close()
    return output


def processFileForBandMode(filepath, partcode, opdir, bandmode, useSolar, noIR):
    """optics_X.nc4 -> opticsBands_X.<MODE>.nc4 (bandaverage.py:327-338)."""
    data = ncio.Dataset(filepath, 'r')
    fn = os.path.basename(filepath)
    opfn = fn.replace('.nc4', '.%s.nc4' % bandmode).replace('optics_', 'opticsBands_')
    return fun(data, partcode, os.path.join(opdir, opfn), bandmode, useSolar, noIR)
