"""Hydrophobic-bin post-pass (mirror of src/geosmie/hydrophobic.py:18-124).

For species with `"hydrophobic": true` the table gets a new first bin that holds the RH-index-0 values of the
(hydrophilic) computed bin at every RH; the computed bin becomes index 1.  Works on files (like the reference, so
runoptics.py keeps its rename / convert / remove sequence) and on in-memory arrays (`convert_arrays`).
"""
import os

import numpy as np

from . import ncio


def _convert_var(name, data, dims, radius_name):
    """New-array for one variable: the reference's per-rank branches (hydrophobic.py:58-113) expressed as one rule --
    the radius axis grows to 2, entry 0 takes the RH-index-0 slice broadcast over RH, entry 1 the original bin 0."""
    if radius_name not in dims:
        return np.array(data)
    ax = dims.index(radius_name)
    if data.shape[ax] == 2:
        return np.array(data)                                # shape unchanged: copied through as it is (hydrophobic.py:58, :121-122)
    if data.ndim == 1:
        return np.array([1, 2], dtype=data.dtype)            # the 1-indexed bin coordinate (:61-63)
    src = np.take(data, 0, axis=ax)                         # only bin 0 of the input is used, as in the reference
    rdims = [d for d in dims if d != radius_name]
    rax = rdims.index('rh')
    dry = np.take(src, [0], axis=rax)
    phobic = np.broadcast_to(dry, src.shape)
    return np.stack([phobic, src], axis=ax)


def convert_arrays(variables, oppclassic):
    """variables: {name: (data, dims)} -> same mapping with the hydrophobic bin prepended."""
    radius_name = 'radius' if oppclassic else 'bin'
    return {k: (_convert_var(k, d, dims, radius_name), dims) for k, (d, dims) in variables.items()}


def doConversion(infn, outfn, pfx, oppclassic):
    """File-to-file conversion with the reference's signature (hydrophobic.py:18)."""
    f = ncio.Dataset(os.path.join(pfx, infn), 'r')
    g = ncio.Dataset(os.path.join(pfx, outfn), 'w')
    for att in f.ncattrs():
        setattr(g, att, getattr(f, att))
    radius_name = 'radius' if oppclassic else 'bin'
    for dimname, dim in list(f.dimensions.items()):
        g.createDimension(dimname, 2 if dimname == radius_name else len(dim))
    for varname, ncvar in list(f.variables.items()):
        var = g.createVariable(varname, ncvar.dtype, ncvar.dimensions, compression='zlib')
        for att in ncvar.ncattrs():
            setattr(var, att, getattr(ncvar, att))
        var[:] = _convert_var(varname, np.array(ncvar[:]), tuple(ncvar.dimensions), radius_name)
    f.close()
    g.close()
