"""NetCDF I/O shim with the subset of the netCDF4-python API the GEOSmie drivers use.

If the `netCDF4` package is importable it is used unchanged (real NetCDF-4/HDF5 files with zlib, i8 variables --
byte-layout identical to the reference).  This image has neither netCDF4 nor HDF5, so the fallback keeps the dataset in
memory and (de)serialises it with scipy.io.netcdf_file (NetCDF-3, 64-bit offset): same dimensions, variables,
attributes and creation order; differences forced by the classic format: no compression, i8 variables stored as i4.
"""
import os

import numpy as np

try:  # pragma: no cover - not available in this image
    import netCDF4 as _nc4
    HAVE_NETCDF4 = True
except Exception:  # ModuleNotFoundError or a broken HDF5
    _nc4 = None
    HAVE_NETCDF4 = False


class _Dim(object):
    def __init__(self, name, size):
        self.name, self.size = name, size

    def __len__(self):
        return self.size


class _Var(object):
    def __init__(self, name, dtype, dims, shape):
        object.__setattr__(self, "_attrs", {})
        object.__setattr__(self, "name", name)
        object.__setattr__(self, "dimensions", tuple(dims))
        dt = np.dtype("S1") if dtype in ("c", "S1") else np.dtype(dtype)
        object.__setattr__(self, "dtype", dt)
        object.__setattr__(self, "data", np.zeros(shape, dtype=dt))

    shape = property(lambda self: self.data.shape)

    def __getitem__(self, k):
        return self.data[k]

    def __setitem__(self, k, v):
        if self.data.dtype.kind == "S" and isinstance(v, str):
            v = v.encode()
        self.data[k] = v

    def __len__(self):
        return len(self.data)

    def __iter__(self):
        return iter(self.data)

    def __setattr__(self, k, v):
        self._attrs[k] = v

    def __getattr__(self, k):
        try:
            return object.__getattribute__(self, "_attrs")[k]
        except KeyError:
            raise AttributeError(k)

    def ncattrs(self):
        return list(self._attrs.keys())


class _MemDataset(object):
    def __init__(self, path, mode="r", **kw):
        object.__setattr__(self, "_path", path)
        object.__setattr__(self, "_mode", mode)
        object.__setattr__(self, "dimensions", {})
        object.__setattr__(self, "variables", {})
        object.__setattr__(self, "_gattrs", {})
        if mode in ("r", "r+", "a"):
            self._load()

    # -- netCDF4-like API
    def createDimension(self, name, size):
        self.dimensions[name] = _Dim(name, size)
        return self.dimensions[name]

    def createVariable(self, name, dtype, dims=(), **kw):
        if isinstance(dims, str):
            dims = (dims,)
        shape = tuple(len(self.dimensions[d]) for d in dims)
        v = _Var(name, dtype, dims, shape)
        self.variables[name] = v
        return v

    def ncattrs(self):
        return list(self._gattrs.keys())

    def __setattr__(self, k, v):
        self._gattrs[k] = v

    def __getattr__(self, k):
        try:
            return object.__getattribute__(self, "_gattrs")[k]
        except KeyError:
            raise AttributeError(k)

    def close(self):
        if self._mode in ("w", "r+", "a"):
            self._store()

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    # -- classic-format (de)serialisation
    def _store(self):
        from scipy.io import netcdf_file
        f = netcdf_file(self._path, "w", version=2)
        for k, v in self._gattrs.items():
            setattr(f, k, v)
        for name, d in self.dimensions.items():
            f.createDimension(name, len(d))
        for name, v in self.variables.items():
            dt = v.dtype
            if dt == np.dtype("i8"):
                dt = np.dtype("i4")
            code = "c" if dt.kind == "S" else dt
            fv = f.createVariable(name, code, v.dimensions)
            if v.data.ndim == 0:
                fv.assignValue(v.data.astype(dt))
            else:
                fv[:] = v.data.astype(dt) if dt.kind != "S" else v.data
            for ak, av in v._attrs.items():
                setattr(fv, ak, av)
        f.close()

    def _load(self):
        from scipy.io import netcdf_file
        if not os.path.exists(self._path):
            raise FileNotFoundError(self._path)
        f = netcdf_file(self._path, "r", mmap=False)
        for k, v in f._attributes.items():
            self._gattrs[k] = v.decode() if isinstance(v, bytes) else v
        for name, size in f.dimensions.items():
            self.dimensions[name] = _Dim(name, size)
        for name, fv in f.variables.items():
            arr = np.array(fv.data)
            dt = arr.dtype.newbyteorder("=")
            v = _Var(name, dt if dt.kind != "S" else "c", fv.dimensions, arr.shape)
            v.data[...] = arr
            for ak, av in fv._attributes.items():
                v._attrs[ak] = av.decode() if isinstance(av, bytes) else av
            self.variables[name] = v
        f.close()


def Dataset(path, mode="r", **kw):
    """netCDF4.Dataset when available, the in-memory classic-format shim otherwise."""
    if HAVE_NETCDF4:
        return _nc4.Dataset(path, mode, **kw)
    return _MemDataset(path, mode, **kw)
