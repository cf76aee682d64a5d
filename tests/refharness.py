"""Harness that imports the UNMODIFIED reference (GEOS-ESM/GEOSmie at /root/reference) in this container.

Test infrastructure only.  It is used (a) by tests/golden/make_golden.py to generate the committed golden
fixtures and (b) by "live" parity tests that are skipped when /root/reference is absent (e.g. on the GPU
box, which only receives /root/repo).

The reference imports `netCDF4` (src/geosmie/dointegration.py:1, hydrophobic.py:18, bandaverage.py:4) and,
through carma_utils.py:6, `matplotlib`; neither is installed in this image, so two stub modules are
injected into sys.modules: an in-memory `netCDF4.Dataset` with a path-keyed registry (so that the
'w' -> 'r+' -> 'r' re-opens of hydrophobic.doConversion see earlier writes) and an empty matplotlib.
"""
import contextlib
import os
import sys
import tempfile
import types

import numpy as np

REF = os.environ.get("GEOSMIE_REFERENCE", "/root/reference")


def have_reference():
    return os.path.isdir(os.path.join(REF, "src", "pymiecoated", "pymiecoated"))


# ----------------------------------------------------------------------------- in-memory netCDF4 stub
class _Dim:
    def __init__(self, name, size):
        self.name, self.size = name, size

    def __len__(self):
        return self.size


class _Var:
    def __init__(self, name, dtype, dims, shape):
        object.__setattr__(self, "_attrs", {})
        object.__setattr__(self, "name", name)
        object.__setattr__(self, "dimensions", dims)
        dt = np.dtype("S1") if dtype == "c" else np.dtype(dtype)
        object.__setattr__(self, "dtype", dt)
        object.__setattr__(self, "data", np.zeros(shape, dtype=dt))

    @property
    def shape(self):
        return self.data.shape

    def __getitem__(self, k):
        return self.data[k]

    def __setitem__(self, k, v):
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


class _Store:
    def __init__(self):
        self.dimensions, self.variables, self.attrs = {}, {}, {}


_REGISTRY = {}


class Dataset:
    """Subset of netCDF4.Dataset used by the reference (see module docstring)."""

    def __init__(self, path, mode="r", **kw):
        path = os.path.abspath(path)
        if mode == "w":
            _REGISTRY[path] = _Store()
        elif path not in _REGISTRY:
            raise FileNotFoundError(path)
        object.__setattr__(self, "_s", _REGISTRY[path])
        object.__setattr__(self, "_path", path)

    @property
    def dimensions(self):
        return self._s.dimensions

    @property
    def variables(self):
        return self._s.variables

    def createDimension(self, name, size):
        self._s.dimensions[name] = _Dim(name, size)
        return self._s.dimensions[name]

    def createVariable(self, name, dtype, dims=(), **kw):
        if isinstance(dims, str):
            dims = (dims,)
        dims = tuple(dims)
        shape = tuple(len(self._s.dimensions[d]) for d in dims)
        v = _Var(name, dtype, dims, shape)
        self._s.variables[name] = v
        return v

    def ncattrs(self):
        return list(self._s.attrs.keys())

    def __setattr__(self, k, v):
        self._s.attrs[k] = v

    def __getattr__(self, k):
        try:
            return object.__getattribute__(self, "_s").attrs[k]
        except KeyError:
            raise AttributeError(k)

    def close(self):
        pass


def registry():
    return _REGISTRY


def install_stubs():
    if "netCDF4" not in sys.modules:
        m = types.ModuleType("netCDF4")
        m.Dataset = Dataset
        sys.modules["netCDF4"] = m
    if "matplotlib" not in sys.modules:
        mpl = types.ModuleType("matplotlib")
        plt = types.ModuleType("matplotlib.pyplot")
        mpl.pyplot = plt
        sys.modules["matplotlib"] = mpl
        sys.modules["matplotlib.pyplot"] = plt


_MODS = {}


def reference():
    """Import the reference modules (cached).  Returns a namespace with the modules as attributes."""
    if not have_reference():
        raise RuntimeError("reference tree not available at %s" % REF)
    if _MODS:
        return types.SimpleNamespace(**_MODS)
    install_stubs()
    for p in (os.path.join(REF, "src", "pymiecoated"), os.path.join(REF, "src", "geosmie")):
        if p not in sys.path:
            sys.path.append(p)
    import importlib

    for name in ("pymiecoated", "pymiecoated.mie_coated", "pymiecoated.mie_coeffs", "pymiecoated.mie_props",
                 "particleparams", "dointegration", "hydrophobic", "bandaverage"):
        _MODS[name.replace(".", "_")] = importlib.import_module(name)
    return types.SimpleNamespace(**_MODS)


@contextlib.contextmanager
def reference_cwd(extra_json=None):
    """chdir into a scratch dir laid out like the reference's flattened run directory
    (src/scripts/geosmie_setup.py:110-136): data/ and geosparticles/ next to the drivers."""
    old = os.getcwd()
    with tempfile.TemporaryDirectory() as d:
        os.symlink(os.path.join(REF, "src", "geosmie", "data"), os.path.join(d, "data"))
        os.symlink(os.path.join(REF, "src", "config", "geosparticles"), os.path.join(d, "geosparticles"))
        for name, text in (extra_json or {}).items():
            with open(os.path.join(d, name), "w") as fp:
                fp.write(text)
        os.chdir(d)
        try:
            yield d
        finally:
            os.chdir(old)
