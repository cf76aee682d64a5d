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


# ---- classic-format constants (NetCDF-3, CDF-2 = 64-bit offsets)
_NC_DIMENSION, _NC_VARIABLE, _NC_ATTRIBUTE = 10, 11, 12
_NC_TYPES = {1: ">i1", 2: "S1", 3: ">i2", 4: ">i4", 5: ">f4", 6: ">f8"}
_NC_CODE = {"i1": 1, "S1": 2, "i2": 3, "i4": 4, "f4": 5, "f8": 6}


def _pad4(n):
    return (n + 3) // 4 * 4


class _Var(object):
    """Variable of the in-memory dataset.  The array is created on first use: a variable that was read from a file stays a
    (path, offset) reference until somebody looks at it -- rungsf touches 7 of the 36 variables of a table, the other 29 are copied
    from file to file as raw bytes -- and `var[:] = array` adopts a matching array instead of copying it into a zero-filled one."""

    def __init__(self, name, dtype, dims, shape, source=None):
        object.__setattr__(self, "_attrs", {})
        object.__setattr__(self, "name", name)
        object.__setattr__(self, "dimensions", tuple(dims))
        dt = np.dtype("S1") if dtype in ("c", "S1") else np.dtype(dtype)
        object.__setattr__(self, "dtype", dt)
        object.__setattr__(self, "_shape", tuple(shape))
        object.__setattr__(self, "_data", None)
        object.__setattr__(self, "_source", source)      # (path, offset, big-endian dtype string) of a variable read from a file
        object.__setattr__(self, "_dirty", False)

    shape = property(lambda self: self._shape)

    @property
    def data(self):
        if self._data is None:
            if self._source is not None:
                path, off, be = self._source
                n = int(np.prod(self._shape)) if self._shape else 1
                raw = np.fromfile(path, dtype=be, count=n, offset=off)
                arr = (raw.astype(self.dtype) if self.dtype.kind != "S" else raw).reshape(self._shape)
                arr.flags.writeable = False      # reading does not make the variable dirty; writes go through __setitem__ (copy on write)
                object.__setattr__(self, "_data", arr)
            else:
                object.__setattr__(self, "_data", np.zeros(self._shape, dtype=self.dtype))
        return self._data

    def read_into(self, dst):
        """Decode the variable straight into `dst` (same shape, any strides, native byte order): one pass from the file's pages, no
        intermediate native copy.  The variable itself stays un-decoded (and is still copied file to file as raw bytes)."""
        if self._data is not None or self._source is None or self.dtype.kind == "S":
            np.copyto(dst, self.data)
            return
        path, off, be = self._source
        mm = np.memmap(path, dtype=be, mode="r", offset=off, shape=self._shape)
        np.copyto(dst, mm)
        del mm

    def untouched_source(self):
        """(path, offset, type) of the bytes in the file this variable was read from, as long as nothing has been assigned to it."""
        return None if self._dirty else self._source

    def __getitem__(self, k):
        return self.data[k]

    def __setitem__(self, k, v):
        if self.dtype.kind == "S" and isinstance(v, str):
            v = v.encode()
        object.__setattr__(self, "_dirty", True)
        if self._data is not None and not self._data.flags.writeable:
            object.__setattr__(self, "_data", self._data.copy())
        if (isinstance(k, slice) and k == slice(None) and isinstance(v, np.ndarray) and v.shape == self._shape and v.dtype == self.dtype
                and self.dtype.kind != "S"):
            object.__setattr__(self, "_data", v)           # adopt: the caller's array is the variable (nothing here writes into it later)
            return
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

    def renameVariable(self, old, new):
        self.variables = {(new if k == old else k): v for k, v in self.variables.items()}
        object.__setattr__(self.variables[new], "name", new)

    def ncattrs(self):
        return list(self._gattrs.keys())

    def __setattr__(self, k, v):
        if k == "variables":
            object.__setattr__(self, k, v)
        else:
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

    # -- classic-format (de)serialisation: own reader / writer (the file is one header + the variables back to back, big-endian);
    #    scipy.io.netcdf_file reads what is written here and vice versa (tests/test_host_logic.py)
    @staticmethod
    def _att_bytes(attrs):
        import struct
        if not attrs:
            return struct.pack(">ii", 0, 0)
        out = [struct.pack(">ii", _NC_ATTRIBUTE, len(attrs))]
        for k, v in attrs.items():
            kb = k.encode()
            out.append(struct.pack(">i", len(kb)) + kb.ljust(_pad4(len(kb)), b"\0"))
            if isinstance(v, (str, bytes)):
                vb = v.encode() if isinstance(v, str) else v
                out.append(struct.pack(">ii", 2, len(vb)) + vb.ljust(_pad4(len(vb)), b"\0"))
            else:
                arr = np.atleast_1d(np.asarray(v))
                if arr.dtype.kind in "iu" and arr.dtype.itemsize > 4:
                    arr = arr.astype("i4")
                if arr.dtype.kind == "b":
                    arr = arr.astype("i1")
                code = _NC_CODE[arr.dtype.newbyteorder("=").str[1:]]
                vb = arr.astype(_NC_TYPES[code]).tobytes()
                out.append(struct.pack(">ii", code, arr.size) + vb.ljust(_pad4(len(vb)), b"\0"))
        return b"".join(out)

    def _store(self):
        import struct
        dimid = {name: i for i, name in enumerate(self.dimensions)}
        hdr = [b"CDF\x02", struct.pack(">i", 0)]
        if self.dimensions:
            hdr.append(struct.pack(">ii", _NC_DIMENSION, len(self.dimensions)))
            for name, d in self.dimensions.items():
                nb = name.encode()
                hdr.append(struct.pack(">i", len(nb)) + nb.ljust(_pad4(len(nb)), b"\0") + struct.pack(">i", len(d)))
        else:
            hdr.append(struct.pack(">ii", 0, 0))
        hdr.append(self._att_bytes(self._gattrs))
        metas = []
        for name, v in self.variables.items():
            dt = v.dtype
            if dt == np.dtype("i8"):
                dt = np.dtype("i4")                         # the classic format has no 64-bit integers
            key = "S1" if dt.kind == "S" else dt.newbyteorder("=").str[1:]
            code = _NC_CODE[key]
            n = int(np.prod(v.shape)) if v.shape else 1
            nbytes = n * np.dtype(_NC_TYPES[code]).itemsize
            nb = name.encode()
            head = (struct.pack(">i", len(nb)) + nb.ljust(_pad4(len(nb)), b"\0") + struct.pack(">i", len(v.dimensions)) +
                    b"".join(struct.pack(">i", dimid[d]) for d in v.dimensions) + self._att_bytes(v._attrs) + struct.pack(">i", code) +
                    struct.pack(">I", min(_pad4(nbytes), 0xFFFFFFFF)))
            metas.append((v, code, nbytes, head))
        varlist = struct.pack(">ii", _NC_VARIABLE, len(metas)) if metas else struct.pack(">ii", 0, 0)
        fixed = sum(len(x) for x in hdr) + len(varlist) + sum(len(m[3]) + 8 for m in metas)
        off = fixed
        tmp = self._path + ".tmp%d" % os.getpid()
        with open(tmp, "wb") as f:
            f.write(b"".join(hdr) + varlist)
            begins = []
            for v, code, nbytes, head in metas:
                f.write(head + struct.pack(">q", off))
                begins.append(off)
                off += _pad4(nbytes)
            for (v, code, nbytes, head), begin in zip(metas, begins):
                assert f.tell() == begin
                src = v.untouched_source()
                if src is not None and src[2] == _NC_TYPES[code]:
                    with open(src[0], "rb") as g:           # a variable nobody looked at: raw bytes from file to file
                        left, spos = nbytes, src[1]
                        if hasattr(os, "copy_file_range") and not os.environ.get("GEOSMIE_NO_COPY_FILE_RANGE"):   # inside the kernel, no bounce through user space
                            f.flush()
                            try:
                                while left:
                                    n = os.copy_file_range(g.fileno(), f.fileno(), left, spos, begin + nbytes - left)
                                    if n <= 0:
                                        break
                                    left -= n
                                    spos += n
                            except OSError:
                                pass
                            f.seek(begin + nbytes - left)
                        g.seek(spos)
                        while left:
                            chunk = g.read(min(left, 1 << 24))
                            if not chunk:
                                raise IOError("%s: unexpected end of file while copying variable %r" % (src[0], v.name))
                            f.write(chunk)
                            left -= len(chunk)
                else:
                    arr = np.ascontiguousarray(v.data)
                    if arr.dtype.kind == "S":
                        f.write(arr.tobytes())
                    else:
                        # byte-swapping pass in cache-sized pieces through one small buffer (no second full-size array)
                        flat = arr.reshape(-1)
                        step = 1 << 19
                        buf = np.empty(min(step, max(flat.size, 1)), dtype=_NC_TYPES[code])
                        for i in range(0, flat.size, step):
                            m = min(step, flat.size - i)
                            buf[:m] = flat[i:i + m]
                            f.write(memoryview(buf[:m]).cast("B"))
                if _pad4(nbytes) != nbytes:
                    f.write(b"\0" * (_pad4(nbytes) - nbytes))
        os.replace(tmp, self._path)

    def _load(self):
        import struct
        if not os.path.exists(self._path):
            raise FileNotFoundError(self._path)
        with open(self._path, "rb") as f:
            buf = f.read(1 << 20)
            if buf[:3] != b"CDF" or buf[3] not in (1, 2):
                raise ValueError("%s is not a classic-format NetCDF file" % self._path)
            big = buf[3] == 2
            pos = [8]

            def need(n):
                nonlocal buf
                while pos[0] + n > len(buf):
                    more = f.read(1 << 20)
                    if not more:
                        raise ValueError("truncated NetCDF header in %s" % self._path)
                    buf += more

            def i4():
                need(4)
                v = struct.unpack_from(">i", buf, pos[0])[0]
                pos[0] += 4
                return v

            def name():
                n = i4()
                need(_pad4(n))
                s_ = buf[pos[0]:pos[0] + n].decode()
                pos[0] += _pad4(n)
                return s_

            def atts():
                tag, n = i4(), i4()
                out = {}
                for _ in range(n if tag == _NC_ATTRIBUTE else 0):
                    k = name()
                    code, cnt = i4(), i4()
                    dt = np.dtype(_NC_TYPES[code])
                    nb = cnt * dt.itemsize
                    need(_pad4(nb))
                    raw = buf[pos[0]:pos[0] + nb]
                    pos[0] += _pad4(nb)
                    if code == 2:
                        out[k] = raw.decode()
                    else:
                        a_ = np.frombuffer(raw, dtype=dt).astype(dt.newbyteorder("="))
                        out[k] = a_[0] if a_.size == 1 else a_
                return out

            tag, n = i4(), i4()
            dims = []
            for _ in range(n if tag == _NC_DIMENSION else 0):
                nm = name()
                dims.append((nm, i4()))
                self.dimensions[nm] = _Dim(nm, dims[-1][1])
            self._gattrs.update(atts())
            tag, n = i4(), i4()
            for _ in range(n if tag == _NC_VARIABLE else 0):
                nm = name()
                nd = i4()
                dn = [dims[i4()][0] for _ in range(nd)]
                va = atts()
                code = i4()
                i4()                                        # vsize
                need(8)
                if big:
                    begin = struct.unpack_from(">q", buf, pos[0])[0]
                    pos[0] += 8
                else:
                    begin = i4()
                be = _NC_TYPES[code]
                dt = np.dtype(be).newbyteorder("=") if code != 2 else "c"
                v = _Var(nm, dt, dn, tuple(self.dimensions[d].size for d in dn), source=(self._path, begin, be))
                v._attrs.update(va)
                self.variables[nm] = v
        if self._mode in ("r+", "a"):
            # the file is rewritten on close: keep reading lazily from a name that survives the replace
            pass


def Dataset(path, mode="r", **kw):
    """netCDF4.Dataset when available, the in-memory classic-format shim otherwise."""
    if HAVE_NETCDF4:
        return _nc4.Dataset(path, mode, **kw)
    return _MemDataset(path, mode, **kw)


def open_copy(infile, outfile):
    """A dataset that starts as a copy of `infile` and is written to `outfile` on close (what `shutil.copyfile` + Dataset(outfile, 'r+')
    do in the reference, convertncdf.py:331-332).  With the classic-format shim nothing is copied up front: variables that are never
    assigned to travel from `infile` to `outfile` as raw bytes when the dataset is closed."""
    if HAVE_NETCDF4 or os.path.abspath(infile) == os.path.abspath(outfile):
        if os.path.abspath(infile) != os.path.abspath(outfile):
            import shutil
            shutil.copyfile(infile, outfile)
        return Dataset(outfile, "r+")
    ds = _MemDataset(infile, "r")
    object.__setattr__(ds, "_path", outfile)
    object.__setattr__(ds, "_mode", "r+")
    return ds
