"""ctypes binding of libgeosmie_b200.so (C ABI: include/geosmie_b200.h).  No CPU fallback -- fails loudly."""
import ctypes as C
import os
import threading

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("GEOSMIE_LIB", os.path.join(_HERE, "libgeosmie_b200.so"))

GM_NSCAL = 11
S_W, S_X2W, S_X3W, S_X4W, S_QEXT, S_QSCA, S_QABS, S_QB, S_G, S_CSCA, S_CEXT = range(11)
F_ELIDE_ZERO_WEIGHT = 1
F_NO_GRAM = 2
F_PHASE_ON_DEVICE = 4
PSD_LOGNORM, PSD_SS, PSD_DU, PSD_NPAR = 1, 2, 3, 4

_lib = None
_lock = threading.Lock()

c_dp = C.POINTER(C.c_double)
c_ip = C.POINTER(C.c_int32)
c_lp = C.POINTER(C.c_int64)
vp = C.c_void_p

# every symbol include/geosmie_b200.h declares: name -> (restype, argtypes)
SIGNATURES = {
    "gm_version": (C.c_int, []),
    "gm_last_error": (C.c_char_p, []),
    "gm_init": (C.c_int, [C.c_int, C.POINTER(vp)]),
    "gm_destroy": (C.c_int, [vp]),
    "gm_set_stream": (C.c_int, [vp, vp]),
    "gm_sync": (C.c_int, [vp]),
    "gm_launch_count": (C.c_int64, [vp]),
    "gm_mie_eval": (C.c_int, [vp, C.c_int, vp, vp, vp, vp, C.c_int, vp, vp, vp, vp, C.c_int, vp, vp, vp, vp]),
    "gm_table_create": (C.c_int, [vp, C.c_int, vp, vp, C.c_int, vp, C.POINTER(vp)]),
    "gm_table_destroy": (C.c_int, [vp]),
    "gm_table_nx": (C.c_int, [vp]),
    "gm_table_nang": (C.c_int, [vp]),
    "gm_table_set_bessel": (C.c_int, [vp, vp, vp, vp]),
    "gm_table_run": (C.c_int, [vp, C.c_int, vp, vp, C.c_int, vp, vp, C.c_int, vp, vp]),
    "gm_table_run_dev": (C.c_int, [vp, C.c_int, vp, vp, C.c_int, vp, vp, C.c_int, vp, vp]),
    "gm_table_run_coated": (C.c_int, [vp, C.c_int, vp, vp, vp, C.c_int, vp, vp, C.c_int, vp, vp]),
    "gm_table_set_gsf": (C.c_int, [vp, vp, C.c_int, C.c_int, vp, vp]),
    "gm_table_gsf_device": (C.c_int, [vp, C.POINTER(vp), C.POINTER(vp)]),
    "gm_table_set_dr": (C.c_int, [vp, vp]),
    "gm_table_run_psd": (C.c_int, [vp, C.c_int, vp, vp, C.c_int, C.c_int, vp, vp, C.c_int, vp, vp]),
    "gm_table_get_weights": (C.c_int, [vp, C.c_int, C.c_int, vp]),
    "gm_table_device_outputs": (C.c_int, [vp, C.POINTER(vp), C.POINTER(vp)]),
    "gm_table_particles": (C.c_int, [vp, C.c_int, vp, vp, vp, vp]),
    "gm_table_last_stats": (C.c_int, [vp, vp]),
    "gm_table_set_timing": (C.c_int, [vp, C.c_int]),
    "gm_table_fetch_normalized": (C.c_int, [vp, C.c_int, vp, vp, vp, vp, vp, vp, vp]),
    "gm_table_last_kernel_ms": (C.c_int, [vp, c_dp, c_dp, c_dp]),
    "gm_table_last_kernel_ms_ex": (C.c_int, [vp, vp, vp]),
    "gm_gsf_expand": (C.c_int, [vp, C.c_int, C.c_int, vp, vp, C.c_int, vp, vp, C.c_int]),
    "gm_gsf_expand_dev": (C.c_int, [vp, C.c_int, C.c_int, vp, vp, C.c_int, vp, vp, C.c_int]),
    "gm_gsf_expand_phase4_dev": (C.c_int, [vp, C.c_int, C.c_int, vp, vp, C.c_int, vp, vp, C.c_int]),
    "gm_gsf_diagnose": (C.c_int, [vp, C.c_int, C.c_int, vp, vp, C.c_int, vp, vp, C.c_int, vp, vp]),
    "gm_gsf_diagnose_dev": (C.c_int, [vp, C.c_int, C.c_int, vp, vp, C.c_int, C.c_int, vp, vp, C.c_int, vp, vp]),
    "gm_band_average": (C.c_int, [vp, C.c_int, C.c_int, vp, vp, C.c_int, vp, vp, C.c_int, vp]),
    "gm_peer_alloc": (C.c_int, [vp, C.c_size_t, C.POINTER(vp), vp]),
    "gm_peer_free": (C.c_int, [vp, vp]),
    "gm_peer_open": (C.c_int, [vp, vp, C.POINTER(vp)]),
    "gm_peer_close": (C.c_int, [vp, vp]),
    "gm_peer_put": (C.c_int, [vp, vp, vp, C.c_size_t]),
    "gm_peer_get2d": (C.c_int, [vp, vp, C.c_size_t, vp, C.c_size_t, C.c_size_t, C.c_size_t]),
    "gm_table_normalize_device": (C.c_int, [vp, C.c_int, vp, vp, C.POINTER(vp)]),
    "gm_peer_join": (C.c_int, [vp]),
    "gm_peer_sync": (C.c_int, [vp]),
    "gm_peer_mark": (C.c_int, [vp, C.c_int]),
    "gm_peer_wait": (C.c_int, [vp, C.c_int]),
    "gm_table_set_mirror": (C.c_int, [vp, vp, vp]),
}
IPC_HANDLE_BYTES = 64


class GeosmieError(RuntimeError):
    pass


def load():
    """dlopen the library and bind every symbol of the header.  Works without a GPU (no compute is run)."""
    global _lib
    with _lock:
        if _lib is not None:
            return _lib
        path = os.environ.get("GEOSMIE_B200_LIB", LIB_PATH)   # another build of the same library (kernel tuning experiments)
        if not os.path.exists(path):
            raise GeosmieError("%s not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` or "
                               "`make -C geosmie_b200/csrc` -- there is no CPU fallback" % path)
        lib = C.CDLL(path)
        for name, (res, args) in SIGNATURES.items():
            try:
                fn = getattr(lib, name)
            except AttributeError:
                if path != LIB_PATH:
                    continue          # an experiment build made before the entry point existed; the in-tree library must have them all
                raise
            fn.restype = res
            fn.argtypes = args
        _lib = lib
        return lib


def check(rc):
    if rc != 0:
        raise GeosmieError("libgeosmie_b200 error %d: %s" % (rc, load().gm_last_error().decode()))


def f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def i64(a):
    return np.ascontiguousarray(a, dtype=np.int64)


def ptr(a):
    return None if a is None else a.ctypes.data_as(vp)


class Handle:
    """One per GPU.  `Handle.get(device)` returns a cached instance."""
    _cache = {}

    def __init__(self, device=0):
        self.lib = load()
        h = vp()
        check(self.lib.gm_init(int(device), C.byref(h)))
        self.h = h
        self.device = device

    @classmethod
    def get(cls, device=None):
        if device is None:
            device = int(os.environ.get("GEOSMIE_DEVICE", os.environ.get("LOCAL_RANK", "0")))
        if device not in cls._cache:
            cls._cache[device] = cls(device)
        return cls._cache[device]

    def set_stream(self, stream_ptr):
        check(self.lib.gm_set_stream(self.h, vp(stream_ptr)))

    def sync(self):
        check(self.lib.gm_sync(self.h))

    def launch_count(self):
        return int(self.lib.gm_launch_count(self.h))

    # ---- multi-GPU exchange over peer memory (gm_peer.cu) ------------------------------------------------------------
    def peer_alloc(self, nbytes):
        """-> (device pointer, 64-byte CUDA IPC handle) of a new exchange buffer on this GPU."""
        p = vp()
        hd = C.create_string_buffer(IPC_HANDLE_BYTES)
        check(self.lib.gm_peer_alloc(self.h, int(nbytes), C.byref(p), C.cast(hd, vp)))
        return p.value, hd.raw

    def peer_free(self, dptr):
        check(self.lib.gm_peer_free(self.h, vp(dptr)))

    def peer_open(self, ipc_handle):
        p = vp()
        hd = C.create_string_buffer(bytes(ipc_handle), IPC_HANDLE_BYTES)
        check(self.lib.gm_peer_open(self.h, C.cast(hd, vp), C.byref(p)))
        return p.value

    def peer_close(self, dptr):
        check(self.lib.gm_peer_close(self.h, vp(dptr)))

    def peer_put(self, dst_ptr, src_ptr, nbytes):
        """Copy-engine transfer on the exchange stream, ordered after the compute stream's current work."""
        check(self.lib.gm_peer_put(self.h, vp(dst_ptr), vp(src_ptr), int(nbytes)))

    def peer_get2d(self, dst_host_ptr, dpitch, src_dev_ptr, spitch, width, height):
        """Strided device / peer -> host copy on the exchange stream (see gm_peer_get2d)."""
        check(self.lib.gm_peer_get2d(self.h, vp(dst_host_ptr), int(dpitch), vp(src_dev_ptr), int(spitch), int(width), int(height)))

    def peer_join(self):
        check(self.lib.gm_peer_join(self.h))

    def peer_sync(self):
        check(self.lib.gm_peer_sync(self.h))

    def peer_mark(self, idx):
        """Remember 'all puts issued so far' under mark idx (0..3)."""
        check(self.lib.gm_peer_mark(self.h, int(idx)))

    def peer_wait(self, idx):
        """The compute stream waits for mark idx (no-op if it was never recorded)."""
        check(self.lib.gm_peer_wait(self.h, int(idx)))

    # ---- per-particle Mie ------------------------------------------------------------------------------------------
    def mie_eval(self, x, mz, mrel, nmax, u=None, xcore=None, ajv=None, ayv=None, want_s12=True, want_ab=False):
        x = f64(np.atleast_1d(x))
        n = x.size
        nmax = i32(np.atleast_1d(nmax))
        mz = np.ascontiguousarray(np.atleast_1d(mz), dtype=np.complex128)
        mrel = np.ascontiguousarray(np.atleast_1d(mrel), dtype=np.complex128)
        stride = 1 if mz.size > 1 else 0
        if stride and (mz.size != n or mrel.size != n):
            raise ValueError("per-particle materials must have one entry per particle")
        q = np.empty((n, 6))
        nang = 0
        uu = None
        s12 = None
        if u is not None and want_s12:
            uu = f64(np.atleast_1d(u))
            nang = uu.size
            s12 = np.empty((n, nang, 4))
        ab = np.empty((int(nmax.sum()), 4)) if want_ab else None
        xc = f64(np.atleast_1d(xcore)) if xcore is not None else None
        off = jv = yv = None
        if ajv is not None and ayv is not None:
            jv, yv = f64(ajv), f64(ayv)
            off = i64(np.concatenate([[0], np.cumsum(nmax)[:-1]]))
        check(self.lib.gm_mie_eval(self.h, n, ptr(x), ptr(xc), ptr(mz), ptr(mrel), stride, ptr(nmax), ptr(off), ptr(jv), ptr(yv),
                                   nang, ptr(uu), ptr(q), ptr(s12), ptr(ab)))
        return q, s12, ab

    # ---- GSF / bands -----------------------------------------------------------------------------------------------
    def gsf_expand(self, ang_deg, F, ng=129, quantize10=False):
        ang = f64(ang_deg)
        F = f64(F)
        ncell = F.shape[0]
        assert F.shape[1:] == (6, ang.size)
        coef = np.empty((ncell, 6, ng))
        cnorm = np.empty(ncell)
        check(self.lib.gm_gsf_expand(self.h, ncell, ang.size, ptr(ang), ptr(F), ng, ptr(coef), ptr(cnorm), int(bool(quantize10))))
        return coef, cnorm

    def gsf_diagnose(self, ang_deg, F, ng=129, quantize10=False):
        """gsf_expand + the diagnostics of spher_expan.f: -> (coef, cnorm, fout [ncell][6][nang] = the .expan_matr columns,
        fiterr [ncell])."""
        ang = f64(ang_deg)
        F = f64(F)
        ncell = F.shape[0]
        assert F.shape[1:] == (6, ang.size)
        coef = np.empty((ncell, 6, ng))
        cnorm = np.empty(ncell)
        fout = np.empty((ncell, 6, ang.size))
        fiterr = np.empty(ncell)
        check(self.lib.gm_gsf_diagnose(self.h, ncell, ang.size, ptr(ang), ptr(F), ng, ptr(coef), ptr(cnorm), int(bool(quantize10)),
                                       ptr(fout), ptr(fiterr)))
        return coef, cnorm, fout, fiterr

    def gsf_expand_phase4_dev(self, ang_deg, ncell, p4_ptr, coef_ptr, cnorm_ptr, ng=129, quantize10=False):
        """Device-pointer GSF expansion of gm_table_run's out_phase (asynchronous except for the small table upload)."""
        ang = f64(ang_deg)
        check(self.lib.gm_gsf_expand_phase4_dev(self.h, ncell, ang.size, ptr(ang), vp(p4_ptr), ng, vp(coef_ptr),
                                                vp(cnorm_ptr) if cnorm_ptr else None, int(bool(quantize10))))

    def band_average(self, lam, v, lo, hi, use_wavenum):
        lam, v, lo, hi = f64(lam), f64(v), f64(lo), f64(hi)
        ncol = v.shape[0]
        out = np.empty((ncol, lo.size))
        check(self.lib.gm_band_average(self.h, ncol, lam.size, ptr(lam), ptr(v), lo.size, ptr(lo), ptr(hi), int(bool(use_wavenum)), ptr(out)))
        return out


class Table:
    """Per-bin device object (x grid, nmax, Riccati-Bessel and pi/tau tables); wraps gm_table_*."""

    def __init__(self, x, nmax, cos_theta, handle=None):
        self.handle = handle or Handle.get()
        self.lib = self.handle.lib
        self.x = f64(x)
        self.nmax = i32(nmax)
        self.cost = f64(cos_theta)
        t = vp()
        check(self.lib.gm_table_create(self.handle.h, self.x.size, ptr(self.x), ptr(self.nmax), self.cost.size, ptr(self.cost), C.byref(t)))
        self.t = t
        self.nx = self.x.size
        self.nang = self.cost.size
        # GM_F_NO_GRAM: force the per-angle contraction for every particle group (ablation / cross-check of the Gram path)
        self.no_gram = os.environ.get("GEOSMIE_NO_GRAM", "0") not in ("", "0")


    def _flags(self, elide):
        return (F_ELIDE_ZERO_WEIGHT if elide else 0) | (F_NO_GRAM if self.no_gram else 0)

    def close(self):
        if getattr(self, "t", None):
            self.lib.gm_table_destroy(self.t)
            self.t = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_bessel(self, jv_half, yv_half):
        off = i64(np.concatenate([[0], np.cumsum(self.nmax + 1)[:-1]]))
        check(self.lib.gm_table_set_bessel(self.t, ptr(off), ptr(f64(jv_half)), ptr(f64(yv_half))))

    def run(self, mz, mrel, w_phase, w_scal=None, elide=False):
        """Host-buffer call.  Returns (scal [ntask][nmode][11], phase [ntask][4][nang])."""
        mz = np.ascontiguousarray(np.atleast_1d(mz), dtype=np.complex128)
        mrel = np.ascontiguousarray(np.atleast_1d(mrel), dtype=np.complex128)
        ntask = mz.size
        wp = f64(w_phase).reshape(ntask, self.nx)
        nmode = 1
        ws = None
        if w_scal is not None:
            ws = f64(w_scal)
            nmode = ws.shape[1]
            assert ws.shape == (ntask, nmode, self.nx)
        scal = np.empty((ntask, nmode, GM_NSCAL))
        phase = np.empty((ntask, 4, self.nang))
        if wp.min() < 0.0:
            # Signed number weights (the reference's non-monotonic 'du' grid makes getDR negative in places): the device folds
            # sqrt(w_phase) into the coefficients, so w_phase must be >= 0 there (gm_table_run answers GM_EINVAL otherwise).  The
            # sums are linear in w: phase = phase(w+) - phase(w-); the scalar sums take the signed per-mode weights as they are.
            neg = wp < 0.0
            wpos = np.where(neg, 0.0, wp)
            check(self.lib.gm_table_run(self.t, ntask, ptr(mz), ptr(mrel), nmode, ptr(wpos), ptr(ws if ws is not None else wp),
                                        self._flags(elide), ptr(scal), ptr(phase)))
            wneg = np.where(neg, -wp, 0.0)
            scal_n, phase_n = np.empty((ntask, 1, GM_NSCAL)), np.empty((ntask, 4, self.nang))
            check(self.lib.gm_table_run(self.t, ntask, ptr(mz), ptr(mrel), 1, ptr(wneg), None, self._flags(True), ptr(scal_n), ptr(phase_n)))
            return scal, phase - phase_n
        check(self.lib.gm_table_run(self.t, ntask, ptr(mz), ptr(mrel), nmode, ptr(wp), ptr(ws), self._flags(elide),
                                    ptr(scal), ptr(phase)))
        return scal, phase

    def run_coated(self, m1, m2, core_ratio, w_phase, w_scal=None, elide=False):
        """Coated spheres on the table's grid (= shell size parameter).  m1/m2 [ntask] complex indices, core_ratio [ntask]."""
        m1 = np.ascontiguousarray(np.atleast_1d(m1), dtype=np.complex128)
        m2 = np.ascontiguousarray(np.atleast_1d(m2), dtype=np.complex128)
        ntask = m1.size
        ratio = f64(np.broadcast_to(core_ratio, (ntask,)))
        wp = f64(w_phase).reshape(ntask, self.nx)
        nmode, ws = 1, None
        if w_scal is not None:
            ws = f64(w_scal)
            nmode = ws.shape[1]
        scal = np.empty((ntask, nmode, GM_NSCAL))
        phase = np.empty((ntask, 4, self.nang))
        check(self.lib.gm_table_run_coated(self.t, ntask, ptr(m1), ptr(m2), ptr(ratio), nmode, ptr(wp), ptr(ws),
                                           self._flags(elide), ptr(scal), ptr(phase)))
        return scal, phase

    def set_gsf(self, ang_deg, ng=129, quantize10=False, coef_out=None, cnorm_out=None):
        """Fuse the GSF moment expansion into the following run calls; coef_out [ntask][6][ng] (ideally pinned) receives
        the moments.  set_gsf(None) switches the stage off."""
        if ang_deg is None:
            check(self.lib.gm_table_set_gsf(self.t, None, 0, 0, None, None))
            self._gsf_keep = None
            return
        ang = f64(ang_deg)
        assert ang.size == self.nang
        self._gsf_keep = (ang, coef_out, cnorm_out)
        check(self.lib.gm_table_set_gsf(self.t, ptr(ang), int(ng), int(bool(quantize10)), ptr(coef_out), ptr(cnorm_out)))

    def set_mirror(self, scal_ptr, phase_ptr):
        """k_finalize also stores its results through these (peer) device pointers; (None, None) switches it off."""
        check(self.lib.gm_table_set_mirror(self.t, vp(scal_ptr) if scal_ptr else None, vp(phase_ptr) if phase_ptr else None))

    def set_dr(self, dr):
        check(self.lib.gm_table_set_dr(self.t, ptr(f64(dr))))

    def run_psd(self, mz, mrel, kind, params, frac, elide=False, out=None, phase_on_device=False):
        """Weights generated on the device from per-(task, mode) parameters.  params [ntask][nmode][4], frac [ntask][nmode].
        `out` = (scal, phase) caller-provided (e.g. pinned) arrays.  phase_on_device=True: the raw phase sums are not downloaded
        (returns (scal, None)); fetch_normalized() delivers the normalised phase matrix instead."""
        mz = np.ascontiguousarray(np.atleast_1d(mz), dtype=np.complex128)
        mrel = np.ascontiguousarray(np.atleast_1d(mrel), dtype=np.complex128)
        ntask = mz.size
        params = f64(params)
        nmode = params.shape[1]
        assert params.shape == (ntask, nmode, PSD_NPAR)
        frac = f64(np.broadcast_to(frac, (ntask, nmode)))
        if phase_on_device:
            scal = out[0] if out is not None else np.empty((ntask, nmode, GM_NSCAL))
            check(self.lib.gm_table_run_psd(self.t, ntask, ptr(mz), ptr(mrel), nmode, int(kind), ptr(params), ptr(frac),
                                            self._flags(elide) | F_PHASE_ON_DEVICE, ptr(scal), None))
            self._last_psd_shape = (ntask, nmode)
            self._last_ntask = ntask
            return scal, None
        scal, phase = out if out is not None else (np.empty((ntask, nmode, GM_NSCAL)), np.empty((ntask, 4, self.nang)))
        check(self.lib.gm_table_run_psd(self.t, ntask, ptr(mz), ptr(mrel), nmode, int(kind), ptr(params), ptr(frac),
                                        self._flags(elide), ptr(scal), ptr(phase)))
        self._last_psd_shape = (ntask, nmode)
        self._last_ntask = ntask
        return scal, phase

    def normalize_device(self, ang_deg):
        """k_phase_norm without the download: device pointer of the block [4][ntask][nang] (p11, p12, p33, p34) + pback4 [ntask][4]
        and its size in bytes (valid until the next call on this table)."""
        ntask = self._last_ntask
        theta = np.radians(f64(ang_deg))
        sint = np.sin(theta)
        blk = vp()
        check(self.lib.gm_table_normalize_device(self.t, ntask, ptr(theta), ptr(sint), C.byref(blk)))
        return blk.value, (4 * ntask * self.nang + 4 * ntask) * 8

    def fetch_pback(self, ang_deg):
        """Only the backscatter values [ntask][4] (p11, p12, p33, p34 at the last angle) of the normalised phase matrix."""
        ntask = self._last_ntask
        theta = np.radians(f64(ang_deg))
        sint = np.sin(theta)
        pback4 = np.empty((ntask, 4))
        check(self.lib.gm_table_fetch_normalized(self.t, ntask, ptr(theta), ptr(sint), None, None, None, None, ptr(pback4)))
        return pback4

    def fetch_normalized(self, ang_deg, out=None):
        """dointegration.py:977-988 on the device-resident phase sums of the last run call: returns (p11, p12, p33, p34) [ntask][nang]
        normalised to 2 / trapz(p11 sin(theta), theta) and pback4 [ntask][4].  `out` = four caller arrays (each C-contiguous,
        ntask * nang doubles: e.g. slices of the final table arrays) to write the planes into."""
        ntask = self._last_ntask
        theta = np.radians(f64(ang_deg))
        sint = np.sin(theta)
        planes = out if out is not None else [np.empty((ntask, self.nang)) for _ in range(4)]
        for a in planes:
            assert a.flags["C_CONTIGUOUS"] and a.dtype == np.float64 and a.size == ntask * self.nang
        pback4 = np.empty((ntask, 4))
        check(self.lib.gm_table_fetch_normalized(self.t, ntask, ptr(theta), ptr(sint), ptr(planes[0]), ptr(planes[1]), ptr(planes[2]),
                                                 ptr(planes[3]), ptr(pback4)))
        return planes, pback4

    def get_weights(self):
        ntask, nmode = self._last_psd_shape
        w = np.empty((ntask, nmode, self.nx))
        check(self.lib.gm_table_get_weights(self.t, ntask, nmode, ptr(w)))
        return w

    def run_into(self, ntask, mz, mrel, w_phase, scal_out, phase_out, elide=False):
        """Host-buffer call writing into caller-provided (ideally pinned) numpy arrays; single-mode weights."""
        check(self.lib.gm_table_run(self.t, ntask, ptr(mz), ptr(mrel), 1, ptr(w_phase), None, self._flags(elide),
                                    ptr(scal_out), ptr(phase_out)))

    def gsf_device(self):
        """Device pointers (coef [ntask][6][ng], cnorm [ntask]) of the fused GSF stage's results."""
        a, b = vp(), vp()
        check(self.lib.gm_table_gsf_device(self.t, C.byref(a), C.byref(b)))
        return a.value, b.value

    def device_outputs(self):
        a, b = vp(), vp()
        check(self.lib.gm_table_device_outputs(self.t, C.byref(a), C.byref(b)))
        return a.value, b.value

    def run_dev(self, ntask, mz_ptr, mrel_ptr, nmode, wphase_ptr, wscal_ptr, out_scal_ptr, out_phase_ptr, elide=False):
        """Device-pointer call (asynchronous on the handle's stream); pointers are integers (tensor.data_ptr())."""
        check(self.lib.gm_table_run_dev(self.t, ntask, vp(mz_ptr), vp(mrel_ptr), nmode, vp(wphase_ptr),
                                        vp(wscal_ptr) if wscal_ptr else None, self._flags(elide),
                                        vp(out_scal_ptr), vp(out_phase_ptr)))

    def particles(self, mz, mrel, want_s12=True):
        mz = np.ascontiguousarray(np.atleast_1d(mz), dtype=np.complex128)
        mrel = np.ascontiguousarray(np.atleast_1d(mrel), dtype=np.complex128)
        ntask = mz.size
        q = np.empty((ntask, self.nx, 6))
        s12 = np.empty((ntask, self.nx, self.nang, 4)) if want_s12 else None
        check(self.lib.gm_table_particles(self.t, ntask, ptr(mz), ptr(mrel), ptr(q), ptr(s12)))
        return q, s12

    def last_stats(self):
        s = np.zeros(8)
        check(self.lib.gm_table_last_stats(self.t, ptr(s)))
        return {"evals": s[0], "sum_nmax": s[1], "sum_nmx": s[2], "k4_groups": s[3], "launches": s[4]}

    def set_timing(self, on=True):
        check(self.lib.gm_table_set_timing(self.t, int(on)))

    def last_kernel_ms(self):
        a, b, c = C.c_double(), C.c_double(), C.c_double()
        check(self.lib.gm_table_last_kernel_ms(self.t, C.byref(a), C.byref(b), C.byref(c)))
        ms, n = np.zeros(8), np.zeros(8, dtype=np.int32)
        check(self.lib.gm_table_last_kernel_ms_ex(self.t, ptr(ms), ptr(n)))
        names = ("k_coeff", "k_contract", "k_finalize", "k_gram", "k_gram_sum_eval", "k_small")
        out = {"coeff": a.value, "contract": b.value, "finalize": c.value}
        out.update({k: float(v) for k, v in zip(names, ms)})
        out["launches"] = dict(zip(names, n.tolist()))
        return out
