#!/usr/bin/env python3
"""Complete optics_SS golden table from the UNMODIFIED reference, farmed over processes (run in this container).

Every one of the 5 x 61 x 36 = 10 980 cells of src/config/geosparticles/ss.json is evaluated DENSELY (all 5608-6244
grid points of the bin, like dointegration.fun does) with the reference's own MultipleMie.preCalculate + rawMie +
integratePSD + the post-processing statements of fun (dointegration.py:950-1001), one (bin, wavelength) column of 36
RH cells per work item.  A work item writes tests/golden/_ss_parts/b<bin>_l<lam>.npz, so the run can be interrupted
and resumed; `merge` packs the parts into full_ss.npz: every scalar variable for every cell, pback, and the six
phase-matrix elements for a stratified subset of cells (the complete phase matrices would be 195 MB).

    python tests/golden/make_golden_ss_full.py run [nproc]      # ~7 h of one core in total
    python tests/golden/make_golden_ss_full.py merge
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
PARTS = os.path.join(HERE, "_ss_parts")
PHASE_LAM = (0, 7, 15, 22, 30, 38, 45, 53, 60)
PHASE_RH = (0, 10, 20, 26, 31, 35)
SCAL = ("qext", "qsca", "qabs", "qb", "g", "csca", "cext", "bsca", "bext", "bbck", "lidar_ratio", "area", "volume", "mass",
        "rEff", "rMass")
_STATE = {}


def _setup():
    if _STATE:
        return _STATE
    os.environ.setdefault("OMP_NUM_THREADS", "1")
    import refharness as rh
    R = rh.reference()
    from scipy.interpolate import interp1d
    with rh.reference_cwd():
        params = R.particleparams.getParticleParams("geosparticles/ss.json", "json")
        water = R.particleparams.getWaterM()
    ang = np.concatenate([np.linspace(0., 1., 100, endpoint=False), np.linspace(1., 10., 100, endpoint=False),
                          np.linspace(10., 180., 171, endpoint=True)])
    _STATE.update(R=R, params=params, water=water, ang=ang, cost=np.cos(np.radians(ang)), interp1d=interp1d, mm={})
    return _STATE


def _bin_mie(b):
    s = _setup()
    if b not in s["mm"]:
        DI = s["R"].dointegration
        lam_all = s["params"]["mList"][0][0]
        xx, dr = DI.initializeXarr(s["params"], b, lam_all[0], lam_all[-1])
        mm = s["R"].pymiecoated_mie_coated.MultipleMie(xx, None, s["cost"])
        mm.preCalculate()
        s["mm"] = {b: (mm, xx, dr)}          # keep one bin's tables per process
    return s["mm"][b]


def column(item):
    """All 36 RH cells of (bin b, wavelength index li), statements of dointegration.fun:811-1001 in order."""
    b, li = item
    path = os.path.join(PARTS, "b%d_l%02d.npz" % (b, li))
    if os.path.exists(path):
        return path
    s = _setup()
    DI, params, water, interp1d = s["R"].dointegration, s["params"], s["water"], s["interp1d"]
    mm, xx, dr = _bin_mie(b)
    ml = params["mList"]
    lam = ml[0][0][li]
    # statements of dointegration.fun:707-710, 813-832 (interp1d at the table's own wavelengths, sign flip, RH cap)
    mr0 = [interp1d(ml[i][0], ml[i][1])(lam) for i in range(len(ml))]
    mi0 = [-interp1d(ml[i][0], ml[i][2])(lam) for i in range(len(ml))]
    nref0 = [complex(mr0[i], mi0[i]) for i in range(len(mr0))]
    nw = complex(interp1d(water[0], water[1])(lam), interp1d(water[0], water[2])(lam))
    rh_used = np.array(params["rh"])
    rh_used[np.where(rh_used > params["maxrh"])[0]] = params["maxrh"]
    nrh = len(rh_used)
    out = {k: np.zeros(nrh) for k in SCAL + ("ssa", "refreal", "refimag", "growth_factor", "rhop", "rLow", "rUp")}
    out["pback"] = np.zeros((nrh, 6))
    phase = {}
    rhop0 = params["rhop0"]
    theta = np.radians(s["ang"])
    _, _, _, rrat0 = DI.getHumidRefractiveIndex(params, b, 0, rh_used, nref0, nw)
    _, reff_mass0, _, _ = DI.calculatePSD(params, b, 0., rh_used, xx, dr, rrat0, lam)
    mass0 = None
    for rhi, onerh in enumerate(rh_used):
        mr, mi, gf, rrat = DI.getHumidRefractiveIndex(params, b, rhi, rh_used, nref0, nw)
        psd, ref, rlow, rup = DI.calculatePSD(params, b, onerh, rh_used, xx, dr, rrat, lam)
        rhop = rrat ** 3. * rhop0 + (1. - rrat ** 3.) * 1000.
        raw = DI.rawMie(mm, DI.scatkeys, DI.scalarkeys, lam, mr[0], mi[0], None, s["cost"])
        allret = [raw for i in range(len(psd))]
        ret = DI.integratePSD(mm.xArr, allret, psd, params["psd"]["params"]["fracs"][b], lam, reff_mass0, rhop0, rhop)
        # post-processing of fun (dointegration.py:950-1001), the same statements
        qsca = np.array(ret["qsca"])
        qext = np.array(ret["qext"])
        qb = np.array(ret["qb"])
        ret["lidar_ratio"] = qext / qb * 4 * np.pi
        ret["ssa"] = qsca / qext
        p11n = 2. * ret["p11"] / np.trapz(ret["p11"] * np.sin(theta), theta)
        for k in ("p12", "p22", "p33", "p34", "p44"):
            ret[k] = ret[k] * p11n / ret["p11"]
        ret["p11"] = p11n
        ret["pback"] = np.array([ret[k][-1] for k in ("p11", "p12", "p33", "p34", "p22", "p44")])
        if rhi == 0.0:
            mass0 = ret["volume"] * rhop0
        ret["area"] = ret["area"] / mass0
        ret["volume"] = ret["volume"] / mass0
        for k in SCAL + ("ssa",):
            out[k][rhi] = ret[k]
        out["pback"][rhi] = ret["pback"]
        out["refreal"][rhi] = mr[0]
        out["refimag"][rhi] = -np.abs(mi[0])
        out["growth_factor"][rhi] = gf
        out["rhop"][rhi] = rhop
        out["rLow"][rhi] = rlow
        out["rUp"][rhi] = rup
        if li in PHASE_LAM and rhi in PHASE_RH:
            for k in ("p11", "p12", "p22", "p33", "p34", "p44"):
                phase["%s_r%d" % (k, rhi)] = np.array(ret[k])
    os.makedirs(PARTS, exist_ok=True)
    np.savez(path + ".tmp.npz", **out, **phase)
    os.replace(path + ".tmp.npz", path)
    return path


def run(nproc):
    import multiprocessing as mp
    # largest bins first, and consecutive items of a worker share the bin (tables are rebuilt per bin change)
    items = [(b, li) for b in (4, 3, 2, 1, 0) for li in range(61)]
    items = [it for it in items if not os.path.exists(os.path.join(PARTS, "b%d_l%02d.npz" % it))]
    print("%d columns to do on %d processes" % (len(items), nproc), flush=True)
    with mp.get_context("fork").Pool(nproc) as pool:
        for n, p in enumerate(pool.imap_unordered(column, items, chunksize=1)):
            print(n + 1, os.path.basename(p), flush=True)


def merge():
    vals = {}
    keys = SCAL + ("ssa", "refreal", "refimag", "growth_factor", "rhop", "rLow", "rUp")
    NL = ("mass", "volume", "area", "rEff", "rMass", "rUp", "rLow", "rhop", "growth_factor")   # (bin, rh) in the file
    for k in keys:
        vals["var__" + k] = np.zeros((5, 61, 36))
    vals["var__pback"] = np.zeros((5, 61, 36, 6))
    cells = [(li, rhi) for li in PHASE_LAM for rhi in PHASE_RH]
    for k in ("p11", "p12", "p22", "p33", "p34", "p44"):
        vals["phase__" + k] = np.zeros((5, len(cells), 371))
    for b in range(5):
        for li in range(61):
            d = np.load(os.path.join(PARTS, "b%d_l%02d.npz" % (b, li)))
            for k in keys:
                vals["var__" + k][b, li] = d[k]
            vals["var__pback"][b, li] = d["pback"]
            for ci, (cl, cr) in enumerate(cells):
                if cl == li:
                    for k in ("p11", "p12", "p22", "p33", "p34", "p44"):
                        vals["phase__" + k][b, ci] = d["%s_r%d" % (k, cr)]
    for k in NL:
        vals["perlam__" + k] = vals["var__" + k]              # what every wavelength pass computed
        vals["var__" + k] = vals["var__" + k][:, -1, :].copy()  # what survives in the file: the last wavelength's write
    st = _setup()
    vals["var__wavelength"] = np.asarray(st["params"]["mList"][0][0], dtype=float)
    vals["var__rh"] = np.asarray(st["params"]["rh"], dtype=float)          # the file's coordinate is the un-capped list
    vals["phase_cells"] = np.array(cells)
    np.savez_compressed(os.path.join(HERE, "full_ss.npz"), **vals)
    print("wrote full_ss.npz")


if __name__ == "__main__":
    if sys.argv[1] == "run":
        run(int(sys.argv[2]) if len(sys.argv) > 2 else max(1, (os.cpu_count() or 2) - 2))
    else:
        merge()
