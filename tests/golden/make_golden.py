#!/usr/bin/env python3
"""Generate the golden fixtures under tests/golden/ by RUNNING THE UNMODIFIED REFERENCE in this container.

    python tests/golden/make_golden.py

Reads /root/reference (never copied into the repo), writes small .npz/.json fixtures next to this file.
The reference cannot travel to the GPU box, the fixtures can.  Every fixture records the inputs it was
made from, so the tests re-feed the same inputs to the oracle (CPU) and to the CUDA path (GPU).

Reference entry points exercised (paths relative to /root/reference):
  src/pymiecoated/pymiecoated/mie_coated.py   Mie, MultipleMie.preCalculate / calculateS12SizeRange
  src/geosmie/dointegration.py                initializeXarr, calculatePSD, getHumidRefractiveIndex, rawMie,
                                              integratePSD, fun
  src/geosmie/hydrophobic.py                  doConversion
  src/geosmie/bandaverage.py                  doAverage, getBands
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import refharness as rh  # noqa: E402

R = rh.reference()
Mie = R.pymiecoated.Mie
MultipleMie = R.pymiecoated_mie_coated.MultipleMie
DI = R.dointegration


def table_angles():
    # dointegration.py:739-742
    ang1 = np.linspace(0., 1., 100, endpoint=False)
    ang2 = np.linspace(1., 10., 100, endpoint=False)
    ang3 = np.linspace(10., 180., 171, endpoint=True)
    return np.concatenate([ang1, ang2, ang3])


# --------------------------------------------------------------------------- mini species configs
RI_SU = "0.30 1.469 -1.0e-8\n0.55 1.430 -1.0e-8\n1.00 1.422 -1.53e-6\n10.0 1.89 -0.455\n"
RI_SS = "0.30 1.510 -2.0e-6\n0.55 1.500 -1.0e-8\n3.00 1.610 -1.0e-2\n"
RI_BC = "0.35 1.75 -0.465\n0.55 1.75 -0.44\n2.00 1.80 -0.50\n"
RI_DU = "0.40 1.53 -0.0085\n0.55 1.53 -0.0055\n1.00 1.53 -0.0043\n"
RI_B2 = "0.35 1.45 -0.01\n0.55 1.44 -0.005\n2.00 1.40 -0.001\n"

MINI = {
    "su_mini": ({
        "rhop0": 1700.0, "rh": [0.0, 0.5, 0.9, 0.99],
        "rhDep": {"type": "simple", "params": {"gf": [1.0, 1.39, 1.77, 2.16]}},
        "psd": {"type": "lognorm", "params": {"r0": [[0.0695e-6]], "rmin0": [[0.005e-6]], "rmax0": [[0.3e-6]],
                                              "sigma": [[2.03]], "numperdec": [40], "fracs": [[1.0]]}},
        "ri": {"format": "wsv", "path": ["ri-su_mini.wsv"]}}, {"ri-su_mini.wsv": RI_SU}),
    "ss_mini": ({
        "rhop0": 2200.0, "rh": [0.0, 0.8, 0.95, 0.99], "maxrh": 0.95,
        "rhDep": {"type": "ss", "params": {"c1": 0.7674, "c2": 3.079, "c3": 2.573e-11, "c4": -1.424}},
        "psd": {"type": "ss", "params": {"rMinMaj": [0.1e-6, 1.5e-6], "rMaxMaj": [0.5e-6, 5.0e-6],
                                         "fracs": [[1.0], [1.0]], "numperdec": [40, 30]}},
        "ri": {"format": "wsv", "path": ["ri-ss_mini.wsv"]}}, {"ri-ss_mini.wsv": RI_SS}),
    "bc_mini": ({
        "rhop0": 1000.0, "rh": [0.0, 0.7, 0.95],
        "rhDep": {"type": "simple", "params": {"gf": [1.0, 1.03, 1.55]}},
        "psd": {"type": "lognorm", "params": {"r0": [[0.0118e-6]], "rmax0": [[0.3e-6]], "rmin0": [[1e-10]],
                                              "sigma": [[2.0]], "numperdec": [25], "fracs": [[1.0]]}},
        "ri": {"format": "wsv", "path": ["ri-bc_mini.wsv"]}, "hydrophobic": True}, {"ri-bc_mini.wsv": RI_BC}),
    # two lognormal modes with two different refractive-index files, rhop0 per bin, two bins
    "mm_mini": ({
        "rhop0": [1500.0, 1800.0], "rh": [0.0, 0.9],
        "rhDep": {"type": "simple", "params": {"gf": [1.0, 1.5]}},
        "psd": {"type": "lognorm", "params": {"r0": [[0.05e-6, 0.2e-6], [0.4e-6, 0.5e-6]],
                                              "rmin0": [[0.005e-6, 0.02e-6], [0.05e-6, 0.05e-6]],
                                              "rmax0": [[0.3e-6, 0.8e-6], [1.0e-6, 1.5e-6]],
                                              "sigma": [[1.8, 1.6], [1.5, 1.7]], "numperdec": [30, 30],
                                              "fracs": [[0.7, 0.3], [0.4, 0.6]]}},
        "ri": {"format": "wsv", "path": ["ri-bc_mini.wsv", "ri-b2_mini.wsv"]}},
        {"ri-bc_mini.wsv": RI_BC, "ri-b2_mini.wsv": RI_B2}),
    # sulfate with CARMA growth (rhDep type 'su', particleparams.py:101-108 -> carma_utils.grow_v75)
    "sucarma_mini": ({
        "rhop0": 1923.0, "rh": [0.0, 0.3, 0.8, 0.95],
        "rhDep": {"type": "su", "params": {"temp": 220.0}},
        "psd": {"type": "lognorm", "params": {"r0": [[0.08e-6]], "rmin0": [[0.01e-6]], "rmax0": [[0.5e-6]],
                                              "sigma": [[1.6]], "numperdec": [30], "fracs": [[1.0]]}},
        "ri": {"format": "wsv", "path": ["ri-su_mini.wsv"]}}, {"ri-su_mini.wsv": RI_SU}),
}


def _ref_data(name):
    with open(os.path.join(rh.REF, "src", "geosmie", "data", name)) as fp:
        return fp.read()


def real_species():
    """Shipped species configs (src/config/geosparticles/{oc,brc,ni}.json) with their REAL refractive-index files (OPAC/GADS
    waso00 through the 'gads' reader, ri-brc.wsv, ri-nitrate.wsv: 61 wavelengths each), thinned in RH and grid density so that
    the reference finishes in minutes.  ni keeps its three bins with per-bin rhop0 and size parameters up to ~7500."""
    out = {}
    for sp, keep_rh, npd in (("oc", [0, 30], 30), ("brc", [0, 35], 30), ("ni", [0, 26], 12),
                             ("du-mie", [0, 20, 35], None),              # 'du' r^-4 sub-bin PSD on its fixed 1000-point grid, trivial rhDep
                             ("v2.0.1/ss.v2.0.1", [0, 24, 35], 16),      # Gong sea salt, 5 bins, maxrh, OPAC sscm00 (gads)
                             ("v2.0.1/bc.v5_7", [0, 33], 20),            # OPAC soot00 (gads), hydrophobic
                             ("experimental/su_ht_reff04_sig16_single", [0, 35], 40)):
        with open(os.path.join(rh.REF, "src", "config", "geosparticles", sp + ".json")) as fp:
            cfg = json.load(fp)
        cfg["rh"] = [cfg["rh"][i] for i in keep_rh]
        if "gf" in cfg["rhDep"]["params"] and len(cfg["rhDep"]["params"]["gf"]) > 1:
            cfg["rhDep"]["params"]["gf"] = [cfg["rhDep"]["params"]["gf"][i] for i in keep_rh]
        if npd is not None:
            cfg["psd"]["params"]["numperdec"] = [npd] * len(cfg["psd"]["params"]["numperdec"])
        files = {path: _ref_data(os.path.basename(path)) for path in cfg["ri"]["path"]}
        name = os.path.basename(sp).replace(".", "_").replace("-", "_")
        out[name + "_real"] = (cfg, files)
    return out


THIN = [0, 13, 29, 44, -1]     # wavelength indices kept for the angle-resolved variables of the real-species fixtures


def dump_dataset(store, thin=None):
    out = {}
    for k, v in store.variables.items():
        a = np.array(v.data)
        if thin is not None and a.ndim == 4 and "ang" in v.dimensions:
            a = a[:, thin]
        out["var__" + k] = a
        out["dims__" + k] = np.array("|".join(v.dimensions))
    for k, d in store.dimensions.items():
        out["dim__" + k] = np.array(len(d))
    return out


def gen_fun():
    """Full dointegration.fun (+ hydrophobic.doConversion) tables for the mini configs, new and legacy layout."""
    allcfg = dict(MINI)
    allcfg.update(real_species())
    for name, (cfg, files) in allcfg.items():
        if ONLY and name not in ONLY:
            continue
        thin = THIN if name.endswith("_real") else None
        extra = {name + ".json": json.dumps(cfg)}
        extra.update({k: v for k, v in files.items() if not k.startswith("data/")})   # data/ is the reference's own directory
        for classic in (False, True):
            if classic and name not in ("bc_mini", "su_mini"):
                continue
            with rh.reference_cwd(extra) as d:
                DI.fun(name + ".json", "json", d, classic)
                fn = "optics_%s.nomom%s.nc4" % (name, ".legacy" if classic else "")
                store = rh.registry()[os.path.join(d, fn)]
                out = dump_dataset(store, thin)
                if cfg.get("hydrophobic"):
                    # runoptics.py:113-121 renames the file first; the stub registry is keyed by path
                    rh.registry()[os.path.join(d, fn + ".nohp")] = store
                    R.hydrophobic.doConversion(fn + ".nohp", fn, d, classic)
                    hp = dump_dataset(rh.registry()[os.path.join(d, fn)], thin)
                    out.update({"hp__" + k: v for k, v in hp.items()})
            if thin is not None:
                out["thin_idx"] = np.array(thin)
            out["config_json"] = np.array(json.dumps(cfg))
            out["files_json"] = np.array(json.dumps(files))
            np.savez_compressed(os.path.join(HERE, "fun_%s%s.npz" % (name, "_legacy" if classic else "")), **out)
            print("wrote fun_%s%s" % (name, "_legacy" if classic else ""))


def gen_single():
    """Single-particle API: efficiencies + S12 for homogeneous, magnetic and coated spheres."""
    rng = np.random.default_rng(0)
    rows = []
    us = np.array([-1.0, -0.6, 0.0, 0.3, 0.999, 1.0])

    def record(kw):
        m = Mie(**kw)
        vals = [m.qext(), m.qsca(), m.qabs(), m.qb(), m.asy(), m.qratio()]
        s = []
        for u in us:
            s1, s2 = m.S12(float(u))
            s += [s1.real, s1.imag, s2.real, s2.imag]
        rows.append((kw, vals, s))

    # BASELINE config 1 + the three test_mie.py cases (test_mie.py:45-109)
    record(dict(x=10.0, m=1.53 + 0.008j))
    record(dict(x=2.5, m=1.5 + 0.5j))
    record(dict(x=1.5, y=5.0, m=1.5 + 0.5j, m2=1.2 + 0.2j))
    record(dict(x=4.0, eps=2.2 + 0.8j, mu=1.6 + 1.4j))
    # random homogeneous spheres, log-uniform x in [1e-2, 3e3] (SURVEY 8d seeds)
    for _ in range(60):
        x = float(10 ** rng.uniform(-2, np.log10(3e3)))
        n = float(rng.uniform(1.2, 2.0))
        k = float(10 ** rng.uniform(-9, 0))
        record(dict(x=x, m=complex(n, k)))
    # magnetic
    for _ in range(8):
        x = float(10 ** rng.uniform(-1, 2))
        record(dict(x=x, eps=complex(rng.uniform(1.5, 4), rng.uniform(0, 1)), mu=complex(rng.uniform(0.8, 2), rng.uniform(0, 1.5))))
    # coated
    for _ in range(30):
        y = float(10 ** rng.uniform(-1, 2.3))
        x = float(y * rng.uniform(0.05, 0.98))
        record(dict(x=x, y=y, m=complex(rng.uniform(1.5, 2.0), 10 ** rng.uniform(-3, -0.2)),
                    m2=complex(rng.uniform(1.2, 1.5), 10 ** rng.uniform(-8, -1))))
    keys = ["x", "y", "eps", "mu", "eps2"]
    par = np.full((len(rows), 8), np.nan)
    for i, (kw, _, _) in enumerate(rows):
        mm = Mie(**kw)
        par[i, 0] = mm.x
        par[i, 1] = np.nan if mm.y is None else mm.y
        par[i, 2:4] = [complex(mm.eps).real, complex(mm.eps).imag]
        par[i, 4:6] = [complex(mm.mu).real, complex(mm.mu).imag]
        if mm.eps2 is not None:
            par[i, 6:8] = [complex(mm.eps2).real, complex(mm.eps2).imag]
    np.savez_compressed(os.path.join(HERE, "mie_single.npz"), par=par, us=us,
                        q=np.array([r[1] for r in rows]), s12=np.array([r[2] for r in rows]).reshape(len(rows), len(us), 4),
                        par_cols=np.array("x,y,eps_re,eps_im,mu_re,mu_im,eps2_re,eps2_im"))
    print("wrote mie_single (%d cases)" % len(rows))


def gen_size_range():
    """MultipleMie.calculateS12SizeRange on small x-grids (the batch API, mie_coated.py:61-89)."""
    ang = table_angles()[::15]
    cost = np.cos(np.radians(ang))
    cases = []
    xs = [np.geomspace(2e-3, 40.0, 48), np.geomspace(0.5, 600.0, 24), np.array([1e-5, 1e-4, 1e-3, 0.01, 0.1, 1.0, 10.0])]
    ms = [(1.43, 1e-8), (1.75, 0.44), (1.33, 0.0), (1.95, 0.79)]
    out = {"ang": ang, "cost": cost, "ncase": np.array(len(xs) * len(ms))}
    ci = 0
    for x in xs:
        mm = MultipleMie(x, None, cost)
        mm.preCalculate()
        for (mr, mi) in ms:
            ret = mm.calculateS12SizeRange(mr, mi)
            s12 = np.array([[[s[0].real, s[0].imag, s[1].real, s[1].imag] for s in row] for row in ret["s12"]])
            q = np.array([ret[k] for k in ("qext", "qsca", "qabs", "qb", "asy", "qratio")]).T
            out["x_%d" % ci] = x
            out["m_%d" % ci] = np.array([mr, mi])
            out["q_%d" % ci] = q
            out["s12_%d" % ci] = s12
            ci += 1
    np.savez_compressed(os.path.join(HERE, "size_range.npz"), **out)
    print("wrote size_range (%d cases)" % ci)


def gen_cells():
    """rawMie + integratePSD (dointegration.py:1211-1254, :1064-1209) for hand-made cells incl. two modes."""
    ang = table_angles()
    cost = np.cos(np.radians(ang))
    rng = np.random.default_rng(1)
    x = np.geomspace(5e-3, 60.0, 96)
    mm = MultipleMie(x, None, cost)
    mm.preCalculate()
    lam = 0.55e-6
    out = {"ang": ang, "x": x, "lam": np.array(lam)}
    cases = [
        dict(m=[(1.43, 1e-8)], nmode=1, fracs=[1.0]),
        dict(m=[(1.75, 0.44)], nmode=2, fracs=[0.7, 0.3]),           # one RI replicated over two PSD modes
        dict(m=[(1.53, 0.006), (1.40, 0.1)], nmode=2, fracs=[0.25, 0.75]),  # one RI per mode
    ]
    for ci, c in enumerate(cases):
        psd = []
        for k in range(c["nmode"]):
            w = np.exp(-0.5 * ((np.log(x) - np.log(2.0 + 3 * k)) / 0.7) ** 2)
            w[x < 0.02] = 0.0
            w[x > 40.0] = 0.0
            psd.append(w / w.sum())
        allret = [DI.rawMie(mm, DI.scatkeys, DI.scalarkeys, lam, mr, mi, None, cost) for (mr, mi) in c["m"]]
        if len(allret) == 1:
            allret = [allret[0] for _ in range(c["nmode"])]
        reff0 = [0.3e-6 * (1 + k) for k in range(c["nmode"])]
        rhop0, rhop = 1700.0, 1300.0
        ret = DI.integratePSD(mm.xArr, allret, psd, c["fracs"], lam, reff0, rhop0, rhop)
        out["m_%d" % ci] = np.array(c["m"])
        out["fracs_%d" % ci] = np.array(c["fracs"])
        out["psd_%d" % ci] = np.array(psd)
        out["reff0_%d" % ci] = np.array(reff0)
        out["rhop_%d" % ci] = np.array([rhop0, rhop])
        for k, v in ret.items():
            out["ret_%d__%s" % (ci, k)] = np.array(v)
    out["ncase"] = np.array(len(cases))
    np.savez_compressed(os.path.join(HERE, "cells.npz"), **out)
    print("wrote cells")


def gen_hostlogic():
    """Grid / PSD / RH bookkeeping of the full BASELINE configs (su, ss, bc): sizes, nmax, weights."""
    out = {}
    for sp in ("su", "ss", "bc"):
        with rh.reference_cwd() as d:
            params = R.particleparams.getParticleParams("geosparticles/%s.json" % sp, "json")
            water = R.particleparams.getWaterM()
        lam_all = params["mList"][0][0]
        out[sp + "__lam"] = lam_all
        out[sp + "__mlist"] = np.array(params["mList"][0])
        out[sp + "__water"] = np.array(water)
        psdtype = params["psd"]["type"]
        nb = len(params["psd"]["params"]["r0"] if psdtype == "lognorm" else params["psd"]["params"]["rMinMaj"])
        rhl = np.array(params["rh"])
        if "maxrh" in params:
            rhl = rhl.copy()
            rhl[rhl > params["maxrh"]] = params["maxrh"]
        for b in range(nb):
            xx, dr = DI.initializeXarr(params, b, lam_all[0], lam_all[-1])
            out["%s__x_%d" % (sp, b)] = xx
            out["%s__dr_%d" % (sp, b)] = dr
            out["%s__nmax_%d" % (sp, b)] = np.round(2 + xx + 4 * xx ** (1.0 / 3.0)).astype(np.int32)
            # a few cells of PSD weights + humid refractive index
            for (li, rhi) in ((0, 0), (20, 16), (60, 35)):
                lam = lam_all[li]
                nref0 = [complex(params["mList"][0][1][li], -params["mList"][0][2][li])]
                nw = complex(np.interp(lam, water[0], water[1]), np.interp(lam, water[0], water[2]))
                mr, mi, gf, rrat = DI.getHumidRefractiveIndex(params, b, rhi, rhl, nref0, nw)
                psd, ref, rlow, rup = DI.calculatePSD(params, b, rhl[rhi], rhl, xx, dr, rrat, lam)
                key = "%s__cell_%d_%d_%d" % (sp, b, li, rhi)
                out[key + "__m"] = np.array([mr[0], mi[0], gf, rrat])
                out[key + "__psd"] = np.array(psd)
                out[key + "__ref"] = np.array(ref + [rlow, rup])
    np.savez_compressed(os.path.join(HERE, "hostlogic.npz"), **out)
    print("wrote hostlogic")


def gen_carma():
    """carma_utils.wtpct / dens / grow_v75 (carma_utils.py:136-314) on a sweep of (RH, dry radius, temperature)."""
    import carma_utils as cu
    rng = np.random.default_rng(4)
    rh = np.concatenate([[1e-7, 0.01, 0.049, 0.05, 0.5, 0.85, 0.851, 0.99, 1.0], rng.uniform(0, 1, 60)])
    rd = 10 ** rng.uniform(-8.5, -5.5, rh.size)
    temp = np.concatenate([np.full(9, 220.0), rng.uniform(190, 260, 60)])
    out = np.array([[cu.wtpct(a, temp=t), cu.dens(a, temp=t), float(cu.grow_v75(a, r, temp=t))] for a, r, t in zip(rh, rd, temp)])
    np.savez_compressed(os.path.join(HERE, "carma_growth.npz"), rh=rh, rd=rd, temp=temp, out=out)
    print("wrote carma_growth")


def gen_bands():
    """bandaverage.doAverage / getBands (bandaverage.py:18-50, :73-124) on a synthetic spectrum."""
    BA = R.bandaverage
    rng = np.random.default_rng(2)
    lam = np.geomspace(0.25e-6, 40e-6, 61)
    v = rng.uniform(0.1, 3.0, size=(5, 61))
    out = {"lam": lam, "v": v}
    for mode in ("GEOS5", "RRTMG", "RRTMGP", "PURDUE"):
        lo, up, mean, usewn, nb = BA.getBands(mode)
        res = np.zeros((5, len(lo)))
        for i in range(5):
            for b in range(len(lo)):
                res[i, b] = BA.doAverage(lam, v[i], lo[b], up[b], usewn, None)
        out[mode + "__lo"] = np.array(lo, dtype=float)
        out[mode + "__up"] = np.array(up, dtype=float)
        out[mode + "__mean"] = np.array(mean, dtype=float)
        out[mode + "__usewn"] = np.array(usewn)
        out[mode + "__nb"] = np.array(nb)
        out[mode + "__avg"] = res
    np.savez_compressed(os.path.join(HERE, "bands.npz"), **out)
    print("wrote bands")


ONLY = [a[4:] for a in sys.argv[1:] if a.startswith("fun:")]

if __name__ == "__main__":
    if ONLY:
        gen_fun()
        sys.exit(0)
    which = sys.argv[1:] or ["single", "size_range", "cells", "hostlogic", "bands", "carma", "fun"]
    for w in which:
        globals()["gen_" + w]()
