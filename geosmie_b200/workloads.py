"""Benchmark / test workloads: the BASELINE species tables (optics_SU / SS / BC) as in-memory configurations.

The numeric parameters are those of src/config/geosparticles/{su,ss,bc}.json; the refractive-index tables (OPAC
suso00 / sscm00 / soot00, HITRAN water) come from a fixture recorded from the reference's own readers
(tests/golden/hostlogic.npz: `<sp>__mlist`, `<sp>__water`), because the reference tree does not travel to the GPU box.
"""
import json
import os

import numpy as np
from scipy.interpolate import interp1d

RH36 = [0.00, 0.05, 0.10, 0.15, 0.20, 0.25, 0.30, 0.35, 0.40, 0.45, 0.50, 0.55, 0.60, 0.65, 0.70, 0.75, 0.80, 0.81, 0.82, 0.83,
        0.84, 0.85, 0.86, 0.87, 0.88, 0.89, 0.90, 0.91, 0.92, 0.93, 0.94, 0.95, 0.96, 0.97, 0.98, 0.99]

SPECIES = {
    "su": {"rhop0": 1700.0, "rh": RH36,
           "rhDep": {"type": "simple", "params": {"gf": [1.00, 1.04, 1.08, 1.12, 1.16, 1.20, 1.23, 1.27, 1.31, 1.35, 1.39, 1.43, 1.46,
                                                         1.50, 1.54, 1.59, 1.64, 1.65, 1.66, 1.67, 1.68, 1.69, 1.71, 1.72, 1.74, 1.75,
                                                         1.77, 1.79, 1.82, 1.84, 1.87, 1.91, 1.94, 1.99, 2.05, 2.16]}},
           "psd": {"type": "lognorm", "params": {"r0": [[0.0695e-6]], "rmin0": [[0.005e-6]], "rmax0": [[0.3e-6]], "sigma": [[2.03]],
                                                 "numperdec": [1000], "fracs": [[1.0]]}}},
    "ss": {"rhop0": 2200.0, "rh": RH36, "maxrh": 0.95,
           "rhDep": {"type": "ss", "params": {"c1": 0.7674, "c2": 3.079, "c3": 2.573e-11, "c4": -1.424}},
           "psd": {"type": "ss", "params": {"rMinMaj": [0.03e-6, 0.1e-6, 0.5e-6, 1.5e-6, 5.0e-6],
                                            "rMaxMaj": [0.1e-6, 0.5e-6, 1.5e-6, 5.0e-6, 10.0e-6],
                                            "fracs": [[1.0], [1.0], [1.0], [1.0], [1.0]], "numperdec": [1600] * 5}}},
    "bc": {"rhop0": 1000.0, "rh": RH36,
           "rhDep": {"type": "simple", "params": {"gf": [1.00, 1.00, 1.00, 1.00, 1.00, 1.00, 1.00, 1.00, 1.00, 1.00, 1.00, 1.00, 1.01,
                                                         1.01, 1.03, 1.10, 1.19, 1.21, 1.23, 1.25, 1.27, 1.30, 1.32, 1.34, 1.36, 1.38,
                                                         1.41, 1.43, 1.46, 1.48, 1.52, 1.55, 1.59, 1.65, 1.72, 1.89]}},
           "psd": {"type": "lognorm", "params": {"r0": [[0.0118e-6]], "rmax0": [[0.3e-6]], "rmin0": [[1e-10]], "sigma": [[2.0]],
                                                 "numperdec": [100], "fracs": [[1.0]]}}, "hydrophobic": True},
}

DEFAULT_FIXTURE = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "hostlogic.npz")


def species_inputs(sp, fixture=DEFAULT_FIXTURE):
    """(params, lambarr, part_m, water_m, rh_used) for dointegration.BinPlan, as dointegration.fun derives them."""
    g = np.load(fixture)
    ml, water = g[sp + "__mlist"], g[sp + "__water"]
    params = json.loads(json.dumps(SPECIES[sp]))
    part_m = [(interp1d(ml[0], ml[1]), interp1d(ml[0], ml[2]))]
    water_m = (interp1d(water[0], water[1]), interp1d(water[0], water[2]))
    rh_used = params["rh"]
    if "maxrh" in params:
        rh_used = np.array(rh_used)
        rh_used[np.where(rh_used > params["maxrh"])[0]] = params["maxrh"]
    return params, ml[0], part_m, water_m, rh_used


def bin_plan(sp, radind=0, cells=None, fixture=DEFAULT_FIXTURE, device_psd=False):
    from . import dointegration as DI
    params, lambarr, part_m, water_m, rh_used = species_inputs(sp, fixture)
    return DI.BinPlan(params, radind, lambarr, rh_used, part_m, water_m, cells=cells, device_psd=device_psd)


def n_bins(sp):
    from . import dointegration as DI
    return len(DI.bins_of(SPECIES[sp])[0])


def species_files(sp, fixture=DEFAULT_FIXTURE):
    """Files of a run directory that reproduce a shipped species config (su / ss / bc): <sp>.json with the parameters of
    src/config/geosparticles/<sp>.json, its refractive-index table re-written in the 'wsv' format of particleparams.py:70-77, and
    data/refrac.water.txt (13 header lines, columns wavelength [um] n k: particleparams.py:79-82).  {relative path: text}."""
    g = np.load(fixture)
    ml, w = g[sp + "__mlist"], g["su__water"]

    def rows(t):
        out = []
        for i in range(t.shape[1]):
            um = float("%.10g" % (t[0, i] * 1e6))       # the table's own decimal value, so that um * 1e-6 is bit-exact
            assert um * 1e-6 == t[0, i]
            out.append("%.17g %.17g %.17g" % (um, t[1, i], t[2, i]))
        return "\n".join(out) + "\n"

    cfg = json.loads(json.dumps(SPECIES[sp]))
    cfg["ri"] = {"format": "wsv", "path": ["ri-%s.wsv" % sp]}
    return {sp + ".json": json.dumps(cfg), "ri-%s.wsv" % sp: rows(ml), os.path.join("data", "refrac.water.txt"): "# header\n" * 13 + rows(w)}


def write_run_dir(d, sp, fixture=DEFAULT_FIXTURE):
    """Write species_files(sp) below directory d (the reference's CWD-relative layout).  Returns the config file name."""
    for name, text in species_files(sp, fixture).items():
        path = os.path.join(d, name)
        os.makedirs(os.path.dirname(path), exist_ok=True)
        with open(path, "w") as fp:
            fp.write(text)
    return sp + ".json"


def fine_grid_files(sp, nlam=2048, fixture=DEFAULT_FIXTURE):
    """BASELINE config 5: the species table on a FINE spectral grid -- its refractive-index spectrum resampled to `nlam` log-spaced
    wavelengths by linear interpolation (what interp1d would return).  {relative path: text} for a run directory and the
    name of the configuration file."""
    g = np.load(fixture)
    ml, w = g[sp + "__mlist"], g["su__water"]
    lam = np.geomspace(ml[0][0], ml[0][-1], nlam)
    lam[0], lam[-1] = ml[0][0], ml[0][-1]
    n, k = np.interp(lam, ml[0], ml[1]), np.interp(lam, ml[0], ml[2])
    cfg = json.loads(json.dumps(SPECIES[sp]))
    cfg.pop("hydrophobic", None)
    cfg["ri"] = {"format": "wsv", "path": ["ri-%s-fine.wsv" % sp]}
    files = species_files(sp, fixture)
    files = {os.path.join("data", "refrac.water.txt"): files[os.path.join("data", "refrac.water.txt")],
             sp + "_fine.json": json.dumps(cfg),
             "ri-%s-fine.wsv" % sp: "\n".join("%.17g %.17g %.17g" % (l * 1e6, a, b) for l, a, b in zip(lam, n, k)) + "\n"}
    return files, sp + "_fine.json"
