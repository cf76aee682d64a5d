#!/usr/bin/env python3
"""Full-size golden vectors from the UNMODIFIED reference (run in this container, ~1 h of one core):

  full_su.npz / full_bc.npz   dointegration.fun on the shipped su.json / bc.json (all 61 x 36 cells, 4459 / 615 sizes):
                              every scalar variable for every cell, pback, and the six phase-matrix elements for a
                              stratified subset of cells (the complete phase matrices would be 39 MB per table);
  full_ss_cells.npz           rawMie + integratePSD + the post-processing of fun for 6 cells of each of the 5 sea-salt
                              bins of ss.json (the whole SS table takes 6-15 h on one core).
    python tests/golden/make_golden_full.py [su] [bc] [ss]
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import refharness as rh  # noqa: E402

R = rh.reference()
DI = R.dointegration
PHASE_CELLS = [(li, rhi) for li in (0, 7, 15, 22, 30, 38, 45, 53, 60) for rhi in (0, 10, 20, 26, 31, 35)]


def full_table(sp):
    with rh.reference_cwd() as d:
        DI.fun("geosparticles/%s.json" % sp, "json", d, False)
        store = rh.registry()[os.path.join(d, "optics_%s.nomom.nc4" % sp)]
        out = {}
        for k, v in store.variables.items():
            a = np.array(v.data)
            if a.ndim == 4 and a.shape[-1] == 371:
                out["phase__" + k] = np.stack([a[:, li, rhi, :] for li, rhi in PHASE_CELLS], axis=1)
            else:
                out["var__" + k] = a
        out["phase_cells"] = np.array(PHASE_CELLS)
    np.savez_compressed(os.path.join(HERE, "full_%s.npz" % sp), **out)
    print("wrote full_%s" % sp)


def ss_cells():
    from scipy.interpolate import interp1d
    MultipleMie = R.pymiecoated_mie_coated.MultipleMie
    with rh.reference_cwd() as d:
        params = R.particleparams.getParticleParams("geosparticles/ss.json", "json")
        water = R.particleparams.getWaterM()
    ml = params["mList"][0]
    lam_all = ml[0]
    rh_used = np.array(params["rh"])
    rh_used[rh_used > params["maxrh"]] = params["maxrh"]
    ang = np.concatenate([np.linspace(0., 1., 100, endpoint=False), np.linspace(1., 10., 100, endpoint=False),
                          np.linspace(10., 180., 171, endpoint=True)])
    cost = np.cos(np.radians(ang))
    out = {"cells": []}
    cells = [(0, 0), (0, 31), (12, 16), (25, 26), (40, 35), (60, 20)]
    for b in range(5):
        xx, dr = DI.initializeXarr(params, b, lam_all[0], lam_all[-1])
        mm = MultipleMie(xx, None, cost)
        mm.preCalculate()
        for (li, rhi) in cells:
            lam = lam_all[li]
            nref0 = [complex(ml[1][li], -ml[2][li])]
            nw = complex(float(interp1d(water[0], water[1])(lam)), float(interp1d(water[0], water[2])(lam)))
            _, _, _, rrat0 = DI.getHumidRefractiveIndex(params, b, 0, rh_used, nref0, nw)
            _, reff_mass0, _, _ = DI.calculatePSD(params, b, 0., rh_used, xx, dr, rrat0, lam)
            mr, mi, gf, rrat = DI.getHumidRefractiveIndex(params, b, rhi, rh_used, nref0, nw)
            psd, ref, rlow, rup = DI.calculatePSD(params, b, rh_used[rhi], rh_used, xx, dr, rrat, lam)
            rhop = rrat ** 3. * params["rhop0"] + (1. - rrat ** 3.) * 1000.
            raw = DI.rawMie(mm, DI.scatkeys, DI.scalarkeys, lam, mr[0], mi[0], None, cost)
            ret = DI.integratePSD(mm.xArr, [raw], psd, params["psd"]["params"]["fracs"][b], lam, reff_mass0, params["rhop0"], rhop)
            key = "b%d_l%d_r%d" % (b, li, rhi)
            for k, v in ret.items():
                out[key + "__" + k] = np.array(v)
            print("ss", key, float(ret["qext"]), flush=True)
    np.savez_compressed(os.path.join(HERE, "full_ss_cells.npz"), cells=np.array(cells), **{k: v for k, v in out.items() if k != "cells"})
    print("wrote full_ss_cells")


if __name__ == "__main__":
    which = sys.argv[1:] or ["bc", "ss", "su"]
    for w in which:
        if w == "ss":
            ss_cells()
        else:
            full_table(w)
